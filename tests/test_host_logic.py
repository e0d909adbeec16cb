"""Host-side logic on CPU (`-m "not gpu"`): the autograd composition (incl. the GP
double-backward), the tflib registry, the step assembly of the three scripts and FlatAdam,
with the kernel launchers replaced by the TEST-ONLY torch stand-ins of tests/fake_backend.py.
The arithmetic of the real kernels is checked by the `-m gpu` tests."""
import numpy as np
import pytest
import torch

from tests import parity


@pytest.mark.parametrize('script,B', [('mnist', 4), ('cifar', 4), ('resnet', 4)])
def test_step_parity_fake_kernels_fp32(fake_kernels, script, B):
    tr, om = parity.build_pair(script, 'cpu', torch.float32, B)
    parity.perturb_params(tr, om)
    rep = parity.critic_parity(script, tr, om, parity.make_inputs(script, B, 11))
    v, k = parity.worst({k: v for k, v in rep.items() if not k.startswith('adam.')})
    assert v < 1e-3, 'critic: ' + parity.format_report(rep)
    # 'adam.*' compares the parameter UPDATE (~1e-4) recovered from fp32 parameters of size O(1): the
    # subtraction itself carries ~ulp(1)/1e-4 = 6e-4 of rounding, hence 2e-3 here
    assert parity.worst(rep, 'adam.')[0] < 2e-3, 'critic: ' + parity.format_report(rep)
    rep = parity.gen_parity(script, tr, om)
    # fp32 product vs fp64 oracle: activations differ by ~1e-6, which flips the ReLU mask of the odd
    # element sitting at zero; ONE flip among the ~1e6 activations of a generator layer moves a
    # norm-relative gradient error to ~1e-3 (1/sqrt(numel)).  Loss terms and the optimizer are held to
    # 1e-3; generator gradients to 1e-2 (a logic error shows up as >= 1e-1).
    grads = sorted(v for k, v in rep.items() if k.startswith('grad.'))
    assert parity.worst(rep, 'loss.')[0] < 1e-3 and parity.worst(rep, 'adam.')[0] < 2e-3, 'gen: ' + parity.format_report(rep)
    assert grads[-1] < 1e-2, 'gen: ' + parity.format_report(rep)


@pytest.mark.parametrize('script,B', [('mnist', 4), ('cifar', 4), ('resnet', 4)])
def test_step_parity_pattern_conditioned(fake_kernels, script, B):
    """With the device's activation patterns handed to the oracle (same linear region) the ReLU-tie
    caveat disappears and every gradient agrees to accumulation-order precision."""
    tr, om = parity.build_pair(script, 'cpu', torch.float32, B)
    parity.perturb_params(tr, om)
    rep = parity.critic_parity(script, tr, om, parity.make_inputs(script, B, 11), conditioned=True)
    assert parity.worst({k: v for k, v in rep.items() if not k.startswith('adam.')})[0] < 5e-4, parity.format_report(rep)
    rep = parity.gen_parity(script, tr, om, conditioned=True)
    assert parity.worst({k: v for k, v in rep.items() if not k.startswith('adam.')})[0] < 5e-4, parity.format_report(rep)


def test_second_step_uses_updated_weights(fake_kernels):
    """Two consecutive critic steps stay in parity (optimizer state + weight-cache invalidation)."""
    tr, om = parity.build_pair('cifar', 'cpu', torch.float32, 4)
    for it in range(2):
        rep = parity.critic_parity('cifar', tr, om, parity.make_inputs('cifar', 4, 20 + it), iteration=it)
        v, k = parity.worst(rep)
        assert v < 2e-3, 'step %d: %s' % (it, parity.format_report(rep))


def test_param_registry_semantics(fake_kernels):
    import ctgan_b200.tflib as lib
    a = lib.param('Discriminator.1.Filters', np.ones((1, 1, 2, 2), dtype='float32'))
    b = lib.param('Discriminator.1.Filters', np.zeros((1, 1, 2, 2), dtype='float32'))
    assert a is b and float(a.sum()) == 4.0                      # create-or-reuse
    lib.param('Generator.Input.W', np.zeros((2, 2), dtype='float32'))
    lib.param('Generator.BN.moving_mean', np.zeros(2, dtype='float32'), trainable=False)
    assert len(lib.params_with_name('Discriminator.')) == 1
    assert len(lib.params_with_name('Generator')) == 2            # includes the non-trainable stat
    assert list(lib.named_params_with_name('Generator')) == ['Generator.Input.W']
    lib.delete_all_params()
    assert lib.params_with_name('Generator') == []


def test_unsupported_options_raise(fake_kernels):
    from ctgan_b200.tflib.ops import conv2d, deconv2d, linear, cond_batchnorm
    x = torch.zeros(1, 4, 4, 4).contiguous(memory_format=torch.channels_last)
    with pytest.raises(Exception, match='Unsupported configuration'):
        conv2d.Conv2D('c', 4, 4, 3, x, mask_type=('a', 1))
    with pytest.raises(Exception, match='Unsupported configuration'):
        conv2d.Conv2D('c', 4, 4, 3, x, weightnorm=True)
    with pytest.raises(Exception, match='Unsupported configuration'):
        deconv2d.Deconv2D('d', 4, 4, 5, x, mask_type=('a', 1))
    with pytest.raises(Exception, match='unsupported'):
        cond_batchnorm.Batchnorm('b', [0], x, labels=torch.zeros(1, dtype=torch.int32), n_labels=10)
    with pytest.raises(Exception, match='Invalid initialization'):
        linear.Linear('l', 4, 4, torch.zeros(1, 4), initialization='nope')


def test_init_formulas_match_reference_statistics(fake_kernels):
    """He / Glorot uniform ranges of TG/tflib/ops/conv2d.py:55-86, deconv2d.py:41-67, linear.py:55-60."""
    import ctgan_b200.tflib as lib
    from ctgan_b200.tflib.ops import conv2d, deconv2d, linear
    np.random.seed(0)
    x = torch.zeros(1, 8, 4, 4).contiguous(memory_format=torch.channels_last)
    conv2d.Conv2D('A', 8, 16, 3, x, stride=1)
    conv2d.Conv2D('B', 8, 16, 5, x, stride=2, he_init=False)
    deconv2d.Deconv2D('C', 8, 16, 5, x)
    linear.Linear('D', 12, 20, torch.zeros(1, 12))
    def bound(name):
        return float(lib._params[name].abs().max())
    assert lib._params['A.Filters'].shape == (3, 3, 8, 16)
    assert lib._params['C.Filters'].shape == (5, 5, 16, 8)
    exp_a = np.sqrt(3) * np.sqrt(4. / (8 * 9 + 16 * 9))
    exp_b = np.sqrt(3) * np.sqrt(2. / (8 * 25 + 16 * 25 / 4.))
    exp_c = np.sqrt(3) * np.sqrt(4. / (8 * 25 / 4. + 16 * 25))
    exp_d = np.sqrt(3) * np.sqrt(2. / (12 + 20))
    for got, exp in [(bound('A.Filters'), exp_a), (bound('B.Filters'), exp_b), (bound('C.Filters'), exp_c), (bound('D.W'), exp_d)]:
        assert 0.9 * exp < got <= exp * (1 + 1e-6)
    assert float(lib._params['A.Biases'].abs().max()) == 0.0


@pytest.mark.parametrize('script,critic,gen', [('resnet', (23, 12), (11, 11)), ('cifar', (8, 4), (4, 4)), ('mnist', (8, 4), (4, 4))])
def test_no_discarded_parameter_gradient_launches(fake_kernels, script, critic, gen):
    """A step launches exactly the filter- / bias-gradient kernels whose results are applied: per critic step one per
    critic layer for the stacked pass + one filter gradient per layer for the gradient penalty's double backward (none in
    its FIRST backward, which only wants d/dx^); per generator step the generator's own (the critic is differentiated
    through, not updated).  (ResNet: 12 critic layers, 11 of them under the GP; 11 generator layers.)"""
    import collections
    import importlib
    import ctgan_b200.kernels as K
    counts = collections.Counter()
    for name in ('conv_wgrad', 'bias_grad'):
        def wrap(f=getattr(K, name), name=name):
            def g(*a, **k):
                counts[name] += 1
                return f(*a, **k)
            return g
        setattr(K, name, wrap())
    try:
        mod = importlib.import_module(parity.SCRIPTS[script][0])
        np.random.seed(0)
        tr = mod.Trainer(device='cpu', seed=1, act_dtype=torch.float32, batch_size=4)
        counts.clear()
        tr.critic_step(*parity.make_inputs(script, 4, 3))
        assert (counts['conv_wgrad'], counts['bias_grad']) == critic
        counts.clear()
        tr.gen_step()
        assert (counts['conv_wgrad'], counts['bias_grad']) == gen
    finally:
        from tests import fake_backend
        K.conv_wgrad, K.bias_grad = fake_backend.conv_wgrad, fake_backend.bias_grad     # the fixture restores the real ones


def test_tower_scopes_draw_the_slices_of_the_stacked_batch(fake_kernels):
    """DeviceRandom.scope_tower: site k of tower i reads part i of the Philox slice that the stacked pass (scope_parts) reserves
    for site k -- the two-branch generator step draws the same numbers as the one-batch generator step."""
    from ctgan_b200.runtime import DeviceRandom
    like_tower = [torch.empty(4, 8, 2, 2), torch.empty(4, 6)]
    like_stack = [torch.empty(8, 8, 2, 2), torch.empty(8, 6)]
    a = DeviceRandom(3, 'cpu')
    a.scope_parts([('drop.0', 4), ('drop.1', 4)])
    stacked = [a.dropout_stream(t)[1] for t in like_stack]
    b = DeviceRandom(3, 'cpu')
    got = []
    for i in range(2):
        b.scope_tower('drop', i, 2)
        got.append([b.dropout_stream(t)[1] for t in like_tower])
    assert got[0] == stacked
    assert got[1] == [o + t.numel() for o, t in zip(stacked, like_tower)]
    assert a.offset == b.offset
    b.scope('x')
    assert b._tower is None


def test_decoupled_penalty_schedule_matches_the_fused_loss(fake_kernels):
    """kernels.config.decouple_gp: the stacked pass differentiates its half of the critic loss (WGAN + CT + ACGAN) before the
    gradient-penalty pass has run, the penalty is differentiated by a second backward call -- same loss terms and the same
    accumulated parameter gradient as one backward pass over the fused loss."""
    import ctgan_b200.kernels as K
    from tests import parity
    res = {}
    for on in (False, True):
        K.config.decouple_gp = on
        try:
            tr, om = parity.build_pair('resnet', 'cpu', torch.float32, 4)
            parity.perturb_params(tr, om)
            rep = parity.critic_parity('resnet', tr, om, parity.make_inputs('resnet', 4, 11), conditioned=True)
            res[on] = rep
            assert parity.worst({k: v for k, v in rep.items() if not k.startswith('adam.')})[0] < 5e-4, parity.format_report(rep)
        finally:
            K.config.decouple_gp = False
    assert abs(res[True]['gradall'] - res[False]['gradall']) < 1e-5
