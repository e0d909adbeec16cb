"""GPU (`-m gpu`): the CUDA path against golden vectors produced by THE REFERENCE'S OWN CODE
(tests/golden/*.npz, see tests/golden/make_golden.py) -- no oracle in between.  The fixtures' parameters are
loaded into the product by reference name, the fixtures' random draws (z, dropout uniforms, alpha, dequant
noise, labels) are injected through DeviceRandom.replay, and disc_cost / gen_cost / the GP gradient / every
parameter gradient are compared.  fp32 path, tolerance 1e-3 (measured ~1e-6; the bound leaves room for a
ReLU tie flip, see tests/test_step_parity_gpu.py)."""
import importlib

import numpy as np
import pytest
import torch

from tests import parity
from tests.test_oracle_vs_reference import load_golden

pytestmark = pytest.mark.gpu
DIM_NAMES = {'mnist': ['DIM'], 'cifar': ['DIM'], 'resnet': ['DIM_G', 'DIM_D'], '64x64': ['DIM']}
# (LS/wgan_LSUN_Bedrooms128.py: tests/test_lsun128_gpu.py against tests/golden/lsun128q.npz -- the small lsun128.npz fixture that pins
# the oracle has 2-channel layers, below the kernels' 4-channel granularity)


@pytest.mark.parametrize('script', ['mnist', 'cifar', 'resnet', '64x64'])
def test_product_matches_reference_golden(script):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    g = load_golden(script)
    prod = importlib.import_module(parity.SCRIPTS[script][0])
    saved = {k: getattr(prod, k) for k in DIM_NAMES[script]}
    for k in saved:
        setattr(prod, k, g['dim'])
    try:
        np.random.seed(0)
        tr = prod.Trainer(device='cuda', seed=1, act_dtype=torch.float32, batch_size=g['B'])
        import ctgan_b200.tflib as lib
        assert set(lib._params) == set(g['params'])
        with torch.no_grad():
            for n, p in lib._params.items():
                p.copy_(torch.from_numpy(g['params'][n]).to(p.device).reshape(p.shape))
        inputs = tuple(torch.from_numpy(a).cuda() for a in g['inputs'])
        ref = g['ref']
        # ---- critic
        tr.rng.replay = g['tape_disc']
        tr.disc_opt.zero_grad()
        res = tr.critic_forward_backward(*inputs)
        cost = float(res['out'][0])
        assert abs(cost - float(ref['disc_cost'])) <= 1e-3 * max(1.0, abs(float(ref['disc_cost']))), (cost, float(ref['disc_cost']))
        gp_rows = len(ref['gp_gradients'])       # 64x64: the reference's `gradients` is the LAST tower's (TG/CT_gan_64x64.py:505)
        assert parity.rel_err(res['gradients'][-gp_rows:], torch.from_numpy(ref['gp_gradients'])) < 1e-3
        floor = 1e-4 * max(float(np.linalg.norm(v)) for v in ref['disc_grads'].values())
        for n, q in tr.disc_opt.params.items():
            assert parity.rel_err(q.grad, torch.from_numpy(ref['disc_grads'][n]), floor) < 1e-3, n
        # ---- generator
        tr.rng.replay = g['tape_gen']
        tr.gen_opt.zero_grad()
        res = tr.gen_forward_backward()
        assert abs(float(res['cost']) - float(ref['gen_cost'])) <= 1e-3 * max(1.0, abs(float(ref['gen_cost'])))
        floor = 1e-4 * max(float(np.linalg.norm(v)) for v in ref['gen_grads'].values())
        for n, q in tr.gen_opt.params.items():
            if n in ref['gen_grads']:
                # 64x64 at DIM 4: 4-element tensors, one ReLU tie flip moves them by a few 1e-3
                assert parity.rel_err(q.grad, torch.from_numpy(ref['gen_grads'][n]), floor) < (5e-3 if script == '64x64' else 2e-3), n
    finally:
        for k, v in saved.items():
            setattr(prod, k, v)


# ---------------------------------------------------------------------------------------------------------------------
# Full width (DIM 128: the shapes the tcgen05 kernels run), every arithmetic path of the product, against vectors the
# REFERENCE'S OWN CODE produced in float64 (tests/golden/resnet128.npz, script tests/golden/make_golden_full.py).
# The comparison is direct -- no oracle, no activation patterns handed over -- so gradient bounds are those of two
# precisions of a ReLU network deciding their own patterns (tests/test_step_parity_gpu.py, "independent" mode);
# the loss terms carry north_star's bar.
FULL_TOL = {   # path: (loss terms, GP gradient / whole parameter gradient, worst single parameter tensor)
    'fp32': (1e-3, 5e-3, 1e-2),
    'tf32': (1e-2, 5e-2, 0.1),
    'bf16': (1e-2, 0.2, 0.25),
}


def _load_full():
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'resnet128.npz'))
    tapes = {'disc': {}, 'gen': {}}
    for k in z.files:
        for kind in ('disc', 'gen'):
            if k.startswith('tape_%s.' % kind):
                tapes[kind][k[len('tape_%s.' % kind):]] = torch.from_numpy(z[k])
            elif k.startswith('keepbits_%s.' % kind):
                tag = k[len('keepbits_%s.' % kind):]
                shape = tuple(int(v) for v in z['keepshape_%s.%s' % (kind, tag)])
                bits = np.unpackbits(z[k])[:int(np.prod(shape))].reshape(shape).astype(bool)
                tapes[kind][tag] = torch.from_numpy(np.where(bits, np.float32(0.999), np.float32(0.0)))   # floor(keep + u)
    return z, tapes


@pytest.mark.parametrize('path', ['bf16', 'tf32', 'fp32'])
def test_full_width_product_matches_reference_golden(path):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from tests.golden.det_params import det_param
    from tests.test_step_parity_gpu import _path
    import ctgan_b200.gan_cifar_resnet as R
    import ctgan_b200.tflib as lib
    import ctgan_b200.kernels as K
    z, tapes = _load_full()
    B, seed, stride = int(z['meta.B']), int(z['meta.seed']), int(z['meta.stride'])
    assert R.DIM_G == int(z['meta.dim']) == R.DIM_D
    tol_loss, tol_all, tol_one = FULL_TOL[path]
    with _path(path) as dtype:
        np.random.seed(0)
        tr = R.Trainer(device='cuda', seed=1, act_dtype=dtype, batch_size=B)
        with torch.no_grad():
            for n, p in lib._params.items():
                p.copy_(torch.from_numpy(det_param(n, p.detach().cpu().numpy(), seed)).to(p.device))
        K.invalidate_weight_cache(tr.gen_opt._ptrs | tr.disc_opt._ptrs)
        tr.gen_opt.refresh_packs(); tr.disc_opt.refresh_packs()
        inputs = (torch.from_numpy(z['input.0']).cuda(), torch.from_numpy(z['input.1']).cuda())

        def compare(opt, kind, cost, ref_cost):
            assert abs(cost - ref_cost) <= tol_loss * max(1.0, abs(ref_cost)), (kind, cost, ref_cost)
            num = den = 0.0
            worst = (0.0, '')
            floor = 1e-2 * max(float(z['gnorm_%s.%s' % (kind, n)]) for n in opt.params if 'gnorm_%s.%s' % (kind, n) in z.files)
            for n, q in opt.params.items():
                if 'gsample_%s.%s' % (kind, n) not in z.files:
                    continue
                got = q.grad.detach().double().cpu().reshape(-1)
                want = torch.from_numpy(z['gsample_%s.%s' % (kind, n)]).double()
                sel = got if want.numel() == got.numel() else got[::stride]          # small tensors are stored in full
                d2, w2 = float(((sel - want) ** 2).sum()), float((want ** 2).sum())
                num += d2; den += w2
                # one tensor: sampled elements, scaled to the full tensor's norm (floor: 1 % of the largest gradient)
                gn = float(z['gnorm_%s.%s' % (kind, n)])
                e = (d2 / max(w2, 1e-300)) ** 0.5 * gn / max(gn, floor)
                worst = max(worst, (e, n))
                assert abs(float(got.norm()) - gn) <= tol_one * max(gn, floor), (kind, n, float(got.norm()), gn)
            print('%s/%s: cost %.6f (ref %.6f)  sampled gradient %.2e  worst tensor %.2e %s' % (kind, path, cost, ref_cost,
                                                                                        (num / den) ** 0.5, worst[0], worst[1]))
            assert (num / den) ** 0.5 < tol_all, (kind, (num / den) ** 0.5)
            assert worst[0] < tol_one, (kind, worst)

        tr.rng.replay = tapes['disc']
        tr.disc_opt.zero_grad()
        res = tr.critic_forward_backward(*inputs)
        out = res['out'].cpu()
        # out = {cost, wgan term, ct, gp, acgan}: the reference's CT_ and gradient_penalty (= 10 * gp, :286) are separate fetches
        assert abs(float(out[2]) - float(z['CT_'])) <= tol_loss * max(1.0, abs(float(z['CT_'])))
        assert abs(10.0 * float(out[3]) - float(z['gradient_penalty'])) <= tol_loss * max(1.0, abs(float(z['gradient_penalty'])))
        assert parity.rel_err(res['gradients'], torch.from_numpy(z['gp_gradients'])) < tol_all
        compare(tr.disc_opt, 'disc', float(out[0]), float(z['disc_cost']))
        tr.rng.replay = tapes['gen']
        tr.gen_opt.zero_grad()
        res = tr.gen_forward_backward()
        compare(tr.gen_opt, 'gen', float(res['cost']), float(z['gen_cost']))
