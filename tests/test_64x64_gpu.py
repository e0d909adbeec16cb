"""GPU tests (`-m gpu`) of SURVEY.md 8(f) row N4: the layer-norm kernel family (csrc/layernorm.cu) and the CT_gan_64x64.py
step against the oracle.  (Written at the end of round 1 without a GPU; validated on a B200 in round 2, where the step
parity exposed an allocator-reuse race of the side-stream parameter-gradient kernel -- functional.LayerNorm.backward.)
Host logic and math on the stand-in backend: tests/test_layernorm_host.py, tests/test_gan_64x64_host.py."""
import pytest
import torch

from tests import parity

pytestmark = pytest.mark.gpu

CL = torch.channels_last


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def act(shape, dtype, seed, shift=0.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * 1.3 + shift).to(dtype).contiguous(memory_format=CL)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('shape', [(32, 64, 64, 64), (5, 128, 16, 16), (3, 512, 4, 4), (2, 24, 7, 5)])
def test_layer_norm_kernels(shape, dtype):
    _need_gpu()
    import ctgan_b200.kernels as K
    from tests import fake_backend as fb
    N, C, H, W = shape
    x, v, c = act(shape, dtype, 1, 0.7), act(shape, dtype, 2), act(shape, dtype, 3)
    g = torch.Generator().manual_seed(4)
    gamma, beta = torch.randn(C, generator=g) * 0.5 + 1.0, torch.randn(C, generator=g) * 0.2
    tol = 1e-4 if dtype == torch.float32 else 1e-2
    y, mean, rstd = K.ln_fwd(x.cuda(), gamma.cuda(), beta.cuda(), 1e-5)
    ry, rmean, rrstd = fb.ln_fwd(x, gamma, beta, 1e-5)
    assert rel(mean, rmean) < 1e-5 and rel(rstd, rrstd) < 1e-5 and rel(y, ry) < tol
    for pre, post in ((1, 0), (0, 1), (0, 0)):
        got = K.ln_core(v.cuda(), x.cuda(), gamma.cuda(), mean, rstd, pre, post)
        assert rel(got, fb.ln_core(v, x, gamma, rmean, rrstd, pre, post)) < tol, (pre, post)
    dg, db = torch.ones(C, device='cuda'), torch.ones(C, device='cuda')
    K.ln_param_grad(v.cuda(), x.cuda(), mean, rstd, dg, db)
    rdg, rdb = torch.ones(C), torch.ones(C)
    fb.ln_param_grad(v, x, rmean, rrstd, rdg, rdb)
    assert rel(dg, rdg) < 1e-4 and rel(db, rdb) < 1e-4
    got = K.ln_bwd2_x(c.cuda(), v.cuda(), x.cuda(), gamma.cuda(), mean, rstd)
    assert rel(got, fb.ln_bwd2_x(c, v, x, gamma, rmean, rrstd)) < tol


@pytest.mark.parametrize('path', ['fp32', 'bf16'])
def test_64x64_step_parity(path):
    _need_gpu()
    dtype = torch.float32 if path == 'fp32' else torch.bfloat16
    B = 8
    tr, om = parity.build_pair('64x64', 'cuda', dtype, B, dim=64, oracle_dtype=torch.float32 if path == 'bf16' else torch.float64)
    parity.perturb_params(tr, om)
    floor = 1e-4 if path == 'fp32' else 1e-2
    rep = parity.critic_parity('64x64', tr, om, parity.make_inputs('64x64', B, 11), conditioned=True, floor_frac=floor)
    print('critic', parity.format_report(rep, 8))
    assert parity.worst(rep, 'loss.')[0] < (1e-3 if path == 'fp32' else 1e-2)
    assert parity.worst(rep, 'grad.')[0] < (1e-3 if path == 'fp32' else 3e-2)
    rep = parity.gen_parity('64x64', tr, om, conditioned=True, floor_frac=floor)
    print('gen', parity.format_report(rep, 8))
    assert parity.worst(rep, 'loss.')[0] < (1e-3 if path == 'fp32' else 1e-2)
    assert parity.worst(rep, 'grad.')[0] < (1e-3 if path == 'fp32' else 3e-2)
