"""Layer norm on the stand-in backend: the twice-differentiable composition in
ctgan_b200/functional.py (LayerNorm / LayerNormBwd) against PyTorch autograd of the plain formula -- first order (dx, dgamma,
dbeta) and the second-order terms the gradient penalty needs (d<c, dx>/d{gy, x, gamma}), including the closed form of the
x-derivative documented in csrc/layernorm.cu."""
import pytest
import torch

CL = torch.channels_last


def _ref_ln(x, gamma, beta, eps=1e-5):
    mu = x.mean(dim=(1, 2, 3), keepdim=True)
    var = ((x - mu) ** 2).mean(dim=(1, 2, 3), keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * gamma.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize('shape', [(3, 8, 4, 4), (2, 5, 3, 7)])
def test_layer_norm_first_and_second_order(fake_kernels, shape):
    import ctgan_b200.functional as F
    torch.manual_seed(0)
    N, C, H, W = shape
    x0 = (torch.randn(shape) * 1.5 + 0.3).contiguous(memory_format=CL)
    g0, b0 = torch.randn(C) * 0.5 + 1.0, torch.randn(C) * 0.2
    gy0, c0 = torch.randn(shape).contiguous(memory_format=CL), torch.randn(shape).contiguous(memory_format=CL)

    def run(fn, dtype):
        x = x0.to(dtype).clone().requires_grad_(True)
        gamma, beta = g0.to(dtype).clone().requires_grad_(True), b0.to(dtype).clone().requires_grad_(True)
        gy = gy0.to(dtype).clone().requires_grad_(True)
        y = fn(x, gamma, beta)
        dx, dgamma, dbeta = torch.autograd.grad(y, (x, gamma, beta), gy, create_graph=True, allow_unused=True)
        ggy, gx, ggamma = torch.autograd.grad(dx, (gy, x, gamma), c0.to(dtype), allow_unused=True)
        return [t.detach() for t in (y, dx, dgamma, dbeta, ggy, gx, ggamma)]

    ref = run(_ref_ln, torch.float64)
    got = run(lambda x, g, b: F.layer_norm(x, g, b, 1e-5), torch.float32)
    for name, a, b in zip(('y', 'dx', 'dgamma', 'dbeta', 'ggy', 'gx', 'ggamma'), got, ref):
        assert _rel(a, b) < 2e-5, name


def test_layer_norm_inside_a_gradient_penalty(fake_kernels):
    """conv -> layer norm -> relu -> conv -> sum, penalty on ||d out / d x||: the weight gradients of BOTH convs (the one
    before the norm receives the second-order x-term) against autograd of the plain formulas."""
    import numpy as np
    import ctgan_b200.functional as F
    import ctgan_b200.kernels as K
    import torch.nn.functional as TF
    torch.manual_seed(1)
    N, C, H = 3, 4, 6
    x0 = torch.randn(N, C, H, H)
    w1, w2 = torch.randn(3, 3, C, C) * 0.3, torch.randn(3, 3, C, C) * 0.3
    g0, b0 = torch.rand(C) + 0.5, torch.randn(C) * 0.1

    def penalty(conv, ln, relu, dtype, leaf):
        x = x0.to(dtype).clone().requires_grad_(True)
        a, b = leaf(w1.to(dtype)), leaf(w2.to(dtype))
        gamma, beta = leaf(g0.to(dtype)), leaf(b0.to(dtype))
        out = conv(relu(ln(conv(x, a), gamma, beta)), b)
        (gx,) = torch.autograd.grad(out.sum(), x, create_graph=True)
        loss = ((gx.reshape(N, -1).norm(dim=1) - 1.0) ** 2).mean() + out.mean()
        return torch.autograd.grad(loss, (a, b, gamma, beta), allow_unused=True)

    ref = penalty(lambda x, w: TF.conv2d(x, w.permute(3, 2, 0, 1), padding=1), _ref_ln, torch.relu, torch.float64,
                  lambda t: t.clone().requires_grad_(True))

    def conv(x, w):
        return F.conv2d(F.ensure_nhwc(x), w, None, 3, 1)
    got = penalty(conv, lambda x, g, b: F.layer_norm(x, g, b, 1e-5), F.relu, torch.float32,
                  lambda t: t.clone().contiguous().requires_grad_(True))
    for name, a, b in zip(('w1', 'w2', 'gamma', 'beta'), got, ref):
        assert a is not None and _rel(a, b) < 1e-4, (name, _rel(a, b))
