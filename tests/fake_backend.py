"""TEST-ONLY stand-in for ctgan_b200.kernels on a machine without a GPU.

`install()` monkeypatches the launcher functions of `ctgan_b200.kernels` with PyTorch-CPU
implementations of the SAME contracts (layouts, dtypes, Philox streams), so the host-side
logic -- the twice-differentiable autograd composition in functional.py, the tflib
registry, the Trainer step assembly, FlatAdam, the gloo data-parallel path -- can be
exercised by `pytest -m "not gpu"`.  The product never imports this file and has no CPU
path of its own; GPU parity is proven by the `-m gpu` tests through the real C ABI.
"""
import math

import numpy as np
import torch
import torch.nn.functional as TF

from . import philox_ref

CL = torch.channels_last


def _f(x):
    return x.detach().to(torch.float32)


def _out(t, dtype, four_d=True):
    t = t.to(dtype)
    if t.dim() == 4:
        return t.contiguous(memory_format=CL)
    return t.contiguous()


def _as4(x, g, which):
    if x.dim() == 4:
        return _f(x)
    C = g.Cin if which == 'x' else g.Cout
    return _f(x).reshape(x.shape[0], C, 1, 1)


def _conv(x4, w, g):
    pb = max((g.Ho - 1) * g.stride + g.kh - g.H - g.pad_t, 0)
    pr = max((g.Wo - 1) * g.stride + g.kw - g.W - g.pad_l, 0)
    xp = TF.pad(x4, (g.pad_l, pr, g.pad_t, pb))
    w4 = w.to(torch.float32).reshape(g.kh, g.kw, g.Cin, g.Cout).permute(3, 2, 0, 1)
    y = TF.conv2d(xp, w4, stride=g.stride)
    return y[:, :, :g.Ho, :g.Wo]


def conv_fprop(x, w, bias, g, relu=False, residual=None, out_dtype=None, w_is_param=False, col=None, res_up2=False):
    if res_up2 and residual is not None:
        residual = upsample2x(residual, 1.0)
    two_d = x.dim() == 2
    y = _conv(_as4(x, g, 'x'), w.detach(), g)
    if bias is not None:
        y = y + _f(bias).view(1, -1, 1, 1)
    if residual is not None:
        y = y + _as4(residual, g, 'y')
    if relu:
        y = torch.relu(y)
    y = y.reshape(g.N, g.Cout) if two_d else y
    return _out(y, out_dtype or x.dtype)


def conv_dgrad(dy, w, g, out_dtype=None, w_is_param=False, col=None, relu_mask=None, out_s2d=False):
    if relu_mask is not None:
        return mul_relu_mask(conv_dgrad(dy, w, g, out_dtype=out_dtype), relu_mask)
    two_d = dy.dim() == 2
    x = torch.zeros(g.N, g.Cin, g.H, g.W, requires_grad=True)
    with torch.enable_grad():
        y = _conv(x, w.detach(), g)
        (dx,) = torch.autograd.grad(y, x, _as4(dy, g, 'y'))
    dx = dx.reshape(g.N, g.Cin) if two_d else dx
    return _out(dx, out_dtype or dy.dtype)


def thin_col(t, g, role):
    return None


def conv_wgrad(x, dy, g, w_shape, accumulate_into=None, col=None):
    w = torch.zeros(g.kh, g.kw, g.Cin, g.Cout, requires_grad=True)
    with torch.enable_grad():
        y = _conv(_as4(x, g, 'x'), w, g)
        (dw,) = torch.autograd.grad(y, w, _as4(dy, g, 'y'))
    dw = dw.reshape(w_shape).contiguous()
    if accumulate_into is not None:
        accumulate_into.add_(dw)
        return accumulate_into
    return dw


def _cols(x):
    return _f(x).reshape(x.shape[0], x.shape[1], -1) if x.dim() == 4 else _f(x).unsqueeze(-1)


def bias_grad(dy, accumulate_into=None):
    db = _cols(dy).sum(dim=(0, 2))
    if accumulate_into is not None:
        accumulate_into.add_(db.reshape(accumulate_into.shape))
        return accumulate_into
    return db


def bias_add(x, b):
    shape = (1, -1, 1, 1) if x.dim() == 4 else (1, -1)
    return _out(_f(x) + _f(b).view(*shape), x.dtype)


def add(a, b):
    return _out(_f(a) + _f(b), a.dtype)


def mul(a, b):
    return _out(_f(a) * _f(b), a.dtype)


def scale(a, s):
    return _out(_f(a) * s, a.dtype)


def cast(x, dtype):
    return x if x.dtype == dtype else _out(_f(x), dtype)


def _phys_index(x):
    """index of every logical element in the physical (dense) buffer of x"""
    idx = torch.zeros(x.shape, dtype=torch.int64)
    for d, (n, s) in enumerate(zip(x.shape, x.stride())):
        shape = [1] * x.dim()
        shape[d] = n
        idx = idx + (torch.arange(n) * s).view(shape)
    return idx


def _uniform_like(x, seed, offset, dyn):
    base = int(offset) + (int(dyn[0]) if dyn is not None else 0)
    u = philox_ref.uniform(seed, base, x.numel())
    return torch.from_numpy(u)[_phys_index(x).reshape(-1)].reshape(x.shape)


def act_dropout(x, slope, keep, u=None, seed=0, offset=0, want_mask=True, dyn=None):
    xf = _f(x)
    m = torch.where(xf > 0, torch.ones_like(xf), torch.full_like(xf, float(np.float32(slope))))
    if keep < 1.0:
        uu = _f(u) if u is not None else _uniform_like(x, seed, offset, dyn)
        m = m * torch.floor(torch.tensor(keep, dtype=torch.float32) + uu) * float(np.float32(1.0) / np.float32(keep))
    m = _out(m, x.dtype)
    y = _out(xf * _f(m), x.dtype)
    return y, (m if want_mask else None)


def fork_dropout_relu(x, keep, u=None, seed=0, offset=0, dyn=None):
    xf = _f(x)
    uu = _f(u) if u is not None else _uniform_like(x, seed, offset, dyn)
    md = _out(torch.floor(torch.tensor(keep, dtype=torch.float32) + uu) * float(np.float32(1.0) / np.float32(keep)), x.dtype)
    mdr = _out(torch.where(xf > 0, _f(md), torch.zeros_like(xf)), x.dtype)
    return _out(xf * _f(md), x.dtype), _out(xf * _f(mdr), x.dtype), md, mdr


def mask_sum2(a, ma, b, mb):
    av = _f(a) * _f(ma) if ma is not None else _f(a)
    return _out(av + _f(b) * _f(mb), a.dtype)


def mask_fork2(c, ma, mb):
    return _out(_f(c) * _f(ma), c.dtype), _out(_f(c) * _f(mb), c.dtype)


def mul_relu_mask(g, y):
    return _out(torch.where(_f(y) > 0, _f(g), torch.zeros_like(_f(g))), g.dtype)


def pool_add_fork(y, s, keep=1.0, u=None, seed=0, offset=0, dyn=None, masks=None):
    yf = _f(y)
    xf = (yf[:, :, ::2, ::2] + yf[:, :, 1::2, ::2] + yf[:, :, ::2, 1::2] + yf[:, :, 1::2, 1::2]) * 0.25 + _f(s)
    x = _out(xf, s.dtype)
    if masks is None:
        m1 = None
        if keep < 1.0:
            uu = _f(u) if u is not None else _uniform_like(x, seed, offset, dyn)
            m1 = _out(torch.floor(torch.tensor(keep, dtype=torch.float32) + uu) * float(np.float32(1.0) / np.float32(keep)), s.dtype)
        f1 = _f(m1) if m1 is not None else torch.ones_like(_f(x))
        m2 = _out(torch.where(_f(x) > 0, f1, torch.zeros_like(f1)), s.dtype)
    else:
        m1, m2 = masks
        f1 = _f(m1) if m1 is not None else torch.ones_like(_f(x))
    return _out(_f(x) * f1, s.dtype), _out(_f(x) * _f(m2), s.dtype), m1, m2


def mask_sum2_up(a, m1, b, m2):
    gx = mask_sum2(a, m1, b, m2)
    return upsample2x(gx, 0.25), gx


def unary_fwd(x, kind):
    return _out(torch.tanh(_f(x)) if kind == 0 else torch.sigmoid(_f(x)), x.dtype)


def unary_bwd(y, dy, kind):
    yf, g = _f(y), _f(dy)
    return _out(g * (1 - yf * yf) if kind == 0 else g * yf * (1 - yf), y.dtype)


def pool2x2(x, scale_):
    xf = _f(x)
    y = (xf[:, :, ::2, ::2] + xf[:, :, 1::2, ::2] + xf[:, :, ::2, 1::2] + xf[:, :, 1::2, 1::2]) * scale_
    return _out(y, x.dtype)


def upsample2x(x, scale_):
    y = _f(x).repeat_interleave(2, dim=2).repeat_interleave(2, dim=3) * scale_
    return _out(y, x.dtype)


def spatial_sum(x, scale_):
    return _out(_f(x).sum(dim=(2, 3)) * scale_, x.dtype)


def spatial_bcast(y, H, W, scale_):
    return _out((_f(y) * scale_)[:, :, None, None].expand(-1, -1, H, W), y.dtype)


def nchw_to_nhwc(x, N, C, H, W, out_dtype):
    return _out(_f(x).reshape(N, C, H, W), out_dtype)


def nhwc_to_nchw(x, out_dtype, out_shape):
    return _f(x).contiguous().reshape(out_shape).to(out_dtype)


def crop(x, h, w):
    return _out(_f(x)[:, :, :h, :w], x.dtype)


def crop_bwd(dy, H, W):
    N, C, h, w = dy.shape
    out = torch.zeros(N, C, H, W)
    out[:, :, :h, :w] = _f(dy)
    return _out(out, dy.dtype)


def prep_real(x_int, denom, noise_hi=0., seed=0, offset=0, dyn=None, out=None, out2=None):
    y = 2 * ((x_int.to(torch.float32) / np.float32(denom)) - 0.5)
    if noise_hi > 0:
        y = y + np.float32(noise_hi) * _uniform_like(x_int, seed, offset, dyn)
    if out2 is not None:
        out2.copy_(y.reshape(out2.shape))
    if out is not None:
        out.copy_(y.reshape(out.shape))
        return out
    return y


def interpolate(real, fake, alpha):
    return real + alpha * (fake - real)


def _bn_fwd1(x, gamma, beta, labels, eps, relu):
    xf = _f(x)
    dims = (0, 2, 3) if x.dim() == 4 else (0,)
    mean = xf.mean(dim=dims)
    var = ((xf - mean.view(1, -1, *([1] * (x.dim() - 2)))) ** 2).mean(dim=dims)
    invstd = torch.rsqrt(var + eps)
    C = xf.shape[1]
    g2, b2 = _f(gamma).reshape(-1, C), _f(beta).reshape(-1, C)
    idx = labels.long() if labels is not None else torch.zeros(x.shape[0], dtype=torch.long)
    sh = (x.shape[0], C) + (1,) * (x.dim() - 2)
    bc = (1, C) + (1,) * (x.dim() - 2)
    y = (xf - mean.view(bc)) * invstd.view(bc) * g2[idx].view(sh) + b2[idx].view(sh)
    if relu:
        y = torch.relu(y)
    return y, mean, invstd


def bn_fused_ok(x, groups=1):
    return False


def bn_fwd(x, gamma, beta, labels, eps, relu, groups=1, up2=False):
    if up2:
        y, m, i = bn_fwd(x, gamma, beta, labels, eps, relu, groups)
        return upsample2x(y, 1.0), m, i
    n = x.shape[0] // groups
    parts = [_bn_fwd1(x[g * n:(g + 1) * n], gamma, beta, None if labels is None else labels[g * n:(g + 1) * n], eps, relu)
             for g in range(groups)]
    y = torch.cat([p[0] for p in parts], 0)
    return _out(y, x.dtype), torch.stack([p[1] for p in parts]), torch.stack([p[2] for p in parts])


def bn_bwd(dy, x, y, gamma, beta, labels, mean, invstd, relu, groups=1, up2=False, accumulate_into=None):
    if up2:
        dy = pool2x2(dy, 1.0)
        y = y[:, :, ::2, ::2] if y is not None else None
    if accumulate_into is not None:
        dx, dg, db = bn_bwd(dy, x, y, gamma, beta, labels, mean, invstd, relu, groups)
        accumulate_into[0].add_(dg); accumulate_into[1].add_(db)
        return dx, accumulate_into[0], accumulate_into[1]
    n = x.shape[0] // groups
    outs = [_bn_bwd1(dy[g * n:(g + 1) * n], x[g * n:(g + 1) * n], None if y is None else y[g * n:(g + 1) * n], gamma,
                     None if labels is None else labels[g * n:(g + 1) * n], mean.reshape(groups, -1)[g],
                     invstd.reshape(groups, -1)[g], relu) for g in range(groups)]
    dx = torch.cat([o[0] for o in outs], 0)
    return _out(dx, x.dtype), sum(o[1] for o in outs), sum(o[2] for o in outs)


def _bn_bwd1(dy, x, y, gamma, labels, mean, invstd, relu):
    xf, g = _f(x), _f(dy)
    C = xf.shape[1]
    if relu:
        g = g * (_f(y) > 0)
    bc = (1, C) + (1,) * (x.dim() - 2)
    sh = (x.shape[0], C) + (1,) * (x.dim() - 2)
    xh = (xf - mean.view(bc)) * invstd.view(bc)
    g2 = _f(gamma).reshape(-1, C)
    nl = g2.shape[0]
    idx = labels.long() if labels is not None else torch.zeros(x.shape[0], dtype=torch.long)
    red = tuple(range(2, x.dim()))
    s1 = g.sum(dim=red) if red else g
    s2 = (g * xh).sum(dim=red) if red else g * xh
    dgamma = torch.zeros(nl, C).index_add_(0, idx, s2)
    dbeta = torch.zeros(nl, C).index_add_(0, idx, s1)
    dxh = g * g2[idx].view(sh)
    dims = (0,) + red
    R = xf.numel() // C
    dx = invstd.view(bc) * (dxh - dxh.sum(dim=dims).view(bc) / R - xh * (dxh * xh).sum(dim=dims).view(bc) / R)
    return dx, dgamma.reshape(gamma.shape), dbeta.reshape(gamma.shape)


def _ln_parts(x, mean, rstd):
    N = x.shape[0]
    xh = (_f(x) - mean.view(N, 1, 1, 1)) * rstd.view(N, 1, 1, 1)
    smean = lambda t: t.mean(dim=(1, 2, 3), keepdim=True)
    return xh, smean, rstd.view(N, 1, 1, 1)


def ln_fwd(x, gamma, beta, eps):
    xf = _f(x)
    mean = xf.mean(dim=(1, 2, 3))
    var = ((xf - mean.view(-1, 1, 1, 1)) ** 2).mean(dim=(1, 2, 3))
    rstd = torch.rsqrt(var + eps)
    xh, _, _ = _ln_parts(x, mean, rstd)
    return _out(xh * _f(gamma).view(1, -1, 1, 1) + _f(beta).view(1, -1, 1, 1), x.dtype), mean, rstd


def ln_core(v, x, gamma, mean, rstd, pre_scale, post_scale):
    xh, smean, r = _ln_parts(x, mean, rstd)
    g = _f(gamma).view(1, -1, 1, 1)
    u = _f(v) * g if pre_scale else _f(v)
    out = r * (u - smean(u) - xh * smean(u * xh))
    return _out(out * g if post_scale else out, x.dtype)


def ln_param_grad(v, x, mean, rstd, dgamma, dbeta):
    xh, _, _ = _ln_parts(x, mean, rstd)
    dgamma.add_((_f(v) * xh).sum(dim=(0, 2, 3)))
    if dbeta is not None:
        dbeta.add_(_f(v).sum(dim=(0, 2, 3)))


def ln_bwd2_x(c, gy, x, gamma, mean, rstd):
    xh, smean, r = _ln_parts(x, mean, rstd)
    a, b = _f(c), _f(gy) * _f(gamma).view(1, -1, 1, 1)
    abar, bbar, A, B = smean(a), smean(b), smean(a * xh), smean(b * xh)
    Q = smean(a * b) - abar * bbar - A * B
    return _out(-r * r * (xh * Q + B * (a - abar) + A * (b - bbar) - 2 * xh * A * B), x.dtype)


def _loss_terms(desc, d_real, d_real2, d_fake, f1, f2, grad, logits, labels):
    diff = d_real - d_real2
    ct_i = desc.lambda2 * diff ** 2 + desc.lambda2 * 0.1 * ((_f(f1) - _f(f2)) ** 2).mean(dim=1) - desc.factor_m
    s = torch.sqrt((grad ** 2).sum(dim=1))
    ce = (torch.logsumexp(logits, 1) - logits.gather(1, labels.long().view(-1, 1)).view(-1)) if logits is not None \
        else torch.zeros_like(s)
    return ct_i, s, ce


def ct_gp_loss_fwd(desc, d_real, d_real2, d_fake, f1, f2, grad, logits, labels):
    if d_real is None:                                    # the penalty term only
        s = torch.sqrt((grad ** 2).sum(1))
        gp = ((s - 1) ** 2).mean()
        out = torch.zeros(8)
        out[0], out[3] = desc.lambda_gp * gp, gp
        per = torch.zeros(desc.B, 4)
        per[:, 0], per[:, 1] = -1.0, s
        return out, per.reshape(-1)
    if grad is None:                                      # no penalty term: slopes recorded as 1
        out, per = ct_gp_loss_fwd(desc, d_real, d_real2, d_fake, f1, f2, torch.zeros(desc.B, 1), logits, labels)
        per = per.reshape(-1, 4).clone()
        out = out.clone()
        out[0] = out[0] - desc.lambda_gp * out[3]
        out[3] = 0.0
        per[:, 1] = 1.0
        return out, per.reshape(-1)
    ct_i, s, ce = _loss_terms(desc, d_real, d_real2, d_fake, f1, f2, grad, logits, labels)
    wgan = d_fake.mean() - d_real.mean()
    ct, gp = torch.clamp(ct_i, min=0).mean(), ((s - 1) ** 2).mean()
    acgan = ce.mean() if logits is not None else torch.zeros(())
    out = torch.zeros(8)
    out[0] = wgan + ct + desc.lambda_gp * gp + desc.acgan_scale * acgan
    out[1], out[2], out[3], out[4] = wgan, ct, gp, acgan
    per = torch.zeros(desc.B, 4)
    per[:, 0], per[:, 1], per[:, 2] = ct_i, s, ce
    return out, per.reshape(-1)


def ct_gp_loss_bwd(desc, gcost, d_real, d_real2, f1, f2, grad, logits, labels, per_sample, outs=None):
    if outs is not None:
        r = ct_gp_loss_bwd(desc, gcost, d_real, d_real2, f1, f2, grad, logits, labels, per_sample)
        for dst, src in zip(outs, (r[0], r[1], r[2], r[3], r[4], r[6])):
            if dst is not None:
                dst.copy_(src)
        return r
    per = per_sample.reshape(-1, 4)
    g, B = gcost[0], desc.B
    if d_real is None:                                    # the penalty term only
        s = per[:, 1]
        cg = torch.where(s > 0, g * desc.lambda_gp / B * 2 * (s - 1) / s, torch.zeros_like(s))
        return None, None, None, None, None, cg.view(-1, 1) * grad, None
    active = (per[:, 0] >= 0).float()
    gct = g * active / B
    diff = d_real - d_real2
    t = gct * 2 * desc.lambda2 * diff
    g_real, g_real2 = -g / B + t, -t
    g_fake = torch.full((desc.NF,), 1.0) * g / desc.NF
    a = _f(f1) - _f(f2)
    gf1 = (gct * 0.1 * desc.lambda2 * 2 / desc.F).view(-1, 1) * a
    s = per[:, 1]
    cg = torch.where(s > 0, g * desc.lambda_gp / B * 2 * (s - 1) / s, torch.zeros_like(s))
    g_grad = cg.view(-1, 1) * grad if grad is not None else None
    g_logits = None
    if logits is not None:
        p = torch.softmax(logits, 1)
        oh = torch.zeros_like(p).scatter_(1, labels.long().view(-1, 1), 1.0)
        g_logits = g * desc.acgan_scale / B * (p - oh)
    return g_real, g_real2, g_fake, gf1.to(f1.dtype), (-gf1).to(f2.dtype), g_grad, g_logits


def mean_fwd(d, sign):
    return (sign * d.mean()).reshape(1)


def mean_bwd(gcost, n, sign):
    return torch.full((n,), 1.0) * sign * gcost[0] / n


def softmax_ce_fwd(logits, labels):
    return (torch.logsumexp(logits, 1) - logits.gather(1, labels.long().view(-1, 1)).view(-1)).mean().reshape(1)


def softmax_ce_bwd(logits, labels, gcost, scale_):
    p = torch.softmax(logits, 1)
    oh = torch.zeros_like(p).scatter_(1, labels.long().view(-1, 1), 1.0)
    return gcost[0] * scale_ / logits.shape[0] * (p - oh)


def adam_step(p, g, m, v, lr_t, beta1, beta2, eps, grad_scale=1.0, lr_t_dev=None):
    if lr_t_dev is not None:
        lr_t = float(lr_t_dev[0])
    gr = g * grad_scale
    m.mul_(beta1).add_(gr, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(gr, gr, value=1 - beta2)
    p.sub_(np.float32(lr_t) * m / (v.sqrt() + eps))


def philox_uniform(shape, device, seed, offset, lo=0., hi=1., memory_format=None, dyn=None):
    base = int(offset) + (int(dyn[0]) if dyn is not None else 0)
    n = int(math.prod(shape))
    u = torch.from_numpy(philox_ref.uniform(seed, base, n))
    u = np.float32(lo) + np.float32(hi - lo) * u
    out = torch.empty(shape, dtype=torch.float32, memory_format=memory_format) if memory_format is not None \
        else torch.empty(shape, dtype=torch.float32)
    out.as_strided((n,), (1,)).copy_(u)       # element i of the BUFFER == stream element offset+i
    return out


def philox_normal(shape, device, seed, offset, dyn=None):
    base = int(offset) + (int(dyn[0]) if dyn is not None else 0)
    n = int(math.prod(shape))
    return torch.from_numpy(philox_ref.normal(seed, base, n)).reshape(shape)


def philox_labels(n, device, n_labels, seed, offset, dyn=None):
    base = int(offset) + (int(dyn[0]) if dyn is not None else 0)
    u = philox_ref.uniform(seed, base, n)
    return torch.from_numpy((u * np.float32(n_labels)).astype('int32'))


def counter_add(counter, delta):
    counter[0] += int(delta)


def invalidate_weight_cache(ptrs=None, forget=False):
    pass


_NAMES = ['conv_fprop', 'conv_dgrad', 'conv_wgrad', 'bias_grad', 'bias_add', 'add', 'mul', 'scale', 'cast',
          'act_dropout', 'fork_dropout_relu', 'mask_sum2', 'mask_fork2', 'mul_relu_mask', 'pool_add_fork', 'mask_sum2_up', 'unary_fwd', 'unary_bwd', 'pool2x2', 'upsample2x', 'spatial_sum', 'spatial_bcast',
          'nchw_to_nhwc', 'nhwc_to_nchw', 'crop', 'crop_bwd', 'prep_real', 'interpolate', 'bn_fwd', 'bn_bwd', 'bn_fused_ok',
          'ct_gp_loss_fwd', 'ct_gp_loss_bwd', 'mean_fwd', 'mean_bwd', 'softmax_ce_fwd', 'softmax_ce_bwd',
          'ln_fwd', 'ln_core', 'ln_param_grad', 'ln_bwd2_x',
          'adam_step', 'philox_uniform', 'philox_normal', 'philox_labels', 'counter_add', 'invalidate_weight_cache']


def install(monkeypatch):
    """Patch ctgan_b200.kernels for the duration of one test (pytest monkeypatch fixture)."""
    import ctgan_b200.kernels as K
    import ctgan_b200.tflib as lib
    g = globals()
    for n in _NAMES:
        monkeypatch.setattr(K, n, g[n])
    monkeypatch.setattr(lib, '_device', torch.device('cpu'))
