"""CPU test of the space-to-depth route for stride-2 5x5 convolutions (ctgan_b200/kernels.py: s2d_geom, conv_fprop /
conv_dgrad / conv_wgrad branches; csrc/conv_s2d.cu).

The route composes five primitives: space_to_depth, depth_to_space, pack_filter_s2d, the raw stride-1 tensor-core
fprop / wgrad launches, and s2d_filter_grad.  Here each primitive is replaced by an independent PyTorch-CPU
restatement of its documented contract (include/ctgan_sm100.h) -- in particular the raw fprop INTERPRETS the packed
operand layouts -- and the composed stride-2 conv family is compared with a direct evaluation of the stride-2 TF-SAME
convolution (tests/fake_backend.py).  The GPU test (tests/test_kernels_gpu.py::test_s2d_conv_family) checks the CUDA
primitives themselves."""
import pytest
import torch

from tests import fake_backend as fb

CL = torch.channels_last


def _s2d(x, g):
    N, C, H, W = x.shape
    Hs, Ws = (H + 1) // 2, (W + 1) // 2
    xp = torch.zeros(N, C, 2 * Hs, 2 * Ws, dtype=x.dtype)
    xp[:, :, :H, :W] = x
    xs = xp.view(N, C, Hs, 2, Ws, 2).permute(0, 3, 5, 1, 2, 4).reshape(N, 4 * C, Hs, Ws)
    return xs.contiguous(memory_format=CL)


def _d2s(xs, g):
    N, C4, Hs, Ws = xs.shape
    C = C4 // 4
    x = xs.reshape(N, 2, 2, C, Hs, Ws).permute(0, 3, 4, 1, 5, 2).reshape(N, C, 2 * Hs, 2 * Ws)
    return x[:, :, :g.H, :g.W].contiguous(memory_format=CL)


def _embed(w, g):
    """W3[R,S,(dy*2+dx)*C+c,o] = w[2(R-1)+dy+pad_t, 2(S-1)+dx+pad_l, c, o]."""
    k, C, O = g.kh, g.Cin, g.Cout
    w3 = torch.zeros(3, 3, 4, C, O)
    for R in range(3):
        for S in range(3):
            for dy in range(2):
                for dx in range(2):
                    r, s = 2 * (R - 1) + dy + g.pad_t, 2 * (S - 1) + dx + g.pad_l
                    if 0 <= r < k and 0 <= s < k:
                        w3[R, S, dy * 2 + dx] = w[r, s].float()
    return w3.reshape(3, 3, 4 * C, O)


def _pack_launch(w, wp_f, wp_d, g):
    w3 = _embed(w.detach(), g).to(torch.bfloat16)
    wp_f.copy_(w3.permute(0, 1, 3, 2).reshape(-1))                       # [T][O][4C]
    wp_d.copy_(w3.reshape(9, 4 * g.Cin, g.Cout).flip(0).reshape(-1))     # [8-T][4C][O]


def _fprop_packed(x, wp, bias, residual, relu_mask, y, g, flags):
    w = wp.float().view(g.kh, g.kw, g.Cout, g.Cin).permute(0, 1, 3, 2).contiguous()    # packed [taps][Cout][Cin] -> HWIO
    out = fb.conv_fprop(x, w, bias, g, relu=bool(flags & 1), residual=residual)
    if relu_mask is not None:
        out = out * (relu_mask > 0)
    y.copy_(out)
    return y


def _wgrad_raw(x, dy, g, dw):
    dw.add_(fb.conv_wgrad(x, dy, g, tuple(dw.shape)))
    return dw


def _filter_grad_launch(dw3, dw, g, accumulate):
    k, C = g.kh, g.Cin
    out = torch.zeros_like(dw)
    d5 = dw3.view(3, 3, 4, C, g.Cout)
    for r in range(k):
        for s in range(k):
            ar, as_ = r - g.pad_t + 2, s - g.pad_l + 2
            out[r, s] = d5[ar >> 1, as_ >> 1, (ar & 1) * 2 + (as_ & 1)]
    if accumulate:
        dw.add_(out)
    else:
        dw.copy_(out)


def _im2col_strided(x, g):
    """col[p][(r*kw+s)*C + c] = x[n, st*ho + r - pad_t, st*wo + s - pad_l, c], 128 columns per output pixel."""
    import torch.nn.functional as TF
    st, C = g.stride, g.Cin
    need_h, need_w = (g.Ho - 1) * st + g.kh, (g.Wo - 1) * st + g.kw
    xp = TF.pad(x.float(), (g.pad_l, max(need_w - g.W - g.pad_l, 0), g.pad_t, max(need_h - g.H - g.pad_t, 0)))
    col = torch.zeros(g.N, 128, g.Ho, g.Wo)
    for r in range(g.kh):
        for s_ in range(g.kw):
            k0 = (r * g.kw + s_) * C
            col[:, k0:k0 + C] = xp[:, :, r:r + st * g.Ho:st, s_:s_ + st * g.Wo:st]
    return col.to(torch.bfloat16).contiguous(memory_format=CL)


def _col2im_strided(col, bias, g):
    st, C = g.stride, g.Cin
    need_h, need_w = (g.Ho - 1) * st + g.kh, (g.Wo - 1) * st + g.kw
    Hp, Wp = max(need_h, g.H + g.pad_t), max(need_w, g.W + g.pad_l)
    dxp = torch.zeros(g.N, C, Hp, Wp)
    for r in range(g.kh):
        for s_ in range(g.kw):
            k0 = (r * g.kw + s_) * C
            dxp[:, :, r:r + st * g.Ho:st, s_:s_ + st * g.Wo:st] += col[:, k0:k0 + C].float()
    dx = dxp[:, :, g.pad_t:g.pad_t + g.H, g.pad_l:g.pad_l + g.W]
    if bias is not None:
        dx = dx + bias.view(1, -1, 1, 1)
    return dx.to(torch.bfloat16).contiguous(memory_format=CL)


def _padk_launch(w, wp_f, wp_d, g):
    kreal = g.kh * g.kw * g.Cin
    w2 = w.detach().reshape(kreal, g.Cout).to(torch.bfloat16)
    f = torch.zeros(g.Cout, 128, dtype=torch.bfloat16); f[:, :kreal] = w2.t()
    d = torch.zeros(128, g.Cout, dtype=torch.bfloat16); d[:kreal] = w2
    wp_f.copy_(f.reshape(-1)); wp_d.copy_(d.reshape(-1))


def _add_prefix(src, dst, n, accumulate):
    flat = dst.view(-1)
    if accumulate:
        flat[:n] += src.reshape(-1)[:n]
    else:
        flat[:n] = src.reshape(-1)[:n]


@pytest.fixture
def K(monkeypatch):
    import ctgan_b200.kernels as K
    monkeypatch.setattr(K, 'tc_available', lambda: True)
    monkeypatch.setattr(K, '_chk', lambda t, name='tensor': None)
    monkeypatch.setattr(K, '_stream', lambda: None)
    monkeypatch.setattr(K, 'space_to_depth', _s2d)
    monkeypatch.setattr(K, 'depth_to_space', _d2s)
    monkeypatch.setattr(K, '_pack_filter_s2d_launch', _pack_launch)
    monkeypatch.setattr(K, '_fprop_tc_packed', _fprop_packed)
    monkeypatch.setattr(K, '_wgrad_tc_raw', _wgrad_raw)
    monkeypatch.setattr(K, '_s2d_filter_grad_launch', _filter_grad_launch)
    monkeypatch.setattr(K, 'im2col_strided', _im2col_strided)
    monkeypatch.setattr(K, 'col2im_strided', _col2im_strided)
    monkeypatch.setattr(K, '_pack_filter_padk_launch', _padk_launch)
    monkeypatch.setattr(K, '_add_prefix_launch', _add_prefix)
    monkeypatch.setattr(K.config, 'use_s2d', True)
    K.invalidate_weight_cache()
    yield K
    K.invalidate_weight_cache()


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def act(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g).to(torch.bfloat16).contiguous(memory_format=CL)


GEOMS = [
    # N, H, W, Cin, Cout  (5x5, stride 2, TF SAME)
    (2, 16, 16, 32, 128),     # even extent: pads (1, 2)       -- CIFAR D.2 / Deconv 8->16 pattern
    (3, 7, 7, 32, 128),       # odd extent:  pads (2, 2)       -- MNIST D.3
    (2, 14, 14, 16, 64),      # 4*Cin = 64, Cout = 64: fprop/dgrad eligible, wgrad falls through to the SIMT kernel
    (2, 8, 6, 32, 128),       # non-square
]


@pytest.mark.parametrize('geom', GEOMS)
def test_s2d_route_equals_stride2_conv(K, geom, monkeypatch):
    N, H, W, Cin, Cout = geom
    g = K.same_geom(N, H, W, Cin, Cout, 5, 2)
    g3 = K.s2d_geom(g)
    assert g3 is not None and (g3.H, g3.W, g3.Cin, g3.kh, g3.pad_t) == ((H + 1) // 2, (W + 1) // 2, 4 * Cin, 3, 1)
    x, dy = act((N, Cin, H, W), 1), act((N, Cout, g.Ho, g.Wo), 2)
    w = (torch.randn(5, 5, Cin, Cout, generator=torch.Generator().manual_seed(3)) * 0.05).contiguous()
    b = torch.randn(Cout, generator=torch.Generator().manual_seed(4))
    wq = w.to(torch.bfloat16).float()
    res = act((N, Cout, g.Ho, g.Wo), 5)
    # any call that reaches the SIMT C entry points would need the GPU: make it fail loudly, except the wgrad fall-through case
    simt = []
    monkeypatch.setattr(K, 'call', lambda name, *a: simt.append(name))

    y = K.conv_fprop(x, w, b, g)
    assert rel(y, fb.conv_fprop(x, wq, b, g)) < 1e-2
    y = K.conv_fprop(x, w, b, g, relu=True, residual=res, col=K.thin_col(x, g, 'x'))
    assert rel(y, fb.conv_fprop(x, wq, b, g, relu=True, residual=res)) < 1e-2
    dx = K.conv_dgrad(dy, w, g)
    assert tuple(dx.shape) == (N, Cin, H, W) and rel(dx, fb.conv_dgrad(dy, wq, g)) < 1e-2
    assert not simt
    if g3.Cin % 128 == 0 and g3.Cout % 128 == 0:
        dw = K.conv_wgrad(x, dy, g, tuple(w.shape))
        ref = fb.conv_wgrad(x, dy, g, tuple(w.shape))
        assert rel(dw, ref) < 1e-5
        acc = torch.ones_like(w)
        K.conv_wgrad(x, dy, g, tuple(w.shape), accumulate_into=acc, col=K.thin_col(x, g, 'x'))
        assert rel(acc - 1, ref) < 1e-4
        assert not simt
    else:
        K.conv_wgrad(x, dy, g, tuple(w.shape))
        assert simt == ['ctgan_conv_wgrad']


THIN_GEOMS = [
    # N, H, W, Cin, Cout, k  (stride 2, TF SAME)
    (2, 32, 32, 3, 128, 5),    # CIFAR Discriminator.1 / (as dgrad) Generator.5
    (3, 28, 28, 1, 64, 5),     # MNIST Discriminator.1 / Generator.5: Cout = 64 -> wgrad falls through
    (2, 14, 10, 3, 128, 5),    # non-square
    (2, 9, 9, 2, 128, 3),      # odd extent, 3x3
]


@pytest.mark.parametrize('geom', THIN_GEOMS)
def test_thin_strided_route_equals_stride2_conv(K, geom, monkeypatch):
    """Stride-2 convs with a thin input as im2col (128 columns per output pixel) + 1x1 tensor-core GEMMs."""
    N, H, W, Cin, Cout, k = geom
    g = K.same_geom(N, H, W, Cin, Cout, k, 2)
    assert K.thin_s2_ok(g) and K.s2d_geom(g) is None
    x, dy = act((N, Cin, H, W), 1), act((N, Cout, g.Ho, g.Wo), 2)
    w = (torch.randn(k, k, Cin, Cout, generator=torch.Generator().manual_seed(3)) * 0.1).contiguous()
    b = torch.randn(Cout, generator=torch.Generator().manual_seed(4))
    wq = w.to(torch.bfloat16).float()
    simt = []
    monkeypatch.setattr(K, 'call', lambda name, *a: simt.append(name))
    y = K.conv_fprop(x, w, b, g, col=K.thin_col(x, g, 'x'))
    assert rel(y, fb.conv_fprop(x, wq, b, g)) < 1e-2
    dx = K.conv_dgrad(dy, w, g)
    assert tuple(dx.shape) == (N, Cin, H, W) and rel(dx, fb.conv_dgrad(dy, wq, g)) < 1.5e-2
    assert not simt
    ref = fb.conv_wgrad(x, dy, g, tuple(w.shape))
    if Cout % 128 == 0:
        assert rel(K.conv_wgrad(x, dy, g, tuple(w.shape)), ref) < 1e-5
        acc = torch.ones_like(w)
        K.conv_wgrad(x, dy, g, tuple(w.shape), accumulate_into=acc, col=K.thin_col(x, g, 'x'))
        assert rel(acc - 1, ref) < 1e-4 and not simt
    else:
        K.conv_wgrad(x, dy, g, tuple(w.shape))
        assert simt == ['ctgan_conv_wgrad']


def test_s2d_param_packs_are_persistent_and_refreshed(K):
    """Parameter filters: the operand pair is packed once, republished in place after an optimizer step."""
    g = K.same_geom(2, 8, 8, 32, 128, 5, 2)
    w = torch.randn(5, 5, 32, 128) * 0.05
    p0 = K.pack_filter_s2d(w, g, 0, cacheable=True)
    assert K.pack_filter_s2d(w, g, 0, cacheable=True) is p0
    d0 = K.pack_filter_s2d(w, g, 1, cacheable=True)
    before = p0.clone()
    w.mul_(2.0)
    K.invalidate_weight_cache({w.data_ptr()})
    K.refresh_lazy_packs([w.data_ptr()])
    assert K.pack_filter_s2d(w, g, 0, cacheable=True) is p0 and K.pack_filter_s2d(w, g, 1, cacheable=True) is d0
    assert torch.equal(p0.float(), (before.float() * 2).to(torch.bfloat16).float())


def test_s2d_route_is_not_taken_when_ineligible(K):
    assert K.s2d_geom(K.same_geom(2, 32, 32, 3, 128, 5, 2)) is None          # 4*Cin = 12: not a multiple of 64
    assert K.s2d_geom(K.same_geom(2, 16, 16, 128, 128, 3, 1)) is None        # stride 1
    assert K.s2d_geom(K.same_geom(2, 16, 16, 128, 128, 5, 2), torch.zeros(1, 1, 1, 1)) is None   # fp32 activations
    K.config.use_s2d = False
    assert K.s2d_geom(K.same_geom(2, 16, 16, 128, 256, 5, 2)) is None


def test_tensor_core_branches_bind_their_entry_points(K, monkeypatch):
    """The stride-1 tensor-core branches and the space-to-depth primitives reach the C ABI with the argument counts
    include/ctgan_sm100.h declares (recorded, not executed: there is no GPU here)."""
    import ctgan_b200.kernels as KK
    from ctgan_b200 import _lib
    monkeypatch.undo()                                    # the real launch helpers, not the restatements above
    monkeypatch.setattr(KK, 'tc_available', lambda: True)
    monkeypatch.setattr(KK, '_chk', lambda t, name='tensor': None)
    monkeypatch.setattr(KK, '_stream', lambda: None)
    monkeypatch.setattr(KK.config, 'use_s2d', True)
    calls = []
    monkeypatch.setattr(KK, 'call', lambda name, *a: calls.append((name, len(a))))
    KK.invalidate_weight_cache()
    g = KK.same_geom(2, 8, 8, 128, 128, 3, 1)
    x, dy = act((2, 128, 8, 8), 1), act((2, 128, 8, 8), 2)
    w, b = torch.randn(3, 3, 128, 128), torch.zeros(128)
    KK.conv_fprop(x, w, b, g, relu=True, residual=dy)
    KK.conv_dgrad(dy, w, g)
    KK.conv_dgrad(dy, w, g, relu_mask=x)
    KK.conv_wgrad(x, dy, g, tuple(w.shape))
    g2 = KK.same_geom(2, 16, 16, 32, 128, 5, 2)
    x2, dy2, w2 = act((2, 32, 16, 16), 1), act((2, 128, 8, 8), 2), torch.randn(5, 5, 32, 128)
    KK.conv_fprop(x2, w2, b, g2)
    KK.conv_dgrad(dy2, w2, g2)
    KK.conv_wgrad(x2, dy2, g2, tuple(w2.shape))
    g3 = KK.same_geom(2, 32, 32, 3, 128, 5, 2)
    x3, dy3, w3 = act((2, 3, 32, 32), 1), act((2, 128, 16, 16), 2), torch.randn(5, 5, 3, 128)
    KK.conv_fprop(x3, w3, b, g3)
    KK.conv_dgrad(dy3, w3, g3)
    KK.conv_wgrad(x3, dy3, g3, tuple(w3.shape))
    names = [n for n, _ in calls]
    assert names == ['ctgan_pack_filter_bf16', 'ctgan_conv_fprop_tc',
                     'ctgan_pack_filter_bf16', 'ctgan_conv_fprop_tc',
                     'ctgan_pack_filter_bf16', 'ctgan_conv_fprop_tc_masked',
                     'ctgan_conv_wgrad_tc',
                     'ctgan_space_to_depth', 'ctgan_pack_filter_s2d', 'ctgan_conv_fprop_tc',
                     'ctgan_pack_filter_s2d', 'ctgan_conv_fprop_tc', 'ctgan_depth_to_space',
                     'ctgan_space_to_depth', 'ctgan_conv_wgrad_tc', 'ctgan_s2d_filter_grad',
                     'ctgan_im2col_strided', 'ctgan_pack_filter_padk', 'ctgan_conv_fprop_tc',
                     'ctgan_pack_filter_padk', 'ctgan_conv_fprop_tc', 'ctgan_col2im_strided',
                     'ctgan_im2col_strided', 'ctgan_conv_wgrad_tc', 'ctgan_add_prefix'], names
    for name, nargs in calls:
        assert nargs == len(_lib._PROTOS[name][1]), name
    KK.invalidate_weight_cache()
