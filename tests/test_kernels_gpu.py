"""GPU tests (`-m gpu`): every kernel of libctgan_sm100, called through the C ABI, against an
independent CPU evaluation of the same contract (tests/fake_backend.py = PyTorch-CPU,
tests/philox_ref.py = numpy Philox4x32-10).  Tolerances: float path 1e-4 relative
(accumulation-order only), BF16 path 1e-2 (one bf16 rounding of the output)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CL = torch.channels_last


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')


@pytest.fixture(scope='module')
def K():
    _need_gpu()
    import ctgan_b200.kernels as K
    return K


def FB():
    from tests import fake_backend
    return fake_backend


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def act(shape, dtype, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    t = (torch.randn(shape, generator=g) * scale).to(dtype)
    return t.contiguous(memory_format=CL) if len(shape) == 4 else t.contiguous()


def filt(shape, seed, scale=0.05):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).contiguous()      # HWIO, plain contiguous


def to_dev(t):
    if t.dim() == 4:
        return t.cuda().contiguous(memory_format=CL)
    return t.cuda()


GEOMS = [
    # N, H, W, Cin, Cout, k, stride
    (3, 32, 32, 3, 128, 5, 2),     # CIFAR D.1   (pad 1,2)
    (2, 16, 16, 128, 256, 5, 2),   # CIFAR D.2
    (3, 28, 28, 1, 64, 5, 2),      # MNIST D.1
    (2, 7, 7, 128, 256, 5, 2),     # MNIST D.3   (pad 2,2)
    (2, 32, 32, 3, 128, 3, 1),     # ResNet D.1.Conv1
    (2, 16, 16, 128, 128, 3, 1),   # ResNet body
    (3, 8, 8, 128, 128, 1, 1),     # shortcut
    (2, 32, 32, 128, 3, 3, 1),     # Generator.Output
    (5, 1, 1, 128, 2048, 1, 1),    # Generator.Input (linear)
    (7, 1, 1, 4096, 1, 1, 1),      # Discriminator.Output (GEMV)
]


@pytest.mark.parametrize('use_tc', [False, True])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('geom', GEOMS)
def test_conv_family(K, geom, dtype, use_tc):
    N, H, W, Cin, Cout, k, s = geom
    if use_tc and dtype != torch.bfloat16:
        pytest.skip('tensor-core path is BF16 only')
    g = K.same_geom(N, H, W, Cin, Cout, k, s)
    two_d = (H == 1 and W == 1)
    x = act((N, Cin) if two_d else (N, Cin, H, W), dtype, 1)
    dy = act((N, Cout) if two_d else (N, Cout, g.Ho, g.Wo), dtype, 2)
    w = filt((k, k, Cin, Cout), 3)
    b = act((Cout,), torch.float32, 4)
    K.config.use_tc = use_tc
    try:
        y = K.conv_fprop(to_dev(x), w.cuda(), b.cuda(), g)
        dx = K.conv_dgrad(to_dev(dy), w.cuda(), g)
        dw = K.conv_wgrad(to_dev(x), to_dev(dy), g, tuple(w.shape))
        db = K.bias_grad(to_dev(dy))
    finally:
        K.config.use_tc = True
    on_tc = use_tc and dtype == torch.bfloat16 and (K._tc_geom_ok(g) or K.s2d_geom(g) is not None)
    wq = w.to(torch.bfloat16).float() if on_tc else w     # the tensor-core routes round the filter to bf16
    fb = FB()
    tol = 1e-4 if dtype == torch.float32 else 1e-2
    assert y.shape == ((N, Cout) if two_d else (N, Cout, g.Ho, g.Wo)) and y.dtype == dtype
    assert rel(y, fb.conv_fprop(x, wq, b, g)) < tol
    assert rel(dx, fb.conv_dgrad(dy, wq, g)) < tol
    assert rel(dw, fb.conv_wgrad(x, dy, g, tuple(w.shape))) < 1e-4 + (2e-3 if dtype == torch.bfloat16 else 0)
    assert rel(db, fb.bias_grad(dy)) < 1e-4


def test_conv_mixed_dtype_heads(K):
    """Critic heads: BF16 features in, float logits out, and the matching backward dtypes."""
    g = K.ConvGeom(6, 1, 1, 128, 1, 1, 10, 1, 1, 1, 0, 0)
    x, w, b = act((6, 128), torch.bfloat16, 1), filt((1, 1, 128, 10), 2, 0.1), act((10,), torch.float32, 3)
    dy = act((6, 10), torch.float32, 4)
    fb = FB()
    y = K.conv_fprop(x.cuda(), w.cuda(), b.cuda(), g, out_dtype=torch.float32)
    assert y.dtype == torch.float32 and rel(y, fb.conv_fprop(x, w, b, g, out_dtype=torch.float32)) < 1e-5
    dx = K.conv_dgrad(dy.cuda(), w.cuda(), g, out_dtype=torch.bfloat16)
    assert dx.dtype == torch.bfloat16 and rel(dx, fb.conv_dgrad(dy, w, g, out_dtype=torch.bfloat16)) < 1e-2
    dw = K.conv_wgrad(x.cuda(), dy.cuda(), g, tuple(w.shape))
    assert rel(dw, fb.conv_wgrad(x, dy, g, tuple(w.shape))) < 1e-5


@pytest.mark.parametrize('shape', [(24, 32, 32, 3), (80, 16, 16, 3), (20, 32, 32, 5), (300, 8, 8, 3), (3, 16, 16, 1),
                                   (200, 4, 4, 3), (37, 32, 32, 3), (64, 32, 32, 3), (200, 16, 16, 3), (60, 20, 32, 3),
                                   (41, 28, 16, 3), (301, 8, 8, 3), (1190, 4, 4, 3)])
def test_tc_fprop_variants_match(K, shape):
    """The fprop_tc kernel family -- 256-pixel work items / persistent grouped stages (default) vs one-tile-per-CTA, with
    and without the halo-reuse A pipeline -- all compute the same convolution (up to the bf16 rounding of a different accumulation order)
    and match the CPU reference; also as dgrad (flipped filter pack) and with the residual/ReLU epilogue."""
    from ctgan_b200 import _lib
    N, H, W, k = shape
    g = K.same_geom(N, H, W, 128, 128, k, 1)
    x, dy, r = act((N, 128, H, W), torch.bfloat16, 1), act((N, 128, H, W), torch.bfloat16, 2), act((N, 128, H, W), torch.bfloat16, 5)
    w, b = filt((k, k, 128, 128), 3), act((128,), torch.float32, 4)
    wq = w.to(torch.bfloat16).float()
    ref_f, ref_d = FB().conv_fprop(x, wq, b, g), FB().conv_dgrad(dy, wq, g)
    ref_r = FB().conv_fprop(x, wq, b, g, relu=True, residual=r)
    try:
        for variant in (4, 3, 1):
            for halo in (1, 0):
                _lib.lib.ctgan_set_fprop_variant(variant)
                _lib.lib.ctgan_set_fprop_halo(halo)
                yf = K.conv_fprop(to_dev(x), w.cuda(), b.cuda(), g)
                yd = K.conv_dgrad(to_dev(dy), w.cuda(), g)
                yr = K.conv_fprop(to_dev(x), w.cuda(), b.cuda(), g, relu=True, residual=to_dev(r))
                assert rel(yf, ref_f) < 1e-2 and rel(yd, ref_d) < 1e-2 and rel(yr, ref_r) < 1e-2, (variant, halo)
    finally:
        _lib.lib.ctgan_set_fprop_variant(4)
        _lib.lib.ctgan_set_fprop_halo(1)


@pytest.mark.parametrize('shape', [(64, 8, 8, 3, 128, 128), (128, 8, 8, 3, 128, 128), (40, 8, 8, 3, 128, 128), (30, 4, 4, 3, 128, 128),
                                   (32, 8, 8, 3, 128, 256), (9, 8, 8, 3, 256, 128), (60, 8, 8, 1, 512, 128), (24, 4, 4, 3, 1024, 512)])
def test_tc_splitk_cluster_kernel(K, shape):
    """Layers with fewer output tiles than half the SMs: K split over a cluster of 2 / 4 CTAs, partial accumulators
    reduce-scattered through distributed shared memory (csrc/conv_splitk.cu) -- against the CPU reference and the
    one-CTA-per-tile kernel, as fprop (+bias), dgrad, with the residual + ReLU epilogue and with the ReLU-backward mask."""
    from ctgan_b200 import _lib
    N, H, W, k, Cin, Cout = shape
    g = K.same_geom(N, H, W, Cin, Cout, k, 1)
    x, dy = act((N, Cin, H, W), torch.bfloat16, 1), act((N, Cout, H, W), torch.bfloat16, 2)
    r, mk = act((N, Cout, H, W), torch.bfloat16, 5), act((N, Cin, H, W), torch.bfloat16, 6)
    w, b = filt((k, k, Cin, Cout), 3), act((Cout,), torch.float32, 4)
    wq = w.to(torch.bfloat16).float()
    ref_f, ref_d = FB().conv_fprop(x, wq, b, g), FB().conv_dgrad(dy, wq, g)
    ref_r = FB().conv_fprop(x, wq, b, g, relu=True, residual=r)
    ref_m = FB().conv_dgrad(dy, wq, g, relu_mask=mk)
    outs = {}
    try:
        for on in (1, 0):
            _lib.lib.ctgan_set_splitk(on)
            yf = K.conv_fprop(to_dev(x), w.cuda(), b.cuda(), g)
            yd = K.conv_dgrad(to_dev(dy), w.cuda(), g)
            yr = K.conv_fprop(to_dev(x), w.cuda(), b.cuda(), g, relu=True, residual=to_dev(r))
            ym = K.conv_dgrad(to_dev(dy), w.cuda(), g, relu_mask=to_dev(mk))
            assert rel(yf, ref_f) < 1e-2 and rel(yd, ref_d) < 1e-2 and rel(yr, ref_r) < 1e-2 and rel(ym, ref_m) < 1e-2, on
            outs[on] = (yf, yd, yr, ym)
        for a, c in zip(outs[1], outs[0]):
            assert rel(a, c) < 4e-3              # same products, different summation order, one bf16 rounding
        # deterministic: no atomics anywhere on this path
        assert torch.equal(K.conv_fprop(to_dev(x), w.cuda(), b.cuda(), g), outs[0][0])
        _lib.lib.ctgan_set_splitk(1)
        assert torch.equal(K.conv_fprop(to_dev(x), w.cuda(), b.cuda(), g), outs[1][0])
    finally:
        _lib.lib.ctgan_set_splitk(1)


@pytest.mark.parametrize('shape', [(6, 32, 32, 3, 128, 128), (20, 16, 16, 3, 128, 256), (70, 8, 8, 3, 256, 128), (9, 16, 16, 1, 128, 128),
                                   (130, 1, 1, 1, 128, 384), (37, 8, 8, 3, 128, 128), (33, 4, 4, 3, 128, 128)])
def test_tf32_conv_family(K, shape):
    """The fp32-storage path on the tensor cores (tcgen05 kind::tf32, csrc/conv_tf32.cu): fprop (+bias, +residual, ReLU),
    dgrad (+ReLU-backward mask) and the multi-job wgrad against the PyTorch-CPU fp32 reference.  Tolerance 3e-3: every
    product is formed from operands rounded to TF32 (2^-11), accumulation is fp32."""
    N, H, W, k, Cin, Cout = shape
    g = K.same_geom(N, H, W, Cin, Cout, k, 1)
    two_d = H == 1
    xs, ys = ((N, Cin), (N, Cout)) if two_d else ((N, Cin, H, W), (N, Cout, H, W))
    x, dy, r, mk = act(xs, torch.float32, 1), act(ys, torch.float32, 2), act(ys, torch.float32, 5), act(xs, torch.float32, 6)
    w, b = filt((k, k, Cin, Cout), 3), act((Cout,), torch.float32, 4)
    K.config.tf32 = True
    try:
        assert K._tf32_geom_ok(g)
        yf = K.conv_fprop(to_dev(x), w.cuda(), b.cuda(), g)
        assert yf.dtype == torch.float32 and rel(yf, FB().conv_fprop(x, w, b, g)) < 3e-3
        if not two_d:
            yr = K.conv_fprop(to_dev(x), w.cuda(), b.cuda(), g, relu=True, residual=to_dev(r))
            assert rel(yr, FB().conv_fprop(x, w, b, g, relu=True, residual=r)) < 3e-3
            ym = K.conv_dgrad(to_dev(dy), w.cuda(), g, relu_mask=to_dev(mk))
            assert rel(ym, FB().conv_dgrad(dy, w, g, relu_mask=mk)) < 3e-3
        yd = K.conv_dgrad(to_dev(dy), w.cuda(), g)
        assert rel(yd, FB().conv_dgrad(dy, w, g)) < 3e-3
        ref_w = FB().conv_wgrad(x, dy, g, (k, k, Cin, Cout))
        if K._wgrad_tf32_ok(g):
            dw = K.conv_wgrad(to_dev(x), to_dev(dy), g, (k, k, Cin, Cout))
            assert rel(dw, ref_w) < 3e-3
            acc = torch.full((k, k, Cin, Cout), 0.5, device='cuda')
            xd, dyd = to_dev(x), to_dev(dy)
            K.conv_wgrad(xd, dyd, g, (k, k, Cin, Cout), accumulate_into=acc, defer=True)
            assert len(K._wgrad_queue32) == 1 and float((acc - 0.5).abs().max()) == 0.0
            K.join_side()
            assert not K._wgrad_queue32 and rel(acc - 0.5, ref_w) < 3e-3
        else:
            assert H == 4                                    # several images per 64-pixel chunk: SIMT fallback
            assert rel(K.conv_wgrad(to_dev(x), to_dev(dy), g, (k, k, Cin, Cout)), ref_w) < 1e-4
        # the SIMT fp32 kernels (config.tf32 off) agree to tf32 rounding
        K.config.tf32 = False
        assert rel(K.conv_fprop(to_dev(x), w.cuda(), b.cuda(), g), yf) < 3e-3
    finally:
        K.config.tf32 = False


def test_tc_residual_relu_epilogue(K):
    g = K.same_geom(3, 8, 8, 128, 128, 3, 1)
    x, r = act((3, 128, 8, 8), torch.bfloat16, 1), act((3, 128, 8, 8), torch.bfloat16, 2)
    w, b = filt((3, 3, 128, 128), 3), act((128,), torch.float32, 4)
    y = K.conv_fprop(to_dev(x), w.cuda(), b.cuda(), g, relu=True, residual=to_dev(r))
    ref = FB().conv_fprop(x, w.to(torch.bfloat16).float(), b, g, relu=True, residual=r)
    assert rel(y, ref) < 1e-2


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_elementwise_and_layout(K, dtype):
    fb = FB()
    tol = 1e-6 if dtype == torch.float32 else 1e-2
    a, b = act((3, 16, 6, 6), dtype, 1), act((3, 16, 6, 6), dtype, 2)
    A, B = to_dev(a), to_dev(b)
    assert rel(K.add(A, B), fb.add(a, b)) < tol
    assert rel(K.mul(A, B), fb.mul(a, b)) < tol
    assert rel(K.scale(A, 0.37), fb.scale(a, 0.37)) < tol
    bias = act((16,), torch.float32, 3)
    assert rel(K.bias_add(A, bias.cuda()), fb.bias_add(a, bias)) < tol
    assert rel(K.pool2x2(A, 0.25), fb.pool2x2(a, 0.25)) < tol
    assert rel(K.upsample2x(A, 1.0), fb.upsample2x(a, 1.0)) < tol
    assert rel(K.spatial_sum(A, 1 / 36.), fb.spatial_sum(a, 1 / 36.)) < tol
    y2 = act((3, 16), dtype, 4)
    assert rel(K.spatial_bcast(y2.cuda(), 6, 6, 0.5), fb.spatial_bcast(y2, 6, 6, 0.5)) < tol
    assert rel(K.crop(A, 5, 4), fb.crop(a, 5, 4)) < tol
    c = act((3, 16, 5, 4), dtype, 5)
    assert rel(K.crop_bwd(to_dev(c), 6, 6), fb.crop_bwd(c, 6, 6)) < tol
    for kind in (0, 1):
        yk = K.unary_fwd(A, kind)
        assert rel(yk, fb.unary_fwd(a, kind)) < max(tol, 1e-5)
        assert rel(K.unary_bwd(yk, B, kind), fb.unary_bwd(yk.cpu(), b, kind)) < max(tol, 1e-5)
    flat = act((3, 16 * 36), torch.float32, 6)
    nh = K.nchw_to_nhwc(flat.cuda(), 3, 16, 6, 6, dtype)
    assert nh.is_contiguous(memory_format=CL) and rel(nh, fb.nchw_to_nhwc(flat, 3, 16, 6, 6, dtype)) < tol
    back = K.nhwc_to_nchw(nh, torch.float32, (3, 16 * 36))
    assert rel(back, flat.to(dtype).float()) < 1e-7
    assert rel(K.cast(A, torch.float32), a.float()) == 0.0


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('slope,keep', [(0.0, 1.0), (0.2, 0.5), (1.0, 0.8)])
def test_act_dropout_matches_philox_reference(K, dtype, slope, keep):
    """The in-kernel Philox stream == ctgan_philox_uniform's materialisation == the numpy reference."""
    fb = FB()
    x = act((4, 8, 5, 5), dtype, 1)
    seed, off = 0x1234ABCD5678, 4096
    y, m = K.act_dropout(to_dev(x), slope, keep, seed=seed, offset=off)
    yr, mr = fb.act_dropout(x, slope, keep, seed=seed, offset=off)
    assert torch.equal(m.cpu().float(), mr.float())
    assert rel(y, yr) < (1e-6 if dtype == torch.float32 else 1e-2)
    u = K.philox_uniform((4, 8, 5, 5), 'cuda', seed, off, memory_format=CL)
    y2, m2 = K.act_dropout(to_dev(x), slope, keep, u=u)
    assert torch.equal(m2, m) and torch.equal(y2, y)
    dyn = torch.tensor([off], dtype=torch.int64, device='cuda')
    y3, _ = K.act_dropout(to_dev(x), slope, keep, seed=seed, offset=0, dyn=dyn)
    assert torch.equal(y3, y)


@pytest.mark.parametrize('geom', [(6, 16, 16, 128, 128, 3, 1), (70, 8, 8, 128, 256, 3, 1), (5, 32, 32, 3, 128, 5, 2),
                                  (9, 16, 16, 128, 256, 5, 2), (7, 14, 14, 64, 128, 5, 2), (11, 8, 8, 256, 512, 5, 2),
                                  (300, 8, 8, 128, 128, 3, 1), (200, 8, 8, 256, 512, 5, 2)])   # last two: [h][n][w] halo boxes (no split-K)
@pytest.mark.parametrize('slope,keep', [(0.2, 0.5), (0.2, 1.0), (1.0, 0.8)])
def test_conv_actdrop_epilogue(K, geom, slope, keep):
    """Conv2D -> LeakyReLU -> dropout in the tcgen05 conv epilogue (stride-1, space-to-depth and strided-im2col routes):
    the multiplier is (v > 0 ? 1 : slope) * floor(keep + u) / keep with u = numpy Philox at the element's NHWC index
    (offset + device counter), y = v * m, also when written in the space-to-depth layout of the next stride-2 layer."""
    from tests import philox_ref
    N, H, W, Cin, Cout, k, stride = geom
    g = K.same_geom(N, H, W, Cin, Cout, k, stride)
    x, w, b = act((N, Cin, H, W), torch.bfloat16, 1), filt((k, k, Cin, Cout), 3), act((Cout,), torch.float32, 4)
    xd = to_dev(x)
    assert K.conv_actdrop_route(xd, g) == ('tc' if stride == 1 else ('padk' if Cin == 3 else 's2d'))
    v = FB().conv_fprop(x, w.to(torch.bfloat16).float(), b, g, out_dtype=torch.float32)      # [N, Cout, Ho, Wo]
    seed, off, dynv = 0xABCDEF0123, 8192, 1028
    n_el = N * g.Ho * g.Wo * Cout
    u = torch.from_numpy(philox_ref.uniform(seed, off + dynv, n_el)).reshape(N, g.Ho, g.Wo, Cout).permute(0, 3, 1, 2)
    mult = torch.where(v > 0, torch.ones_like(v), torch.full_like(v, slope))
    if keep < 1.0:
        mult = mult * torch.floor(keep + u) / keep
    mr = mult.to(torch.bfloat16).float()
    dyn = torch.tensor([dynv], dtype=torch.int64, device='cuda')
    y, m = K.conv_fprop_actdrop(xd, w.cuda(), b.cuda(), g, slope, keep, seed, off, dyn=dyn)
    sure = v.abs() > 2e-2 * v.abs().mean()                    # away from the sign boundary the multiplier is exact
    assert torch.equal(m.cpu().float()[sure], mr[sure])
    assert float((m.cpu().float() != mr).float().mean()) < 2e-2
    assert rel(y.cpu().float()[sure], (v * mr)[sure]) < 1e-2
    assert rel(y.float(), v.cuda() * m.float()) < 1e-2        # y == v * m for the stored multiplier
    if g.Ho % 2 == 0 and g.Wo % 2 == 0:
        ys, ms = K.conv_fprop_actdrop(xd, w.cuda(), b.cuda(), g, slope, keep, seed, off, dyn=dyn, out_s2d=True)
        og = K.ConvGeom(N, g.Ho, g.Wo, Cout, g.Ho, g.Wo, Cout, 1, 1, 1, 0, 0)
        assert tuple(ys.shape) == (N, 4 * Cout, g.Ho // 2, g.Wo // 2)
        assert torch.equal(K.depth_to_space(ys, og), y) and torch.equal(K.depth_to_space(ms, og), m)
        # layout kernels with the multiplier: the backward / double backward of the fused activation
        gq = to_dev(act((N, 4 * Cout, g.Ho // 2, g.Wo // 2), torch.bfloat16, 9))
        d = K.depth_to_space(gq, og, mul=ms)
        assert rel(d, K.depth_to_space(gq, og).float() * m.float()) < 1e-2
        c = to_dev(act((N, Cout, g.Ho, g.Wo), torch.bfloat16, 10))
        sq = K.space_to_depth(c, og, mul=ms)
        assert rel(sq, K.space_to_depth(c, og).float() * ms.float()) < 1e-2


def test_philox_streams_match_reference(K):
    from tests import philox_ref
    seed = 987654321987
    u = K.philox_uniform((1000,), 'cuda', seed, 12, lo=-1.0, hi=3.0).cpu().numpy()
    np.testing.assert_allclose(u, -1.0 + 4.0 * philox_ref.uniform(seed, 12, 1000), rtol=0, atol=1e-6)
    z = K.philox_normal((501,), 'cuda', seed, 40).cpu().numpy()
    np.testing.assert_allclose(z, philox_ref.normal(seed, 40, 501), rtol=1e-4, atol=1e-5)
    assert abs(float(z.mean())) < 0.2 and 0.8 < float(z.std()) < 1.2
    lab = K.philox_labels(777, 'cuda', 10, seed, 8).cpu().numpy()
    np.testing.assert_array_equal(lab, (philox_ref.uniform(seed, 8, 777) * np.float32(10)).astype('int32'))
    ctr = torch.zeros(1, dtype=torch.int64, device='cuda')
    K.counter_add(ctr, 12)
    u2 = K.philox_uniform((1000,), 'cuda', seed, 0, lo=-1.0, hi=3.0, dyn=ctr).cpu().numpy()
    np.testing.assert_array_equal(u, u2)


def test_prep_real_and_interpolate(K):
    fb = FB()
    xi = torch.randint(0, 256, (5, 3072), dtype=torch.int32)
    np.testing.assert_allclose(K.prep_real(xi.cuda(), 255.).cpu().numpy(), fb.prep_real(xi, 255.).numpy(), atol=1e-7)
    got = K.prep_real(xi.cuda(), 256., 1. / 128, seed=5, offset=16).cpu()
    ref = fb.prep_real(xi, 256., 1. / 128, seed=5, offset=16)
    np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=1e-7)
    assert float(got.min()) >= -1.0 and float(got.max()) < 1.0
    got8 = K.prep_real(xi.to(torch.uint8).cuda(), 256., 1. / 128, seed=5, offset=16).cpu()       # the loaders' uint8 pixels
    assert torch.equal(got8, got)
    r, f, a = torch.randn(5, 64), torch.randn(5, 64), torch.rand(5, 1)
    np.testing.assert_allclose(K.interpolate(r.cuda(), f.cuda(), a.cuda()).cpu().numpy(),
                               fb.interpolate(r, f, a).numpy(), atol=1e-6)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('shape,cond', [((32, 128, 8, 8), True), ((6, 128, 32, 32), False), ((64, 8192), False),
                                        ((5, 256, 8, 8), False), ((64, 128, 32, 32), True), ((10, 128, 4, 4), True)])
@pytest.mark.parametrize('groups', [1, 2])
@pytest.mark.parametrize('relu', [False, True])
@pytest.mark.parametrize('fused', [True, False])
def test_batch_norm(K, dtype, shape, cond, relu, groups, fused):
    """Both batch-norm paths (two-kernel BF16 'fused' path with red.global sums; generic three-kernel Welford path)
    against the PyTorch-CPU definition; mean offset 0.5 with std 2 exercises the shifted sums."""
    fb = FB()
    if shape[0] % groups:
        pytest.skip('batch not divisible by groups')
    if fused and dtype != torch.bfloat16:
        pytest.skip('the fused path is BF16 only')
    K.config.use_bn_fused = fused
    try:
        C = shape[1]
        x = act(shape, dtype, 1, 2.0) + 0.5
        x = x.contiguous(memory_format=CL) if len(shape) == 4 else x
        nl = 10 if cond else 1
        gamma, beta = act((nl, C), torch.float32, 2) + 1.0, act((nl, C), torch.float32, 3)
        labels = torch.randint(0, 10, (shape[0],), dtype=torch.int32) if cond else None
        dy = act(shape, dtype, 4)
        lab = labels.cuda() if cond else None
        if fused and not K.bn_fused_ok(to_dev(x), groups):
            pytest.skip('shape not eligible for the fused path')
        assert K.bn_fused_ok(to_dev(x), groups) == fused
        y, mean, invstd = K.bn_fwd(to_dev(x), gamma.cuda(), beta.cuda(), lab, 1e-5, relu, groups)
        yr, mr, ir = fb.bn_fwd(x, gamma, beta, labels, 1e-5, relu, groups)
        tol = 2e-5 if dtype == torch.float32 else 1e-2
        assert rel(mean, mr) < 1e-5 and rel(invstd, ir) < 1e-5 and rel(y, yr) < tol
        dx, dg, db = K.bn_bwd(to_dev(dy), to_dev(x), y, gamma.cuda(), beta.cuda(), lab, mean, invstd, relu, groups)
        dxr, dgr, dbr = fb.bn_bwd(dy, x, y.cpu(), gamma, beta, labels, mean.cpu(), invstd.cpu(), relu, groups)
        assert rel(dx, dxr) < (1e-4 if dtype == torch.float32 else 2e-2)
        assert rel(dg, dgr) < 1e-4 and rel(db, dbr) < 1e-4
        # parameter gradients accumulated in place (the flat gradient bucket)
        ag, ab = torch.full_like(dg, 0.5), torch.full_like(db, -0.25)
        dx2, _, _ = K.bn_bwd(to_dev(dy), to_dev(x), y, gamma.cuda(), beta.cuda(), lab, mean, invstd, relu, groups, accumulate_into=(ag, ab))
        assert rel(dx2, dxr) < (1e-4 if dtype == torch.float32 else 2e-2)
        assert rel(ag - 0.5, dgr) < 1e-4 and rel(ab + 0.25, dbr) < 1e-4
        if len(shape) == 4:
            # output written 2x nearest-neighbour upsampled; the gradient arrives with that shape
            N, _, H, W = shape
            yu, mean_u, invstd_u = K.bn_fwd(to_dev(x), gamma.cuda(), beta.cuda(), lab, 1e-5, relu, groups, up2=True)
            assert tuple(yu.shape) == (N, C, 2 * H, 2 * W) and yu.is_contiguous(memory_format=CL)
            assert rel(yu, fb.upsample2x(yr.to(dtype), 1.0)) < tol
            dyu = act((N, C, 2 * H, 2 * W), dtype, 5)
            dxu, dgu, dbu = K.bn_bwd(to_dev(dyu), to_dev(x), yu, gamma.cuda(), beta.cuda(), lab, mean_u, invstd_u, relu, groups, up2=True)
            dxr, dgr, dbr = fb.bn_bwd(dyu, x, fb.upsample2x(y.cpu(), 1.0), gamma, beta, labels, mean.cpu(), invstd.cpu(), relu, groups, up2=True)
            assert rel(dxu, dxr) < (1e-4 if dtype == torch.float32 else 2e-2)
            # (the bf16 reference rounds the 2x2-pooled gradient to bf16 before reducing; the kernel sums the four loads in fp32)
            assert rel(dgu, dgr) < (1e-4 if dtype == torch.float32 else 1e-2) and rel(dbu, dbr) < (1e-4 if dtype == torch.float32 else 1e-2)
    finally:
        K.config.use_bn_fused = True


@pytest.mark.parametrize('feat_dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('with_logits', [False, True])
def test_ct_gp_loss(K, feat_dtype, with_logits):
    from ctgan_b200._lib import LossDesc, F32, BF16
    fb = FB()
    B, NF, Fd, Pd = 16, 16, 640, 3072
    g = torch.Generator().manual_seed(0)
    d_real, d_real2, d_fake = torch.randn(B, generator=g), torch.randn(B, generator=g), torch.randn(NF, generator=g)
    f1, f2 = act((B, Fd), feat_dtype, 1), act((B, Fd), feat_dtype, 2)
    grad = torch.randn(B, Pd, generator=g) * 0.02
    logits = torch.randn(B, 10, generator=g) if with_logits else None
    labels = torch.randint(0, 10, (B,), dtype=torch.int32) if with_logits else None
    desc = LossDesc(B, NF, Fd, Pd, 10 if with_logits else 0, BF16 if feat_dtype == torch.bfloat16 else F32,
                    10.0, 2.0, 0.05, 1.0 if with_logits else 0.0)
    c = lambda t: t.cuda() if t is not None else None
    out, per = K.ct_gp_loss_fwd(desc, c(d_real), c(d_real2), c(d_fake), c(f1), c(f2), c(grad), c(logits), c(labels))
    outr, perr = fb.ct_gp_loss_fwd(desc, d_real, d_real2, d_fake, f1, f2, grad, logits, labels)
    np.testing.assert_allclose(out.cpu().numpy()[:5], outr.numpy()[:5], rtol=2e-5, atol=1e-6)
    gcost = torch.tensor([0.7])
    gs = K.ct_gp_loss_bwd(desc, c(gcost), c(d_real), c(d_real2), c(f1), c(f2), c(grad), c(logits), c(labels), per)
    gr = fb.ct_gp_loss_bwd(desc, gcost, d_real, d_real2, f1, f2, grad, logits, labels, perr)
    for a, b in zip(gs, gr):
        if b is None:
            assert a is None
        else:
            assert rel(a, b) < (1e-2 if b.dtype == torch.bfloat16 else 2e-5)


def test_generator_loss_pieces(K):
    fb = FB()
    d = torch.randn(64)
    lg, lab = torch.randn(64, 10), torch.randint(0, 10, (64,), dtype=torch.int32)
    gc = torch.tensor([1.3])
    np.testing.assert_allclose(K.mean_fwd(d.cuda(), -1.0).cpu().numpy(), fb.mean_fwd(d, -1.0).numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(K.mean_bwd(gc.cuda(), 64, -1.0).cpu().numpy(), fb.mean_bwd(gc, 64, -1.0).numpy(), rtol=1e-6)
    np.testing.assert_allclose(K.softmax_ce_fwd(lg.cuda(), lab.cuda()).cpu().numpy(), fb.softmax_ce_fwd(lg, lab).numpy(), rtol=1e-5)
    assert rel(K.softmax_ce_bwd(lg.cuda(), lab.cuda(), gc.cuda(), 0.1), fb.softmax_ce_bwd(lg, lab, gc, 0.1)) < 1e-5


def test_adam_tf_semantics(K):
    from oracle.tf_ops import TFAdam
    n = 100003
    g0 = torch.Generator().manual_seed(0)
    p = torch.randn(n, generator=g0)
    ref_p = {'w': p.clone().double()}
    opt = TFAdam(0.5, 0.9)
    P, M, V = p.cuda(), torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
    lr = 1e-4
    for t in range(1, 4):
        g = torch.randn(n, generator=g0) * 10 ** float(-t)
        opt.apply(ref_p, {'w': g.double()}, lr)
        lr_t = lr * np.sqrt(1 - 0.9 ** t) / (1 - 0.5 ** t)
        K.adam_step(P, g.cuda(), M, V, lr_t, 0.5, 0.9, 1e-8)
    assert rel(P.cpu() - p, ref_p['w'] - p.double()) < 1e-3     # update recovered from fp32 params
    assert rel(M, opt.m['w']) < 1e-6 and rel(V, opt.v['w']) < 1e-6
    # device-resident lr and gradient pre-scale
    P2, M2, V2 = p.cuda(), torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
    g = torch.randn(n, generator=g0)
    K.adam_step(P2, (2 * g).cuda(), M2, V2, 123.0, 0.0, 0.9, 1e-8, grad_scale=0.5, lr_t_dev=torch.tensor([1e-3], device='cuda'))
    P3, M3, V3 = p.cuda(), torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
    K.adam_step(P3, g.cuda(), M3, V3, 1e-3, 0.0, 0.9, 1e-8)
    assert torch.equal(P2, P3)


def test_bad_descriptors_fail_loudly(K):
    from ctgan_b200._lib import CtganError
    g = K.ConvGeom(1, 4, 4, 8, 4, 4, 8, 3, 3, 1, 1, 1)
    x = act((1, 8, 4, 4), torch.float32).cuda().contiguous(memory_format=CL)
    with pytest.raises(CtganError):
        K.conv_fprop(x, torch.zeros(3, 3, 8, 8, device='cuda'), None, g._replace(stride=0))
    with pytest.raises(RuntimeError):
        K.conv_fprop(x.cpu(), torch.zeros(3, 3, 8, 8), None, g)           # no CPU path
    with pytest.raises(CtganError):
        K.act_dropout(x, 0.2, 0.0)


@pytest.mark.parametrize('shape', [(24, 32, 32, 3, 128, 128), (80, 16, 16, 3, 128, 256), (300, 8, 8, 3, 256, 128),
                                   (37, 32, 32, 3, 128, 128), (200, 4, 4, 3, 128, 128), (9, 16, 16, 1, 128, 128)])
def test_tc_wgrad_variants_match(K, shape):
    """wgrad_tc: filter-column CTAs sharing one x halo box (default for 3x3) vs one box pair per tap; both match
    the CPU reference, also when accumulating into a pre-filled gradient buffer."""
    from ctgan_b200 import _lib
    N, H, W, k, Cin, Cout = shape
    g = K.same_geom(N, H, W, Cin, Cout, k, 1)
    x, dy = act((N, Cin, H, W), torch.bfloat16, 1), act((N, Cout, H, W), torch.bfloat16, 2)
    ref = FB().conv_wgrad(x, dy, g, (k, k, Cin, Cout))
    try:
        for variant in (2, 1):
            _lib.lib.ctgan_set_wgrad_variant(variant)
            dw = K.conv_wgrad(to_dev(x), to_dev(dy), g, (k, k, Cin, Cout))
            assert rel(dw, ref) < 2e-3, variant
            acc = torch.full((k, k, Cin, Cout), 0.5, device='cuda')
            K.conv_wgrad(to_dev(x), to_dev(dy), g, (k, k, Cin, Cout), accumulate_into=acc)
            assert rel(acc - 0.5, ref) < 2e-3, variant
    finally:
        _lib.lib.ctgan_set_wgrad_variant(2)


@pytest.mark.parametrize('balance', [5, 0, 3])
@pytest.mark.parametrize('chunk', [0, 64, 128])
@pytest.mark.parametrize('items_per_sm', [2, 1, 5])
def test_tc_wgrad_multi_job_launch(K, items_per_sm, chunk, balance):
    """Deferred filter gradients (csrc/conv_wgrad_multi.cu): a mix of layers -- 3x3 at 32x32 / 16x16 / 8x8 / 4x4, 1x1, a Linear,
    ragged batch sizes, wide Cin / Cout, two jobs adding into the SAME gradient -- queued and run as one launch, against
    the CPU reference of each job; the queue is empty afterwards and ineligible jobs launch immediately."""
    from ctgan_b200 import _lib
    shapes = [(64, 8, 8, 3, 128, 128), (37, 8, 8, 3, 128, 128), (20, 16, 16, 3, 128, 256), (7, 32, 32, 3, 256, 128),
              (50, 8, 8, 1, 128, 128), (130, 1, 1, 1, 128, 384), (40, 4, 4, 3, 128, 128), (33, 4, 4, 3, 256, 128),
              (20, 16, 16, 3, 64, 64), (9, 8, 8, 3, 64, 128), (12, 8, 8, 1, 192, 64),      # 64 (mod 128) channels
              (3, 64, 64, 3, 64, 128),                                                  # 64-pixel-wide image: 24 KB halo boxes
              (64, 8, 8, 3, 128, 128)]       # 4x4: four images per 64-pixel chunk, [h][n][w] halo boxes
    _lib.lib.ctgan_set_wgrad_multi_items_per_sm(items_per_sm)
    _lib.lib.ctgan_set_wgrad_multi_chunk(chunk)            # pixels per pipeline stage: 64 / 128 / by image width
    _lib.lib.ctgan_set_wgrad_multi_balance(balance, -1)    # item -> CTA assignment: longest-first / round robin / rotated
    try:
        refs, accs, keep = [], [], []
        for i, (N, H, W, k, Cin, Cout) in enumerate(shapes):
            g = K.same_geom(N, H, W, Cin, Cout, k, 1)
            two_d = H == 1
            x = act((N, Cin) if two_d else (N, Cin, H, W), torch.bfloat16, 10 + i)
            dy = act((N, Cout) if two_d else (N, Cout, H, W), torch.bfloat16, 30 + i)
            ref = FB().conv_wgrad(x, dy, g, (k, k, Cin, Cout))
            if i == len(shapes) - 1:                       # second contribution to job 0's gradient
                acc, ref = accs[0], None
                refs[0] = refs[0] + FB().conv_wgrad(x, dy, g, (k, k, Cin, Cout))
            else:
                acc = torch.full((k, k, Cin, Cout), 0.25, device='cuda')
            xd, dyd = to_dev(x), to_dev(dy)
            assert K.wgrad_deferrable(xd, dyd, g)
            K.conv_wgrad(xd, dyd, g, (k, k, Cin, Cout), accumulate_into=acc, defer=True)
            keep.append((xd, dyd))
            if ref is not None:
                refs.append(ref); accs.append(acc)
        assert len(K._wgrad_queue) == len(shapes)
        for acc in accs:
            assert float((acc - 0.25).abs().max()) == 0.0        # nothing has run yet
        K.join_side()
        assert not K._wgrad_queue
        for i, (acc, ref) in enumerate(zip(accs, refs)):
            assert rel(acc - 0.25, ref) < 2e-3, (i, shapes[i])
        # a 5x5 stride-1 filter is not a job of the multi-launch kernel (3x3 / 1x1 only) -> not deferrable, runs at once
        g = K.same_geom(40, 8, 8, 128, 128, 5, 1)
        x, dy = act((40, 128, 8, 8), torch.bfloat16, 1), act((40, 128, 8, 8), torch.bfloat16, 2)
        assert not K.wgrad_deferrable(to_dev(x), to_dev(dy), g)
        acc = torch.zeros((5, 5, 128, 128), device='cuda')
        K.conv_wgrad(to_dev(x), to_dev(dy), g, (5, 5, 128, 128), accumulate_into=acc, defer=True)
        assert not K._wgrad_queue
        assert rel(acc, FB().conv_wgrad(x, dy, g, (5, 5, 128, 128))) < 2e-3
    finally:
        _lib.lib.ctgan_set_wgrad_multi_items_per_sm(2)
        _lib.lib.ctgan_set_wgrad_multi_chunk(0)
        _lib.lib.ctgan_set_wgrad_multi_balance(1, -1)


@pytest.mark.parametrize('geom', [(5, 32, 32, 3, 128, 3), (3, 16, 16, 3, 128, 1), (4, 32, 32, 128, 3, 3), (70, 8, 8, 3, 256, 3),
                                  (3, 64, 64, 3, 64, 3), (2, 64, 64, 64, 3, 3), (5, 16, 16, 3, 192, 3),      # wide side 64 (mod 128)
                                  (2, 16, 16, 256, 4, 3), (3, 12, 20, 3, 128, 3),
                                  (2, 32, 32, 64, 3, 5), (3, 24, 16, 3, 128, 5)])      # 75 im2col columns: two filter-row groups
def test_thin_tc_conv_family(K, geom):
    """3-channel-side convs (Discriminator.1.*, Generator.Output) through the im2col tensor-core path: fprop, dgrad,
    wgrad (fresh and accumulating) against the CPU reference, and against the SIMT thin kernels."""
    N, H, W, Cin, Cout, k = geom
    g = K.same_geom(N, H, W, Cin, Cout, k, 1)
    x, dy = act((N, Cin, H, W), torch.bfloat16, 1), act((N, Cout, g.Ho, g.Wo), torch.bfloat16, 2)
    w, b = filt((k, k, Cin, Cout), 3), act((Cout,), torch.float32, 4)
    wq = w.to(torch.bfloat16).float()
    fb = FB()
    if k * k * min(Cin, Cout) <= 64:
        assert K._thin_side(g, to_dev(x)) == ('in' if Cin < Cout else 'out')
    else:                                            # LSUN Generator.Output: the route takes the filter in row groups
        assert K._thin_side(g, to_dev(x)) is None and len(K._thin_split(g, to_dev(x))) == 2
    res = {}
    for thin in (True, False):
        K.config.use_thin_tc = thin
        try:
            y = K.conv_fprop(to_dev(x), w.cuda(), b.cuda(), g)
            dx = K.conv_dgrad(to_dev(dy), w.cuda(), g)
            dw = K.conv_wgrad(to_dev(x), to_dev(dy), g, tuple(w.shape))
            acc = torch.full(tuple(w.shape), 0.25, device='cuda')
            K.conv_wgrad(to_dev(x), to_dev(dy), g, tuple(w.shape), accumulate_into=acc)
        finally:
            K.config.use_thin_tc = True
        res[thin] = (y, dx, dw)
        wr = wq if thin else w
        assert rel(y, fb.conv_fprop(x, wr, b, g)) < 1e-2, thin
        assert rel(dx, fb.conv_dgrad(dy, wr, g)) < 1e-2, thin
        assert rel(dw, fb.conv_wgrad(x, dy, g, tuple(w.shape))) < 2e-3, thin
        assert rel(acc - 0.25, fb.conv_wgrad(x, dy, g, tuple(w.shape))) < 2e-3, thin
    for a, b_ in zip(res[True], res[False]):
        assert rel(a, b_.float().cpu()) < 2e-2


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_fused_mask_kernels(K, dtype):
    """fork_dropout_relu == act_dropout(slope 1) followed by relu (same Philox slice); mask_sum2 / mask_fork2 /
    mul_relu_mask against their definitions."""
    fb = FB()
    x = act((6, 128, 8, 8), dtype, 1)
    d, r, md, mdr = K.fork_dropout_relu(to_dev(x), 0.5, seed=9, offset=64)
    d0, m0 = K.act_dropout(to_dev(x), 1.0, 0.5, seed=9, offset=64)
    r0, m1 = K.act_dropout(d0, 0.0, 1.0)
    assert torch.equal(d, d0) and torch.equal(md, m0) and torch.equal(r, r0)
    assert torch.equal(mdr.float(), (m0.float() * m1.float()))
    dr, rr, mdr_, mdrr = fb.fork_dropout_relu(x, 0.5, seed=9, offset=64)
    assert torch.equal(d.cpu(), dr) and torch.equal(r.cpu(), rr) and torch.equal(md.cpu(), mdr_) and torch.equal(mdr.cpu(), mdrr)
    a, b = act((6, 128, 8, 8), dtype, 2), act((6, 128, 8, 8), dtype, 3)
    tol = 1e-6 if dtype == torch.float32 else 8e-3
    assert rel(K.mask_sum2(to_dev(a), md, to_dev(b), mdr), fb.mask_sum2(a, md.cpu(), b, mdr.cpu())) < tol
    assert rel(K.mask_sum2(to_dev(a), None, to_dev(b), mdr), fb.mask_sum2(a, None, b, mdr.cpu())) < tol
    o1, o2 = K.mask_fork2(to_dev(a), md, mdr)
    assert torch.equal(o1, K.mul(to_dev(a), md)) and torch.equal(o2, K.mul(to_dev(a), mdr))
    assert torch.equal(K.mul_relu_mask(to_dev(a), r), K.mul(to_dev(a), m1))


def test_fused_activation_graph_matches_unfused():
    """The critic with fused activation nodes (relu in conv epilogues, dropout/relu forks, skip add in the epilogue) vs
    the literal op sequence of the reference: same losses and gradients for the critic and the generator step."""
    import numpy as np
    import ctgan_b200.gan_cifar_resnet as R
    rs = np.random.RandomState(3)
    x = torch.from_numpy(rs.randint(0, 256, (8, 3072)).astype('int32')).cuda()
    y = torch.from_numpy(rs.randint(0, 10, (8,)).astype('int32')).cuda()
    res = {}
    for fused in (True, False):
        R.FUSE_D_ACT = R.FUSE_SKIP_ADD = R.COMMUTE_1X1 = R.FUSE_RELU_BWD = R.FUSE_POOL_FORK = fused
        try:
            np.random.seed(1234)
            tr = R.Trainer(device='cuda', seed=5, act_dtype=torch.float32, batch_size=8)
            tr.disc_opt.zero_grad()
            out = tr.critic_forward_backward(x, y)
            tr.gen_opt.zero_grad()
            gc = tr.gen_forward_backward()['cost']
            res[fused] = (out['out'].clone(), out['gradients'].clone(), tr.disc_opt.flat_g.clone(), gc.clone(), tr.gen_opt.flat_g.clone())
        finally:
            R.FUSE_D_ACT = R.FUSE_SKIP_ADD = R.COMMUTE_1X1 = R.FUSE_POOL_FORK = R.FUSE_RELU_BWD = True     # the module defaults
    for a, b in zip(res[True], res[False]):
        assert rel(a, b) < 2e-3


@pytest.mark.parametrize('shape', [(64, 32, 32, 3), (24, 16, 16, 3), (10, 8, 8, 3), (6, 16, 16, 1)])
def test_dgrad_relu_mask_epilogue(K, shape):
    """conv_dgrad(..., relu_mask=x) == conv_dgrad(...) * [x > 0] on every tensor-core kernel family (pair, lean<1>, lean<0>)
    and on the SIMT path."""
    N, H, W, k = shape
    g = K.same_geom(N, H, W, 128, 128, k, 1)
    dy, xm = act((N, 128, H, W), torch.bfloat16, 2), act((N, 128, H, W), torch.bfloat16, 7)
    w = filt((k, k, 128, 128), 3)
    for use_tc in (True, False):
        K.config.use_tc = use_tc
        try:
            plain = K.conv_dgrad(to_dev(dy), w.cuda(), g)
            masked = K.conv_dgrad(to_dev(dy), w.cuda(), g, relu_mask=to_dev(xm))
        finally:
            K.config.use_tc = True
        want = torch.where(to_dev(xm) > 0, plain, torch.zeros_like(plain))
        assert torch.equal(masked == 0, want == 0) and rel(masked, want) < 4e-3, use_tc     # (a different kernel accumulates)


@pytest.mark.parametrize('shape', [(64, 32, 32), (40, 16, 16), (12, 8, 8)])
def test_residual_upsampled_in_epilogue(K, shape):
    """conv_fprop(..., residual=low_res, res_up2=True) == conv_fprop(..., residual=upsample2x(low_res)), bit for bit."""
    N, H, W = shape
    g = K.same_geom(N, H, W, 128, 128, 3, 1)
    x, r = act((N, 128, H, W), torch.bfloat16, 1), act((N, 128, H // 2, W // 2), torch.bfloat16, 5)
    w, b = filt((3, 3, 128, 128), 3), act((128,), torch.float32, 4)
    full = K.conv_fprop(to_dev(x), w.cuda(), b.cuda(), g, residual=K.upsample2x(to_dev(r), 1.0))
    fused = K.conv_fprop(to_dev(x), w.cuda(), b.cuda(), g, residual=to_dev(r), res_up2=True)
    assert torch.equal(full, fused)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('keep', [1.0, 0.5])
def test_pool_add_fork_kernels(K, dtype, keep):
    """pool_add_fork / mask_sum2_up against the CPU stand-ins (and so against pool -> add -> dropout -> relu)."""
    fb = FB()
    y, s = act((6, 128, 16, 16), dtype, 1), act((6, 128, 8, 8), dtype, 2)
    o1, o2, m1, m2 = K.pool_add_fork(to_dev(y), to_dev(s), keep, seed=7, offset=128)
    r1, r2, q1, q2 = fb.pool_add_fork(y, s, keep, seed=7, offset=128)
    tol = 1e-6 if dtype == torch.float32 else 8e-3
    assert (m1 is None) == (keep == 1.0)
    if m1 is not None:
        assert torch.equal(m1.cpu(), q1)
    # a pooled value within rounding of zero may land on the other side of the ReLU: compare away from it
    far = (r1.float().abs() > 1e-2)
    assert torch.equal(m2.cpu()[far], q2[far]) and rel(o1, r1) < tol and rel(o2.cpu() * far, r2 * far) < tol
    p1, p2, _, _ = K.pool_add_fork(to_dev(y), to_dev(s), masks=(m1, m2))
    assert torch.equal(p1, o1) and torch.equal(p2, o2)
    a, b = act((6, 128, 8, 8), dtype, 3), act((6, 128, 8, 8), dtype, 4)
    gy, gx = K.mask_sum2_up(to_dev(a), m1, to_dev(b), m2)
    hy, hx = fb.mask_sum2_up(a, None if m1 is None else m1.cpu(), b, m2.cpu())
    assert rel(gx, hx) < tol and rel(gy, hy) < tol


# ---------------------------------------------------------------- stride-2 5x5 convs on the tensor cores (csrc/conv_s2d.cu)
S2D_GEOMS = [
    # N, H, W, Cin, Cout  (5x5, stride 2, TF SAME)
    (64, 16, 16, 128, 256),    # CIFAR Discriminator.2 / (as dgrad) Generator.3
    (64, 8, 8, 256, 512),      # CIFAR Discriminator.3 / Generator.2
    (192, 16, 16, 128, 256),   # ... on the stacked critic pass
    (5, 8, 8, 128, 256),       # MNIST Generator.2 (as dgrad), ragged batch
    (50, 14, 14, 64, 128),     # MNIST Discriminator.2 / Generator.3: 7x7 space-to-depth image
    (50, 7, 7, 128, 256),      # MNIST Discriminator.3: odd extent, pads (2, 2)
    (3, 12, 20, 32, 128),      # non-square, 4*Cin = 128
]


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('shape', [(3, 16, 16, 128), (2, 7, 7, 64), (4, 9, 6, 3), (1, 32, 32, 8)])
def test_s2d_layout_kernels(K, shape, dtype):
    """space_to_depth / depth_to_space against the test-side restatement (exact: pure data movement)."""
    from tests import test_s2d_host as H
    N, Hh, W, C = shape
    g = K.same_geom(N, Hh, W, C, 64, 5, 2)
    x = act((N, C, Hh, W), dtype, 1)
    xs = K.space_to_depth(to_dev(x), g)
    ref = H._s2d(x, g)
    assert xs.shape == ref.shape and xs.is_contiguous(memory_format=CL) and torch.equal(xs.cpu(), ref)
    back = K.depth_to_space(xs, g)
    assert back.shape == x.shape and torch.equal(back.cpu(), x)


@pytest.mark.parametrize('geom', [(5, 8, 16, 1), (5, 32, 64, 2), (5, 3, 5, 1), (3, 16, 32, 0)])
def test_s2d_filter_kernels(K, geom):
    """pack_filter_s2d (both operand layouts) and its adjoint s2d_filter_grad against the restatement."""
    from tests import test_s2d_host as H
    k, Cin, Cout, pad = geom
    g = K.ConvGeom(1, 8, 8, Cin, 4, 4, Cout, k, k, 2, pad, pad)
    w = filt((k, k, Cin, Cout), 3, 1.0)
    n = 36 * Cin * Cout
    wp_f = torch.empty(n, dtype=torch.bfloat16, device='cuda')
    wp_d = torch.empty(n, dtype=torch.bfloat16, device='cuda')
    K._pack_filter_s2d_launch(w.cuda(), wp_f, wp_d, g)
    rf, rd = torch.empty(n, dtype=torch.bfloat16), torch.empty(n, dtype=torch.bfloat16)
    H._pack_launch(w, rf, rd, g)
    assert torch.equal(wp_f.cpu(), rf) and torch.equal(wp_d.cpu(), rd)
    dw3 = filt((3, 3, 4 * Cin, Cout), 5, 1.0)
    for accumulate in (0, 1):
        dw = torch.ones(k, k, Cin, Cout, device='cuda')
        K._s2d_filter_grad_launch(dw3.cuda(), dw, g, accumulate)
        ref = torch.ones(k, k, Cin, Cout)
        H._filter_grad_launch(dw3, ref, g, accumulate)
        assert torch.equal(dw.cpu(), ref)


@pytest.mark.parametrize('geom', S2D_GEOMS)
def test_s2d_conv_family(K, geom):
    """Stride-2 5x5 SAME conv family on the space-to-depth tensor-core route vs the direct CPU evaluation, and vs the
    SIMT kernels the same call takes with the route off."""
    N, H, W, Cin, Cout = geom
    g = K.same_geom(N, H, W, Cin, Cout, 5, 2)
    x, dy = act((N, Cin, H, W), torch.bfloat16, 1), act((N, Cout, g.Ho, g.Wo), torch.bfloat16, 2)
    w, b = filt((5, 5, Cin, Cout), 3), act((Cout,), torch.float32, 4)
    K.config.use_s2d = True
    try:
        assert K.s2d_geom(g, to_dev(x)) is not None
        y = K.conv_fprop(to_dev(x), w.cuda(), b.cuda(), g)
        dx = K.conv_dgrad(to_dev(dy), w.cuda(), g)
        dw = K.conv_wgrad(to_dev(x), to_dev(dy), g, tuple(w.shape))
        acc = torch.ones(5, 5, Cin, Cout, device='cuda')
        K.conv_wgrad(to_dev(x), to_dev(dy), g, tuple(w.shape), accumulate_into=acc, col=K.thin_col(to_dev(x), g, 'x'))
        K.config.use_s2d = False
        ys = K.conv_fprop(to_dev(x), w.cuda(), b.cuda(), g)           # SIMT route
    finally:
        K.config.use_s2d = True
    wq = w.to(torch.bfloat16).float()
    fb = FB()
    assert rel(y, fb.conv_fprop(x, wq, b, g)) < 1e-2 and rel(y, ys) < 1.5e-2
    assert rel(dx, fb.conv_dgrad(dy, wq, g)) < 1e-2
    ref_w = fb.conv_wgrad(x, dy, g, tuple(w.shape))
    assert rel(dw, ref_w) < 2e-3 and rel(acc - 1, ref_w) < 2e-3


THIN_S2_GEOMS = [
    # N, H, W, Cin, Cout, k  (stride 2, TF SAME)
    (192, 32, 32, 3, 128, 5),   # CIFAR Discriminator.1 on the stacked critic pass / (as dgrad) Generator.5
    (64, 32, 32, 3, 128, 5),
    (50, 28, 28, 1, 64, 5),     # MNIST Discriminator.1 / Generator.5 (Cout = 64: wgrad on the SIMT kernel)
    (3, 14, 10, 3, 128, 5),     # non-square, ragged pixel count
    (2, 9, 9, 2, 256, 3),       # odd extent, 3x3
]


@pytest.mark.parametrize('geom', THIN_S2_GEOMS)
def test_thin_strided_conv_family(K, geom):
    """Stride-2 convs with a <= 8-channel input (im2col_strided / col2im_strided / pack_filter_padk / add_prefix + the
    1x1 tcgen05 GEMMs): the layout kernels exactly against the test-side restatement, the family against the direct
    CPU evaluation."""
    from tests import test_s2d_host as H
    N, Hh, W, Cin, Cout, k = geom
    g = K.same_geom(N, Hh, W, Cin, Cout, k, 2)
    x, dy = act((N, Cin, Hh, W), torch.bfloat16, 1), act((N, Cout, g.Ho, g.Wo), torch.bfloat16, 2)
    w, b = filt((k, k, Cin, Cout), 3, 0.1), act((Cout,), torch.float32, 4)
    assert K.thin_s2_ok(g, to_dev(x))
    col = K.im2col_strided(to_dev(x), g)
    assert torch.equal(col.cpu(), H._im2col_strided(x, g))
    dcol = act((N, 128, g.Ho, g.Wo), torch.bfloat16, 6)
    bc = act((Cin,), torch.float32, 7)
    got, ref = K.col2im_strided(to_dev(dcol), bc.cuda(), g), H._col2im_strided(dcol, bc, g)
    assert rel(got, ref) < 4e-3                       # fp32 sums of <= 9 bf16 terms, one bf16 rounding
    n = 128 * Cout
    wp = [torch.empty(n, dtype=torch.bfloat16, device='cuda') for _ in range(2)]
    K._pack_filter_padk_launch(w.cuda(), wp[0], wp[1], g)
    rp = [torch.empty(n, dtype=torch.bfloat16) for _ in range(2)]
    H._padk_launch(w, rp[0], rp[1], g)
    assert torch.equal(wp[0].cpu(), rp[0]) and torch.equal(wp[1].cpu(), rp[1])

    y = K.conv_fprop(to_dev(x), w.cuda(), b.cuda(), g, col=col)
    dx = K.conv_dgrad(to_dev(dy), w.cuda(), g)
    dw = K.conv_wgrad(to_dev(x), to_dev(dy), g, tuple(w.shape))
    acc = torch.ones(k, k, Cin, Cout, device='cuda')
    K.conv_wgrad(to_dev(x), to_dev(dy), g, tuple(w.shape), accumulate_into=acc, col=col)
    wq = w.to(torch.bfloat16).float()
    fb = FB()
    assert rel(y, fb.conv_fprop(x, wq, b, g)) < 1e-2
    assert rel(dx, fb.conv_dgrad(dy, wq, g)) < 1.5e-2
    ref_w = fb.conv_wgrad(x, dy, g, tuple(w.shape))
    assert rel(dw, ref_w) < 2e-3 and rel(acc - 1, ref_w) < 2e-3


def test_head_fprop_long_k(K):
    """Discriminator.Output of the DCGAN critics: Linear(8192 -> 1) on BF16 features, float logits."""
    for M, Kd in ((192, 8192), (50, 4096), (7, 1024)):
        g = K.ConvGeom(M, 1, 1, Kd, 1, 1, 1, 1, 1, 1, 0, 0)
        x, w, b = act((M, Kd), torch.bfloat16, 1), filt((1, 1, Kd, 1), 2, 0.05), act((1,), torch.float32, 3)
        y = K.conv_fprop(x.cuda(), w.cuda(), b.cuda(), g, out_dtype=torch.float32)
        assert rel(y, FB().conv_fprop(x, w, b, g, out_dtype=torch.float32)) < 1e-5


# ---- ConvMeanPool(3x3) as one stride-2 4x4 conv: zero-block skipping, layout-changing epilogues, the box filter
POOLCONV_GEOMS = [
    # N, H, W, C (Cin = Cout = C)
    (192, 32, 32, 128),     # ResNet Discriminator.1.Conv2 on the stacked pass: pair kernel (one 16x16 image per item)
    (64, 32, 32, 128),      # ... on the gradient-penalty pass: lean kernel, halo boxes
    (64, 16, 16, 128),      # Discriminator.2.Conv2: 8x8 space-to-depth images, [h][n][w] halo boxes
    (3, 32, 32, 256),       # two channel blocks per phase, ragged tile count
]


@pytest.mark.parametrize('k', [4, 5])
@pytest.mark.parametrize('geom', POOLCONV_GEOMS)
def test_s2d_zero_block_skipping(K, geom, k):
    """CTGAN_EPI_S2D_SKIP: the (tap, phase) blocks of the embedded 3x3 filter without a filter element are neither loaded
    nor multiplied -- fprop and dgrad are BIT-IDENTICAL to the launches that multiply the zeros."""
    N, H, W, C = geom
    g = K.same_geom(N, H, W, C, C, k, 2)
    x, dy = to_dev(act((N, C, H, W), torch.bfloat16, 1)), to_dev(act((N, C, g.Ho, g.Wo), torch.bfloat16, 2))
    w, b = filt((k, k, C, C), 3).cuda(), act((C,), torch.float32, 4).cuda()
    from ctgan_b200 import _lib
    outs = {}
    max_k, K.config.s2d_skip_max_k = K.config.s2d_skip_max_k, 7       # the policy (k <= 4 by default) is not under test
    _lib.lib.ctgan_set_splitk(0)        # sub-wave layers keep the split-K kernel, which ignores the flag: compare lean with lean
    try:
        assert K.s2d_geom(g, x) is not None and K._s2d_skip_flags(g, 0) and K._s2d_skip_flags(g, 1)
        for skip in (True, False):
            K.config.s2d_skip = skip
            outs[skip] = (K.conv_fprop(x, w, b, g), K.conv_dgrad(dy, w, g), K.conv_dgrad(dy, w, g, out_s2d=True))
    finally:
        K.config.s2d_skip, K.config.s2d_skip_max_k = True, max_k
        _lib.lib.ctgan_set_splitk(1)
    for a, c in zip(outs[True], outs[False]):
        assert torch.equal(a, c)
    wq = w.cpu().to(torch.bfloat16).float()
    if N <= 64:
        fb = FB()
        assert rel(outs[True][0], fb.conv_fprop(x.cpu(), wq, b.cpu(), g)) < 1e-2
        assert rel(outs[True][1], fb.conv_dgrad(dy.cpu(), wq, g)) < 1e-2


@pytest.mark.parametrize('geom', [(64, 32, 32, 3, 128), (6, 16, 16, 128, 128), (192, 16, 16, 128, 128), (2, 8, 8, 128, 256)])
def test_conv_epilogue_space_to_depth_output(K, geom):
    """CTGAN_EPI_OUT_S2D: relu(conv + b) written by the epilogue in the space-to-depth layout == the layout kernel applied to
    the plain result, bit for bit (thin-input GEMM route, lean and pair kernels)."""
    N, H, W, Cin, Cout = geom
    g = K.same_geom(N, H, W, Cin, Cout, 3, 1)
    x = to_dev(act((N, Cin, H, W), torch.bfloat16, 1))
    w, b = filt((3, 3, Cin, Cout), 3).cuda(), act((Cout,), torch.float32, 4).cuda()
    assert K.conv_fprop_s2d_out_ok(x, g)
    from ctgan_b200 import _lib
    ys = K.conv_fprop(x, w, b, g, relu=True, out_s2d=True)
    _lib.lib.ctgan_set_splitk(0)        # the layout-changing launch stays on the lean / pair kernels: same kernel for the reference
    try:
        y = K.conv_fprop(x, w, b, g, relu=True)
    finally:
        _lib.lib.ctgan_set_splitk(1)
    assert tuple(ys.shape) == (N, 4 * Cout, H // 2, W // 2)
    assert torch.equal(ys, K.space_to_depth(y, K.ConvGeom(N, H, W, Cout, H, W, Cout, 1, 1, 1, 0, 0)))


@pytest.mark.parametrize('geom', POOLCONV_GEOMS)
def test_s2d_dgrad_masked_plain_output(K, geom):
    """CTGAN_EPI_OUT_D2S (+ relu_mask in the space-to-depth layout): the dgrad of the stride-2 conv, masked and written as
    the plain tensor by the epilogue == mask multiply + depth_to_space of the space-to-depth result, bit for bit."""
    N, H, W, C = geom
    g = K.same_geom(N, H, W, C, C, 4, 2)
    dy = to_dev(act((N, C, g.Ho, g.Wo), torch.bfloat16, 2))
    w = filt((4, 4, C, C), 3).cuda()
    hs = to_dev(act((N, 4 * C, H // 2, W // 2), torch.bfloat16, 5)).relu()       # the conv's input (a ReLU output), s2d layout
    dx = K.conv_dgrad(dy, w, g, relu_mask=hs)
    assert tuple(dx.shape) == (N, C, H, W)
    dxs = K.conv_dgrad(dy, w, g, out_s2d=True)
    ref = K.depth_to_space(K.mul_relu_mask(dxs, hs), g)
    assert torch.equal(dx, ref)
    # the adjoint used by the double backward: space_to_depth of a plain tensor, masked by the same pattern
    c = to_dev(act((N, C, H, W), torch.bfloat16, 6))
    pg = K.ConvGeom(N, H, W, C, H, W, C, 1, 1, 1, 0, 0)
    assert torch.equal(K.space_to_depth_mask(c, hs, pg), K.mul_relu_mask(K.space_to_depth(c, pg), hs))


def test_box_filter_is_conv_then_mean_pool(K):
    """ctgan_box_filter: the stride-2 4x4 conv with the box-summed filter == mean_pool_2x2(conv3x3) (float, CPU reference);
    ctgan_box_filter_grad is its adjoint and clears the scratch gradient."""
    import torch.nn.functional as TF
    C, O = 16, 24
    w3 = filt((3, 3, C, O), 1, 1.0)
    w4 = K.box_filter(w3.cuda(), torch.empty(4, 4, C, O, device='cuda')).cpu()
    x = torch.randn(2, C, 12, 10, generator=torch.Generator().manual_seed(2))
    ref = TF.avg_pool2d(TF.conv2d(x, w3.permute(3, 2, 0, 1), padding=1), 2)
    got = TF.conv2d(x, w4.permute(3, 2, 0, 1), stride=2, padding=1)
    assert rel(got, ref) < 1e-6
    # adjoint: <box(w3), d4> == <w3, box^T(d4)>
    d4 = filt((4, 4, C, O), 3, 1.0)
    d3 = torch.ones(3, 3, C, O, device='cuda')
    d4d = d4.cuda()
    K.box_filter_grad(d4d, d3)
    assert float(d4d.abs().max()) == 0.0
    lhs = float((w4.double() * d4.double()).sum())
    rhs = float((w3.double() * (d3.cpu().double() - 1)).sum())
    assert abs(lhs - rhs) < 1e-6 * max(1.0, abs(lhs))


@pytest.mark.parametrize('N', [64, 6])
def test_conv_mean_pool_fused_block_matches_unfused(K, N):
    """functional level: [conv3x3 + relu -> ConvMeanPool(3x3) + skip] through the fused route (space-to-depth hand-over, one
    stride-2 conv, masked plain-layout dgrad, box-filter gradient fold) against the unfused ops: outputs, input gradient,
    filter / bias gradients and the gradient-penalty style double backward."""
    import ctgan_b200.functional as F
    C, H, W = 128, 16, 16
    gen = torch.Generator().manual_seed(0)

    def leaf(shape, scale, dtype=torch.float32):
        t = (torch.randn(shape, generator=gen) * scale).to(dtype).cuda()
        if t.dim() == 4 and dtype == torch.bfloat16:
            t = t.contiguous(memory_format=CL)
        return t.requires_grad_(True)

    w1, b1 = leaf((3, 3, C, C), 0.03), leaf((C,), 0.1)
    w2, b2 = leaf((3, 3, C, C), 0.03), leaf((C,), 0.1)
    x0 = leaf((N, C, H, W), 1.0, torch.bfloat16)
    sc0 = leaf((N, C, H // 2, W // 2), 1.0, torch.bfloat16)
    for p in (w1, b1, w2, b2):
        F.register_param(p)

    def block(x, sc, fused):
        if fused:
            hs = F.conv2d(x, w1, b1, 3, 1, relu=True, relu_bwd_fused=True, out_s2d=True)
            return F.conv_mean_pool_s2d(hs, w2, b2, residual=sc)
        h = F.conv2d(x, w1, b1, 3, 1, relu=True, relu_bwd_fused=True)
        full = F.conv2d(h, w2, b2, 3, 1, in_relu=True)
        return F.add(sc, F.mean_pool_2x2(full))

    saved = (K.config.pool_conv_s2d, K.config.pool_conv_min_tiles)
    K.config.pool_conv_s2d, K.config.pool_conv_min_tiles = True, 1               # the route's switch / size policy are not under test
    try:
        assert F.conv2d_s2d_out_ok(x0, C, 3) and F.conv_mean_pool_s2d_ok(x0, C, C)
    finally:
        K.config.pool_conv_s2d, K.config.pool_conv_min_tiles = saved
    res = {}
    for fused in (True, False):
        for p in (w1, b1, w2, b2):
            p.grad = torch.zeros_like(p)
        x = x0.detach().clone(memory_format=torch.preserve_format).requires_grad_(True)
        sc = sc0.detach().clone(memory_format=torch.preserve_format).requires_grad_(True)
        y = block(x, sc, fused)
        # first derivative as the gradient penalty takes it: input gradient only, differentiable
        with F.no_param_grads():
            gx, = torch.autograd.grad(y.float().square().sum() * 1e-3, [x], create_graph=True)
        loss = y.float().sum() * 1e-2 + gx.float().square().sum()
        # the trainers' final backward: filter / bias gradients are accumulated in place by the kernels
        loss.backward(inputs=[x, sc, w1, b1, w2, b2])
        K.join_side()
        torch.cuda.synchronize()
        res[fused] = dict(y=y.detach().float(), gx=gx.detach().float(), dx=x.grad.float(), dsc=sc.grad.float(),
                          w1=w1.grad.clone(), b1=b1.grad.clone(), w2=w2.grad.clone(), b2=b2.grad.clone())
    tol = dict(y=1e-2, gx=1.5e-2, dx=2e-2, dsc=1e-2, w1=2e-2, b1=2e-2, w2=2e-2, b2=2e-2)
    for k_, t in tol.items():
        assert rel(res[True][k_], res[False][k_]) < t, (k_, rel(res[True][k_], res[False][k_]))
