"""Tensor-level launchers over the C ABI (no autograd here; see functional.py).

PyTorch supplies device memory (caching allocator) and the current CUDA stream; every
arithmetic op is a kernel of libctgan_sm100.so.  Activations are torch tensors that are
*logically* NCHW (the reference's convention, TG/tflib/ops/conv2d.py:22-26) and
*physically* NHWC (torch.channels_last), dtype float32 ("fp32 path") or bfloat16
("BF16 path").  2-D tensors [batch, features] are the H=W=1 case.
"""
import collections
import ctypes

import torch

from . import _lib
from ._lib import F32, BF16, ConvDesc, LossDesc, call

CL = torch.channels_last


class Config:
    use_tc = True            # use the tcgen05 kernels when a call is eligible
    use_thin_tc = True       # route 3-channel-side convs through the im2col tensor-core path
    use_s2d = True           # stride-2 5x5 convs (DCGAN critics / Deconv2D) as 3x3 tensor-core convs over the space-to-depth image
    use_thin_s2 = True       # stride-2 convs with a <= 8 channel input (Discriminator.1, the last Deconv2D) as im2col + 1x1 tensor-core GEMMs
    s2d_min_extent = 1       # (tunable) smallest space-to-depth image side that takes the tensor-core route
    side_stream = True       # run direct-accumulation wgrad / bias-grad launches on a second stream
    branch_streams = True    # run independent sub-graphs of a step (GP pass vs stacked pass) as stream branches
    branch_priority = 0      # CUDA stream priority of the branch stream (lower = higher priority): equal priorities measure best
    critic_splitk = False    # cluster split-K inside the two-branch ResNet critic step (see kernels.splitk)
    s2d_skip = True          # space-to-depth route: skip the (tap, phase) blocks of the embedded 3x3 filter that hold no filter element
    s2d_skip_max_k = 4       # ... for stride-2 filters up to this size: 16 of 36 blocks live for k = 4; with k = 5 (25 of 36) the
                             # shorter MMA stream does not pay for the issue loop's mask arithmetic (CIFAR-DCGAN 248 -> 243 it/s)
    # ConvMeanPool(3x3) as ONE stride-2 4x4 conv on the space-to-depth route (gan_cifar_resnet._pool_conv_fused, functional.
    # conv_mean_pool_s2d): 2.25x fewer multiply-adds for Discriminator.{1,2}.Conv2.  OFF by default: measured
    # +0.9 % per iteration on the same box (critic graph 922.6 -> 913.3 us, generator 1235 -> 1227 us: the filter-gradient launch
    # shrinks by 26 us) -- the forward / dgrad launches stay bound by the shared-memory fill (fewer live taps per halo box, the same filter bytes per MMA)
    # and by one-image work items (192 items on 148 SMs), see profiles/r02_experiments.md.  Both settings are GPU-tested.
    pool_conv_s2d = False
    pool_conv_min_tiles = 96 # ... for layers with at least this many 128-pixel output tiles (smaller ones are latency-bound and lose)
    s2d_embed_wgrad = True   # stride-2 filter gradients: the embedded 3x3 job writes the k x k gradient itself (no scratch + gather)
    decouple_gp = False      # ResNet critic step: the stacked pass runs its own backward as soon as its half of the loss is known,
                             # the gradient penalty is differentiated on its stream branch (two backward calls, gan_cifar_resnet.py).
                             # Measured equal (922-928 vs 919-923 us): the penalty branch's three traversals at batch 64 are the
                             # critical path either way (profiles/r02_experiments.md); kept as a checked option
    gen_towers = False       # ResNet generator step: the reference's per-device towers as two stream branches instead of one stacked
                             # batch (measured: 1303 us vs 1226 us stacked -- twice the launches, no shorter chain; kept as a checked option)
    gen_splitk = False       # cluster split-K inside the two-tower generator step
    branch_stacked = False   # which sub-graph runs on the branch stream: False = the gradient-penalty pass, True = the stacked pass
    tc_min_rows = 1          # (tunable) minimum GEMM rows to prefer the tensor-core path
    use_bn_fused = True      # BF16 batch norm as two kernels per direction (sums with red.global + apply)
    fuse_act_dropout = True  # Conv2D -> LeakyReLU -> dropout of the DCGAN critics in the tcgen05 conv epilogue (Philox in registers)
    peer_update = True       # data parallel: reduce-scatter + Adam + all-gather as one kernel over NVLink peer memory (else NCCL all-reduce)
    tf32 = False             # fp32 activations: stride-1 convs / linears on tcgen05 kind::tf32 instead of the FP32-FMA SIMT kernels
    defer_wgrad = True       # queue the final backward's tensor-core filter gradients and run them as ONE launch at the join


config = Config()
import os as _os
# A/B hook for the benches and tests/graph_times.py: CTGAN_CONFIG="pool_conv_s2d=0,s2d_skip=0" overrides switches at import
for _kv in filter(None, _os.environ.get('CTGAN_CONFIG', '').split(',')):
    _k, _v = _kv.split('=')
    if not hasattr(Config, _k.strip()):
        raise RuntimeError('ctgan_b200: CTGAN_CONFIG names an unknown switch %r' % _k)
    setattr(config, _k.strip(), type(getattr(Config, _k.strip()))(int(_v)))
# A/B switches from the environment (benchmarks): CTGAN_PEER_UPDATE=0 -> NCCL all-reduce + Adam instead of the peer-memory kernel
import os as _os
if _os.environ.get('CTGAN_PEER_UPDATE') is not None:
    config.peer_update = bool(int(_os.environ['CTGAN_PEER_UPDATE']))

_tc_avail = None


def tc_available():
    global _tc_avail
    if _tc_avail is None:
        _tc_avail = bool(_lib.lib.ctgan_tc_available())
    return _tc_avail


def _dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError('ctgan_b200: unsupported dtype %s' % t.dtype)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _chk(t, name='tensor'):
    if not t.is_cuda:
        raise RuntimeError('ctgan_b200: %s must be a CUDA tensor (there is no CPU path)' % name)


def is_nhwc(x):
    return x.dim() != 4 or x.is_contiguous(memory_format=CL)


def require_nhwc(x, name='activation'):
    _chk(x, name)
    if x.dim() == 4:
        if not x.is_contiguous(memory_format=CL):
            raise RuntimeError('ctgan_b200: %s must be channels_last (NHWC) contiguous' % name)
    elif not x.is_contiguous():
        raise RuntimeError('ctgan_b200: %s must be contiguous' % name)
    return x


def empty_act(shape, dtype, device):
    if len(shape) == 4:
        return torch.empty(shape, dtype=dtype, device=device, memory_format=CL)
    return torch.empty(shape, dtype=dtype, device=device)


def nhwc_dims(x):
    """(N, H, W, C) of a logical-NCHW / 2-D activation."""
    if x.dim() == 4:
        N, C, H, W = x.shape
        return N, H, W, C
    if x.dim() == 2:
        return x.shape[0], 1, 1, x.shape[1]
    raise RuntimeError('ctgan_b200: activations must be 2-D or 4-D')


# --------------------------------------------------------------------------- side stream
# Filter / bias gradients are off the critical path of a backward pass: nothing reads them before the optimizer
# step, while the dgrad chain is a strict sequence of (at batch 64, often under-filled) kernels.  The
# direct-accumulation wgrad / bias-grad launches therefore go to a second stream and overlap that chain; inside a
# CUDA-graph capture the fork/join becomes parallel graph branches.
_side_streams = {}
_side_pending = []       # tensors the side stream may still be reading (kept alive until the join)


def on_side(fn, *keep):
    """Run fn() (kernel launches only) on the side stream, ordered after everything queued so far on the current
    stream; `keep` = the tensors those kernels read."""
    if not (config.side_stream and keep and keep[0].is_cuda):
        fn()
        return
    dev = keep[0].device.index
    side = _side_streams.get(dev)
    if side is None:
        side = _side_streams[dev] = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    _side_pending.append((dev, keep))


def join_side():
    """Make the current stream wait for the side stream's work; release the kept tensors.  Deferred filter gradients
    (defer_wgrad) are launched here, on the current stream."""
    flush_wgrads()
    if _side_pending:
        for dev in {d for d, _ in _side_pending}:
            torch.cuda.current_stream(dev).wait_stream(_side_streams[dev])
        _side_pending.clear()
    for hook in _join_hooks:
        hook()


class splitk:
    """with splitk(False): the tcgen05 forward launches issued inside keep one CTA per tile.  The cluster split-K kernel
    (csrc/conv_splitk.cu) shortens a sub-wave layer's latency by occupying up to 4x the SMs; that pays when the layer runs
    alone (generator step, DCGAN steps: -25 us / step) and costs when a second stream branch competes for the SMs (the
    ResNet critic step with its gradient-penalty branch: +30 us) -- measured, profiles/README.md."""
    _on = True

    def __init__(self, on):
        self.on = bool(on)

    def __enter__(self):
        self.prev, splitk._on = splitk._on, self.on
        if self.on != self.prev:
            _lib.lib.ctgan_set_splitk(int(self.on))
        return self

    def __exit__(self, *exc):
        if self.prev != splitk._on:
            _lib.lib.ctgan_set_splitk(int(self.prev))
        splitk._on = self.prev
        return False


# Independent sub-graphs of one step (the gradient-penalty pass vs the stacked critic pass) as two stream branches.
_branch_streams = {}


def fork_branch(t):
    """Mark the point after which a branch may start: returns a handle (None when branching is off / on CPU)."""
    if not (config.branch_streams and t.is_cuda):
        return None
    dev = t.device.index
    if dev not in _branch_streams:
        # high priority: the branch is the longer chain of small kernels (3 traversals at batch B vs 2 at 3B)
        _branch_streams[dev] = torch.cuda.Stream(device=dev, priority=config.branch_priority)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(dev))
    return (dev, ev)


class branch:
    """with branch(fork): launches go to the branch stream, ordered after the fork point only."""

    def __init__(self, fork):
        self.fork, self.ctx = fork, None

    def __enter__(self):
        if self.fork is not None:
            dev, ev = self.fork
            _branch_streams[dev].wait_event(ev)
            self.ctx = torch.cuda.stream(_branch_streams[dev])
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def join_branch(fork):
    """The current stream waits for everything queued on the branch stream."""
    if fork is not None:
        torch.cuda.current_stream(fork[0]).wait_stream(_branch_streams[fork[0]])


# --------------------------------------------------------------------------- conv family
ConvGeom = collections.namedtuple('ConvGeom', 'N H W Cin Ho Wo Cout kh kw stride pad_t pad_l')


def same_geom(N, H, W, Cin, Cout, k, stride):
    """TF 'SAME' geometry (SURVEY.md 8(c) rule 1)."""
    Ho, Wo = -(-H // stride), -(-W // stride)
    pt = max((Ho - 1) * stride + k - H, 0) // 2
    pl = max((Wo - 1) * stride + k - W, 0) // 2
    return ConvGeom(N, H, W, Cin, Ho, Wo, Cout, k, k, stride, pt, pl)


def _desc(g, xdt, ydt):
    return ConvDesc(g.N, g.H, g.W, g.Cin, g.Ho, g.Wo, g.Cout, g.kh, g.kw, g.stride, g.pad_t, g.pad_l, xdt, ydt)


def _x_shape(g, two_d):
    return (g.N, g.Cin) if two_d else (g.N, g.Cin, g.H, g.W)


def _y_shape(g, two_d):
    return (g.N, g.Cout) if two_d else (g.N, g.Cout, g.Ho, g.Wo)


def _tc_geom_ok(g):
    return (config.use_tc and tc_available() and g.stride == 1 and g.Ho == g.H and g.Wo == g.W
            and g.Cin % 64 == 0 and g.Cout % 64 == 0 and 0 <= g.pad_t < g.kh and 0 <= g.pad_l < g.kw)


def _tf32_geom_ok(g, cin_mult=32):
    """float activations on the tensor cores (kind::tf32): opt-in, stride 1, Cin % cin_mult == 0, Cout % 128 == 0."""
    return (config.tf32 and config.use_tc and tc_available() and g.stride == 1 and g.Ho == g.H and g.Wo == g.W
            and g.Cin % cin_mult == 0 and g.Cout % 128 == 0 and 0 <= g.pad_t < g.kh and 0 <= g.pad_l < g.kw)


# packed-filter cache for PARAMETERS only (keyed by storage address, shape, layout)
_pack_cache = {}


def invalidate_weight_cache(ptrs=None, forget=False):
    """Called by an optimizer after ITS parameters changed in place (ptrs = their data pointers; None = all).
    forget: the tensors at `ptrs` no longer exist (moved into a flat buffer): also drop their registered lazy packs."""
    if ptrs is None:
        _pack_cache.clear()
        _lazy_packs.clear()
        _derived_filters.clear()
        return
    for k in [k for k in _pack_cache if k[0] in ptrs]:
        del _pack_cache[k]
    if forget:
        for ptr in ptrs:
            _lazy_packs.pop(ptr, None)
            for d, _ in _derived_filters.pop(ptr, ()):
                _lazy_packs.pop(d.data_ptr(), None)


# Filters COMPUTED from a parameter by a kernel of their own (ConvMeanPool's box-summed 4x4 filter, functional.box_filter):
# {parameter ptr: [(derived tensor, launch that recomputes it in place)]}.  FilterPacker.refresh() re-runs the launches and
# re-packs the derived tensors' operands after every optimizer step; _join_hooks fold their scratch gradients back.
_derived_filters = {}
_join_hooks = []


def register_derived_filter(src, derived, refresh):
    _derived_filters.setdefault(src.data_ptr(), []).append((derived, refresh))


def box_filter(w3, w4):
    """w4 [4,4,Cin,Cout] = the 'SAME' 4x4 / stride-2 filter equal to mean-pooling the 'SAME' 3x3 conv with w3 over 2x2 windows."""
    call('ctgan_box_filter', _p(w3), _p(w4), w3.shape[2], w3.shape[3], _stream())
    return w4


def box_filter_grad(dw4, dw3):
    """dw3 += the adjoint of box_filter applied to dw4; dw4 is cleared."""
    call('ctgan_box_filter_grad', _p(dw4), _p(dw3), dw3.shape[2], dw3.shape[3], _stream())


def pack_filter(w, transpose_flip, cacheable=False):
    """float HWIO [kh,kw,Cin,Cout] (or [in,out]) -> bf16 operand of the tcgen05 kernels."""
    if w.dim() == 4:
        taps, cin, cout = w.shape[0] * w.shape[1], w.shape[2], w.shape[3]
    else:
        taps, cin, cout = 1, w.shape[0], w.shape[1]
    if w.dtype != torch.float32 or not w.is_contiguous():
        raise RuntimeError('ctgan_b200: filters must be contiguous float32 HWIO')
    wd = w.detach()
    return _lazy_pack(w, (w.data_ptr(), tuple(w.shape), transpose_flip),
                      lambda: torch.empty(taps * cin * cout, dtype=torch.bfloat16, device=w.device),
                      lambda wp: call('ctgan_pack_filter_bf16', _p(wd), _p(wp), taps, cin, cout, int(transpose_flip), _stream()),
                      cacheable)


def pack_filter_f32(w, transpose_flip, cacheable=False):
    """float HWIO [kh,kw,Cin,Cout] (or [in,out]) -> float operand of the kind::tf32 kernels (same layouts as pack_filter)."""
    if w.dim() == 4:
        taps, cin, cout = w.shape[0] * w.shape[1], w.shape[2], w.shape[3]
    else:
        taps, cin, cout = 1, w.shape[0], w.shape[1]
    if w.dtype != torch.float32 or not w.is_contiguous():
        raise RuntimeError('ctgan_b200: filters must be contiguous float32 HWIO')
    wd = w.detach()
    return _lazy_pack(w, (w.data_ptr(), tuple(w.shape), 'f32', transpose_flip),
                      lambda: torch.empty(taps * cin * cout, dtype=torch.float32, device=w.device),
                      lambda wp: call('ctgan_pack_filter_f32', _p(wd), _p(wp), taps, cin, cout, int(transpose_flip), _stream()),
                      cacheable)


def _fprop_tf32(x, wp, bias, residual, relu_mask, y, g, flags):
    d = _desc(g, F32, F32)
    call('ctgan_conv_fprop_tf32', ctypes.byref(d), _p(x), _p(wp), _p(bias), _p(residual), _p(relu_mask), _p(y), flags, _stream())
    return y


def _check_filter(w, g):
    _chk(w, 'filter')
    if w.dtype != torch.float32 or not w.is_contiguous() or w.numel() != g.kh * g.kw * g.Cin * g.Cout:
        raise RuntimeError('ctgan_b200: filter must be a contiguous float32 HWIO tensor matching the geometry')


# ---- thin-channel (3-channel image side) convolutions on the tensor cores: see include/ctgan_sm100.h
def _thin_side(g, t):
    """'in' / 'out' when the conv has a thin (<= 8 channel) input / output side the im2col GEMM path handles."""
    if not (config.use_tc and config.use_thin_tc and tc_available()) or t.dim() != 4 or t.dtype != torch.bfloat16:
        return None
    # (pad_t may fall outside [0, kh): the row groups of a filter split by _thin_split keep the whole filter's offsets)
    if not (g.stride == 1 and g.Ho == g.H and g.Wo == g.W and -8 < g.pad_t < 8 and 0 <= g.pad_l < g.kw and g.H * g.W >= 64):
        return None
    taps = g.kh * g.kw
    if g.Cin <= 8 and taps * g.Cin <= 64 and g.Cout % 64 == 0:
        return 'in'
    if g.Cout <= 8 and taps * g.Cout <= 64 and g.Cin % 64 == 0:
        return 'out'
    return None


def _thin_split(g, t):
    """[(r0, r1), ...]: filter-row groups for a stride-1 conv with a thin (<= 8 channel) side whose taps * C exceed the 64
    im2col columns of the thin tensor-core route (LSUN Generator.Output: 5x5, 64 -> 3: 75 columns) -- each group is a conv
    of its own (kh = r1 - r0, pad_t shifted) that fits; the family is linear in the filter, so fprop / dgrad are sums over
    the groups and wgrad fills the groups' row blocks of dw.  None when the call needs no split (or cannot use the route)."""
    if not (config.use_tc and config.use_thin_tc and tc_available()) or t.dim() != 4 or t.dtype != torch.bfloat16:
        return None
    if not (g.stride == 1 and g.Ho == g.H and g.Wo == g.W and 0 <= g.pad_t < g.kh and 0 <= g.pad_l < g.kw and g.H * g.W >= 64):
        return None
    C, Cw = min(g.Cin, g.Cout), max(g.Cin, g.Cout)
    if C > 8 or Cw % 64 or g.kh * g.kw * C <= 64 or g.kw * C > 64 or g.kh > 8:
        return None
    rows = 64 // (g.kw * C)
    return [(r, min(r + rows, g.kh)) for r in range(0, g.kh, rows)]


def _sub_geom(g, r0, r1):
    return g._replace(kh=r1 - r0, pad_t=g.pad_t - r0)


def pack_filter_thin(w, kind, cacheable=False):
    """float HWIO filter with one thin side -> zero-padded 64 x Cw bf16 operand (kinds: conv_tc.cu)."""
    key = (w.data_ptr(), tuple(w.shape), 2 + kind) if cacheable else None
    if key is not None and key in _pack_cache:
        return _pack_cache[key]
    taps, cin, cout = w.shape[0] * w.shape[1], w.shape[2], w.shape[3]
    wp = torch.empty(64 * max(cin, cout), dtype=torch.bfloat16, device=w.device)
    call('ctgan_pack_filter_thin', _p(w), _p(wp), taps, min(cin, cout), max(cin, cout), kind, _stream())
    if key is not None:
        _pack_cache[key] = wp
    return wp


def thin_col(t, g, role):
    """The im2col matrix conv_fprop/conv_wgrad (role 'x') or conv_dgrad/conv_wgrad (role 'dy') would build from t, or
    None when the call does not take the thin path -- lets a caller build it once and pass it to both."""
    if role == 'x' and s2d_geom(g, t) is not None:
        return space_to_depth(t, g)                     # shared by fprop and wgrad of a stride-2 conv
    if role == 'x' and thin_s2_ok(g, t):
        return im2col_strided(t, g)
    side = _thin_side(g, t)
    if side == 'in' and role == 'x':
        return im2col_thin(t, g, g.Cin, 1)
    if side == 'out' and role == 'dy':
        return im2col_thin(t, g, g.Cout, -1)
    return None


def im2col_thin(src, g, C, sign):
    col = empty_act((g.N, 64, g.H, g.W), torch.bfloat16, src.device)
    d = _desc(g, BF16, BF16)
    call('ctgan_im2col_thin', ctypes.byref(d), C, sign, _p(src), _p(col), _stream())
    return col


def _gemm1x1_tc(x, wp, bias, g, cin, cout, residual=None, flags=0):
    """1x1 tensor-core conv over the pixels of g: [P x cin] x packed filter -> [P x cout]."""
    g1 = ConvGeom(g.N, g.H, g.W, cin, g.H, g.W, cout, 1, 1, 1, 0, 0)
    s2d = bool(flags & _lib.EPI_OUT_S2D)
    y = empty_act((g.N, 4 * cout, g.H // 2, g.W // 2) if s2d else (g.N, cout, g.H, g.W), torch.bfloat16, x.device)
    d = _desc(g1, BF16, BF16)
    call('ctgan_conv_fprop_tc', ctypes.byref(d), _p(x), _p(wp), _p(bias), _p(residual), _p(y), flags, _stream())
    return y


def _col2im_thin(col, bias, g, C, sign):
    dst = empty_act((g.N, C, g.H, g.W), torch.bfloat16, col.device)
    d = _desc(g, BF16, BF16)
    call('ctgan_col2im_thin', ctypes.byref(d), C, sign, _p(col), _p(bias), _p(dst), _stream())
    return dst


class FilterPacker:
    """BF16 operand copies (fprop layout + tap-flipped dgrad layout) of every tensor-core-eligible filter of one
    flat parameter buffer, refreshed with ONE kernel launch after each optimizer step."""

    def __init__(self, flat_p, params, offsets):
        import numpy as np
        self.flat_p = flat_p
        self.param_ptrs = [p.data_ptr() for p in params.values() if p.dim() in (2, 4)]
        rows, self.views, dst = [], [], 0
        for name, p in params.items():
            if p.dim() == 4:
                taps, cin, cout = p.shape[0] * p.shape[1], p.shape[2], p.shape[3]
            elif p.dim() == 2 and name.endswith('.W'):
                taps, cin, cout = 1, p.shape[0], p.shape[1]
            else:
                continue
            if p.dim() == 4 and taps * min(cin, cout) <= 64 and min(cin, cout) <= 8 and max(cin, cout) % 64 == 0:
                for kind in ((0, 3) if cin < cout else (2, 1)):       # thin operands: 64 x Cw each
                    rows.append((offsets[name], dst, taps, cin, cout, 2 + kind))
                    self.views.append((p, 2 + kind, dst, 64 * max(cin, cout)))
                    dst += 64 * max(cin, cout)
                continue
            if cin % 64 or cout % 64 or taps > 9:       # k > 3: stride-2 layers, packed lazily in the layout their route needs
                continue
            for flip in (0, 1):
                rows.append((offsets[name], dst, taps, cin, cout, flip))
                self.views.append((p, flip, dst, taps * cin * cout))
                dst += (taps * cin * cout + 63) // 64 * 64
        # float operand packs of the same filters for the kind::tf32 path (kernels.config.tf32)
        rows32, self.views32, dst32 = [], [], 0
        if config.tf32:
            for name, p in params.items():
                if p.dim() == 4:
                    taps, cin, cout = p.shape[0] * p.shape[1], p.shape[2], p.shape[3]
                elif p.dim() == 2 and name.endswith('.W'):
                    taps, cin, cout = 1, p.shape[0], p.shape[1]
                else:
                    continue
                if cin % 32 or cout % 32 or taps > 9 or (cin % 128 and cout % 128):
                    continue
                for flip in (0, 1):
                    rows32.append((offsets[name], dst32, taps, cin, cout, flip))
                    self.views32.append((p, flip, dst32, taps * cin * cout))
                    dst32 += (taps * cin * cout + 63) // 64 * 64
        self.n32 = len(rows32)
        if self.n32:
            tab = np.zeros(self.n32, dtype=[('src', '<i8'), ('dst', '<i8'), ('taps', '<i4'), ('cin', '<i4'), ('cout', '<i4'), ('flip', '<i4')])
            for i, r in enumerate(rows32):
                tab[i] = r
            self.table32 = torch.from_numpy(tab.view(np.uint8).copy()).to(flat_p.device)
            self.packs32 = torch.empty(dst32, dtype=torch.float32, device=flat_p.device)
        self.n = len(rows)
        if self.n:
            tab = np.zeros(self.n, dtype=[('src', '<i8'), ('dst', '<i8'), ('taps', '<i4'), ('cin', '<i4'), ('cout', '<i4'), ('flip', '<i4')])
            for i, r in enumerate(rows):
                tab[i] = r
            self.table = torch.from_numpy(tab.view(np.uint8).copy()).to(flat_p.device)
            self.packs = torch.empty(dst, dtype=torch.bfloat16, device=flat_p.device)

    def refresh(self):
        """Re-pack everything (one launch) and publish the views in the pack cache of the current epoch."""
        if not (config.use_tc and tc_available()):
            return
        # the per-filter packs (space-to-depth / thin-channel operands) run on the side stream next to the one-launch pack
        derived = [e for ptr in self.param_ptrs for e in _derived_filters.get(ptr, ())]
        for _, recompute in derived:
            recompute()
        lazy = [ptr for ptr in self.param_ptrs + [d.data_ptr() for d, _ in derived] if _lazy_packs.get(ptr)]
        if lazy:
            on_side(lambda: refresh_lazy_packs(lazy), self.flat_p)
        if self.n:
            call('ctgan_pack_filters_multi', _p(self.flat_p), _p(self.packs), _p(self.table), self.n, _stream())
            for p, flip, dst, numel in self.views:
                _pack_cache[(p.data_ptr(), tuple(p.shape), flip)] = self.packs[dst:dst + numel]
        if self.n32:
            call('ctgan_pack_filters_multi_f32', _p(self.flat_p), _p(self.packs32), _p(self.table32), self.n32, _stream())
            for p, flip, dst, numel in self.views32:
                _pack_cache[(p.data_ptr(), tuple(p.shape), 'f32', flip)] = self.packs32[dst:dst + numel]
        if lazy:
            join_side()


# ---- raw tensor-core launches (operands already in kernel layout)
def _fprop_tc_packed(x, wp, bias, residual, relu_mask, y, g, flags):
    """y = conv(x, packed filter) on the stride-1 tcgen05 kernels; g is the stride-1 geometry the kernel sees."""
    d = _desc(g, BF16, BF16)
    if relu_mask is None:
        call('ctgan_conv_fprop_tc', ctypes.byref(d), _p(x), _p(wp), _p(bias), _p(residual), _p(y), flags, _stream())
    else:
        call('ctgan_conv_fprop_tc_masked', ctypes.byref(d), _p(x), _p(wp), _p(bias), _p(residual), _p(relu_mask), _p(y), flags, _stream())
    return y


def _wgrad_tc_raw(x, dy, g, dw):
    """dw (float HWIO of g, pre-initialised) += wgrad(x, dy) on the tcgen05 kernels."""
    d = _desc(g, BF16, BF16)
    call('ctgan_conv_wgrad_tc', ctypes.byref(d), _p(x), _p(dy), _p(dw), _stream())
    return dw


# ---- stride-2 5x5 'SAME' convs as 3x3 stride-1 tensor-core convs over the space-to-depth image (csrc/conv_s2d.cu)
def s2d_geom(g, t=None):
    """The 3x3 / stride-1 geometry (over [N, 4*Cin, ceil(H/2), ceil(W/2)]) equivalent to the stride-2 conv g, or None
    when the call does not take that route.  t: an activation of the call (dtype check)."""
    if not (config.use_tc and config.use_s2d and tc_available()):
        return None
    if t is not None and (t.dim() != 4 or t.dtype != torch.bfloat16):
        return None
    if g.stride != 2 or g.kh != g.kw or not (0 <= g.pad_t <= 2 and 0 <= g.pad_l <= 2):
        return None
    if g.kh - 1 - g.pad_t > 3 or g.kw - 1 - g.pad_l > 3 or g.kh < 3:      # block offsets of the taps must be {-1, 0, +1}
        return None
    Hs, Ws = (g.H + 1) // 2, (g.W + 1) // 2
    if Hs != g.Ho or Ws != g.Wo or min(Hs, Ws) < config.s2d_min_extent:
        return None
    if (4 * g.Cin) % 64 or g.Cout % 64:
        return None
    return ConvGeom(g.N, Hs, Ws, 4 * g.Cin, Hs, Ws, g.Cout, 3, 3, 1, 1, 1)


def _s2d_skip_flags(g, mode):
    """CTGAN_EPI_S2D_SKIP flags for the 3x3 launch that stands for the stride-2 conv g (mode 0: its fprop, 1: its dgrad)."""
    if not config.s2d_skip or g.Cin % (128 if mode else 64) or not (3 <= g.kh <= config.s2d_skip_max_k):
        return 0
    return _lib.EPI_S2D_SKIP | (mode << 9) | (g.kh << 10) | (g.pad_t << 14) | (g.pad_l << 16)


def space_to_depth_mask(x, pattern, g):
    """xs = space_to_depth(x) where `pattern` (a ReLU output in the space-to-depth layout) is positive, else 0."""
    require_nhwc(x, 'x'); require_nhwc(pattern, 'pattern')
    xs = empty_act((g.N, 4 * g.Cin, (g.H + 1) // 2, (g.W + 1) // 2), x.dtype, x.device)
    _same_layout(pattern, xs)
    call('ctgan_space_to_depth_mask', _p(x), _p(pattern), _p(xs), g.N, g.H, g.W, g.Cin, _dt(x), _stream())
    return xs


def space_to_depth(x, g, mul=None):
    """x [N,Cin,H,W] -> xs [N,4*Cin,ceil(H/2),ceil(W/2)], channel (dy*2+dx)*Cin + c <- x[.., 2i+dy, 2j+dx] (zero outside).
    mul: a multiplier in the layout of xs applied on the way (xs = space_to_depth(x) * mul)."""
    require_nhwc(x, 'x')
    xs = empty_act((g.N, 4 * g.Cin, (g.H + 1) // 2, (g.W + 1) // 2), x.dtype, x.device)
    if mul is None:
        call('ctgan_space_to_depth', _p(x), _p(xs), g.N, g.H, g.W, g.Cin, _dt(x), _stream())
    else:
        require_nhwc(mul, 'mul'); _same_layout(mul, xs)
        call('ctgan_space_to_depth_mul', _p(x), _p(mul), _p(xs), g.N, g.H, g.W, g.Cin, _dt(x), _stream())
    return xs


def depth_to_space(xs, g, mul=None):
    """Inverse of space_to_depth, cropped to the [N,Cin,H,W] of g.  mul (layout of xs): x = depth_to_space(xs * mul)."""
    require_nhwc(xs, 'xs')
    x = empty_act((g.N, g.Cin, g.H, g.W), xs.dtype, xs.device)
    if mul is None:
        call('ctgan_depth_to_space', _p(xs), _p(x), g.N, g.H, g.W, g.Cin, _dt(xs), _stream())
    else:
        require_nhwc(mul, 'mul'); _same_layout(mul, xs)
        call('ctgan_depth_to_space_mul', _p(xs), _p(mul), _p(x), g.N, g.H, g.W, g.Cin, _dt(xs), _stream())
    return x


def _pack_filter_s2d_launch(w, wp_f, wp_d, g):
    call('ctgan_pack_filter_s2d', _p(w), _p(wp_f), _p(wp_d), g.kh, g.Cin, g.Cout, g.pad_t, g.pad_l, _stream())


def _s2d_filter_grad_launch(dw3, dw, g, accumulate):
    call('ctgan_s2d_filter_grad', _p(dw3), _p(dw), g.kh, g.Cin, g.Cout, g.pad_t, g.pad_l, int(accumulate), _stream())


# Persistent operand packs of PARAMETER filters that the one-launch FilterPacker does not cover: {param ptr: {key: (launch, bufs)}}.
# They are created at a filter's first use (the layout can depend on the call: the space-to-depth pads follow the
# input extent) and re-packed IN PLACE after every optimizer step by FilterPacker.refresh(), so a captured CUDA
# graph always reads current weights from the same addresses.
_lazy_packs = {}


def _lazy_pack(w, key, make_bufs, launch, cacheable):
    if cacheable and key in _pack_cache:
        return _pack_cache[key]
    reg = _lazy_packs.get(w.data_ptr(), {}) if cacheable else {}
    bufs = reg[key][1] if key in reg else make_bufs()     # registered but invalidated by an update: re-pack in place
    launch(bufs)
    if cacheable:
        _pack_cache[key] = bufs
        _lazy_packs.setdefault(w.data_ptr(), {})[key] = (launch, bufs)
    return bufs


def refresh_lazy_packs(ptrs):
    """Re-pack (in place) every registered operand of the parameters at `ptrs`; publish them in the cache."""
    for ptr in ptrs:
        for key, (launch, bufs) in _lazy_packs.get(ptr, {}).items():
            launch(bufs)
            _pack_cache[key] = bufs


def pack_filter_s2d(w, g, flip, cacheable=False):
    """bf16 operands of the embedded 3x3 filter: flip 0 -> [9][Cout][4Cin] (fprop), 1 -> tap-flipped [9][4Cin][Cout] (dgrad)."""
    wd, n = w.detach(), 36 * g.Cin * g.Cout
    return _lazy_pack(w, (w.data_ptr(), tuple(w.shape), 's2d', g.pad_t, g.pad_l),
                      lambda: tuple(torch.empty(n, dtype=torch.bfloat16, device=w.device) for _ in range(2)),
                      lambda bufs: _pack_filter_s2d_launch(wd, bufs[0], bufs[1], g), cacheable)[flip]


# ---- stride-2 convs with a thin (<= 8 channel) input as im2col + 1x1 tensor-core GEMMs (csrc/conv_s2d.cu)
def thin_s2_ok(g, t=None):
    if not (config.use_tc and config.use_thin_s2 and tc_available()):
        return False
    if t is not None and (t.dim() != 4 or t.dtype != torch.bfloat16):
        return False
    return (g.stride == 2 and g.Cin <= 8 and g.kh * g.kw * g.Cin <= 128 and g.Cout % 64 == 0 and g.kh < 256 and g.kw < 256
            and g.Ho * g.Wo >= 16)


def _out_pixels_geom(g, cin, cout):
    """1x1 geometry over the OUTPUT pixels of g."""
    return ConvGeom(g.N, g.Ho, g.Wo, cin, g.Ho, g.Wo, cout, 1, 1, 1, 0, 0)


def im2col_strided(x, g):
    """col [N, 128, Ho, Wo] (NHWC: one 128-column row per output pixel), column (r*kw+s)*Cin + c."""
    require_nhwc(x, 'x')
    col = empty_act((g.N, 128, g.Ho, g.Wo), torch.bfloat16, x.device)
    d = _desc(g, BF16, BF16)
    call('ctgan_im2col_strided', ctypes.byref(d), g.Cin, _p(x), _p(col), _stream())
    return col


def col2im_strided(col, bias, g):
    dx = empty_act((g.N, g.Cin, g.H, g.W), torch.bfloat16, col.device)
    d = _desc(g, BF16, BF16)
    call('ctgan_col2im_strided', ctypes.byref(d), g.Cin, _p(col), _p(bias), _p(dx), _stream())
    return dx


def _pack_filter_padk_launch(w, wp_f, wp_d, g):
    call('ctgan_pack_filter_padk', _p(w), _p(wp_f), _p(wp_d), g.kh * g.kw * g.Cin, g.Cout, _stream())


def _add_prefix_launch(src, dst, n, accumulate):
    call('ctgan_add_prefix', _p(src), _p(dst), int(n), int(accumulate), _stream())


def pack_filter_padk(w, g, flip, cacheable=False):
    """bf16 operands of the filter as a [128 -> Cout] matrix (rows >= taps*Cin zero): flip 0 -> [Cout][128], 1 -> [128][Cout]."""
    wd, n = w.detach(), 128 * g.Cout
    return _lazy_pack(w, (w.data_ptr(), tuple(w.shape), 'padk'),
                      lambda: tuple(torch.empty(n, dtype=torch.bfloat16, device=w.device) for _ in range(2)),
                      lambda bufs: _pack_filter_padk_launch(wd, bufs[0], bufs[1], g), cacheable)[flip]


def conv_fprop_s2d_out_ok(x, g):
    """True when conv_fprop(x, ., ., g, out_s2d=True) has a route (the tcgen05 epilogue writes the space-to-depth layout)."""
    if not (config.use_tc and tc_available() and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4):
        return False
    if g.stride != 1 or g.Ho % 2 or g.Wo % 2 or g.Cout % 128:
        return False
    return _tc_geom_ok(g) or (_thin_split(g, x) is None and _thin_side(g, x) == 'in')


def conv_fprop(x, w, bias, g, relu=False, residual=None, out_dtype=None, w_is_param=False, col=None, res_up2=False,
               out_s2d=False):
    """y = conv(x, w) [+ bias] [+ residual] [relu].  x NHWC/2-D act, w float HWIO.
    res_up2: residual is [N, Cout, Ho/2, Wo/2] and is added nearest-neighbour upsampled (in the tensor-core epilogue).
    out_s2d (see conv_fprop_s2d_out_ok): y is returned in the space-to-depth layout [N, 4*Cout, Ho/2, Wo/2] of the stride-2
    conv that consumes it."""
    if out_s2d:
        if not conv_fprop_s2d_out_ok(x, g) or residual is not None:
            raise RuntimeError('ctgan_b200: conv_fprop(out_s2d=True) needs a tensor-core route and no residual')
        require_nhwc(x, 'x')
        _check_filter(w, g)
        flags = (_lib.EPI_RELU if relu else 0) | _lib.EPI_OUT_S2D
        if _tc_geom_ok(g):
            y = empty_act((g.N, 4 * g.Cout, g.Ho // 2, g.Wo // 2), torch.bfloat16, x.device)
            return _fprop_tc_packed(x, pack_filter(w, 0, cacheable=w_is_param), bias, None, None, y, g, flags)
        col = col if col is not None else im2col_thin(x, g, g.Cin, 1)
        return _gemm1x1_tc(col, pack_filter_thin(w, 0, cacheable=w_is_param), bias, g, 64, g.Cout, None, flags)
    if res_up2 and residual is not None and not (_dt(x) == BF16 and (out_dtype or x.dtype) == torch.bfloat16 and _tc_geom_ok(g)):
        residual, res_up2 = upsample2x(residual, 1.0), False        # other paths: materialise the upsampled residual
    require_nhwc(x, 'x')
    _check_filter(w, g)
    parts = _thin_split(g, x) if (not relu and residual is None and (out_dtype or x.dtype) == torch.bfloat16) else None
    if parts:                                         # thin side with > 64 im2col columns: the sum over filter-row groups
        y = None
        for i, (r0, r1) in enumerate(parts):
            yi = conv_fprop(x, w[r0:r1], bias if i == 0 else None, _sub_geom(g, r0, r1))
            y = yi if y is None else add(y, yi)
        return y
    two_d = x.dim() == 2
    out_dtype = out_dtype or x.dtype
    y = empty_act(_y_shape(g, two_d), out_dtype, x.device)
    xdt, ydt = _dt(x), _dt(y)
    flags = (_lib.EPI_RELU if relu else 0) | (_lib.EPI_RES_UP2 if (res_up2 and residual is not None) else 0)
    if xdt == BF16 and ydt == BF16 and _tc_geom_ok(g):
        wp = pack_filter(w, 0, cacheable=w_is_param)
        if residual is not None:
            require_nhwc(residual, 'residual')
        return _fprop_tc_packed(x, wp, bias, residual, None, y, g, flags)
    if xdt == F32 and ydt == F32 and _tf32_geom_ok(g):
        if residual is not None:
            require_nhwc(residual, 'residual')
        return _fprop_tf32(x, pack_filter_f32(w, 0, cacheable=w_is_param), bias, residual, None, y, g, flags)
    g3 = s2d_geom(g, x) if (xdt == BF16 and ydt == BF16) else None
    if g3 is not None:                                # stride 2: 3x3 conv over the space-to-depth image
        if residual is not None:
            require_nhwc(residual, 'residual')
        xs = col if (col is not None and tuple(col.shape) == (g3.N, g3.Cin, g3.H, g3.W)) else space_to_depth(x, g)
        return _fprop_tc_packed(xs, pack_filter_s2d(w, g, 0, cacheable=w_is_param), bias, residual, None, y, g3,
                                flags | _s2d_skip_flags(g, 0))
    if xdt == BF16 and ydt == BF16 and thin_s2_ok(g, x):      # y = im2col_strided(x) x W128
        if residual is not None:
            require_nhwc(residual, 'residual')
        col = col if (col is not None and tuple(col.shape) == (g.N, 128, g.Ho, g.Wo)) else im2col_strided(x, g)
        return _fprop_tc_packed(col, pack_filter_padk(w, g, 0, cacheable=w_is_param), bias, residual, None, y,
                                _out_pixels_geom(g, 128, g.Cout), flags)
    side = _thin_side(g, x) if (xdt == BF16 and ydt == BF16) else None
    if side == 'in':                                  # y = im2col(x) x w[(t,c)][Cout]
        if residual is not None:
            require_nhwc(residual, 'residual')
        col = col if col is not None else im2col_thin(x, g, g.Cin, 1)
        return _gemm1x1_tc(col, pack_filter_thin(w, 0, cacheable=w_is_param), bias, g, 64, g.Cout, residual, flags)
    if side == 'out' and not relu and residual is None:   # ycol[(t,o)] = x x w, y = col2im(ycol) + bias
        ycol = _gemm1x1_tc(x, pack_filter_thin(w, 2, cacheable=w_is_param), None, g, g.Cin, 64)
        return _col2im_thin(ycol, bias, g, g.Cout, 1)
    d = _desc(g, xdt, ydt)
    call('ctgan_conv_fprop', ctypes.byref(d), _p(x), _p(w), _p(bias), _p(y), flags, _stream())
    if residual is not None:
        if relu:
            raise RuntimeError('ctgan_b200: residual+relu fusion needs the tensor-core path')
        call('ctgan_add', _p(y), _p(residual), _p(y), y.numel(), ydt, _stream())
    return y


def conv_actdrop_route(x, g):
    """The tensor-core route whose epilogue can apply bias + LeakyReLU + dropout for the conv g on activation x
    ('tc' | 's2d' | 'padk'), or None (then conv_fprop + act_dropout run as two kernels)."""
    if not (config.use_tc and config.fuse_act_dropout and tc_available() and x.is_cuda and x.dtype == torch.bfloat16):
        return None
    if g.Cout % 128 or x.dim() != 4:
        return None
    if _tc_geom_ok(g):
        return 'tc'
    if s2d_geom(g, x) is not None:
        return 's2d'
    if thin_s2_ok(g, x):
        return 'padk'
    return None


def conv_fprop_actdrop(x, w, bias, g, slope, keep, seed, offset, dyn=None, out_s2d=False, w_is_param=False, col=None):
    """(y, m): y = v * m with v = conv(x, w) + bias and m = (v > 0 ? 1 : slope) * floor(keep + u) / keep, both from the
    conv epilogue (ctgan_conv_fprop_tc_actdrop).  out_s2d: y and m come in the space-to-depth layout
    [N, 4*Cout, Ho/2, Wo/2] of a following stride-2 layer.  Call only when conv_actdrop_route(x, g) is not None."""
    require_nhwc(x, 'x')
    _check_filter(w, g)
    route = conv_actdrop_route(x, g)
    if route is None:
        raise RuntimeError('ctgan_b200: conv_fprop_actdrop needs a tensor-core route (see conv_actdrop_route)')
    shape = (g.N, 4 * g.Cout, g.Ho // 2, g.Wo // 2) if out_s2d else (g.N, g.Cout, g.Ho, g.Wo)
    if out_s2d and (g.Ho % 2 or g.Wo % 2):
        raise RuntimeError('ctgan_b200: space-to-depth output needs even Ho, Wo')
    y, m = empty_act(shape, torch.bfloat16, x.device), empty_act(shape, torch.bfloat16, x.device)
    if route == 'tc':
        src, wp, gk = x, pack_filter(w, 0, cacheable=w_is_param), g
    elif route == 's2d':
        g3 = s2d_geom(g, x)
        src = col if (col is not None and tuple(col.shape) == (g3.N, g3.Cin, g3.H, g3.W)) else space_to_depth(x, g)
        wp, gk = pack_filter_s2d(w, g, 0, cacheable=w_is_param), g3
    else:
        src = col if (col is not None and tuple(col.shape) == (g.N, 128, g.Ho, g.Wo)) else im2col_strided(x, g)
        wp, gk = pack_filter_padk(w, g, 0, cacheable=w_is_param), _out_pixels_geom(g, 128, g.Cout)
    d = _desc(gk, BF16, BF16)
    call('ctgan_conv_fprop_tc_actdrop', ctypes.byref(d), _p(src), _p(wp), _p(bias), _p(y), _p(m), float(slope), float(keep),
         int(seed), int(offset), _p(dyn), int(bool(out_s2d)), _stream())
    return y, m


def conv_dgrad(dy, w, g, out_dtype=None, w_is_param=False, col=None, relu_mask=None, out_s2d=False):
    """dx = conv^T(dy, w) with the geometry of the FORWARD conv g.  Also Deconv2D forward.
    relu_mask: a tensor of dx's shape (the conv's input, itself a ReLU output): dx is zeroed where it is <= 0.
    out_s2d (space-to-depth route only): return dx in the space-to-depth layout [N, 4*Cin, H/2, W/2] -- the layout the
    conv's input was handed over in -- instead of converting it back."""
    require_nhwc(dy, 'dy')
    if relu_mask is not None:
        require_nhwc(relu_mask, 'relu_mask')
    _check_filter(w, g)
    parts = _thin_split(g, dy) if (not out_s2d and (out_dtype or dy.dtype) == torch.bfloat16) else None
    if parts:                                         # see conv_fprop
        dx = None
        for r0, r1 in parts:
            di = conv_dgrad(dy, w[r0:r1], _sub_geom(g, r0, r1))
            dx = di if dx is None else add(dx, di)
        return mul_relu_mask(dx, relu_mask) if relu_mask is not None else dx
    two_d = dy.dim() == 2
    out_dtype = out_dtype or dy.dtype
    dx = empty_act(_x_shape(g, two_d), out_dtype, dy.device)
    xdt, ydt = _dt(dx), _dt(dy)
    if xdt == BF16 and ydt == BF16 and _tc_geom_ok(g):
        # stride-1 dgrad == fprop of dy with the tap-flipped filter, Cin<->Cout, pad -> k-1-pad
        gt = ConvGeom(g.N, g.H, g.W, g.Cout, g.H, g.W, g.Cin, g.kh, g.kw, 1, g.kh - 1 - g.pad_t, g.kw - 1 - g.pad_l)
        wp = pack_filter(w, 1, cacheable=w_is_param)
        return _fprop_tc_packed(dy, wp, None, None, relu_mask, dx, gt, 0)
    if xdt == F32 and ydt == F32:
        gt = ConvGeom(g.N, g.H, g.W, g.Cout, g.H, g.W, g.Cin, g.kh, g.kw, 1, g.kh - 1 - g.pad_t, g.kw - 1 - g.pad_l)
        if g.stride == 1 and g.Ho == g.H and g.Wo == g.W and _tf32_geom_ok(gt):
            return _fprop_tf32(dy, pack_filter_f32(w, 1, cacheable=w_is_param), None, None, relu_mask, dx, gt, 0)
    g3 = s2d_geom(g, dy) if (xdt == BF16 and ydt == BF16) else None
    if relu_mask is not None and g3 is not None and tuple(relu_mask.shape) == (g3.N, g3.Cin, g3.H, g3.W):
        # the mask is the conv's input in the SPACE-TO-DEPTH layout (a ReLU output written that way by its producer)
        gt = ConvGeom(g3.N, g3.H, g3.W, g3.Cout, g3.H, g3.W, g3.Cin, 3, 3, 1, 1, 1)
        wp = pack_filter_s2d(w, g, 1, cacheable=w_is_param)
        if out_s2d or g.Cin % 128 or g.H % 2 or g.W % 2:
            dxs = empty_act((g3.N, g3.Cin, g3.H, g3.W), torch.bfloat16, dy.device)
            _fprop_tc_packed(dy, wp, None, None, relu_mask, dxs, gt, _s2d_skip_flags(g, 1))
            return dxs if out_s2d else depth_to_space(dxs, g)
        # masked in the epilogue (mask read in its own layout), written as the plain [N, Cin, H, W] tensor
        return _fprop_tc_packed(dy, wp, None, None, relu_mask, dx, gt, _s2d_skip_flags(g, 1) | _lib.EPI_OUT_D2S)
    if relu_mask is not None:                       # other paths: the mask as a separate kernel
        return mul_relu_mask(conv_dgrad(dy, w, g, out_dtype=out_dtype, w_is_param=w_is_param, col=col), relu_mask)
    if g3 is not None:                                # dxs = dgrad of the 3x3 conv (fprop with the flipped pack), dx = depth_to_space
        gt = ConvGeom(g3.N, g3.H, g3.W, g3.Cout, g3.H, g3.W, g3.Cin, 3, 3, 1, 1, 1)
        wp = pack_filter_s2d(w, g, 1, cacheable=w_is_param)
        dxs = empty_act((g3.N, g3.Cin, g3.H, g3.W), torch.bfloat16, dy.device)
        _fprop_tc_packed(dy, wp, None, None, None, dxs, gt, _s2d_skip_flags(g, 1))
        return dxs if out_s2d else depth_to_space(dxs, g)
    if out_s2d:
        raise RuntimeError('ctgan_b200: conv_dgrad(out_s2d=True) needs the space-to-depth route')
    if xdt == BF16 and ydt == BF16 and thin_s2_ok(g, dy):     # dcol = dy x W128^T, dx = col2im_strided(dcol)
        dcol = empty_act((g.N, 128, g.Ho, g.Wo), torch.bfloat16, dy.device)
        _fprop_tc_packed(dy, pack_filter_padk(w, g, 1, cacheable=w_is_param), None, None, None, dcol,
                         _out_pixels_geom(g, g.Cout, 128), 0)
        return col2im_strided(dcol, None, g)
    side = _thin_side(g, dy) if (xdt == BF16 and ydt == BF16) else None
    if side == 'in':                                  # dxcol[(t,ci)] = dy x w^T, dx = col2im(dxcol, -1)
        dxcol = _gemm1x1_tc(dy, pack_filter_thin(w, 3, cacheable=w_is_param), None, g, g.Cout, 64)
        return _col2im_thin(dxcol, None, g, g.Cin, -1)
    if side == 'out':                                 # dx = im2col(dy, -1) x w[(t,o)][Cin]
        col = col if col is not None else im2col_thin(dy, g, g.Cout, -1)
        return _gemm1x1_tc(col, pack_filter_thin(w, 1, cacheable=w_is_param), None, g, 64, g.Cin)
    d = _desc(g, xdt, ydt)
    call('ctgan_conv_dgrad', ctypes.byref(d), _p(dy), _p(w), _p(dx), _stream())
    return dx


# ---- deferred filter gradients: every tensor-core wgrad of a backward pass as ONE launch (csrc/conv_wgrad_multi.cu)
_wgrad_queue = []        # (x, dy, geometry, dw) jobs; the tensors stay alive until the flush
_wgrad_queue32 = []      # the same for float operands (kind::tf32)
_wgrad_post = []         # launches that consume a job's scratch result (space-to-depth gather, im2col prefix add)


def _wgrad_multi_ok(g):
    d = _desc(g, BF16, BF16)
    return bool(_lib.lib.ctgan_conv_wgrad_tc_multi_ok(ctypes.byref(d)))


def _launch_wgrad_jobs(jobs, dt, entry):
    """jobs: (x, dy, geometry, dw[, embed]) -- embed = (k, C, pad_t, pad_l) of the stride-2 filter a space-to-depth job stands for."""
    n = len(jobs)
    descs = (ConvDesc * n)(*[_desc(j[2], dt, dt) for j in jobs])
    xs = (ctypes.c_void_p * n)(*[j[0].data_ptr() for j in jobs])
    dys = (ctypes.c_void_p * n)(*[j[1].data_ptr() for j in jobs])
    dws = (ctypes.c_void_p * n)(*[j[3].data_ptr() for j in jobs])
    if any(len(j) > 4 and j[4] is not None for j in jobs):
        emb = (ctypes.c_int * (4 * n))(*[v for j in jobs for v in (j[4] if len(j) > 4 and j[4] is not None else (0, 0, 0, 0))])
        call(entry + '_embed', n, descs, xs, dys, dws, emb, _stream())
    else:
        call(entry, n, descs, xs, dys, dws, _stream())


def flush_wgrads():
    """Launch the queued filter gradients (one kernel per 24 jobs) and their post-processing on the current stream."""
    if _wgrad_queue32:
        jobs = list(_wgrad_queue32)
        _wgrad_queue32.clear()
        _launch_wgrad_jobs(jobs, F32, 'ctgan_conv_wgrad_tf32_multi')
    if not _wgrad_queue:
        return
    jobs, post = list(_wgrad_queue), list(_wgrad_post)
    _wgrad_queue.clear()
    _wgrad_post.clear()
    _launch_wgrad_jobs(jobs, BF16, 'ctgan_conv_wgrad_tc_multi')
    for fn in post:
        fn()


def _wgrad_tf32_ok(g):
    d = _desc(g, F32, F32)
    return _tf32_geom_ok(g, 128) and bool(_lib.lib.ctgan_conv_wgrad_tf32_multi_ok(ctypes.byref(d)))


def _wgrad_tc(x, dy, g, dw, defer, post=None, embed=None):
    """dw += wgrad(x, dy) on the tensor cores, now or (defer) at the next join_side(); post() consumes dw afterwards.
    embed: see _launch_wgrad_jobs (the job reduces straight into the stride-2 filter's gradient)."""
    if defer and config.defer_wgrad and _wgrad_multi_ok(g):
        _wgrad_queue.append((x, dy, g, dw, embed))
        if post is not None:
            _wgrad_post.append(post)
        return
    if embed is not None or g.Cin % 128 or g.Cout % 128:
        _launch_wgrad_jobs([(x, dy, g, dw, embed)], BF16, 'ctgan_conv_wgrad_tc_multi')
        if post is not None:
            post()
        return
    _wgrad_tc_raw(x, dy, g, dw)
    if post is not None:
        post()


def _wgrad_route(x, dy, g):
    """Which kernel family conv_wgrad takes: ('tc', g) | ('s2d', g3) | ('padk', g1) | ('thin', side) | ('simt', None)."""
    xdt, ydt = _dt(x), _dt(dy)
    bf = xdt == BF16 and ydt == BF16
    if xdt == F32 and ydt == F32 and _wgrad_tf32_ok(g):
        return 'tf32', g
    if bf and _tc_geom_ok(g) and ((g.Cin % 128 == 0 and g.Cout % 128 == 0) or
                                  (g.Cin % 64 == 0 and g.Cout % 64 == 0 and _wgrad_multi_ok(g))):
        return 'tc', g        # 64 (mod 128) channels: only the multi-job kernel (zero-filled upper operand half)
    g3 = s2d_geom(g, x) if bf else None
    if g3 is not None and g3.Cin % 128 == 0 and g3.Cout % 128 == 0:
        return 's2d', g3
    if bf and thin_s2_ok(g, x) and g.Cout % 128 == 0:
        return 'padk', _out_pixels_geom(g, 128, g.Cout)
    side = _thin_side(g, x) if bf else None
    if side is not None:
        return 'thin', side
    return 'simt', None


def wgrad_deferrable(x, dy, g):
    """True when conv_wgrad(x, dy, g, ..., defer=True) only queues work (no launch on the calling stream)."""
    if not (config.defer_wgrad and x.is_cuda):
        return False
    route, gj = _wgrad_route(x, dy, g)
    if route == 'tf32':
        return True
    return route in ('tc', 's2d', 'padk') and _wgrad_multi_ok(gj)


def conv_wgrad(x, dy, g, w_shape, accumulate_into=None, col=None, defer=False):
    """dw (float HWIO, shape w_shape) = sum over pixels of x (shifted) * dy.
    accumulate_into: a float tensor of that shape (e.g. the parameter's slice of the flat gradient
    bucket) to ADD the result to instead of allocating one; returns it.
    defer (with accumulate_into): the tensor-core routes only queue the job; K.join_side() launches the queue."""
    require_nhwc(x, 'x')
    require_nhwc(dy, 'dy')
    xdt, ydt = _dt(x), _dt(dy)
    d = _desc(g, xdt, ydt)
    acc = accumulate_into
    if acc is not None and (acc.dtype != torch.float32 or not acc.is_contiguous() or tuple(acc.shape) != tuple(w_shape)):
        raise RuntimeError('ctgan_b200: accumulate_into must be a contiguous float32 tensor of the filter shape')
    defer = defer and acc is not None
    parts = _thin_split(g, x) if (xdt == BF16 and ydt == BF16) else None
    if parts:                                         # see conv_fprop: every row group fills its own block of dw
        dw = acc if acc is not None else zeros(w_shape, torch.float32, x.device)
        for r0, r1 in parts:
            conv_wgrad(x, dy, _sub_geom(g, r0, r1), (r1 - r0,) + tuple(w_shape[1:]), accumulate_into=dw[r0:r1])
        return dw
    route, gj = _wgrad_route(x, dy, g)
    if route == 'tf32':
        dw = acc if acc is not None else zeros(w_shape, torch.float32, x.device)
        if defer and config.defer_wgrad:
            _wgrad_queue32.append((x, dy, g, dw))
        else:
            _launch_wgrad_jobs([(x, dy, g, dw)], F32, 'ctgan_conv_wgrad_tf32_multi')
        return dw
    if route == 'tc':
        dw = acc if acc is not None else zeros(w_shape, torch.float32, x.device)
        _wgrad_tc(x, dy, g, dw, defer)
        return dw
    if route == 's2d':
        g3 = gj
        xs = col if (col is not None and tuple(col.shape) == (g3.N, g3.Cin, g3.H, g3.W)) else space_to_depth(x, g)
        if config.s2d_embed_wgrad and g.Cin % 128 == 0 and g.kh == g.kw and g.kh <= 5 and _wgrad_multi_ok(g3):
            # only the taps of the embedded filter that carry an element of w, reduced straight into its gradient
            dw = acc if acc is not None else zeros(w_shape, torch.float32, x.device)
            _wgrad_tc(xs, dy, g3, dw, defer, embed=(g.kh, g.Cin, g.pad_t, g.pad_l))
            return dw
        dw3 = zeros((3, 3, g3.Cin, g3.Cout), torch.float32, x.device)
        dw = acc if acc is not None else torch.empty(w_shape, dtype=torch.float32, device=x.device)
        _wgrad_tc(xs, dy, g3, dw3, defer, post=lambda: _s2d_filter_grad_launch(dw3, dw, g, acc is not None))
        return dw
    if route == 'padk':
        col = col if (col is not None and tuple(col.shape) == (g.N, 128, g.Ho, g.Wo)) else im2col_strided(x, g)
        dw128 = zeros((1, 1, 128, g.Cout), torch.float32, x.device)
        dw = acc if acc is not None else torch.empty(w_shape, dtype=torch.float32, device=x.device)
        n_real = g.kh * g.kw * g.Cin * g.Cout                                   # HWIO order == column order
        _wgrad_tc(col, dy, gj, dw128, defer, post=lambda: _add_prefix_launch(dw128, dw, n_real, acc is not None))
        return dw
    if route == 'thin':
        side = gj
        dw = acc if acc is not None else zeros(w_shape, torch.float32, x.device)
        P, taps = g.N * g.H * g.W, g.kh * g.kw
        if side == 'in':
            col = col if col is not None else im2col_thin(x, g, g.Cin, 1)
            call('ctgan_wgrad_thin_tc', _p(dy), _p(col), P, g.Cout, g.Cin, taps, 0, _p(dw), _stream())
        else:
            col = col if col is not None else im2col_thin(dy, g, g.Cout, -1)
            call('ctgan_wgrad_thin_tc', _p(x), _p(col), P, g.Cin, g.Cout, taps, 1, _p(dw), _stream())
        return dw
    dw = acc if acc is not None else torch.empty(w_shape, dtype=torch.float32, device=x.device)
    call('ctgan_conv_wgrad', ctypes.byref(d), _p(x), _p(dy), _p(dw), 1 if acc is not None else 0, _stream())
    return dw


def bias_grad(dy, accumulate_into=None):
    require_nhwc(dy, 'dy')
    N, H, W, C = nhwc_dims(dy)
    db = accumulate_into if accumulate_into is not None else torch.empty(C, dtype=torch.float32, device=dy.device)
    call('ctgan_bias_grad', _p(dy), _p(db), N * H * W, C, _dt(dy), 1 if accumulate_into is not None else 0, _stream())
    return db


def bias_add(x, b):
    require_nhwc(x, 'x')
    N, H, W, C = nhwc_dims(x)
    y = torch.empty_like(x)
    call('ctgan_bias_add', _p(x), _p(b), _p(y), N * H * W, C, _dt(x), _stream())
    return y


# --------------------------------------------------------------------------- element-wise
def _same_layout(a, b):
    if a.shape != b.shape or a.dtype != b.dtype or a.stride() != b.stride():
        raise RuntimeError('ctgan_b200: element-wise operands must share shape, dtype and layout')


def _dense(x, name='tensor'):
    _chk(x, name)
    if not (x.is_contiguous() or (x.dim() == 4 and x.is_contiguous(memory_format=CL))):
        raise RuntimeError('ctgan_b200: %s must be dense' % name)
    return x


def add(a, b):
    _dense(a); _dense(b); _same_layout(a, b)
    out = torch.empty_like(a)
    call('ctgan_add', _p(a), _p(b), _p(out), a.numel(), _dt(a), _stream())
    return out


def mul(a, b):
    _dense(a); _dense(b); _same_layout(a, b)
    out = torch.empty_like(a)
    call('ctgan_mul', _p(a), _p(b), _p(out), a.numel(), _dt(a), _stream())
    return out


def scale(a, s):
    _dense(a)
    out = torch.empty_like(a)
    call('ctgan_scale', _p(a), float(s), _p(out), a.numel(), _dt(a), _stream())
    return out


def cast(x, dtype):
    _dense(x)
    if x.dtype == dtype:
        return x
    out = torch.empty_like(x, dtype=dtype)
    call('ctgan_cast', _p(x), _dt(x), _p(out), _dt(out), x.numel(), _stream())
    return out


def act_dropout(x, slope, keep, u=None, seed=0, offset=0, want_mask=True, dyn=None):
    """y = x * m,  m = (x>0 ? 1 : slope) * (keep<1 ? floor(keep+u)/keep : 1).
    `u` (float32, same logical shape/layout as x) overrides the Philox stream (seed, offset)."""
    _dense(x)
    y = torch.empty_like(x)
    m = torch.empty_like(x) if want_mask else None
    if u is not None:
        _dense(u)
        if u.dtype != torch.float32 or u.shape != x.shape or u.stride() != x.stride():
            raise RuntimeError('ctgan_b200: explicit dropout noise must be float32 with the layout of x')
    call('ctgan_act_dropout_fwd', _p(x), _p(u), _p(y), _p(m), x.numel(), _dt(x), float(slope), float(keep),
         int(seed), int(offset), _p(dyn), _stream())
    return y, m


def fork_dropout_relu(x, keep, u=None, seed=0, offset=0, dyn=None):
    """d = dropout(x), r = relu(d) and their multipliers (md, mdr) in one kernel; same Philox slice as act_dropout."""
    _dense(x)
    d, r, md, mdr = (torch.empty_like(x) for _ in range(4))
    if u is not None:
        _dense(u)
        if u.dtype != torch.float32 or u.shape != x.shape or u.stride() != x.stride():
            raise RuntimeError('ctgan_b200: explicit dropout noise must be float32 with the layout of x')
    call('ctgan_fork_dropout_relu', _p(x), _p(u), _p(d), _p(r), _p(md), _p(mdr), x.numel(), _dt(x), float(keep),
         int(seed), int(offset), _p(dyn), _stream())
    return d, r, md, mdr


def mask_sum2(a, ma, b, mb):
    """a*ma + b*mb (ma None: a + b*mb)."""
    _dense(a); _dense(b); _dense(mb); _same_layout(a, b); _same_layout(b, mb)
    if ma is not None:
        _dense(ma); _same_layout(a, ma)
    out = torch.empty_like(a)
    call('ctgan_mask_sum2', _p(a), _p(ma), _p(b), _p(mb), _p(out), a.numel(), _dt(a), _stream())
    return out


def mask_fork2(c, ma, mb):
    """(c*ma, c*mb)."""
    _dense(c); _dense(ma); _dense(mb); _same_layout(c, ma); _same_layout(c, mb)
    o1, o2 = torch.empty_like(c), torch.empty_like(c)
    call('ctgan_mask_fork2', _p(c), _p(ma), _p(mb), _p(o1), _p(o2), c.numel(), _dt(c), _stream())
    return o1, o2


def mul_relu_mask(g, y):
    """g * [y > 0]: the backward of a ReLU whose output is y."""
    _dense(g); _dense(y); _same_layout(g, y)
    out = torch.empty_like(g)
    call('ctgan_mul_relu_mask', _p(g), _p(y), _p(out), g.numel(), _dt(g), _stream())
    return out


def pool_add_fork(y, s, keep=1.0, u=None, seed=0, offset=0, dyn=None, masks=None):
    """x = meanpool2x2(y) + s; returns (o1, o2, m1, m2) with o1 = x*m1, o2 = x*m2.
    masks=None: m1 = dropout multiplier (None when keep == 1), m2 = m1*[x>0] are computed; masks=(m1, m2): given."""
    require_nhwc(y); require_nhwc(s)
    N, C, H, W = y.shape
    if tuple(s.shape) != (N, C, H // 2, W // 2) or s.dtype != y.dtype:
        raise RuntimeError('ctgan_b200: pool_add_fork operands do not match')
    o1, o2 = torch.empty_like(s), torch.empty_like(s)
    if masks is None:
        m1 = torch.empty_like(s) if keep < 1.0 else None
        m2 = torch.empty_like(s)
        if u is not None:
            _dense(u)
            if u.dtype != torch.float32 or u.shape != s.shape or u.stride() != s.stride():
                raise RuntimeError('ctgan_b200: explicit dropout noise must be float32 with the layout of the output')
        compute = 1
    else:
        m1, m2 = masks
        compute = 0
    call('ctgan_pool_add_fork', compute, _p(y), _p(s), _p(u), _p(m1), _p(m2), _p(o1), _p(o2), N, H, W, C, _dt(y), float(keep),
         int(seed), int(offset), _p(dyn), _stream())
    return o1, o2, m1, m2


def mask_sum2_up(a, m1, b, m2):
    """gx = a*m1 + b*m2 (m1 None: a + b*m2) and gy = 0.25*gx replicated 2x2; returns (gy, gx)."""
    require_nhwc(a); _same_layout(a, b); _same_layout(b, m2)
    N, C, Ho, Wo = a.shape
    gx = torch.empty_like(a)
    gy = empty_act((N, C, 2 * Ho, 2 * Wo), a.dtype, a.device)
    call('ctgan_mask_sum2_up', _p(a), _p(m1), _p(b), _p(m2), _p(gx), _p(gy), N, Ho, Wo, C, _dt(a), _stream())
    return gy, gx


def unary_fwd(x, kind):
    _dense(x)
    y = torch.empty_like(x)
    call('ctgan_unary_fwd', _p(x), _p(y), x.numel(), _dt(x), kind, _stream())
    return y


def unary_bwd(y, dy, kind):
    _dense(y); _dense(dy); _same_layout(y, dy)
    dx = torch.empty_like(y)
    call('ctgan_unary_bwd', _p(y), _p(dy), _p(dx), y.numel(), _dt(y), kind, _stream())
    return dx


def pool2x2(x, scale_):
    require_nhwc(x)
    N, C, H, W = x.shape
    y = empty_act((N, C, H // 2, W // 2), x.dtype, x.device)
    call('ctgan_pool2x2', _p(x), _p(y), N, H, W, C, float(scale_), _dt(x), _stream())
    return y


def upsample2x(x, scale_):
    require_nhwc(x)
    N, C, H, W = x.shape
    y = empty_act((N, C, 2 * H, 2 * W), x.dtype, x.device)
    call('ctgan_upsample2x', _p(x), _p(y), N, H, W, C, float(scale_), _dt(x), _stream())
    return y


def spatial_sum(x, scale_):
    require_nhwc(x)
    N, C, H, W = x.shape
    y = torch.empty((N, C), dtype=x.dtype, device=x.device)
    call('ctgan_spatial_sum', _p(x), _p(y), N, H * W, C, float(scale_), _dt(x), _stream())
    return y


def spatial_bcast(y, H, W, scale_):
    _dense(y)
    N, C = y.shape
    x = empty_act((N, C, H, W), y.dtype, y.device)
    call('ctgan_spatial_bcast', _p(y), _p(x), N, H * W, C, float(scale_), _dt(y), _stream())
    return x


def nchw_to_nhwc(x, N, C, H, W, out_dtype):
    """x: any dense tensor holding N*C*H*W elements in NCHW order -> logical [N,C,H,W] channels_last."""
    _chk(x)
    if not x.is_contiguous():
        raise RuntimeError('ctgan_b200: nchw_to_nhwc input must be contiguous')
    y = empty_act((N, C, H, W), out_dtype, x.device)
    call('ctgan_nchw_to_nhwc', _p(x), _dt(x), _p(y), _dt(y), N, C, H, W, _stream())
    return y


def nhwc_to_nchw(x, out_dtype, out_shape):
    """x: logical [N,C,H,W] channels_last -> contiguous tensor of `out_shape` in NCHW element order."""
    require_nhwc(x)
    N, C, H, W = x.shape
    y = torch.empty(out_shape, dtype=out_dtype, device=x.device)
    call('ctgan_nhwc_to_nchw', _p(x), _dt(x), _p(y), _dt(y), N, C, H, W, _stream())
    return y


def crop(x, h, w):
    require_nhwc(x)
    N, C, H, W = x.shape
    y = empty_act((N, C, h, w), x.dtype, x.device)
    call('ctgan_crop', _p(x), _p(y), N, H, W, C, h, w, _dt(x), _stream())
    return y


def crop_bwd(dy, H, W):
    require_nhwc(dy)
    N, C, h, w = dy.shape
    dx = empty_act((N, C, H, W), dy.dtype, dy.device)
    call('ctgan_crop_bwd', _p(dy), _p(dx), N, H, W, C, h, w, _dt(dy), _stream())
    return dx


def prep_real(x_int, denom, noise_hi=0., seed=0, offset=0, dyn=None, out=None, out2=None):
    """2 * (x / denom - .5) [+ U[0, noise_hi)].  out / out2: contiguous float destinations of x's shape (e.g. two row ranges of
    a stacked critic input) written instead of a fresh tensor."""
    _chk(x_int)
    if x_int.dtype not in (torch.int32, torch.uint8) or not x_int.is_contiguous():
        raise RuntimeError('ctgan_b200: real data must be contiguous int32 or uint8')
    y = out if out is not None else torch.empty(x_int.shape, dtype=torch.float32, device=x_int.device)
    for t in (y, out2):
        if t is not None and (t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != x_int.numel()):
            raise RuntimeError('ctgan_b200: prep_real destinations must be contiguous float32 of the input size')
    call('ctgan_prep_real_dup', _p(x_int), int(x_int.dtype == torch.uint8), _p(y), _p(out2), x_int.numel(), float(denom), float(noise_hi),
         int(seed), int(offset), _p(dyn), _stream())
    return y


def zeros_like(t):
    """torch.zeros_like through a memset node (no fill kernel inside the captured steps)."""
    out = torch.empty_like(t)
    return zero_(out) if (out.is_cuda and out.is_contiguous()) else out.zero_()


def zeros(shape, dtype, device):
    out = torch.empty(shape, dtype=dtype, device=device)
    return zero_(out) if out.is_cuda else out.zero_()


def zero_(t):
    """In-place zero fill as a memset node (no fill kernel)."""
    _chk(t)
    if not t.is_contiguous():
        raise RuntimeError('ctgan_b200: zero_ needs a contiguous tensor')
    call('ctgan_memset_zero', _p(t), t.numel() * t.element_size(), _stream())
    return t


def interpolate(real, fake, alpha):
    for t in (real, fake, alpha):
        _chk(t)
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError('ctgan_b200: interpolate operands must be contiguous float32')
    B, P_ = real.shape
    out = torch.empty_like(real)
    call('ctgan_interpolate', _p(real), _p(fake), _p(alpha), _p(out), B, P_, _stream())
    return out


# --------------------------------------------------------------------------- batch norm
def bn_fused_ok(x, groups=1):
    """True when the two-kernel BF16 batch-norm path (ctgan_bn_fwd_fused / ctgan_bn_bwd_fused) takes this activation."""
    if not (x.is_cuda and x.dtype == torch.bfloat16 and config.use_bn_fused):
        return False
    N, H, W, C = nhwc_dims(x)
    return bool(_lib.lib.ctgan_bn_fused_ok(N, H, W, C, groups, BF16)) and x.data_ptr() % 16 == 0


def bn_fwd(x, gamma, beta, labels, eps, relu, groups=1, up2=False):
    """Training-mode batch norm (+ per-label gamma/beta rows, + ReLU).  up2: the output is written nearest-neighbour
    upsampled 2x ([N, C, 2H, 2W]) -- the UpsampleConv that follows Normalize + relu in the generator blocks."""
    require_nhwc(x)
    N, H, W, C = nhwc_dims(x)
    mean = torch.empty((groups, C), dtype=torch.float32, device=x.device)
    invstd = torch.empty((groups, C), dtype=torch.float32, device=x.device)
    if bn_fused_ok(x, groups):
        y = empty_act((N, C, 2 * H, 2 * W), x.dtype, x.device) if up2 else torch.empty_like(x)
        ws = torch.empty(2 * groups * C, dtype=torch.float32, device=x.device)
        flags = (_lib.BN_RELU if relu else 0) | (_lib.BN_UP2 if up2 else 0)
        call('ctgan_bn_fwd_fused', _p(x), _p(gamma), _p(beta), _p(labels), _p(y), _p(mean), _p(invstd), _p(ws),
             N, H, W, C, float(eps), flags, int(groups), _stream())
        return y, mean, invstd
    y = torch.empty_like(x)
    ws = torch.empty(_lib.lib.ctgan_bn_workspace_floats(N, H * W, C, groups), dtype=torch.float32, device=x.device)
    call('ctgan_bn_fwd', _p(x), _p(gamma), _p(beta), _p(labels), _p(y), _p(mean), _p(invstd), _p(ws),
         N, H * W, C, float(eps), int(relu), int(groups), _dt(x), _stream())
    if up2:
        y = upsample2x(y, 1.0)
    return y, mean, invstd


def bn_bwd(dy, x, y, gamma, beta, labels, mean, invstd, relu, groups=1, up2=False, accumulate_into=None):
    """Returns (dx, dgamma, dbeta).  y: the forward output (only the three-kernel path reads it: the ReLU pattern; with
    up2 it is the upsampled output).  up2: dy has the upsampled shape.  accumulate_into=(dgamma, dbeta): float
    accumulators (slices of the flat gradient bucket) the table gradients are ADDED to; they are returned."""
    require_nhwc(dy); require_nhwc(x)
    N, H, W, C = nhwc_dims(x)
    n_labels = gamma.numel() // C
    dx = torch.empty_like(x)
    if bn_fused_ok(x, groups) and dy.data_ptr() % 16 == 0:
        if accumulate_into is not None:
            dgamma, dbeta = accumulate_into
        else:
            dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(gamma)
        ws = torch.empty(2 * groups * C, dtype=torch.float32, device=x.device)
        flags = (_lib.BN_RELU if relu else 0) | (_lib.BN_UP2 if up2 else 0) | (_lib.BN_ACCUM if accumulate_into is not None else 0)
        call('ctgan_bn_bwd_fused', _p(dy), _p(x), _p(gamma), _p(beta), _p(labels), _p(mean), _p(invstd), _p(dx), _p(dgamma),
             _p(dbeta), _p(ws), N, H, W, C, n_labels, flags, int(groups), _stream())
        return dx, dgamma, dbeta
    if up2:
        dy = pool2x2(dy, 1.0)                       # adjoint of the nearest-neighbour upsample: 2x2 sums
        y = y[:, :, ::2, ::2].contiguous(memory_format=CL) if (relu and y is not None) else y
    dgamma = torch.empty_like(gamma)
    dbeta = torch.empty_like(gamma)
    ws = torch.empty(_lib.lib.ctgan_bn_workspace_floats(N, H * W, C, groups), dtype=torch.float32, device=x.device)
    call('ctgan_bn_bwd', _p(dy), _p(x), _p(y), _p(gamma), _p(labels), _p(mean), _p(invstd), _p(dx), _p(dgamma),
         _p(dbeta), _p(ws), N, H * W, C, n_labels, int(relu), int(groups), _dt(x), _stream())
    if accumulate_into is not None:
        accumulate_into[0].add_(dgamma); accumulate_into[1].add_(dbeta)
        return dx, accumulate_into[0], accumulate_into[1]
    return dx, dgamma, dbeta


# --------------------------------------------------------------------------- layer norm (csrc/layernorm.cu)
def _ln_dims(x):
    require_nhwc(x)
    if x.dim() != 4:
        raise RuntimeError('ctgan_b200: layer norm expects a 4-D activation')
    N, C = x.shape[0], x.shape[1]
    return N, x.numel() // N, C


def _ln_ws(N, M, device):
    return torch.empty(_lib.lib.ctgan_ln_workspace_floats(N, M), dtype=torch.float32, device=device)


def ln_fwd(x, gamma, beta, eps):
    """y = (x - mean_s) * rstd_s * gamma_c + beta_c over each sample's (C,H,W); returns (y, mean [N], rstd [N])."""
    N, M, C = _ln_dims(x)
    y = torch.empty_like(x)
    mean = torch.empty(N, dtype=torch.float32, device=x.device)
    rstd = torch.empty(N, dtype=torch.float32, device=x.device)
    call('ctgan_ln_fwd', _p(x), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd), _p(_ln_ws(N, M, x.device)), N, M, C, float(eps),
         _dt(x), _stream())
    return y, mean, rstd


def ln_core(v, x, gamma, mean, rstd, pre_scale, post_scale):
    """[gamma *] rstd * (u - mean_s(u) - xh * mean_s(u * xh)),  u = [gamma *] v."""
    N, M, C = _ln_dims(x)
    require_nhwc(v); _same_layout(v, x)
    out = torch.empty_like(x)
    call('ctgan_ln_core', _p(v), _p(x), _p(gamma), _p(mean), _p(rstd), _p(out), _p(_ln_ws(N, M, x.device)), N, M, C,
         int(pre_scale), int(post_scale), _dt(x), _stream())
    return out


def ln_param_grad(v, x, mean, rstd, dgamma, dbeta):
    """dgamma[c] += sum v * xh, dbeta[c] += sum v (float accumulators, e.g. slices of the flat gradient bucket)."""
    N, M, C = _ln_dims(x)
    require_nhwc(v); _same_layout(v, x)
    call('ctgan_ln_param_grad', _p(v), _p(x), _p(mean), _p(rstd), _p(dgamma), _p(dbeta), N, M, C, _dt(x), _stream())


def ln_bwd2_x(c, gy, x, gamma, mean, rstd):
    """The x-derivative of <c, core(gamma * gy)>."""
    N, M, C = _ln_dims(x)
    require_nhwc(c); require_nhwc(gy); _same_layout(c, x); _same_layout(gy, x)
    gx = torch.empty_like(x)
    call('ctgan_ln_bwd2_x', _p(c), _p(gy), _p(x), _p(gamma), _p(mean), _p(rstd), _p(gx), _p(_ln_ws(N, M, x.device)), N, M, C,
         _dt(x), _stream())
    return gx


# --------------------------------------------------------------------------- losses
def _f32c(t, name):
    _chk(t, name)
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise RuntimeError('ctgan_b200: %s must be contiguous float32' % name)
    return t


def ct_gp_loss_fwd(desc, d_real, d_real2, d_fake, f1, f2, grad, logits, labels):
    """grad None: no penalty term; d_real .. f2 None: the penalty term only (include/ctgan_sm100.h)."""
    dev = (d_real if d_real is not None else grad).device
    out = torch.empty(8, dtype=torch.float32, device=dev)
    per_sample = torch.empty(4 * desc.B, dtype=torch.float32, device=dev)
    call('ctgan_ct_gp_loss_fwd', ctypes.byref(desc), _p(d_real), _p(d_real2), _p(d_fake), _p(f1), _p(f2), _p(grad),
         _p(logits), _p(labels), _p(out), _p(per_sample), _stream())
    return out, per_sample


def ct_gp_loss_bwd(desc, gcost, d_real, d_real2, f1, f2, grad, logits, labels, per_sample, outs=None):
    """outs: optional (g_real, g_real2, g_fake, g_f1, g_f2, g_logits) contiguous destination views (row ranges of the
    gradients of stacked critic outputs), so that no slice-backward / accumulation kernels are needed."""
    dev = (d_real if d_real is not None else grad).device
    if d_real is None:                                   # penalty only
        g_grad = torch.empty_like(grad)
        call('ctgan_ct_gp_loss_bwd', ctypes.byref(desc), _p(gcost), None, None, None, None, _p(grad), None, None, _p(per_sample),
             None, None, None, None, None, _p(g_grad), None, _stream())
        return None, None, None, None, None, g_grad, None
    if outs is not None:
        g_real, g_real2, g_fake, g_f1, g_f2, g_logits = outs
        for t in outs:
            if t is not None and not t.is_contiguous():
                raise RuntimeError('ctgan_b200: loss gradient destinations must be contiguous')
    else:
        g_real = torch.empty(desc.B, dtype=torch.float32, device=dev)
        g_real2 = torch.empty(desc.B, dtype=torch.float32, device=dev)
        g_fake = torch.empty(desc.NF, dtype=torch.float32, device=dev)
        g_f1 = torch.empty_like(f1)
        g_f2 = torch.empty_like(f2)
        g_logits = torch.empty_like(logits) if logits is not None else None
    g_grad = torch.empty_like(grad) if grad is not None else None
    call('ctgan_ct_gp_loss_bwd', ctypes.byref(desc), _p(gcost), _p(d_real), _p(d_real2), _p(f1), _p(f2), _p(grad),
         _p(logits), _p(labels), _p(per_sample), _p(g_real), _p(g_real2), _p(g_fake), _p(g_f1), _p(g_f2),
         _p(g_grad), _p(g_logits), _stream())
    return g_real, g_real2, g_fake, g_f1, g_f2, g_grad, g_logits


def mean_fwd(d, sign):
    _f32c(d, 'd')
    out = torch.empty(1, dtype=torch.float32, device=d.device)
    call('ctgan_mean_fwd', _p(d), _p(out), d.numel(), float(sign), _stream())
    return out


def mean_bwd(gcost, n, sign):
    g = torch.empty(n, dtype=torch.float32, device=gcost.device)
    call('ctgan_mean_bwd', _p(gcost), _p(g), n, float(sign), _stream())
    return g


def softmax_ce_fwd(logits, labels):
    _f32c(logits, 'logits')
    out = torch.empty(1, dtype=torch.float32, device=logits.device)
    call('ctgan_softmax_ce_fwd', _p(logits), _p(labels), _p(out), logits.shape[0], logits.shape[1], _stream())
    return out


def softmax_ce_bwd(logits, labels, gcost, scale_):
    g = torch.empty_like(logits)
    call('ctgan_softmax_ce_bwd', _p(logits), _p(labels), _p(gcost), float(scale_), _p(g), logits.shape[0],
         logits.shape[1], _stream())
    return g


# --------------------------------------------------------------------------- optimizer / rng
def adam_step(p, g, m, v, lr_t, beta1, beta2, eps, grad_scale=1.0, lr_t_dev=None):
    for t in (p, g, m, v):
        _f32c(t, 'adam buffer')
    call('ctgan_adam_step', _p(p), _p(g), _p(m), _p(v), p.numel(), float(lr_t), float(beta1), float(beta2),
         float(eps), float(grad_scale), _p(lr_t_dev), _stream())


def philox_uniform(shape, device, seed, offset, lo=0., hi=1., memory_format=None, dyn=None):
    out = (torch.empty(shape, dtype=torch.float32, device=device, memory_format=memory_format)
           if memory_format is not None else torch.empty(shape, dtype=torch.float32, device=device))
    call('ctgan_philox_uniform', _p(out), out.numel(), float(lo), float(hi), int(seed), int(offset), _p(dyn), _stream())
    return out


def philox_normal(shape, device, seed, offset, dyn=None):
    out = torch.empty(shape, dtype=torch.float32, device=device)
    call('ctgan_philox_normal', _p(out), out.numel(), int(seed), int(offset), _p(dyn), _stream())
    return out


def philox_labels(n, device, n_labels, seed, offset, dyn=None):
    out = torch.empty(n, dtype=torch.int32, device=device)
    call('ctgan_philox_labels', _p(out), n, int(n_labels), int(seed), int(offset), _p(dyn), _stream())
    return out


def counter_add(counter, delta):
    call('ctgan_counter_add', _p(counter), int(delta), _stream())


# --------------------------------------------------------------------------- data parallel over peer memory (csrc/peer.cu)
class _RawFloats:
    """float32 view of a raw device allocation for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {'shape': (int(n),), 'typestr': '<f4', 'data': (int(ptr), False), 'version': 2}


def peer_alloc_floats(n, device):
    """(tensor, ptr): n zeroed floats in cudaMalloc memory that can be exported over CUDA IPC."""
    out = ctypes.c_void_p()
    with torch.cuda.device(device):
        call('ctgan_peer_alloc', ctypes.byref(out), int(n) * 4)
        t = torch.as_tensor(_RawFloats(out.value, n), device=device)
    return t, out.value


def ipc_handle(ptr):
    buf = ctypes.create_string_buffer(64)
    call('ctgan_ipc_get_handle', ctypes.c_void_p(ptr), buf)
    return buf.raw


def ipc_open(handle):
    out = ctypes.c_void_p()
    call('ctgan_ipc_open_handle', ctypes.create_string_buffer(handle, 64), ctypes.byref(out))
    return out.value


def peer_reduce_adam(world, rank, g_ptrs, p_ptrs, flag_ptrs, m, v, n, lr_t, beta1, beta2, eps, grad_scale, lr_t_dev=None):
    arr = lambda ps: (ctypes.c_void_p * world)(*ps)
    call('ctgan_peer_reduce_adam', world, rank, arr(g_ptrs), arr(p_ptrs), arr(flag_ptrs), _p(m), _p(v), int(n), float(lr_t),
         float(beta1), float(beta2), float(eps), float(grad_scale), _p(lr_t_dev), _stream())
