"""Autograd shell over the CUDA kernels: twice-differentiable primitives.

The gradient penalty differentiates THROUGH `tf.gradients(D(x^), x^)`
(TG/CT_gan_cifar.py:144-154), so every op on the critic path must have a backward that
is itself built from differentiable ops.  The conv family is closed under
differentiation:
    F(x, w)  = fprop            dF/dx -> D(gy, w)      dF/dw -> G(x, gy)
    D(gy, w) = dgrad            dD/dgy -> F(c, w)      dD/dw -> G(c, gy)
    G(x, gy) = wgrad            dG/dx -> D(gy, c)      dG/dgy -> F(x, c)
and all critic non-linearities are multiplications by a constant 0/alpha/(1/keep) mask
(ReLU, LeakyReLU, dropout), whose backward is the same multiplication; pooling /
upsampling / layout changes are linear maps paired with their adjoints.  Backward
functions return None (not zeros) for paths that carry no gradient, so the second-order
pass costs exactly one fprop-shaped and one wgrad-shaped GEMM per layer.

Generator-only ops (batch norm, tanh, sigmoid) and the loss heads are once-differentiable.
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import kernels as K
from ._lib import LossDesc, F32, BF16

CL = torch.channels_last

_param_ptrs = set()

# Parity-test hook: when set (DeviceRandom.begin_recording), every activation site reports the
# 0/1 pattern it applied, in call order, so the oracle can be evaluated on the same linear region.
pattern_recorder = None


def register_param(t):
    _param_ptrs.add(t.data_ptr())


def _is_param(w):
    return w.data_ptr() in _param_ptrs


# When True, the FINAL (non-differentiable) backward pass adds filter / bias gradients straight into
# the parameter's `.grad` (a view of FlatAdam's flat bucket) inside the wgrad / bias-grad kernels and
# returns None to autograd: no temporary, no zero-fill, no AccumulateGrad add per contribution.
direct_param_grads = True


# Which parameter gradients the RUNNING backward pass is asked for.  `ctx.needs_input_grad` only says that a parameter
# requires grad, not that this particular backward wants its gradient: the gradient penalty's first backward
# (tf.gradients(D(x^), [x^]), TG/CT_gan_cifar.py:144) wants d/dx^ only, and the generator step differentiates through the
# critic without updating it (var_list=gen_params, TG/CT_gan_cifar.py:153).  Without these switches every critic layer ran a
# filter-gradient and a bias-gradient kernel in both cases and threw the result away (11 + 11 launches per ResNet critic
# step, 12 + 12 per generator step).
_skip_param_grads = False
_frozen_ptrs = set()


class no_param_grads:
    """with no_param_grads(): the backward passes run inside produce input gradients only."""

    def __enter__(self):
        global _skip_param_grads
        self.prev, _skip_param_grads = _skip_param_grads, True
        return self

    def __exit__(self, *exc):
        global _skip_param_grads
        _skip_param_grads = self.prev
        return False


class frozen:
    """with frozen(params): backward passes run inside skip the gradients of these parameters."""

    def __init__(self, params):
        self.ptrs = {p.data_ptr() for p in params}

    def __enter__(self):
        self.added = self.ptrs - _frozen_ptrs
        _frozen_ptrs.update(self.added)
        return self

    def __exit__(self, *exc):
        _frozen_ptrs.difference_update(self.added)
        return False


def _wants(p, ptr=None):
    """ptr: the parameter a derived operand stands for (derived_from), recorded by the forward pass."""
    return not _skip_param_grads and (ptr if ptr is not None else p.data_ptr()) not in _frozen_ptrs


def derived_from(t, param):
    """Mark t (a differentiable re-ordering of `param`, e.g. the critic head's rows in NHWC order) so that frozen(params)
    also skips the gradient of t in the layer that consumes it."""
    t._ctgan_src_ptr = param.data_ptr()
    return t


def _direct(p):
    return (direct_param_grads and not torch.is_grad_enabled() and p.is_leaf and p.grad is not None
            and p.grad.is_contiguous() and _is_param(p))


def _direct_wgrad(x, gy, g, w, col):
    """w.grad += wgrad(x, gy) in the final backward: tensor-core jobs are queued and run as ONE launch over all layers at the
    step's join (K.flush_wgrads); the other routes launch now on the side stream."""
    ent = _box_by_w4.get(w.data_ptr())
    if ent is not None:
        _box_dirty[w.data_ptr()] = ent           # its scratch gradient is folded into the 3x3 parameter's at the next join
    if K.wgrad_deferrable(x, gy, g):
        K.conv_wgrad(x, gy, g, tuple(w.shape), accumulate_into=w.grad, col=col, defer=True)
    else:
        K.on_side(lambda: K.conv_wgrad(x, gy, g, tuple(w.shape), accumulate_into=w.grad, col=col), x, gy, col)


def _dense_like(g, ref_dim4):
    """Make an incoming gradient dense in the layout the kernels use."""
    if g is None:
        return None
    if g.dim() == 4:
        if not g.is_contiguous(memory_format=CL):
            g = g.contiguous(memory_format=CL)
    elif not g.is_contiguous():
        g = g.contiguous()
    return g


# ------------------------------------------------------------------------- conv family
class S2DAct:
    """An activation [N, C, H, W] (H, W even) held in the SPACE-TO-DEPTH layout [N, 4C, H/2, W/2] of the stride-2 layer
    that consumes it: written that way by the producing conv's fused epilogue (ConvF act=..., out_s2d), read directly as
    the 3x3 tensor-core conv's input -- no layout kernel in between."""

    def __init__(self, t, C, H, W, plain_grad=False):
        self.t, self.C, self.H, self.W = t, C, H, W
        # plain_grad: the producer (ConvF out_s2d) expects the gradient of t with the PLAIN [N, C, H, W] layout inside t's
        # shape -- what the consuming conv's dgrad epilogue writes (OUT_D2S), see conv_mean_pool_s2d
        self.plain_grad = plain_grad

    @property
    def shape(self):
        return (self.t.shape[0], self.C, self.H, self.W)

    def dim(self):
        return 4

    @property
    def dtype(self):
        return self.t.dtype


def _plain_in_s2d_shape(x):
    """The plain NHWC tensor x [N, C, H, W] re-viewed (no copy) with the shape [N, 4C, H/2, W/2] of its space-to-depth image."""
    N, C, H, W = x.shape
    return x.permute(0, 2, 3, 1).reshape(N, H // 2, W // 2, 4 * C).permute(0, 3, 1, 2)


def _s2d_shape_as_plain(t):
    """Inverse of _plain_in_s2d_shape: t [N, 4C, Hs, Ws] holding plain-layout content -> [N, C, 2Hs, 2Ws]."""
    N, C4, Hs, Ws = t.shape
    return t.permute(0, 2, 3, 1).reshape(N, 2 * Hs, 2 * Ws, C4 // 4).permute(0, 3, 1, 2)


# ConvMeanPool(3x3) filters as derived 4x4 / stride-2 filters (csrc/conv_s2d.cu: box_filter_kernel)
_box_filters = {}        # 3x3 parameter ptr -> (w3, w4)
_box_by_w4 = {}          # w4 ptr -> (w3, w4)
_box_dirty = {}          # w4 ptr -> (w3, w4) whose scratch gradient holds contributions of the running backward pass


def box_filter(w3):
    """The 'SAME' 4x4 / stride-2 filter W4 = boxsum(w3) / 4 that makes mean_pool_2x2(conv3x3(x, w3)) ONE conv (2.25x fewer
    multiply-adds, TG/CT_gan_cifar_resnet.py:89-92): a leaf kept current by FilterPacker.refresh() after every optimizer
    step; the filter-gradient kernels accumulate into its scratch .grad, folded into w3.grad at the step's join."""
    ent = _box_filters.get(w3.data_ptr())
    if ent is not None and ent[0] is w3:
        return ent[1]
    w3d = w3.detach()
    w4 = torch.empty((4, 4) + tuple(w3.shape[2:]), dtype=torch.float32, device=w3.device).requires_grad_(True)
    w4.grad = K.zeros(tuple(w4.shape), torch.float32, w3.device)
    w4d = w4.detach()
    K.box_filter(w3d, w4d)
    register_param(w4)
    derived_from(w4, w3)
    K.register_derived_filter(w3, w4, lambda: K.box_filter(w3d, w4d))
    _box_filters[w3.data_ptr()] = _box_by_w4[w4.data_ptr()] = (w3, w4)
    return w4


def _fold_box_grads():
    for w3, w4 in _box_dirty.values():
        if w3.grad is not None:
            K.box_filter_grad(w4.grad, w3.grad)
    _box_dirty.clear()


K._join_hooks.append(_fold_box_grads)


class S2DMask(Function):
    """xs = space_to_depth(x) * [pattern > 0], pattern a ReLU output in the space-to-depth layout: the adjoint of a dgrad
    epilogue that applied that ReLU's backward and wrote the plain layout (the gradient penalty's double backward)."""

    @staticmethod
    def forward(ctx, x, pattern, geom):
        return K.space_to_depth_mask(x, pattern, geom)

    @staticmethod
    def backward(ctx, c):
        raise RuntimeError('ctgan_b200: S2DMask is differentiated once (third-order gradients are not needed)')


def _plain_geom(N, H, W, C):
    """Geometry record for the layout kernels (space_to_depth / depth_to_space read N, H, W, Cin)."""
    return K.ConvGeom(N, H, W, C, H, W, C, 1, 1, 1, 0, 0)


class D2SMul(Function):
    """x = depth_to_space(xs * m), m a constant multiplier stored in the space-to-depth layout: the backward of an
    activation + dropout whose result a fused conv epilogue wrote in that layout (one kernel)."""

    @staticmethod
    def forward(ctx, xs, m, geom):
        ctx.m, ctx.geom = m, geom
        return K.depth_to_space(xs, geom, mul=m)

    @staticmethod
    def backward(ctx, c):
        return S2DMul.apply(_dense_like(c, True), ctx.m, ctx.geom), None, None


class S2DMul(Function):
    """xs = space_to_depth(x) * m: the adjoint of D2SMul (the gradient penalty's double backward)."""

    @staticmethod
    def forward(ctx, x, m, geom):
        ctx.m, ctx.geom = m, geom
        return K.space_to_depth(x, geom, mul=m)

    @staticmethod
    def backward(ctx, c):
        return D2SMul.apply(_dense_like(c, True), ctx.m, ctx.geom), None, None


class ConvF(Function):
    @staticmethod
    def forward(ctx, x, w, b, g, out_dtype, col=None, residual=None, relu=False, in_relu=False, relu_bwd_fused=False,
                res_up2=False, act=None, x_s2d=False, out_s2d=False):
        # res_up2: residual has half the output resolution and is added nearest-neighbour upsampled
        ctx.res_up2 = res_up2
        # in_relu: x is the output of a ReLU whose backward this conv applies in its dgrad epilogue (dx zeroed where
        #          x <= 0); the producer is then built with relu_bwd_fused=True and skips its own mask multiply
        ctx.in_relu, ctx.relu_bwd_fused = in_relu, relu_bwd_fused
        ctx.g = g
        ctx.w_ptr = getattr(w, '_ctgan_src_ptr', None)
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        ctx.bias = b
        # x_s2d: x IS the space-to-depth image of the conv's input (S2DAct); its gradient is returned in that layout
        ctx.x_s2d = x_s2d
        ctx.col = x if x_s2d else (col if col is not None else K.thin_col(x, g, 'x'))   # im2col / s2d of x: built once, reused by wgrad
        ctx.mult = ctx.out_geom = None
        # out_s2d (no act): y = [relu](conv + b) written by the epilogue in the space-to-depth layout of the stride-2 conv that
        # consumes it (conv_mean_pool_s2d); that conv's dgrad applies the ReLU's backward and returns the PLAIN layout
        ctx.out_plain_grad = False
        if out_s2d and act is None:
            if not (relu and relu_bwd_fused) or residual is not None or x_s2d:
                raise RuntimeError('ctgan_b200: ConvF(out_s2d) is the relu-fused producer of a conv_mean_pool_s2d only')
            y = K.conv_fprop(x, w, b, g, out_dtype=out_dtype, w_is_param=_is_param(w), col=ctx.col, relu=True, out_s2d=True)
            ctx.relu_out = y.detach()
            ctx.out_plain_grad = True
            if pattern_recorder is not None:
                pattern_recorder(K.depth_to_space(y.detach(), _plain_geom(g.N, g.Ho, g.Wo, g.Cout)) > 0)
            return y
        if act is not None:
            # act = (slope, keep, seed, offset, dyn, out_s2d): bias + LeakyReLU + dropout in the conv epilogue; the
            # multiplier m it stores makes backward and double backward plain products (MulConst / D2SMul)
            slope, keep, seed, offset, dyn, out_s2d = act
            y, m = K.conv_fprop_actdrop(x, w, b, g, slope, keep, seed, offset, dyn, out_s2d=out_s2d, w_is_param=_is_param(w),
                                        col=ctx.col)
            ctx.mult = m
            ctx.out_geom = _plain_geom(g.N, g.Ho, g.Wo, g.Cout) if out_s2d else None
            ctx.relu_out = None
            if pattern_recorder is not None and slope != 1.0:
                md = m.detach().float()
                if out_s2d:
                    md = K.depth_to_space(m.detach(), ctx.out_geom).float()
                pattern_recorder(md > 0.5 * (1.0 + slope) / keep)      # m in {0, slope/keep, 1/keep}: kept and v > 0
            return y
        # residual: y = conv(x) + b + residual in the conv epilogue (the block's skip connection); its gradient is gy
        # relu: the nonlinearity that follows the conv, in the epilogue; its multiplier is [y > 0]
        y = K.conv_fprop(x, w, b, g, out_dtype=out_dtype, w_is_param=_is_param(w), col=ctx.col, residual=residual, relu=relu,
                         res_up2=res_up2)
        ctx.relu_out = y.detach() if relu else None      # detached: a constant for MulReluMask, never an autograd edge
        if relu and pattern_recorder is not None:
            pattern_recorder(y.detach() > 0)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gy = _dense_like(gy, True)
        if ctx.out_plain_grad:
            gy = _s2d_shape_as_plain(gy)
        if ctx.mult is not None:
            gy = D2SMul.apply(gy, ctx.mult, ctx.out_geom) if ctx.out_geom is not None else MulConst.apply(gy, ctx.mult)
        if ctx.relu_out is not None and not ctx.relu_bwd_fused:
            gy = MulReluMask.apply(gy, ctx.relu_out)
        n_in = len(ctx.needs_input_grad)                 # apply() is called with 5 .. 13 arguments
        gy_res = None
        if n_in > 6 and ctx.needs_input_grad[6] and ctx.needs_input_grad[0] and gy.requires_grad:
            gy, gy_res = Fork2.apply(gy)                 # (dgrad operand, residual gradient): explicit fork for the double backward
        gx = gw = gb = None
        dycol = K.thin_col(gy, ctx.g, 'dy')          # im2col of a 3-channel gy: shared by dgrad and wgrad
        if ctx.needs_input_grad[0]:
            gx = ConvD.apply(gy, w, ctx.g, x.dtype, dycol, x.detach() if ctx.in_relu else None, ctx.x_s2d)
        if ctx.needs_input_grad[1] and _wants(w, ctx.w_ptr):
            if _direct(w):
                col, g = (ctx.col if ctx.col is not None else dycol), ctx.g
                _direct_wgrad(x, gy, g, w, col)
            else:
                if ctx.x_s2d:
                    raise RuntimeError('ctgan_b200: a conv fed in space-to-depth layout supports in-place filter gradients only')
                gw = ConvG.apply(x, gy, ctx.g, tuple(w.shape))
        if ctx.has_bias and ctx.needs_input_grad[2] and _wants(ctx.bias):
            b = ctx.bias
            if _direct(b):
                K.on_side(lambda: K.bias_grad(gy.detach(), accumulate_into=b.grad), gy)
            else:
                gb = K.bias_grad(gy.detach())
        g_res = None
        if n_in > 6 and ctx.needs_input_grad[6]:
            gr = gy_res if gy_res is not None else gy
            g_res = Pool.apply(gr, 1.0) if ctx.res_up2 else gr       # adjoint of the 2x nearest upsample: 2x2 sums
        return (gx, gw, gb, None, None, None, g_res, None, None, None, None, None, None, None)[:n_in]


class ConvD(Function):
    """dx = conv^T(gy, w) for the forward geometry g (== Deconv2D forward).  out_s2d: dx is produced (and its cotangent
    arrives) in the space-to-depth layout the conv's input was handed over in."""

    @staticmethod
    def forward(ctx, gy, w, g, out_dtype, col=None, relu_mask=None, out_s2d=False):
        ctx.g = g
        ctx.out_s2d = out_s2d
        ctx.save_for_backward(gy, w)
        ctx.relu_mask = relu_mask                     # constant: dx = conv^T(gy, w) * [relu_mask > 0]
        # an s2d-fed conv behind a fused ReLU (conv_mean_pool_s2d): relu_mask has the space-to-depth layout, dx is masked
        # and written in the PLAIN layout by the dgrad epilogue, handed back inside the shape of the layer's input
        ctx.plain_grad = bool(out_s2d and relu_mask is not None)
        if ctx.plain_grad:
            dx = K.conv_dgrad(gy, w, g, out_dtype=out_dtype, w_is_param=_is_param(w), col=col, relu_mask=relu_mask)
            return _plain_in_s2d_shape(dx)
        return K.conv_dgrad(gy, w, g, out_dtype=out_dtype, w_is_param=_is_param(w), col=col, relu_mask=relu_mask,
                            out_s2d=out_s2d)

    @staticmethod
    def backward(ctx, c):
        gy, w = ctx.saved_tensors
        c = _dense_like(c, True)
        if ctx.plain_grad:
            g_ = ctx.g
            c = S2DMask.apply(_dense_like(_s2d_shape_as_plain(c), True), ctx.relu_mask, _plain_geom(g_.N, g_.H, g_.W, g_.Cin))
        elif ctx.relu_mask is not None:
            c = MulReluMask.apply(c, ctx.relu_mask)
        ggy = gw = None
        ccol = c if ctx.out_s2d else K.thin_col(c, ctx.g, 'x')   # im2col / s2d of c: shared by fprop and wgrad
        if ctx.needs_input_grad[0]:
            if ctx.out_s2d:
                ggy = ConvF.apply(c, w, None, ctx.g, gy.dtype, None, None, False, False, False, False, None, True)
            else:
                ggy = ConvF.apply(c, w, None, ctx.g, gy.dtype, ccol)
        if ctx.needs_input_grad[1] and _wants(w):
            if _direct(w):
                _direct_wgrad(c, gy, ctx.g, w, ccol)
            else:
                if ctx.out_s2d:
                    raise RuntimeError('ctgan_b200: a conv fed in space-to-depth layout supports in-place filter gradients only')
                gw = ConvG.apply(c, gy, ctx.g, tuple(w.shape))
        return (ggy, gw, None, None, None, None, None)[:len(ctx.needs_input_grad)]


class ConvG(Function):
    """dw = wgrad(x, gy) (float HWIO)."""

    @staticmethod
    def forward(ctx, x, gy, g, w_shape):
        ctx.g = g
        ctx.save_for_backward(x, gy)
        return K.conv_wgrad(x, gy, g, w_shape)

    @staticmethod
    def backward(ctx, c):
        x, gy = ctx.saved_tensors
        c = c.contiguous()
        gx = ggy = None
        if ctx.needs_input_grad[0]:
            gx = ConvD.apply(gy, c, ctx.g, x.dtype)
        if ctx.needs_input_grad[1]:
            ggy = ConvF.apply(x, c, None, ctx.g, gy.dtype)
        return gx, ggy, None, None


def ensure_nhwc(x):
    """Accept an NCHW-contiguous 4-D activation at the op boundary and re-lay it out as NHWC."""
    if x.dim() == 4 and not x.is_contiguous(memory_format=CL):
        N, C, H, W = x.shape
        return ToNHWC.apply(x.reshape(N, C * H * W), C, H, W, x.dtype)
    return x


def conv2d_s2d_out_ok(x, cout, k):
    """True when conv2d(x, ., ., k, 1, relu=True, relu_bwd_fused=True, out_s2d=True) has a route."""
    if isinstance(x, S2DAct) or x.dim() != 4:
        return False
    N, H, W, Cin = K.nhwc_dims(x)
    return K.conv_fprop_s2d_out_ok(x, K.same_geom(N, H, W, Cin, cout, k, 1))


def conv_mean_pool_s2d_ok(x, cin, cout):
    """True when mean_pool_2x2(conv3x3(x)) of a bf16 [N, cin, H, W] activation can run as ONE stride-2 4x4 conv on the
    space-to-depth route (all three of its kernels on the tensor cores, zero blocks skipped)."""
    if not (K.config.pool_conv_s2d and K.config.s2d_embed_wgrad and direct_param_grads) or x.dim() != 4 or not x.is_cuda:
        return False
    N, H, W, _ = K.nhwc_dims(x)
    if H % 2 or W % 2 or cin % 128 or cout % 128 or x.dtype != torch.bfloat16:
        return False
    if N * H * W // 512 < K.config.pool_conv_min_tiles:
        # a quarter of the tiles, each with 16/9 of the k-blocks: a layer that is already below one wave of 128-pixel tiles
        # is latency-bound and only gets slower (Discriminator.2.Conv2 at batch 64: 128 -> 32 tiles)
        return False
    g = K.same_geom(N, H, W, cin, cout, 4, 2)
    g3 = K.s2d_geom(g, x)
    return g3 is not None and K._wgrad_multi_ok(g3)


def conv_mean_pool_s2d(xs, w3, b, residual=None):
    """mean_pool_2x2(conv2d(x, w3, 'SAME') + b) [+ residual] for x = relu(.) handed over as an S2DAct by its producer
    (conv2d(..., relu=True, relu_bwd_fused=True, out_s2d=True)): ONE 'SAME' 4x4 / stride-2 conv with the box-summed filter
    (functional.box_filter), 16 taps at a quarter of the positions.  Its dgrad applies the producer's ReLU backward."""
    if not (isinstance(xs, S2DAct) and xs.plain_grad):
        raise RuntimeError('ctgan_b200: conv_mean_pool_s2d needs the S2DAct of a conv2d(out_s2d=True) producer')
    w4 = box_filter(w3)
    N, Cin, H, W = xs.shape
    g = K.same_geom(N, H, W, Cin, w4.shape[-1], 4, 2)
    if residual is not None:
        residual = ensure_nhwc(residual)
    return ConvF.apply(xs.t, w4, b, g, xs.dtype, None, residual, False, True, False, False, None, True)


def conv2d(x, w, b, k, stride, out_dtype=None, residual=None, relu=False, in_relu=False, relu_bwd_fused=False,
           res_up2=False, out_s2d=False):
    """tf.nn.conv2d(SAME) + bias_add on a logical-NCHW activation (+ an optional residual added in the epilogue).
    out_s2d: return the result as an S2DAct for conv_mean_pool_s2d (see conv2d_s2d_out_ok)."""
    if isinstance(x, S2DAct):
        N, Cin, H, W = x.shape
        g = K.same_geom(N, H, W, Cin, w.shape[-1], k, stride)
        if residual is not None or relu or in_relu:
            raise RuntimeError('ctgan_b200: a space-to-depth input supports the plain conv only')
        return ConvF.apply(x.t, w, b, g, out_dtype or x.dtype, None, None, False, False, False, False, None, True)
    N, H, W, Cin = K.nhwc_dims(x)
    g = K.same_geom(N, H, W, Cin, w.shape[-1], k, stride)
    if out_s2d:
        y = ConvF.apply(x, w, b, g, out_dtype or x.dtype, None, None, relu, in_relu, relu_bwd_fused, False, None, False, True)
        return S2DAct(y, w.shape[-1], g.Ho, g.Wo, plain_grad=True)
    if residual is not None:
        residual = ensure_nhwc(residual)
        want = (N, w.shape[-1], g.Ho // 2, g.Wo // 2) if res_up2 else (N, w.shape[-1], g.Ho, g.Wo)
        if residual.dtype != (out_dtype or x.dtype) or tuple(residual.shape) != want:
            raise RuntimeError('ctgan_b200: residual must have the shape and dtype of the conv output')
        return ConvF.apply(x, w, b, g, out_dtype or x.dtype, None, residual, relu, in_relu, relu_bwd_fused, res_up2)
    if relu or in_relu:
        return ConvF.apply(x, w, b, g, out_dtype or x.dtype, None, None, relu, in_relu, relu_bwd_fused)
    return ConvF.apply(x, w, b, g, out_dtype or x.dtype)


def conv2d_act_dropout(x, w, b, k, stride, slope, keep, rng, next_cout=None, next_k=None):
    """Conv2D -> LeakyReLU(slope) -> tf.nn.dropout(keep) (TG/CT_gan_cifar.py:84-96): in the conv epilogue when the conv
    takes a tensor-core route, else as conv + one activation kernel.  rng: the DeviceRandom whose next dropout site this
    is.  next_cout / next_k: the stride-2 conv that consumes the result -- when that conv takes the space-to-depth route
    the result is returned as an S2DAct (written by the epilogue in that layout)."""
    x_s2d = isinstance(x, S2DAct)
    xt = x.t if x_s2d else x
    if x_s2d:
        N, Cin, H, W = x.shape
    else:
        N, H, W, Cin = K.nhwc_dims(x)
    Cout = w.shape[-1]
    g = K.same_geom(N, H, W, Cin, Cout, k, stride)
    like = torch.empty((N, Cout, g.Ho, g.Wo), dtype=xt.dtype, device='meta').contiguous(memory_format=CL)
    args = rng.dropout_args(like) if keep < 1.0 else dict(seed=0, offset=0, dyn=None)
    if 'u' not in args and K.conv_actdrop_route(xt, g) is not None:
        out_s2d = False
        if next_cout is not None and g.Ho % 2 == 0 and g.Wo % 2 == 0:
            gn = K.same_geom(N, g.Ho, g.Wo, Cout, next_cout, next_k, 2)
            out_s2d = K.s2d_geom(gn, xt) is not None
        act = (slope, keep, args['seed'], args['offset'], args['dyn'], out_s2d)
        y = ConvF.apply(xt, w, b, g, xt.dtype, None, None, False, False, False, False, act, x_s2d)
        return S2DAct(y, Cout, g.Ho, g.Wo) if out_s2d else y
    y = conv2d(x, w, b, k, stride)
    return ActDropout.apply(y, slope, keep, args.get('u'), args.get('seed', 0), args.get('offset', 0), args.get('dyn'))


def conv2d_transpose2(x, w, b):
    """tf.nn.conv2d_transpose(SAME, stride 2), filter [k,k,out,in]: the dgrad of the stride-2
    SAME conv that maps [N,2H,2W,out] -> [N,H,W,in], then bias_add."""
    N, H, W, Cin = K.nhwc_dims(x)
    k, Cout = w.shape[0], w.shape[2]
    g = K.same_geom(N, 2 * H, 2 * W, Cout, Cin, k, 2)
    y = ConvD.apply(x, w, g, x.dtype)
    if b is not None:
        y = BiasAdd.apply(y, b)
    return y


def linear(x, w, b, out_dtype=None):
    """tf.matmul + bias_add on [batch, in] (TG/tflib/ops/linear.py:132-146)."""
    g = K.ConvGeom(x.shape[0], 1, 1, w.shape[0], 1, 1, w.shape[1], 1, 1, 1, 0, 0)
    return ConvF.apply(x, w, b, g, out_dtype or x.dtype)


class BiasAdd(Function):
    """y = x + b[c] (tf.nn.bias_add after conv2d_transpose, TG/tflib/ops/deconv2d.py:105-110)."""

    @staticmethod
    def forward(ctx, x, b):
        ctx.bias = b
        return K.bias_add(x, b)

    @staticmethod
    def backward(ctx, gy):
        gy = _dense_like(gy, True)
        gb = None
        if ctx.needs_input_grad[1] and _wants(ctx.bias):
            if _direct(ctx.bias):
                b = ctx.bias
                K.on_side(lambda: K.bias_grad(gy.detach(), accumulate_into=b.grad), gy)
            else:
                gb = K.bias_grad(gy.detach())
        return gy, gb


# ------------------------------------------------------------------------- masks / element-wise
class MulConst(Function):
    """y = x * m with m a constant (no gradient).  Its own backward."""

    @staticmethod
    def forward(ctx, x, m):
        ctx.save_for_backward(m)
        return K.mul(x, m)

    @staticmethod
    def backward(ctx, gy):
        (m,) = ctx.saved_tensors
        return MulConst.apply(_dense_like(gy, True), m), None


class ActDropout(Function):
    """y = act(x) then tf.nn.dropout: one fused kernel producing y and the multiplier m."""

    @staticmethod
    def forward(ctx, x, slope, keep, u, seed, offset, dyn=None):
        y, m = K.act_dropout(x, slope, keep, u=u, seed=seed, offset=offset, dyn=dyn)
        if pattern_recorder is not None and slope != 1.0:
            pattern_recorder(x.detach() > 0)
        ctx.save_for_backward(m)
        return y

    @staticmethod
    def backward(ctx, gy):
        (m,) = ctx.saved_tensors
        return MulConst.apply(_dense_like(gy, True), m), None, None, None, None, None, None


def relu(x):
    return ActDropout.apply(x, 0.0, 1.0, None, 0, 0)


def leaky_relu_dropout(x, slope, keep, u=None, seed=0, offset=0, dyn=None):
    return ActDropout.apply(x, slope, keep, u, seed, offset, dyn)


def dropout(x, keep, u=None, seed=0, offset=0, dyn=None):
    if keep == 1.0:
        return x
    return ActDropout.apply(x, 1.0, keep, u, seed, offset, dyn)


class MaskSum2(Function):
    """a*ma + b*mb with constant multipliers (ma None = identity): the backward of a fork."""

    @staticmethod
    def forward(ctx, a, b, ma, mb):
        ctx.ma, ctx.mb = ma, mb
        return K.mask_sum2(a, ma, b, mb)

    @staticmethod
    def backward(ctx, c):
        c = _dense_like(c, True)
        if ctx.ma is None:
            return c, MulConst.apply(c, ctx.mb), None, None
        ca, cb = MaskFork2.apply(c, ctx.ma, ctx.mb)
        return ca, cb, None, None


class MaskFork2(Function):
    """c -> (c*ma, c*mb) with constant multipliers: the backward of MaskSum2."""

    @staticmethod
    def forward(ctx, c, ma, mb):
        ctx.ma, ctx.mb = ma, mb
        return K.mask_fork2(c, ma, mb)

    @staticmethod
    def backward(ctx, ga, gb):
        return _mask_sum(ga, gb, ctx.ma, ctx.mb), None, None


def _mask_sum(ga, gb, ma, mb):
    """ga*ma + gb*mb where either incoming gradient may be absent."""
    if ga is None and gb is None:
        return None
    if gb is None:
        ga = _dense_like(ga, True)
        return MulConst.apply(ga, ma) if ma is not None else ga
    if ga is None:
        return MulConst.apply(_dense_like(gb, True), mb)
    return MaskSum2.apply(_dense_like(ga, True), _dense_like(gb, True), ma, mb)


class ForkDropoutRelu(Function):
    """(d, r) = (dropout(x), relu(dropout(x))) in one kernel: the input of a residual block whose skip connection
    takes d and whose first nonlinearity takes r.  backward: gx = gd*md + gr*mdr in one kernel."""

    @staticmethod
    def forward(ctx, x, keep, u, seed, offset, dyn=None):
        d, r, md, mdr = K.fork_dropout_relu(x, keep, u=u, seed=seed, offset=offset, dyn=dyn)
        if pattern_recorder is not None:
            pattern_recorder(x.detach() > 0)
        ctx.md, ctx.mdr = md, mdr
        return d, r

    @staticmethod
    def backward(ctx, gd, gr):
        return _mask_sum(gd, gr, ctx.md, ctx.mdr), None, None, None, None, None


class ForkRelu(Function):
    """(x, relu(x)) for a block input without dropout: backward gx = gd + gr*m in one kernel."""

    @staticmethod
    def forward(ctx, x):
        r, m = K.act_dropout(x, 0.0, 1.0)
        if pattern_recorder is not None:
            pattern_recorder(x.detach() > 0)
        ctx.m = m
        return x.view_as(x), r

    @staticmethod
    def backward(ctx, gd, gr):
        return _mask_sum(gd, gr, None, ctx.m)


def fork_dropout_relu(x, keep, u=None, seed=0, offset=0, dyn=None):
    """-> (dropout(x), relu(dropout(x)))."""
    if keep == 1.0:
        return ForkRelu.apply(x)
    return ForkDropoutRelu.apply(x, keep, u, seed, offset, dyn)


def fork_relu(x):
    """-> (x, relu(x))."""
    return ForkRelu.apply(x)


class PoolAddFork(Function):
    """(o1, o2) = (x*m1, x*m2), x = meanpool2x2(y) + s: the end of a down-sampling critic block (ConvMeanPool + skip add)
    fused with the next block's input fork; masks computed from the data (forward pass) or given (its re-use as the
    adjoint of MaskSumUp)."""

    @staticmethod
    def forward(ctx, y, s, keep, u, seed, offset, dyn, m1, m2):
        if m2 is None:
            o1, o2, m1, m2 = K.pool_add_fork(y, s, keep, u=u, seed=seed, offset=offset, dyn=dyn)
            if pattern_recorder is not None:
                # [x > 0] wherever it matters: a dropped element (m1 == 0) is zero on both sides whatever its pattern
                pattern_recorder((m2.detach() > 0) if m1 is None else ((m2.detach() > 0) | (m1.detach() == 0)))
        else:
            o1, o2, _, _ = K.pool_add_fork(y, s, masks=(m1, m2))
        ctx.m1, ctx.m2 = m1, m2
        return o1, o2

    @staticmethod
    def backward(ctx, g1, g2):
        gy, gs = _mask_sum_up(g1, g2, ctx.m1, ctx.m2)
        return gy, gs, None, None, None, None, None, None, None


class MaskSumUp(Function):
    """(gy, gx): gx = a*m1 + b*m2, gy = 0.25*gx replicated 2x2 -- the adjoint of PoolAddFork with fixed masks."""

    @staticmethod
    def forward(ctx, a, b, m1, m2):
        ctx.m1, ctx.m2 = m1, m2
        return K.mask_sum2_up(a, m1, b, m2)

    @staticmethod
    def backward(ctx, cy, cx):
        if cy is None and cx is None:
            return None, None, None, None
        if cy is None or cx is None:                  # one output unused: c = cx + meanpool(cy) from the unfused pieces
            c = _dense_like(cx, True) if cy is None else Pool.apply(_dense_like(cy, True), 0.25)
            ca = MulConst.apply(c, ctx.m1) if ctx.m1 is not None else c
            return ca, MulConst.apply(c, ctx.m2), None, None
        ca, cb = PoolAddFork.apply(_dense_like(cy, True), _dense_like(cx, True), 1.0, None, 0, 0, None, ctx.m1, ctx.m2)
        return ca, cb, None, None


def _mask_sum_up(g1, g2, m1, m2):
    """-> (gy, gs) for incoming gradients of the two fork outputs (either may be absent)."""
    if g1 is not None and g2 is not None:
        return MaskSumUp.apply(_dense_like(g1, True), _dense_like(g2, True), m1, m2)
    gx = _mask_sum(g1, g2, m1, m2)
    if gx is None:
        return None, None
    return Upsample.apply(gx, 0.25), gx


def pool_add_fork(y, s, keep=1.0, u=None, seed=0, offset=0, dyn=None):
    """-> (d, r): d = dropout(meanpool2x2(y) + s) (no dropout for keep == 1), r = relu(d)."""
    return PoolAddFork.apply(y, s, keep, u, seed, offset, dyn, None, None)


class MulReluMask(Function):
    """g * [y > 0], y = the output of a ReLU fused into a conv epilogue (a constant here)."""

    @staticmethod
    def forward(ctx, g, y):
        ctx.y = y
        return K.mul_relu_mask(g, y)

    @staticmethod
    def backward(ctx, c):
        return MulReluMask.apply(_dense_like(c, True), ctx.y), None


class Add(Function):
    @staticmethod
    def forward(ctx, a, b):
        return K.add(a, b)

    @staticmethod
    def backward(ctx, g):
        return g, g


def add(a, b):
    return Add.apply(a, b)


class Fork2(Function):
    """(x, x): a tensor consumed by two sub-graphs (a block's shortcut and its first layer).  The engine would sum the two
    gradients with an ATen add; here the adjoint is the library's own add kernel (and, being an Add node, differentiable
    again for the gradient penalty)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x), x.view_as(x)

    @staticmethod
    def backward(ctx, g1, g2):
        if g1 is None or g2 is None:
            return g1 if g2 is None else g2
        return Add.apply(_dense_like(g1, True), _dense_like(g2, True))


def fork2(x):
    return Fork2.apply(x) if x.requires_grad else (x, x)


class AddScaled(Function):
    """a + s * b on tensors of one shape (the generator cost: -mean D(G(z)) + 0.1 * CE, TG/CT_gan_cifar_resnet.py:326-330)."""

    @staticmethod
    def forward(ctx, a, b, s):
        ctx.s = s
        return K.add(a, K.scale(b, s))

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        g = g.contiguous()
        return g, K.scale(g, ctx.s), None


class Pool(Function):
    @staticmethod
    def forward(ctx, x, scale):
        ctx.scale = scale
        return K.pool2x2(x, scale)

    @staticmethod
    def backward(ctx, g):
        return Upsample.apply(_dense_like(g, True), ctx.scale), None


class Upsample(Function):
    @staticmethod
    def forward(ctx, x, scale):
        ctx.scale = scale
        return K.upsample2x(x, scale)

    @staticmethod
    def backward(ctx, g):
        return Pool.apply(_dense_like(g, True), ctx.scale), None


def mean_pool_2x2(x):
    return Pool.apply(x, 0.25)


def upsample_2x(x):
    return Upsample.apply(x, 1.0)


class SpatialSum(Function):
    @staticmethod
    def forward(ctx, x, scale):
        ctx.scale, ctx.hw = scale, (x.shape[2], x.shape[3])
        return K.spatial_sum(x, scale)

    @staticmethod
    def backward(ctx, g):
        return SpatialBcast.apply(_dense_like(g, False), ctx.hw[0], ctx.hw[1], ctx.scale), None


class SpatialBcast(Function):
    @staticmethod
    def forward(ctx, y, H, W, scale):
        ctx.scale = scale
        return K.spatial_bcast(y, H, W, scale)

    @staticmethod
    def backward(ctx, g):
        return SpatialSum.apply(_dense_like(g, True), ctx.scale), None, None, None


def spatial_mean(x):
    return SpatialSum.apply(x, 1.0 / (x.shape[2] * x.shape[3]))


class ToNHWC(Function):
    """[N, C*H*W] (NCHW element order, the reference's flat tensors) -> logical [N,C,H,W] NHWC."""

    @staticmethod
    def forward(ctx, x, C, H, W, out_dtype):
        ctx.in_dtype, ctx.in_shape = x.dtype, tuple(x.shape)
        return K.nchw_to_nhwc(x.contiguous(), x.shape[0], C, H, W, out_dtype)

    @staticmethod
    def backward(ctx, g):
        return ToNCHW.apply(_dense_like(g, True), ctx.in_dtype, ctx.in_shape), None, None, None, None


class ToNCHW(Function):
    """logical [N,C,H,W] NHWC -> contiguous `out_shape` in NCHW element order."""

    @staticmethod
    def forward(ctx, x, out_dtype, out_shape):
        ctx.in_dtype, ctx.chw = x.dtype, tuple(x.shape[1:])
        return K.nhwc_to_nchw(x, out_dtype, out_shape)

    @staticmethod
    def backward(ctx, g):
        C, H, W = ctx.chw
        return ToNHWC.apply(g.contiguous(), C, H, W, ctx.in_dtype), None, None


def to_nhwc(x_flat, C, H, W, dtype):
    return ToNHWC.apply(x_flat, C, H, W, dtype)


def flat_nhwc(x):
    """logical [N,C,H,W] NHWC activation -> [N, H*W*C] in its OWN element order: a view, no kernel (cf. to_flat_nchw)."""
    x = ensure_nhwc(x)
    N, C, H, W = x.shape
    return x.permute(0, 2, 3, 1).reshape(N, H * W * C)


def to_flat_nchw(x, dtype=None):
    N, C, H, W = x.shape
    return ToNCHW.apply(x, dtype or x.dtype, (N, C * H * W))


class Crop(Function):
    @staticmethod
    def forward(ctx, x, h, w):
        ctx.hw = (x.shape[2], x.shape[3])
        return K.crop(x, h, w)

    @staticmethod
    def backward(ctx, g):
        return CropBwd.apply(_dense_like(g, True), ctx.hw[0], ctx.hw[1]), None, None


class CropBwd(Function):
    @staticmethod
    def forward(ctx, g, H, W):
        ctx.hw = (g.shape[2], g.shape[3])
        return K.crop_bwd(g, H, W)

    @staticmethod
    def backward(ctx, c):
        return Crop.apply(_dense_like(c, True), ctx.hw[0], ctx.hw[1]), None, None


def crop(x, h, w):
    return Crop.apply(x, h, w)


class Cast(Function):
    @staticmethod
    def forward(ctx, x, dtype):
        ctx.in_dtype = x.dtype
        return K.cast(x, dtype)

    @staticmethod
    def backward(ctx, g):
        return Cast.apply(_dense_like(g, True), ctx.in_dtype), None


def cast(x, dtype):
    return x if x.dtype == dtype else Cast.apply(x, dtype)


# ------------------------------------------------------------------------- generator-only ops
class BatchNormReLU(Function):
    """Training-mode BN (biased var, eps) with optional per-label gamma/beta, fused ReLU and (up2) the nearest-neighbour
    2x upsampling of the UpsampleConv that follows it written directly by the normalisation kernel."""

    @staticmethod
    def forward(ctx, x, gamma, beta, labels, eps, relu, groups=1, up2=False):
        y, mean, invstd = K.bn_fwd(x, gamma, beta, labels, eps, relu, groups, up2)
        ctx.groups, ctx.up2 = groups, up2
        if pattern_recorder is not None and relu:
            yd = y.detach()
            pattern_recorder((yd[:, :, ::2, ::2] if up2 else yd) > 0)
        ctx.relu = relu
        ctx.labels = labels
        ctx.params = (gamma, beta)
        # the two-kernel path recomputes the ReLU pattern from x: y need not be kept alive for the backward
        keep_y = y if (relu and not K.bn_fused_ok(x, groups)) else None
        ctx.save_for_backward(x, keep_y, gamma, beta, mean, invstd)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x, y, gamma, beta, mean, invstd = ctx.saved_tensors
        gy = _dense_like(gy, True)
        pg, pb = ctx.params
        if ctx.needs_input_grad[1] and ctx.needs_input_grad[2] and _direct(pg) and _direct(pb):
            # table gradients are added into the flat gradient bucket by the reduction kernel itself
            dx, _, _ = K.bn_bwd(gy, x, y, gamma, beta, ctx.labels, mean, invstd, ctx.relu, ctx.groups, ctx.up2,
                                accumulate_into=(pg.grad, pb.grad))
            return dx, None, None, None, None, None, None, None
        dx, dgamma, dbeta = K.bn_bwd(gy, x, y, gamma, beta, ctx.labels, mean, invstd, ctx.relu, ctx.groups, ctx.up2)
        return dx, dgamma, dbeta, None, None, None, None, None


def batch_norm(x, gamma, beta, labels=None, eps=1e-5, relu=False, groups=1, up2=False):
    """groups > 1: statistics per block of N/groups consecutive samples (one block per reference device split).
    up2: returns the result nearest-neighbour upsampled 2x."""
    return BatchNormReLU.apply(x, gamma, beta, labels, eps, relu, groups, up2)


class LayerNorm(Function):
    """Layer normalisation over each sample's (C,H,W) with per-channel gamma / beta (TG/tflib/ops/layernorm.py:6-21) -- the
    critic's Normalize in TG/CT_gan_64x64.py:87-93.  Twice differentiable: layer norm is not
    piecewise linear, so the gradient penalty needs the second-order terms of LayerNormBwd below."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        y, mean, rstd = K.ln_fwd(x, gamma, beta, eps)
        ctx.save_for_backward(x, gamma, mean, rstd)
        ctx.beta = beta
        return y

    @staticmethod
    def backward(ctx, gy):
        x, gamma, mean, rstd = ctx.saved_tensors
        beta = ctx.beta
        gy = _dense_like(gy, True)
        dx = dgamma = dbeta = None
        if ctx.needs_input_grad[0]:
            # under create_graph (the penalty's first backward) this node is differentiated again
            dx = LayerNormBwd.apply(gy, x, gamma, mean, rstd)
        if ctx.needs_input_grad[1] and _wants(gamma):
            # parameter gradients are never differentiated further: plain kernels, detached
            if _direct(gamma) and _direct(beta):
                g, b = gamma.grad, beta.grad
                # mean / rstd are 4*N-byte blocks: once this node is released the allocator hands them to the next small
                # allocation of the main stream while the side stream may not have run yet -- keep them until the join
                K.on_side(lambda: K.ln_param_grad(gy.detach(), x.detach(), mean, rstd, g, b), gy, x, mean, rstd)
            else:
                dgamma, dbeta = torch.zeros_like(gamma), torch.zeros_like(gamma)
                K.ln_param_grad(gy.detach(), x.detach(), mean, rstd, dgamma, dbeta)
        return dx, dgamma, dbeta, None


class LayerNormBwd(Function):
    """dx = core(gamma * gy) with core(u) = rstd * (u - mean_s(u) - xh * mean_s(u * xh)); mean / rstd are functions of x.
    Its backward (cotangent c of dx):  ggy = gamma * core(c),  ggamma_c = sum gy * core(c),  gx = K.ln_bwd2_x(...)."""

    @staticmethod
    def forward(ctx, gy, x, gamma, mean, rstd):
        ctx.save_for_backward(gy, x, gamma, mean, rstd)
        return K.ln_core(gy, x, gamma, mean, rstd, 1, 0)

    @staticmethod
    @once_differentiable
    def backward(ctx, c):
        gy, x, gamma, mean, rstd = ctx.saved_tensors
        c = _dense_like(c, True)
        ggy = gx = ggamma = None
        if ctx.needs_input_grad[0]:
            ggy = K.ln_core(c, x, gamma, mean, rstd, 0, 1)
        if ctx.needs_input_grad[1]:
            gx = K.ln_bwd2_x(c, gy, x, gamma, mean, rstd)
        if ctx.needs_input_grad[2] and _wants(gamma):
            t = K.mul(gy, K.ln_core(c, x, gamma, mean, rstd, 0, 0))
            if _direct(gamma):
                g = gamma.grad
                K.on_side(lambda: K.bias_grad(t, accumulate_into=g), t)
            else:
                ggamma = K.bias_grad(t)
        return ggy, gx, ggamma, None, None


def layer_norm(x, gamma, beta, eps=1e-5):
    return LayerNorm.apply(x, gamma, beta, eps)


class Unary(Function):
    @staticmethod
    def forward(ctx, x, kind):
        y = K.unary_fwd(x, kind)
        ctx.kind = kind
        ctx.save_for_backward(y)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        return K.unary_bwd(y, _dense_like(gy, True), ctx.kind), None


def tanh(x):
    return Unary.apply(x, 0)


def sigmoid(x):
    return Unary.apply(x, 1)


# ------------------------------------------------------------------------- losses
class CTGPLoss(Function):
    """Fused critic loss (TG/CT_gan_cifar.py:123-151, TG/CT_gan_cifar_resnet.py:244-300).
    Returns a float[8] tensor {cost, wgan, ct, gp, acgan, ...}; only element 0 is differentiable."""

    @staticmethod
    def forward(ctx, d_real, d_real2, d_fake, f1, f2, grad, logits, labels, hp):
        B, F_ = f1.shape
        desc = LossDesc(B, d_fake.shape[0], F_, grad.shape[1], 0 if logits is None else logits.shape[1],
                        BF16 if f1.dtype == torch.bfloat16 else F32,
                        hp['lambda_gp'], hp['lambda2'], hp['factor_m'], hp.get('acgan_scale', 0.0))
        tensors = [t.contiguous() for t in (d_real, d_real2, d_fake, f1, f2, grad)]
        lg = logits.contiguous() if logits is not None else None
        out, per_sample = K.ct_gp_loss_fwd(desc, *tensors, lg, labels)
        ctx.desc, ctx.labels, ctx.has_logits = desc, labels, logits is not None
        ctx.save_for_backward(*tensors, per_sample, *([lg] if lg is not None else []))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        saved = ctx.saved_tensors
        d_real, d_real2, d_fake, f1, f2, grad, per_sample = saved[:7]
        lg = saved[7] if ctx.has_logits else None
        gcost = gout[0:1].contiguous()
        g = K.ct_gp_loss_bwd(ctx.desc, gcost, d_real, d_real2, f1, f2, grad, lg, ctx.labels, per_sample)
        g_real, g_real2, g_fake, g_f1, g_f2, g_grad, g_logits = g
        return g_real, g_real2, g_fake, g_f1, g_f2, g_grad, g_logits, None, None


class CTGPLossStacked(Function):
    """CTGPLoss on the outputs of ONE stacked critic call: d_all [R] / f_all [R, F] / logits_all [R, L] hold the reference's
    separate critic calls as row ranges; rows = dict(real=(a, b), real2=(a, b), fake=(a, b)).  The backward kernel
    writes straight into row ranges of the full-size gradients (no slice-backward zero-fills, copies and adds)."""

    @staticmethod
    def forward(ctx, d_all, f_all, grad, logits_all, labels, hp, rows):
        (r0, r1), (s0, s1), (k0, k1) = rows['real'], rows['real2'], rows['fake']
        d_all, f_all = d_all.contiguous(), f_all.contiguous()
        la = logits_all.contiguous() if logits_all is not None else None
        d_real, d_real2, d_fake = d_all[r0:r1], d_all[s0:s1], d_all[k0:k1]
        f1, f2 = f_all[r0:r1], f_all[s0:s1]
        lg = la[r0:r1] if la is not None else None
        B, F_ = f1.shape
        # grad None: the loss without its gradient-penalty term (GPLoss evaluates that one on its own stream branch)
        desc = LossDesc(B, d_fake.shape[0], F_, grad.shape[1] if grad is not None else 1, 0 if lg is None else lg.shape[1],
                        BF16 if f1.dtype == torch.bfloat16 else F32,
                        hp['lambda_gp'], hp['lambda2'], hp['factor_m'], hp.get('acgan_scale', 0.0))
        grad = grad.contiguous() if grad is not None else None
        out, per_sample = K.ct_gp_loss_fwd(desc, d_real, d_real2, d_fake, f1, f2, grad, lg, labels)
        ctx.desc, ctx.labels, ctx.rows, ctx.has_grad = desc, labels, rows, grad is not None
        ctx.save_for_backward(d_all, f_all, grad if grad is not None else per_sample, per_sample, *([la] if la is not None else []))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        saved = ctx.saved_tensors
        d_all, f_all, grad, per_sample = saved[:4]
        if not ctx.has_grad:
            grad = None
        la = saved[4] if len(saved) > 4 else None
        (r0, r1), (s0, s1), (k0, k1) = ctx.rows['real'], ctx.rows['real2'], ctx.rows['fake']
        covered = (r1 - r0) + (s1 - s0) + (k1 - k0) == d_all.shape[0]
        g_d = torch.empty_like(d_all) if covered else K.zeros_like(d_all)
        g_f = K.zeros_like(f_all)                                  # the fake rows carry no feature gradient
        g_l = K.zeros_like(la) if la is not None else None
        outs = (g_d[r0:r1], g_d[s0:s1], g_d[k0:k1], g_f[r0:r1], g_f[s0:s1], g_l[r0:r1] if g_l is not None else None)
        g = K.ct_gp_loss_bwd(ctx.desc, gout[0:1].contiguous(), d_all[r0:r1], d_all[s0:s1], f_all[r0:r1], f_all[s0:s1], grad,
                             la[r0:r1] if la is not None else None, ctx.labels, per_sample, outs=outs)
        return g_d, g_f, g[5], g_l, None, None, None


class GPLoss(Function):
    """lambda * mean((||grad_i||_2 - 1)^2): the gradient-penalty term on its own (TG/CT_gan_cifar_resnet.py:284-286), for the
    schedule that differentiates it on the penalty's stream branch while the stacked pass runs its own backward.
    Returns float[8] {lambda*gp, 0, 0, gp, 0, ...}; only element 0 is differentiable."""

    @staticmethod
    def forward(ctx, grad, hp):
        grad = grad.contiguous()
        desc = LossDesc(grad.shape[0], 1, 1, grad.shape[1], 0, F32, hp['lambda_gp'], hp['lambda2'], hp['factor_m'], 0.0)
        out, per_sample = K.ct_gp_loss_fwd(desc, None, None, None, None, None, grad, None, None)
        ctx.desc = desc
        ctx.save_for_backward(grad, per_sample)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        grad, per_sample = ctx.saved_tensors
        g = K.ct_gp_loss_bwd(ctx.desc, gout[0:1].contiguous(), None, None, None, None, grad, None, None, per_sample)
        return g[5], None


class MeanLoss(Function):
    @staticmethod
    def forward(ctx, d, sign):
        ctx.n, ctx.sign = d.numel(), sign
        return K.mean_fwd(d.contiguous(), sign)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        return K.mean_bwd(g.contiguous(), ctx.n, ctx.sign), None


class SoftmaxCE(Function):
    @staticmethod
    def forward(ctx, logits, labels):
        logits = logits.contiguous()
        ctx.labels = labels
        ctx.save_for_backward(logits)
        return K.softmax_ce_fwd(logits, labels)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (logits,) = ctx.saved_tensors
        return K.softmax_ce_bwd(logits, ctx.labels, g.contiguous(), 1.0), None
