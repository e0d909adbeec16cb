/* ctgan_sm100.h -- C ABI of libctgan_sm100.so: hand-written sm_100a kernels for the
 * CT-GAN critic/generator training step.
 *
 * The reference (biuyq/CT-GAN) has no FFI: its boundary is the python op surface
 * tflib.ops.* which hands every tensor op to TensorFlow 1.2.1.  Each entry point
 * below replaces the TF call(s) a reference file makes (cited as TG/<file>:<line>,
 * TG = CT-GANs/tensorflow_generative_model) and is what a `ctypes` binding in the
 * reference's tflib would bind.
 *
 * Conventions
 *  - plain pointers + sizes; no torch types.  All pointers are DEVICE pointers
 *    unless a parameter is documented as host.  The caller owns every buffer
 *    (including workspaces); the library allocates nothing persistent.
 *  - activations are NHWC ("channels last"), element type `dtype`:
 *      CTGAN_F32 = 0 (float), CTGAN_BF16 = 1 (__nv_bfloat16).
 *    parameters, gradients of parameters and optimizer state are always float.
 *  - conv filters are HWIO [kh][kw][Cin][Cout] float, exactly the reference's
 *    `<name>.Filters` layout (TG/tflib/ops/conv2d.py:70-88).
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), is
 *    CUDA-graph capturable (no sync, no allocation) and returns 0 on success,
 *    <0 for a bad/unsupported descriptor, >0 for a cudaError_t.  The message of
 *    the last failure on the calling thread is returned by ctgan_last_error().
 *  - sm_100a only.  There is no CPU fallback.
 */
#ifndef CTGAN_SM100_H
#define CTGAN_SM100_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTGAN_F32  0
#define CTGAN_BF16 1

#define CTGAN_ERR_BAD_DESC    (-1)
#define CTGAN_ERR_UNSUPPORTED (-2)
#define CTGAN_ERR_DRIVER      (-3)

/* ---- library ------------------------------------------------------------- */
int         ctgan_version(void);              /* 100*major + minor */
/* number of GPU kernels this library has launched (or captured into a CUDA graph) so far */
unsigned long long ctgan_kernel_launches(void);
const char* ctgan_last_error(void);
/* 1 when the tcgen05/TMA path is usable on the current device (cc 10.x), else 0 */
int         ctgan_tc_available(void);

/* ---- convolution family --------------------------------------------------
 * One descriptor describes the FORWARD correlation
 *   y[n,p,q,o] = sum_{r,s,c} x[n, p*stride + r - pad_t, q*stride + s - pad_l, c] * w[r,s,c,o]
 * with x [N,H,W,Cin], y [N,Ho,Wo,Cout].  TF 'SAME' padding is
 * pad_t = pad_l = max((Ho-1)*stride + k - H, 0) / 2  (asymmetric: the remainder
 * goes after), see oracle/tf_ops.py:same_pad.
 *   fprop  replaces tf.nn.conv2d            TG/tflib/ops/conv2d.py:106-112 (+bias_add :120)
 *   dgrad  replaces Conv2DBackpropInput and IS tf.nn.conv2d_transpose
 *                                           TG/tflib/ops/deconv2d.py:97-103
 *   wgrad  replaces Conv2DBackpropFilter    (tf.gradients, TG/CT_gan_cifar.py:153-154)
 * tf.matmul in TG/tflib/ops/linear.py:132-136 is the H=W=k=1 case.
 * The GP double-backward (TG/CT_gan_cifar.py:144-154) composes these three.
 */
typedef struct {
    int32_t N, H, W, Cin;          /* input  x  [N,H,W,Cin]   */
    int32_t Ho, Wo, Cout;          /* output y  [N,Ho,Wo,Cout] */
    int32_t kh, kw, stride;
    int32_t pad_t, pad_l;
    int32_t x_dtype;               /* dtype of x / dx */
    int32_t y_dtype;               /* dtype of y / dy */
} ctgan_conv_desc;

#define CTGAN_EPI_RELU 1           /* y = max(y, 0) after bias (+residual) */
#define CTGAN_EPI_RES_UP2 2        /* tensor-core path: residual is [N, H/2, W/2, Cout]; pixel (h, w) adds residual (h/2, w/2)
                                      (nearest-neighbour 2x upsample of a ResidualBlock('up') shortcut, never materialised) */

/* tensor-core path (lean / pair kernels), the space-to-depth route of stride-2 convs (ctgan_pack_filter_s2d):
 * OUT_S2D: y [N,H,W,Cout] (H, W even) is written in the space-to-depth layout [N,H/2,W/2,4*Cout] of the stride-2 conv that
 *          consumes it (residual / relu_mask keep the plain layout);
 * OUT_D2S: the conv's output is the space-to-depth image [N,H,W,4C] of a plain [N,2H,2W,C] tensor (the dgrad of a stride-2
 *          conv); y is written as THAT tensor; relu_mask (if any) has the space-to-depth layout;
 * S2D_SKIP(mode,k,pad_t,pad_l): the packed 3x3 filter embeds a stride-2 k x k filter with those pads; the (tap, phase) blocks
 *          that hold no filter element are skipped (25 of 36 live for k = 5, 16 of 36 for k = 4).  mode 0: the launch is
 *          the fprop over the space-to-depth image (phases = input channel blocks); 1: its dgrad (flipped pack, phases =
 *          output channel blocks).  Same result as without the flag. */
#define CTGAN_EPI_OUT_S2D 4
#define CTGAN_EPI_OUT_D2S 8
#define CTGAN_EPI_S2D_SKIP 256
#define CTGAN_EPI_S2D_SKIP_FLAGS(mode, k, pad_t, pad_l) (256 | ((mode) << 9) | ((k) << 10) | ((pad_t) << 14) | ((pad_l) << 16))

/* generic SIMT implicit-GEMM kernels: any shape, fp32 FMA, float accumulate */
int ctgan_conv_fprop(const ctgan_conv_desc* d, const void* x, const float* w_hwio,
                     const float* bias /*nullable*/, void* y, int flags, void* stream);
int ctgan_conv_dgrad(const ctgan_conv_desc* d, const void* dy, const float* w_hwio,
                     void* dx, void* stream);
/* dw (float, HWIO) is OVERWRITTEN when accumulate==0, else added to. */
int ctgan_conv_wgrad(const ctgan_conv_desc* d, const void* x, const void* dy,
                     float* dw, int accumulate, void* stream);

/* tcgen05/TMEM/TMA implicit-GEMM kernels (BF16 operands, fp32 accumulate in TMEM).
 * Eligibility: x/y BF16, stride 1, Ho==H, Wo==W, Cin % 64 == 0, Cout % 64 == 0.
 * `wp` is a packed BF16 filter [kh*kw][Cout][Cin] made by ctgan_pack_filter_bf16.
 * fprop_tc also serves dgrad of a stride-1 conv: pack with transpose_flip=1 and
 * swap Cin/Cout, pad = k-1-pad in the descriptor.
 * residual (nullable, BF16, same shape as y) is added before the optional ReLU. */
/* test hook: 0 disables the halo-reuse variants of fprop_tc; 1 (default) = wherever eligible, including tiles of several
 * small images ([h][n][w] boxes through a tensor map with N and H exchanged); 2 = one-image tiles only */
void ctgan_set_fprop_halo(int on);
/* 1 (default): kernels are launched with programmatic stream serialization (each begins with griddepcontrol.launch_dependents +
 * griddepcontrol.wait, so launch latency and set-up overlap the predecessor's tail; ordering semantics unchanged); 0: plain launches */
void ctgan_set_pdl(int on);
/* test hook: 4 = 256-pixel work items where eligible, else 3 (default); 3 = persistent grouped-stage kernel; 1 = one tile per CTA */
void ctgan_set_fprop_variant(int v);
/* test / A-B hook: 1 (default) = layers with fewer output tiles than half the SMs split K over a cluster of 2 / 4 CTAs and
 * reduce the partial accumulators through distributed shared memory (csrc/conv_splitk.cu); 0 = one CTA per tile */
void ctgan_set_splitk(int on);
/* test/benchmark hook: 2 (default) = 3x3 wgrad CTAs own one filter column and share the x halo box; 1 = per-tap boxes */
void ctgan_set_wgrad_variant(int v);
int ctgan_conv_fprop_tc(const ctgan_conv_desc* d, const void* x, const void* wp,
                        const float* bias /*nullable*/, const void* residual /*nullable*/,
                        void* y, int flags, void* stream);
/* same, and the output is zeroed wherever relu_mask (bf16, the shape of y; nullable) is <= 0: the dgrad into a tensor
 * that is the output of a ReLU (Conv2DBackpropInput followed by ReluGrad, one kernel) */
int ctgan_conv_fprop_tc_masked(const ctgan_conv_desc* d, const void* x, const void* wp, const float* bias /*nullable*/,
                               const void* residual /*nullable*/, const void* relu_mask /*nullable*/, void* y, int flags,
                               void* stream);
/* Conv2D -> (+bias) -> LeakyReLU(slope) -> tf.nn.dropout(keep) in the conv EPILOGUE (TG/CT_gan_cifar.py:84-96,
 * TG/CT_gan_mnist.py:92-104; bias_add at TG/tflib/ops/conv2d.py:114-120):
 *   v = conv(x, wp) + bias;  m = (v > 0 ? 1 : slope) * (keep < 1 ? floor(keep + u) / keep : 1);  y = v * m
 * u = Philox4x32-10(seed) uniform number offset + *dyn_offset + (NHWC element index of y): the stream ctgan_act_dropout_fwd
 * draws from, generated in registers.  mult (bf16, layout of y) receives m: the backward and the double backward are
 * products with m.  out_s2d != 0 (H, W even): y and mult are written in the space-to-depth layout [N, H/2, W/2, 4*Cout],
 * channel (dy*2+dx)*Cout + c <- pixel (2i+dy, 2j+dx), i.e. directly as the input of the next stride-2 layer's 3x3 conv
 * (ctgan_space_to_depth not needed).  Cout must be a multiple of 128; slope = 1 gives plain dropout, keep = 1 plain LeakyReLU. */
int ctgan_conv_fprop_tc_actdrop(const ctgan_conv_desc* d, const void* x, const void* wp, const float* bias /*nullable*/,
                                void* y, void* mult, float slope, float keep, uint64_t seed, uint64_t offset,
                                const uint64_t* dyn_offset /*nullable*/, int out_s2d, void* stream);
/* dw[r,s,c,o] (float HWIO) = sum_pixels x[.., c] * dy[.., o]; split over pixels with
 * fp32 atomics, so dw must hold the value to accumulate onto (zeros for a plain wgrad). */
int ctgan_conv_wgrad_tc(const ctgan_conv_desc* d, const void* x, const void* dy,
                        float* dw, void* stream);
/* The filter gradients of n layers in ONE persistent launch (csrc/conv_wgrad_multi.cu): job i adds wgrad(xs[i], dys[i]) of
 * geometry descs[i] into dws[i] (float HWIO, pre-initialised) exactly like n calls of ctgan_conv_wgrad_tc, but the work of
 * all jobs is cut into ~2 items per SM, so small layers neither under-fill the GPU nor pay a 49-way split of their pixel
 * range.  descs / xs / dys / dws are HOST arrays (read during the call).  Eligible jobs (ctgan_conv_wgrad_tc_multi_ok):
 * BF16, stride 1, Cin and Cout multiples of 64 (a count that is 64 mod 128 leaves the upper half of one 128-channel operand
 * zero-filled by TMA; those accumulator rows / columns are not stored), 1x1, or 3x3 on images of >= 64 pixels with a power-of-two
 * width >= 8, or 3x3 on 4x4 images (several images per pixel chunk).
 * Replaces Conv2DBackpropFilter of every Conv2D of a backward pass (autodiff of TG/tflib/ops/conv2d.py:106). */
int ctgan_conv_wgrad_tc_multi_ok(const ctgan_conv_desc* d);
/* ..._embed: the same launch with an optional space-to-depth embedding per job (embed = n x 4 ints {k, C, pad_t, pad_l}, k = 0 or
 * embed == NULL: plain job).  An embedded job is the 3x3 pad-1 correlation over the 4C-channel space-to-depth image that stands for a
 * stride-2 k x k (k <= 5) convolution with C input channels (C % 128 == 0) and TF-SAME pads (pad_t, pad_l): descs[i] describes the
 * 3x3 geometry (Cin = 4C), dws[i] is the k x k filter gradient [k][k][C][Cout] itself -- only the row/column taps that carry a filter
 * element are multiplied, and they are reduced straight into it (no 3x3x4C scratch, zero-fill or gather: ctgan_s2d_filter_grad). */
/* tuning / test hook: pixels per pipeline stage of the multi-job kernel: 64, 128, or 0 (default) = 128 when a job's images are
 * at least 64 pixels wide (their halo rows are then shared by twice as many image rows), else 64 */
void ctgan_set_wgrad_multi_chunk(int px);
/* tuning / A-B hook: on = 1 (default) assigns the launch's work items longest-first to the least-loaded CTA, 0 = round robin,
 * 3 = as 1 with each CTA's list rotated, 5 = as 1 even when the predicted gain over round robin is below 10 %;
 * overhead >= 0 sets the fixed per-item cost (in 64-pixel chunks) that assignment assumes */
void ctgan_set_wgrad_multi_balance(int on, int overhead);
/* diagnostic: predicted makespan gain (%) of that assignment over round robin for the most recent launch */
int ctgan_wgrad_multi_last_gain_pct(void);
/* the item -> CTA assignment alone (host code, callable without a GPU; tests): cost[items] > 0 in arbitrary units; order must
 * hold 1024 entries and receives order[b + grid * k] = the k-th item CTA b runs (-1 = none).  Returns the number of slots used
 * (a multiple of grid) or 0 = keep round robin (items <= grid, more than 1024 slots, or a predicted makespan gain below
 * min_gain_pct %); *gain_pct (nullable) receives the predicted gain. */
int ctgan_wgrad_multi_assign(const int* cost, int items, int grid, int min_gain_pct, int rotate, short* order, int* gain_pct);
/* A/B hook: 0 = tensor-core forward launches without a residual use the generic kernels (residual tested at run time) */
void ctgan_set_fprop_nores(int on);
int ctgan_conv_wgrad_tc_multi_embed(int n, const ctgan_conv_desc* descs, const void* const* xs, const void* const* dys,
                                    float* const* dws, const int* embed, void* stream);
int ctgan_conv_wgrad_tc_multi(int n, const ctgan_conv_desc* descs, const void* const* xs, const void* const* dys,
                              float* const* dws, void* stream);
/* tuning hook: work items per SM that launch aims for (default 2) */
void ctgan_set_wgrad_multi_items_per_sm(int v);
/* ---- the fp32-storage path on the tensor cores: tcgen05.mma kind::tf32 (csrc/conv_tf32.cu) -------------------------
 * x, y, residual, relu_mask: FLOAT NHWC tensors; wp: float operand pack from ctgan_pack_filter_f32 / _multi_f32
 * ([taps][Cout][Cin] for fprop, tap-flipped [taps][Cin][Cout] for dgrad).  Products are formed with TF32 operand precision
 * (10-bit mantissa), accumulated in fp32.  Stride 1, Cin a multiple of 32, Cout of 128 (ctgan_conv_tf32_ok).  Epilogue as
 * ctgan_conv_fprop_tc_masked.  The wgrad entry is the float twin of ctgan_conv_wgrad_tc_multi (Cin, Cout multiples of 128).
 * Replaces tf.nn.conv2d / Conv2DBackpropInput / Conv2DBackpropFilter / tf.matmul on float tensors
 * (TG/tflib/ops/conv2d.py:106-112, linear.py:132-136) -- north_star: "TF32/BF16 inputs and fp32 accumulation". */
int ctgan_conv_tf32_ok(const ctgan_conv_desc* d);
int ctgan_conv_fprop_tf32(const ctgan_conv_desc* d, const void* x, const void* wp, const float* bias /*nullable*/,
                          const void* residual /*nullable*/, const void* relu_mask /*nullable*/, void* y, int flags, void* stream);
int ctgan_conv_wgrad_tf32_multi_ok(const ctgan_conv_desc* d);
int ctgan_conv_wgrad_tf32_multi(int n, const ctgan_conv_desc* descs, const void* const* xs, const void* const* dys,
                                float* const* dws, void* stream);
int ctgan_pack_filter_f32(const float* w_hwio, float* wp, int taps, int Cin, int Cout, int transpose_flip, void* stream);
/* table: as ctgan_pack_filters_multi, dst offsets in floats */
int ctgan_pack_filters_multi_f32(const float* flat, float* packs, const void* table, int n_entries, void* stream);
/* w_hwio float [taps][Cin][Cout] ->
 *   transpose_flip==0: wp[t][o][c] = w[t][c][o]                (fprop operand)
 *   transpose_flip==1: wp[t][c][o] = w[taps-1-t][c][o]         (dgrad operand)  */
int ctgan_pack_filter_bf16(const float* w_hwio, void* wp, int taps, int Cin, int Cout,
                           int transpose_flip, void* stream);

/* Packs many filters of one flat float parameter buffer in ONE launch.  table: device array of n_entries records
 * {int64 src_offset (floats), int64 dst_offset (bf16 elements), int32 taps, Cin, Cout, transpose_flip}. */
/* ---- thin-channel convolutions on the tensor cores (the 3-channel image side of the ResNet: Conv2D
 * 'Discriminator.1.Conv1' / '.Shortcut' 3 -> DIM_D and 'Generator.Output' DIM_G -> 3, CT_gan_cifar_resnet.py:125-136, :148-150;
 * op = tflib/ops/conv2d.py:106-112).  The thin side (C <= 8 channels, taps*C <= 64) is expanded to
 * col[pixel][64] bf16, k = tap*C + c, so that fprop / dgrad / wgrad become 1x1 tensor-core GEMMs:
 *   im2col: col[p][(t,c)] = src[p + sign*off(t)][c],  off(t) = (r - pad_t, s - pad_l), zero outside the image
 *   col2im: dst[p][c] = bias[c] + sum_t col[p + sign*off(t)][(t,c)]
 * d gives N,H,W,kh,kw,pads (stride 1).  pack kinds: see conv_tc.cu (0: C->Cw fprop, 1: Cw->C dgrad, 2: Cw->C fprop,
 * 3: C->Cw dgrad); every pack is 64*Cw bf16.  wgrad_thin_tc ADDS wide^T x col into dw: mode 0 = dw [taps][C][Cw]
 * (wide = dy, col = im2col(x, +1)), mode 1 = dw [taps][Cw][C] (wide = x, col = im2col(dy, -1)). */
int ctgan_im2col_thin(const ctgan_conv_desc* d, int C, int sign, const void* src, void* col, void* stream);
int ctgan_col2im_thin(const ctgan_conv_desc* d, int C, int sign, const void* col, const float* bias, void* dst, void* stream);
int ctgan_pack_filter_thin(const float* w, void* wp_bf16, int taps, int C, int Cw, int kind, void* stream);
int ctgan_wgrad_thin_tc(const void* wide, const void* col, long long P, int Cw, int C, int taps, int mode, float* dw, void* stream);
int ctgan_pack_filters_multi(const float* flat_params, void* packs_bf16, const void* table, int n_entries, void* stream);

/* ---- stride-2 5x5 convolutions on the tensor cores (the DCGAN critics' tf.nn.conv2d(stride 2),
 * TG/tflib/ops/conv2d.py:106-112 called from TG/CT_gan_cifar.py:84-94 and TG/CT_gan_mnist.py:92-102, and -- as their
 * dgrad -- the generators' tf.nn.conv2d_transpose, TG/tflib/ops/deconv2d.py:97-103).  A stride-2 'SAME' correlation
 * over x [N,H,W,C] is a stride-1 3x3 pad-1 correlation over the space-to-depth image
 *   xs [N,Hs,Ws,4C],  xs[n,i,j,(dy*2+dx)*C+c] = x[n,2i+dy,2j+dx,c]  (zero outside x),  Hs = ceil(H/2), Ws = ceil(W/2)
 * with the embedded filter W3[R,S,(dy*2+dx)*C+c,o] = w[2(R-1)+dy+pad_t, 2(S-1)+dx+pad_l, c, o] (zero outside k x k),
 * valid when pad <= 2 and k-1-pad <= 3 (k = 5 with TF-SAME pads 1 or 2).  fprop = ctgan_conv_fprop_tc(xs, wp_f),
 * dgrad = depth_to_space(ctgan_conv_fprop_tc(dy, wp_d)), wgrad = s2d_filter_grad(ctgan_conv_wgrad_tc(xs, dy)).
 *   space_to_depth / depth_to_space: H, W, C are those of the full-resolution tensor x; depth_to_space crops to H x W.
 *   pack_filter_s2d: w float HWIO [k][k][Cin][Cout] -> wp_f [9][Cout][4Cin] and the tap-flipped wp_d [9][4Cin][Cout]
 *                    (bf16; either may be null).
 *   s2d_filter_grad: dw [k][k][Cin][Cout] (+)= the tap entries of dw3 [3][3][4Cin][Cout] (both float HWIO). */
int ctgan_space_to_depth(const void* x, void* xs, int N, int H, int W, int C, int dtype, void* stream);
int ctgan_depth_to_space(const void* xs, void* x, int N, int H, int W, int C, int dtype, void* stream);
/* the same moves with an element-wise multiplier ms stored in the SPACE-TO-DEPTH layout (the multiplier written by
 * ctgan_conv_fprop_tc_actdrop with out_s2d): x = depth_to_space(xs * ms) is the backward of that fused activation, and
 * xs = space_to_depth(x) * ms its adjoint (the gradient penalty's double backward) */
int ctgan_space_to_depth_mul(const void* x, const void* ms, void* xs, int N, int H, int W, int C, int dtype, void* stream);
int ctgan_depth_to_space_mul(const void* xs, const void* ms, void* x, int N, int H, int W, int C, int dtype, void* stream);
/* xs = space_to_depth(x) where `pattern` (space-to-depth layout; a ReLU output) is positive, else 0: the adjoint of a dgrad
 * epilogue that applied that ReLU's backward and wrote the plain layout (OUT_D2S + relu_mask), for the double backward */
int ctgan_space_to_depth_mask(const void* x, const void* pattern, void* xs, int N, int H, int W, int C, int dtype, void* stream);
/* ConvMeanPool(3x3) (TG/CT_gan_cifar_resnet.py:89-92) == one 'SAME' 4x4 / stride-2 conv with the box-summed filter
 * w4[u][v] = 1/4 sum_{a,b in {0,1}} w3[u-a][v-b] (HWIO, float): 2.25x fewer multiply-adds.  box_filter_grad is the adjoint:
 * dw3 += fold(dw4), and clears dw4 (the scratch gradient of the derived filter) for the next step. */
int ctgan_box_filter(const float* w3, float* w4, int Cin, int Cout, void* stream);
int ctgan_box_filter_grad(float* dw4, float* dw3, int Cin, int Cout, void* stream);
int ctgan_pack_filter_s2d(const float* w, void* wp_f, void* wp_d, int k, int Cin, int Cout, int pad_t, int pad_l, void* stream);
int ctgan_s2d_filter_grad(const float* dw3, float* dw, int k, int Cin, int Cout, int pad_t, int pad_l, int accumulate, void* stream);

/* ---- strided convolutions with a thin (C <= 8 channel, taps*C <= 128) input on the tensor cores: 'Discriminator.1'
 * (3 -> DIM 5x5/2, TG/CT_gan_cifar.py:84; 1 -> DIM, TG/CT_gan_mnist.py:92) and, as the dgrad of that geometry, the
 * generators' last Deconv2D (TG/CT_gan_cifar.py:75, TG/CT_gan_mnist.py:83; op = TG/tflib/ops/deconv2d.py:97-103).
 * One 128-column im2col row per OUTPUT pixel p = (n, ho, wo), column k = (r*kw + s)*C + c (HWIO order):
 *   im2col_strided: col[p][k] = x[n, stride*ho + r - pad_t, stride*wo + s - pad_l, c]     (zero outside x / k >= taps*C)
 *   col2im_strided: dx[n,h,w,c] = bias[c] + sum over taps with h + pad_t - r = stride*ho (same for w) of col[(n,ho,wo)][k]
 * so fprop = ctgan_conv_fprop_tc(1x1, col, wp_f), dgrad = col2im_strided(ctgan_conv_fprop_tc(1x1, dy, wp_d)) and
 * wgrad = the first taps*C rows of ctgan_conv_wgrad_tc(1x1, col, dy), added with ctgan_add_prefix.
 *   pack_filter_padk: w float [Kreal][Cout] (HWIO flattened) -> wp_f [Cout][128], wp_d [128][Cout] bf16, zero for k >= Kreal.
 *   add_prefix: dst[i] (+)= src[i], i < n  (atomic when accumulating: several streams add into one gradient bucket). */
int ctgan_im2col_strided(const ctgan_conv_desc* d, int C, const void* x, void* col, void* stream);
int ctgan_col2im_strided(const ctgan_conv_desc* d, int C, const void* col, const float* bias /*nullable*/, void* dx, void* stream);
int ctgan_pack_filter_padk(const float* w, void* wp_f, void* wp_d, int Kreal, int Cout, void* stream);
int ctgan_add_prefix(const float* src, float* dst, int64_t n, int accumulate, void* stream);

/* db[c] (float) = sum over rows of dy[rows][C]   (gradient of tf.nn.bias_add) */
int ctgan_bias_grad(const void* dy, float* db, int64_t rows, int C, int dtype,
                    int accumulate, void* stream);

/* y[r,c] = x[r,c] + b[c]  (tf.nn.bias_add after conv2d_transpose, TG/tflib/ops/deconv2d.py:105-110) */
int ctgan_bias_add(const void* x, const float* b, void* y, int64_t rows, int C, int dtype, void* stream);

/* ---- element-wise / data movement --------------------------------------- */
int ctgan_cast(const void* x, int x_dtype, void* y, int y_dtype, int64_t n, void* stream);
int ctgan_add(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream);
int ctgan_mul(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream);
int ctgan_scale(const void* a, float s, void* out, int64_t n, int dtype, void* stream);

/* Fused activation + dropout multiplier (TG/CT_gan_cifar.py:47-48,86 ; tf.nn.dropout):
 *   m = (x > 0 ? 1 : slope) * (keep < 1 ? floor(keep + u) / keep : 1);   y = x * m
 * slope = 0 -> ReLU, 0.2 -> LeakyReLU, 1 -> dropout only.  u is read from `u`
 * (float, nullable) or, when u == NULL and keep < 1, generated as
 * philox_uniform(seed, offset + i) for element i -- identical to what
 * ctgan_philox_uniform(seed, offset) materialises.  m (nullable) gets the multiplier. */
int ctgan_act_dropout_fwd(const void* x, const float* u, void* y, void* m, int64_t n, int dtype,
                          float slope, float keep, uint64_t seed, uint64_t offset,
                          const uint64_t* dyn_offset /*nullable*/, void* stream);
/* Fused forms of the same masks where a block input forks into the skip connection and the block's first ReLU
 * (Discriminator: tf.nn.dropout after ResidualBlock 2/3 followed by ResidualBlock's shortcut + nonlinearity,
 * TG/CT_gan_cifar_resnet.py:113-139,183-190):
 *   fork_dropout_relu: d = x*md, r = x*mdr, md = floor(keep+u)/keep, mdr = md*[x>0]  (same Philox slice as act_dropout_fwd)
 *   mask_sum2:         out = a*ma + b*mb        (ma == NULL: out = a + b*mb)           backward of a fork
 *   mask_fork2:        o1 = c*ma, o2 = c*mb                                            backward of mask_sum2
 *   mul_relu_mask:     out = g*[y>0]            backward of a ReLU fused into a conv epilogue (y = its output) */
int ctgan_fork_dropout_relu(const void* x, const float* u, void* d, void* r, void* md, void* mdr, int64_t n, int dtype,
                            float keep, uint64_t seed, uint64_t offset, const uint64_t* dyn_offset, void* stream);
int ctgan_mask_sum2(const void* a, const void* ma, const void* b, const void* mb, void* out, int64_t n, int dtype, void* stream);
int ctgan_mask_fork2(const void* c, const void* ma, const void* mb, void* o1, void* o2, int64_t n, int dtype, void* stream);
int ctgan_mul_relu_mask(const void* g, const void* y, void* out, int64_t n, int dtype, void* stream);
/* End of a down-sampling critic block fused with the next block's input fork (ConvMeanPool + shortcut add,
 * TG/CT_gan_cifar_resnet.py:89-92,139, then :183-190): x = meanpool2x2(y) + s  (y [N,H,W,C], s [N,H/2,W/2,C]),
 *   compute_masks = 1: m1 = floor(keep+u)/keep (written only if keep < 1), m2 = m1*[x>0]; o1 = x*m1, o2 = x*m2
 *   compute_masks = 0: the same linear map with the given masks (m1 nullable = identity)
 * mask_sum2_up is its adjoint: gx = a*m1 + b*m2 (low resolution), gy = 0.25*gx replicated 2x2. */
int ctgan_pool_add_fork(int compute_masks, const void* y, const void* s, const float* u, void* m1, void* m2, void* o1, void* o2,
                        int N, int H, int W, int C, int dtype, float keep, uint64_t seed, uint64_t offset,
                        const uint64_t* dyn_offset, void* stream);
int ctgan_mask_sum2_up(const void* a, const void* m1, const void* b, const void* m2, void* gx, void* gy, int N, int Ho, int Wo, int C,
                       int dtype, void* stream);

/* unary: kind 0 = tanh, 1 = sigmoid (TG/CT_gan_cifar.py:77, TG/CT_gan_mnist.py:85) */
int ctgan_unary_fwd(const void* x, void* y, int64_t n, int dtype, int kind, void* stream);
int ctgan_unary_bwd(const void* y, const void* dy, void* dx, int64_t n, int dtype, int kind, void* stream);

/* 2x2 mean pool / nearest 2x upsample on NHWC (TG/CT_gan_cifar_resnet.py:91,96,102-105).
 * pool:     y[N,H/2,W/2,C] = scale * (sum of the 2x2 block)      (scale .25 = mean pool)
 * upsample: y[N,2H,2W,C]   = scale * x[n,h/2,w/2,c]
 * Each is the other's adjoint, so they also serve as backward kernels. */
int ctgan_pool2x2(const void* x, void* y, int N, int H, int W, int C, float scale, int dtype, void* stream);
int ctgan_upsample2x(const void* x, void* y, int N, int H, int W, int C, float scale, int dtype, void* stream);
/* y[N,C] = scale * sum_hw x[N,HW,C]  and its adjoint x[N,HW,C] = scale * y[N,C]
 * (tf.reduce_mean(output, axis=[2,3]), TG/CT_gan_cifar_resnet.py:179) */
int ctgan_spatial_sum(const void* x, void* y, int N, int HW, int C, float scale, int dtype, void* stream);
int ctgan_spatial_bcast(const void* y, void* x, int N, int HW, int C, float scale, int dtype, void* stream);

/* NCHW float <-> NHWC dtype (the reference's NCHW<->NHWC transposes, TG/tflib/ops/deconv2d.py:89,112) */
int ctgan_nchw_to_nhwc(const void* x, int x_dtype, void* y, int y_dtype, int N, int C, int H, int W, void* stream);
int ctgan_nhwc_to_nchw(const void* x, int x_dtype, void* y, int y_dtype, int N, int C, int H, int W, void* stream);
/* crop the top-left h x w window of NHWC x[N,H,W,C] (TG/CT_gan_mnist.py:77) and its adjoint (zero pad) */
int ctgan_crop(const void* x, void* y, int N, int H, int W, int C, int h, int w, int dtype, void* stream);
int ctgan_crop_bwd(const void* dy, void* dx, int N, int H, int W, int C, int h, int w, int dtype, void* stream);

/* real-data preparation (TG/CT_gan_cifar.py:102-103, TG/CT_gan_cifar_resnet.py:201-202):
 *   y = 2 * (x_int / denom - 0.5) + noise,  noise = noise_hi * philox_uniform(seed, offset+i)
 * when noise_hi > 0.  y is float [n]. */
int ctgan_prep_real(const int32_t* x_int, float* y, int64_t n, float denom, float noise_hi,
                    uint64_t seed, uint64_t offset, const uint64_t* dyn_offset /*nullable*/, void* stream);
/* the same on the uint8 pixels the reference's loaders yield (TG/tflib/cifar10.py:17-19 `images` is uint8; the
 * feed converts to the int32 placeholder TG/CT_gan_cifar_resnet.py:191): 1 byte/pixel over PCIe instead of 4 */
int ctgan_prep_real_u8(const uint8_t* x_u8, float* y, int64_t n, float denom, float noise_hi,
                       uint64_t seed, uint64_t offset, const uint64_t* dyn_offset /*nullable*/, void* stream);
/* the same with a second destination y2 (nullable) receiving identical values -- the real batch occupies two row ranges of the
 * stacked critic input -- and the input type as a flag (0 = int32, 1 = uint8) */
int ctgan_prep_real_dup(const void* x, int x_is_u8, float* y, float* y2, int64_t n, float denom, float noise_hi,
                        uint64_t seed, uint64_t offset, const uint64_t* dyn_offset, void* stream);
/* cudaMemsetAsync(p, 0, bytes): the zero-fill of the flat gradient bucket as a memset node instead of a fill kernel */
int ctgan_memset_zero(void* p, int64_t bytes, void* stream);
/* out[b,p] = real[b,p] + alpha[b] * (fake[b,p] - real[b,p])   (TG/CT_gan_cifar.py:142-143) */
int ctgan_interpolate(const float* real, const float* fake, const float* alpha, float* out,
                      int B, int P, void* stream);

/* ---- batch norm (+ conditional gamma/beta gather, + fused ReLU) ----------
 * Training-mode BN over rows = N*HW per channel, biased variance, eps
 * (tf.nn.fused_batch_norm TG/tflib/ops/batchnorm.py:29-30; moments+batch_normalization
 * batchnorm.py:77-84 and cond_batchnorm.py:10-16).  x is [N,HW,C].  gamma/beta are
 * float [n_labels][C]; labels (int32 [N], nullable) picks the row per sample
 * (labels == NULL -> row 0).  relu != 0 fuses tf.nn.relu on the output.
 * groups >= 1 (N % groups == 0): statistics are computed separately for each block of N/groups
 * consecutive samples -- the reference normalises per device split (TG/CT_gan_cifar_resnet.py:196-199).
 * ws: float workspace of ctgan_bn_workspace_floats(N, HW, C, groups) floats.
 * save_mean / save_invstd: float [groups][C], consumed by the backward. */
int64_t ctgan_bn_workspace_floats(int N, int HW, int C, int groups);
int ctgan_bn_fwd(const void* x, const float* gamma, const float* beta, const int32_t* labels,
                 void* y, float* save_mean, float* save_invstd, float* ws,
                 int N, int HW, int C, float eps, int relu, int groups, int dtype, void* stream);
/* y is the forward OUTPUT (used for the ReLU mask when relu != 0).  dgamma/dbeta are
 * float [n_labels][C], overwritten. */
int ctgan_bn_bwd(const void* dy, const void* x, const void* y, const float* gamma,
                 const int32_t* labels, const float* save_mean, const float* save_invstd,
                 void* dx, float* dgamma, float* dbeta, float* ws,
                 int N, int HW, int C, int n_labels, int relu, int groups, int dtype, void* stream);

/* BF16 fast path of the two calls above: TWO kernels per direction (statistics with red.global into ws + apply; the finalize
 * launches are gone), optional 2x nearest-neighbour upsampling of the output fused in (the tf.concat x4 + depth_to_space of
 * UpsampleConv, TG/CT_gan_cifar_resnet.py:100-107, which follows Normalize + relu in every generator block :132-137), the
 * ReLU pattern recomputed from x in the backward (y is not an input), and parameter gradients optionally accumulated in place.
 * x: [N,H,W,C] bf16.  flags: CTGAN_BN_RELU; CTGAN_BN_UP2 (fwd: y is [N,2H,2W,C]; bwd: dy has that shape and is summed 2x2 on
 * load); CTGAN_BN_ACCUM (bwd: dgamma / dbeta [n_labels][C] are added to instead of overwritten, e.g. slices of the flat
 * gradient bucket).  ws: >= 2*groups*C floats.  Eligible (ctgan_bn_fused_ok): BF16, C/8 a power of two <= 256, groups | N. */
#define CTGAN_BN_RELU  1
#define CTGAN_BN_UP2   2
#define CTGAN_BN_ACCUM 4
int ctgan_bn_fused_ok(int N, int H, int W, int C, int groups, int dtype);
int ctgan_bn_fwd_fused(const void* x, const float* gamma, const float* beta, const int32_t* labels, void* y,
                       float* save_mean, float* save_invstd, float* ws, int N, int H, int W, int C, float eps,
                       int flags, int groups, void* stream);
int ctgan_bn_bwd_fused(const void* dy, const void* x, const float* gamma, const float* beta, const int32_t* labels,
                       const float* save_mean, const float* save_invstd, void* dx, float* dgamma, float* dbeta,
                       float* ws, int N, int H, int W, int C, int n_labels, int flags, int groups, void* stream);

/* ---- layer normalisation over (C,H,W) per sample, per-channel scale / offset (the critic's
 * Normalize of TG/CT_gan_64x64.py:87-93; op TG/tflib/ops/layernorm.py:6-21, eps 1e-5).  x, y, v, out: NHWC activations
 * [N][M], M = H*W*C; mean, rstd: float [N]; gamma, beta, dgamma, dbeta: float [C]; ws: ctgan_ln_workspace_floats() floats.
 * With xh = (x - mean) * rstd and core(u) = rstd * (u - mean_s(u) - xh * mean_s(u * xh)) over each sample:
 *   ln_fwd         y = xh * gamma + beta (also writes mean, rstd)
 *   ln_core        out = [gamma *] core([gamma *] v)      pre_scale / post_scale select the two products:
 *                  backward dx = core(gamma * gy);  backward-of-backward ggy = gamma * core(c)
 *   ln_param_grad  dgamma[c] += sum v * xh,  dbeta[c] += sum v   (dbeta nullable)
 *   ln_bwd2_x      the x-derivative of <c, dx>:  gx = -rstd^2 * [xh*Q + B*(a - mean a) + A*(b - mean b) - 2*xh*A*B],
 *                  a = c, b = gamma * gy, A = mean_s(a*xh), B = mean_s(b*xh), Q = mean_s(a*b) - mean a * mean b - A*B */
int64_t ctgan_ln_workspace_floats(int N, int64_t M);
int ctgan_ln_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, float* ws,
                 int N, int64_t M, int C, float eps, int dtype, void* stream);
int ctgan_ln_core(const void* v, const void* x, const float* gamma, const float* mean, const float* rstd, void* out, float* ws,
                  int N, int64_t M, int C, int pre_scale, int post_scale, int dtype, void* stream);
int ctgan_ln_param_grad(const void* v, const void* x, const float* mean, const float* rstd, float* dgamma, float* dbeta,
                        int N, int64_t M, int C, int dtype, void* stream);
int ctgan_ln_bwd2_x(const void* c, const void* gy, const void* x, const float* gamma, const float* mean, const float* rstd,
                    void* gx, float* ws, int N, int64_t M, int C, int dtype, void* stream);

/* ---- fused CT + GP + WGAN (+ACGAN) loss -----------------------------------
 * TG/CT_gan_cifar.py:123-151, TG/CT_gan_mnist.py:146-167, TG/CT_gan_cifar_resnet.py:244-300.
 *   wgan  = mean(d_fake[0..NF)) - mean(d_real[0..B))
 *   CT_i  = lambda2*(d_real-d_real2)^2 + 0.1*lambda2*mean_f((f1-f2)^2); ct = mean(max(CT_i - M, 0))
 *   s_i   = ||grad[i,:]||_2 ;  gp = mean((s_i-1)^2)
 *   acgan = mean(logsumexp(logits_i) - logits_i[label_i])      (logits nullable)
 *   cost  = wgan + ct + lambda*gp + acgan_scale*acgan
 * d_* are float; f1,f2 [B,F] have `feat_dtype`; grad [B,P] float; logits [B,10] float.
 * out (float[8]) = {cost, wgan, ct, gp, acgan, 0,0,0}.  per_sample (float [4*B]) keeps
 * {CT_i - M, s_i, 0, 0} for the backward.  One warp-shuffle reduction kernel.
 * The two halves can be evaluated separately (the penalty shares nothing with the critic outputs of the stacked pass, so its
 * branch of the step need not wait for them): grad == NULL -> no penalty term (gp = 0, g_grad untouched);
 * d_real == d_real2 == d_fake == f1 == f2 == NULL (and logits NULL) -> cost = lambda*gp only. */
typedef struct {
    int32_t B, NF, F, P, n_classes, feat_dtype;
    float lambda_gp, lambda2, factor_m, acgan_scale;
} ctgan_loss_desc;
int ctgan_ct_gp_loss_fwd(const ctgan_loss_desc* d, const float* d_real, const float* d_real2,
                         const float* d_fake, const void* f1, const void* f2, const float* grad,
                         const float* logits, const int32_t* labels,
                         float* out, float* per_sample, void* stream);
/* cotangents of `cost` scaled by *gcost (device float): g_d_real [B], g_d_real2 [B],
 * g_d_fake [NF], g_f1/g_f2 [B,F] (feat_dtype), g_grad [B,P] float, g_logits [B,n_classes] float. */
int ctgan_ct_gp_loss_bwd(const ctgan_loss_desc* d, const float* gcost,
                         const float* d_real, const float* d_real2, const void* f1, const void* f2,
                         const float* grad, const float* logits, const int32_t* labels,
                         const float* per_sample,
                         float* g_d_real, float* g_d_real2, float* g_d_fake, void* g_f1, void* g_f2,
                         float* g_grad, float* g_logits, void* stream);
/* generator loss pieces: out[0] = sign * mean(d[0..n)); g[i] = sign * (*gcost) / n */
int ctgan_mean_fwd(const float* d, float* out, int n, float sign, void* stream);
int ctgan_mean_bwd(const float* gcost, float* g, int n, float sign, void* stream);
/* softmax cross-entropy: out[0] = mean_i CE_i ; g_logits = (*gcost) * scale * (softmax - onehot) / B */
int ctgan_softmax_ce_fwd(const float* logits, const int32_t* labels, float* out, int B, int n_classes, void* stream);
int ctgan_softmax_ce_bwd(const float* logits, const int32_t* labels, const float* gcost, float scale,
                         float* g_logits, int B, int n_classes, void* stream);

/* ---- optimizer -------------------------------------------------------------
 * tf.train.AdamOptimizer on one flat float buffer (TG/CT_gan_cifar_resnet.py:333-338):
 *   m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g ; p -= lr_t * m / (sqrt(v) + eps)
 * lr_t = lr*sqrt(1-b2^t)/(1-b1^t) is computed by the host and passed in.
 * grad_scale multiplies g first (1/world_size after a sum all-reduce). */
int ctgan_adam_step(float* p, const float* g, float* m, float* v, int64_t n,
                    float lr_t, float beta1, float beta2, float eps, float grad_scale,
                    const float* lr_t_dev /*nullable: overrides lr_t, for graph replay*/, void* stream);

/* ---- data parallel: the optimizer update as ONE kernel over NVLink peer memory (csrc/peer.cu) --------
 * reduce-scatter of the ranks' flat gradient buckets + Adam on the owned slice + all-gather of the new parameters, with
 * in-kernel barriers over peer flag words: replaces NCCL all-reduce + ctgan_adam_step, needs no host call between two CUDA
 * graphs, and keeps replicas bit-identical (each element is summed and updated by exactly one rank).
 *   ctgan_peer_alloc / _free          cudaMalloc'ed (zeroed) device memory that can be exported over CUDA IPC
 *   ctgan_ipc_get_handle / _open / _close   64-byte cudaIpcMemHandle_t of such an allocation <-> its mapping in another process
 *   ctgan_peer_flag_bytes             size of one rank's flag block (allocate with ctgan_peer_alloc)
 *   ctgan_peer_reduce_adam            g, p, flags: HOST arrays of `world` device pointers (entry `rank` = own buffers, the others
 *                                     the peers' buffers opened over IPC); m, v: own moments; n floats (multiple of 4);
 *                                     update rule and lr_t_dev as ctgan_adam_step.  Every rank must launch it once per step. */
int ctgan_peer_alloc(void** out, int64_t bytes);
int ctgan_peer_free(void* ptr);
int ctgan_ipc_get_handle(const void* ptr, void* handle64);
int ctgan_ipc_open_handle(const void* handle64, void** out);
int ctgan_ipc_close_handle(void* ptr);
int ctgan_peer_flag_bytes(void);
int ctgan_peer_reduce_adam(int world, int rank, float* const* g, float* const* p, void* const* flags, float* m, float* v,
                           int64_t n, float lr_t, float beta1, float beta2, float eps, float grad_scale,
                           const float* lr_t_dev /*nullable*/, void* stream);

/* ---- Philox4x32-10 random numbers -------------------------------------------
 * Element i of a stream (seed, offset) is lane (offset+i)&3 of
 * philox4x32_10(counter = (offset+i)>>2, key = seed); u = (bits >> 8) * 2^-24 in [0,1).
 * uniform:  out = lo + (hi-lo)*u           (tf.random_uniform, TG/CT_gan_cifar.py:138)
 * normal:   Box-Muller on elements (2i,2i+1) of the stream   (tf.random_normal, :60)
 * labels:   (int32)(u * n_labels)          (TG/CT_gan_cifar_resnet.py:319) */
int ctgan_philox_uniform(float* out, int64_t n, float lo, float hi, uint64_t seed, uint64_t offset,
                         const uint64_t* dyn_offset, void* stream);
int ctgan_philox_normal(float* out, int64_t n, uint64_t seed, uint64_t offset, const uint64_t* dyn_offset, void* stream);
int ctgan_philox_labels(int32_t* out, int64_t n, int n_labels, uint64_t seed, uint64_t offset,
                        const uint64_t* dyn_offset, void* stream);
/* CUDA-graph-safe streams: every RNG consumer adds *dyn_offset (device, nullable, a multiple of 4) to
 * its `offset`, so a captured step draws fresh numbers on every replay once the counter is advanced
 * by ctgan_counter_add (itself a kernel, captured at the end of the step). */
int ctgan_counter_add(uint64_t* counter, uint64_t delta, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CTGAN_SM100_H */
