// loss.cu -- the fused CT + GP + WGAN (+ACGAN) critic loss and the generator loss pieces.
//
// One warp-shuffle reduction kernel evaluates, per sample, the consistency term
// d(D(x'),D(x'')) + 0.1 d(D_(x'),D_(x'')) - M', the gradient-penalty slope ||grad_x D||_2 and
// the ACGAN cross-entropy; a one-CTA tail folds the per-sample values into the five
// scalars.  Replaces the ~25 TF ops of TG/CT_gan_cifar.py:123-151,
// TG/CT_gan_mnist.py:146-167 and TG/CT_gan_cifar_resnet.py:244-248,285-300.
// Latency-bound (R config: 852 KB of inputs), so: one launch for the per-sample pass, one
// tiny launch for the tail, no atomics (bitwise reproducible).
#include "common.cuh"

namespace ctgan {

constexpr int LOSS_T = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// sum over the CTA; result valid in every thread
__device__ __forceinline__ float block_sum(float v, float* sh) {
    v = warp_sum(v);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    float t = (l < (blockDim.x >> 5)) ? sh[l] : 0.f;
    t = warp_sum(t);
    return t;
}

// per_sample[i*4 + 0] = CT_i - M,  [1] = slope_i,  [2] = CE_i,  [3] = unused
__global__ void __launch_bounds__(LOSS_T)
ct_gp_per_sample_kernel(ctgan_loss_desc d, const float* __restrict__ d_real, const float* __restrict__ d_real2,
                        const void* __restrict__ f1, const void* __restrict__ f2, const float* __restrict__ grad,
                        const float* __restrict__ logits, const int32_t* __restrict__ labels,
                        float* __restrict__ per_sample) {
    ctgan::pdl_entry();
    __shared__ float sh[32];
    // The two halves of the loss can be evaluated separately (the gradient-penalty term shares nothing with the critic
    // outputs of the stacked pass): grad == nullptr -> no penalty term (slope recorded as 1), d_real == nullptr -> penalty only.
    const int i = blockIdx.x;
    float sq = 0.f;
    if (d_real) {
        for (int f = threadIdx.x; f < d.F; f += LOSS_T) {
            float a = ld_act(f1, (int64_t)i * d.F + f, d.feat_dtype) - ld_act(f2, (int64_t)i * d.F + f, d.feat_dtype);
            sq += a * a;
        }
    }
    sq = block_sum(sq, sh);
    float g2 = 0.f;
    if (grad) {
        for (int p = threadIdx.x; p < d.P; p += LOSS_T) {
            float a = grad[(int64_t)i * d.P + p];
            g2 += a * a;
        }
    }
    g2 = block_sum(g2, sh);
    if (threadIdx.x == 0) {
        float ct = -1.f;                                   // inactive under max(CT - M, 0)
        if (d_real) {
            float diff = d_real[i] - d_real2[i];
            ct = d.lambda2 * diff * diff + d.lambda2 * 0.1f * (sq / (float)d.F) - d.factor_m;
        }
        per_sample[i * 4 + 0] = ct;
        per_sample[i * 4 + 1] = grad ? sqrtf(g2) : 1.f;
        float ce = 0.f;
        if (logits) {
            const float* lg = logits + (int64_t)i * d.n_classes;
            float mx = lg[0];
            for (int k = 1; k < d.n_classes; ++k) mx = fmaxf(mx, lg[k]);
            float se = 0.f;
            for (int k = 0; k < d.n_classes; ++k) se += expf(lg[k] - mx);
            ce = logf(se) + mx - lg[labels[i]];
        }
        per_sample[i * 4 + 2] = ce;
        per_sample[i * 4 + 3] = 0.f;
    }
}

__global__ void __launch_bounds__(LOSS_T)
ct_gp_tail_kernel(ctgan_loss_desc d, const float* __restrict__ d_real, const float* __restrict__ d_fake,
                  const float* __restrict__ per_sample, int has_logits, float* __restrict__ out) {
    ctgan::pdl_entry();
    __shared__ float sh[32];
    float sr = 0.f, sf = 0.f, sct = 0.f, sgp = 0.f, sce = 0.f;
    for (int i = threadIdx.x; i < d.B; i += LOSS_T) {
        if (d_real) sr += d_real[i];
        float c = per_sample[i * 4 + 0];
        sct += fmaxf(c, 0.f);
        float s = per_sample[i * 4 + 1] - 1.f;
        sgp += s * s;
        sce += per_sample[i * 4 + 2];
    }
    if (d_fake) for (int i = threadIdx.x; i < d.NF; i += LOSS_T) sf += d_fake[i];
    sr = block_sum(sr, sh); sf = block_sum(sf, sh); sct = block_sum(sct, sh);
    sgp = block_sum(sgp, sh); sce = block_sum(sce, sh);
    if (threadIdx.x == 0) {
        float wgan = sf / (float)d.NF - sr / (float)d.B;
        float ct = sct / (float)d.B, gp = sgp / (float)d.B;
        float acgan = has_logits ? sce / (float)d.B : 0.f;
        out[0] = wgan + ct + d.lambda_gp * gp + d.acgan_scale * acgan;
        out[1] = wgan; out[2] = ct; out[3] = gp; out[4] = acgan;
        out[5] = 0.f; out[6] = 0.f; out[7] = 0.f;
    }
}

__global__ void __launch_bounds__(LOSS_T)
ct_gp_bwd_kernel(ctgan_loss_desc d, const float* __restrict__ gcost,
                 const float* __restrict__ d_real, const float* __restrict__ d_real2,
                 const void* __restrict__ f1, const void* __restrict__ f2, const float* __restrict__ grad,
                 const float* __restrict__ logits, const int32_t* __restrict__ labels,
                 const float* __restrict__ per_sample,
                 float* __restrict__ g_d_real, float* __restrict__ g_d_real2, float* __restrict__ g_d_fake,
                 void* __restrict__ g_f1, void* __restrict__ g_f2, float* __restrict__ g_grad,
                 float* __restrict__ g_logits) {
    ctgan::pdl_entry();
    const int i = blockIdx.x;
    const float g = gcost[0];
    const float invB = 1.f / (float)d.B;
    // tf.maximum(CT-M, 0*(CT-M)): the gradient flows to the first argument where it is >= the second
    const float active = per_sample[i * 4 + 0] >= 0.f ? 1.f : 0.f;
    const float gct = g * active * invB;
    if (d_real) {
        if (threadIdx.x == 0) {
            float diff = d_real[i] - d_real2[i];
            float t = gct * 2.f * d.lambda2 * diff;
            g_d_real[i] = -g * invB + t;
            g_d_real2[i] = -t;
        }
        for (int j = i * LOSS_T + threadIdx.x; j < d.NF; j += gridDim.x * LOSS_T) g_d_fake[j] = g / (float)d.NF;
        const float cf = gct * 0.1f * d.lambda2 * 2.f / (float)d.F;
        for (int f = threadIdx.x; f < d.F; f += LOSS_T) {
            int64_t o = (int64_t)i * d.F + f;
            float a = ld_act(f1, o, d.feat_dtype) - ld_act(f2, o, d.feat_dtype);
            st_act(g_f1, o, d.feat_dtype, cf * a);
            st_act(g_f2, o, d.feat_dtype, -cf * a);
        }
    }
    // d/dgrad of lambda * mean((s-1)^2), s = ||grad||: lambda * 2 (s-1)/s * grad / B.
    // The reference has no epsilon under the sqrt (TG/CT_gan_cifar.py:148), so s == 0 is NaN there;
    // here a zero slope yields a zero cotangent instead of poisoning the step.
    const float s = per_sample[i * 4 + 1];
    const float cg = s > 0.f ? g * d.lambda_gp * invB * 2.f * (s - 1.f) / s : 0.f;
    if (grad) {
        for (int p = threadIdx.x; p < d.P; p += LOSS_T) {
            int64_t o = (int64_t)i * d.P + p;
            g_grad[o] = cg * grad[o];
        }
    }
    if (logits && threadIdx.x < d.n_classes) {
        const float* lg = logits + (int64_t)i * d.n_classes;
        float mx = lg[0];
        for (int k = 1; k < d.n_classes; ++k) mx = fmaxf(mx, lg[k]);
        float se = 0.f;
        for (int k = 0; k < d.n_classes; ++k) se += expf(lg[k] - mx);
        int k = threadIdx.x;
        float p = expf(lg[k] - mx) / se;
        g_logits[(int64_t)i * d.n_classes + k] = g * d.acgan_scale * invB * (p - (k == labels[i] ? 1.f : 0.f));
    }
}

__global__ void __launch_bounds__(LOSS_T)
mean_kernel(const float* __restrict__ x, float* __restrict__ out, int n, float sign) {
    ctgan::pdl_entry();
    __shared__ float sh[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += LOSS_T) s += x[i];
    s = block_sum(s, sh);
    if (threadIdx.x == 0) out[0] = sign * s / (float)n;
}

__global__ void mean_bwd_kernel(const float* __restrict__ gcost, float* __restrict__ g, int n, float sign) {
    ctgan::pdl_entry();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) g[i] = sign * gcost[0] / (float)n;
}

__global__ void __launch_bounds__(LOSS_T)
softmax_ce_fwd_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels, float* __restrict__ out,
                      int B, int K) {
    ctgan::pdl_entry();
    __shared__ float sh[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < B; i += LOSS_T) {
        const float* lg = logits + (int64_t)i * K;
        float mx = lg[0];
        for (int k = 1; k < K; ++k) mx = fmaxf(mx, lg[k]);
        float se = 0.f;
        for (int k = 0; k < K; ++k) se += expf(lg[k] - mx);
        s += logf(se) + mx - lg[labels[i]];
    }
    s = block_sum(s, sh);
    if (threadIdx.x == 0) out[0] = s / (float)B;
}
__global__ void softmax_ce_bwd_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels,
                                      const float* __restrict__ gcost, float scale, float* __restrict__ g_logits,
                                      int B, int K) {
    ctgan::pdl_entry();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const float* lg = logits + (int64_t)i * K;
    float mx = lg[0];
    for (int k = 1; k < K; ++k) mx = fmaxf(mx, lg[k]);
    float se = 0.f;
    for (int k = 0; k < K; ++k) se += expf(lg[k] - mx);
    float c = gcost[0] * scale / (float)B;
    for (int k = 0; k < K; ++k)
        g_logits[(int64_t)i * K + k] = c * (expf(lg[k] - mx) / se - (k == labels[i] ? 1.f : 0.f));
}

static int check_loss_desc(const ctgan_loss_desc* d) {
    CTGAN_REQUIRE(d != nullptr, CTGAN_ERR_BAD_DESC, "loss: null descriptor");
    CTGAN_REQUIRE(d->B > 0 && d->NF > 0 && d->F > 0 && d->P > 0, CTGAN_ERR_BAD_DESC, "loss: non-positive size");
    CTGAN_REQUIRE(dtype_ok(d->feat_dtype), CTGAN_ERR_BAD_DESC, "loss: bad feat_dtype");
    CTGAN_REQUIRE(d->n_classes >= 0 && d->n_classes <= LOSS_T, CTGAN_ERR_BAD_DESC, "loss: n_classes out of range");
    return 0;
}

}  // namespace ctgan

using namespace ctgan;

extern "C" int ctgan_ct_gp_loss_fwd(const ctgan_loss_desc* d, const float* d_real, const float* d_real2,
                                    const float* d_fake, const void* f1, const void* f2, const float* grad,
                                    const float* logits, const int32_t* labels,
                                    float* out, float* per_sample, void* stream) {
    if (int r = check_loss_desc(d)) return r;
    CTGAN_REQUIRE(out && per_sample && (grad || d_real) && (!d_real || (d_real2 && d_fake && f1 && f2)), CTGAN_ERR_BAD_DESC,
                  "loss_fwd: null pointer (grad may be null: no penalty term; d_real.. may be null together: penalty only)");
    CTGAN_REQUIRE(d_real || !logits, CTGAN_ERR_BAD_DESC, "loss_fwd: the penalty-only form takes no logits");
    CTGAN_REQUIRE(!logits || (labels && d->n_classes > 0), CTGAN_ERR_BAD_DESC, "loss_fwd: logits need labels and n_classes");
    cudaStream_t st = as_stream(stream);
    CTGAN_LAUNCH((ct_gp_per_sample_kernel), d->B, LOSS_T, 0, st, *d, d_real, d_real2, f1, f2, grad, logits, labels, per_sample);
    CTGAN_CHECK_LAUNCH("ct_gp_per_sample");
    CTGAN_LAUNCH((ct_gp_tail_kernel), 1, LOSS_T, 0, st, *d, d_real, d_fake, per_sample, logits != nullptr, out);
    CTGAN_CHECK_LAUNCH("ct_gp_tail");
    return 0;
}

extern "C" int ctgan_ct_gp_loss_bwd(const ctgan_loss_desc* d, const float* gcost,
                                    const float* d_real, const float* d_real2, const void* f1, const void* f2,
                                    const float* grad, const float* logits, const int32_t* labels,
                                    const float* per_sample,
                                    float* g_d_real, float* g_d_real2, float* g_d_fake, void* g_f1, void* g_f2,
                                    float* g_grad, float* g_logits, void* stream) {
    if (int r = check_loss_desc(d)) return r;
    CTGAN_REQUIRE(gcost && per_sample && (grad || d_real) && (!grad || g_grad) &&
                  (!d_real || (d_real2 && f1 && f2 && g_d_real && g_d_real2 && g_d_fake && g_f1 && g_f2)),
                  CTGAN_ERR_BAD_DESC, "loss_bwd: null pointer");
    CTGAN_REQUIRE(d_real || !logits, CTGAN_ERR_BAD_DESC, "loss_bwd: the penalty-only form takes no logits");
    CTGAN_REQUIRE(!logits || (labels && g_logits && d->n_classes > 0), CTGAN_ERR_BAD_DESC, "loss_bwd: logits need labels/g_logits");
    CTGAN_LAUNCH((ct_gp_bwd_kernel), d->B, LOSS_T, 0, as_stream(stream), *d, gcost, d_real, d_real2, f1, f2, grad, logits, labels,
                                                            per_sample, g_d_real, g_d_real2, g_d_fake, g_f1, g_f2,
                                                            g_grad, g_logits);
    CTGAN_CHECK_LAUNCH("ct_gp_bwd");
    return 0;
}

extern "C" int ctgan_mean_fwd(const float* x, float* out, int n, float sign, void* stream) {
    CTGAN_REQUIRE(x && out && n > 0, CTGAN_ERR_BAD_DESC, "mean_fwd: bad args");
    CTGAN_LAUNCH((mean_kernel), 1, LOSS_T, 0, as_stream(stream), x, out, n, sign);
    CTGAN_CHECK_LAUNCH("mean_fwd");
    return 0;
}
extern "C" int ctgan_mean_bwd(const float* gcost, float* g, int n, float sign, void* stream) {
    CTGAN_REQUIRE(gcost && g && n > 0, CTGAN_ERR_BAD_DESC, "mean_bwd: bad args");
    CTGAN_LAUNCH((mean_bwd_kernel), ceil_div(n, 128), 128, 0, as_stream(stream), gcost, g, n, sign);
    CTGAN_CHECK_LAUNCH("mean_bwd");
    return 0;
}
extern "C" int ctgan_softmax_ce_fwd(const float* logits, const int32_t* labels, float* out, int B, int K, void* stream) {
    CTGAN_REQUIRE(logits && labels && out && B > 0 && K > 0, CTGAN_ERR_BAD_DESC, "softmax_ce_fwd: bad args");
    CTGAN_LAUNCH((softmax_ce_fwd_kernel), 1, LOSS_T, 0, as_stream(stream), logits, labels, out, B, K);
    CTGAN_CHECK_LAUNCH("softmax_ce_fwd");
    return 0;
}
extern "C" int ctgan_softmax_ce_bwd(const float* logits, const int32_t* labels, const float* gcost, float scale,
                                    float* g_logits, int B, int K, void* stream) {
    CTGAN_REQUIRE(logits && labels && gcost && g_logits && B > 0 && K > 0, CTGAN_ERR_BAD_DESC, "softmax_ce_bwd: bad args");
    CTGAN_LAUNCH((softmax_ce_bwd_kernel), ceil_div(B, 128), 128, 0, as_stream(stream), logits, labels, gcost, scale, g_logits, B, K);
    CTGAN_CHECK_LAUNCH("softmax_ce_bwd");
    return 0;
}
