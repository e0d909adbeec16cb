// norm.cu -- training-mode batch norm (+ per-label gamma/beta gather, + fused ReLU) and
// the bias-gradient column reduction.  HBM-bound: forward reads x twice (stats, apply) and
// writes y once; backward reads dy,x(,y) twice and writes dx once.
//
// Replaces tf.nn.fused_batch_norm (TG/tflib/ops/batchnorm.py:29-30), tf.nn.moments +
// tf.nn.batch_normalization (batchnorm.py:77-84, cond_batchnorm.py:10-16) and the
// following tf.nn.relu (TG/CT_gan_cifar_resnet.py:135,138,164; TG/CT_gan_cifar.py:64,69,73).
#include "common.cuh"

namespace ctgan {

constexpr int BN_LANES = 32;   // threads along channels, each owning 4 consecutive channels
constexpr int BN_ROWS  = 8;    // threads along rows
constexpr int BN_CCH   = BN_LANES * 4;   // channels per CTA

static inline int bn_fwd_rowblocks(int64_t R) {
    int64_t nb = (R + 63) / 64;
    if (nb > 160) nb = 160;
    if (nb < 1) nb = 1;
    return (int)nb;
}
static inline int bn_bwd_splits(int HW) {
    int s = HW / 64;
    if (s < 1) s = 1;
    if (s > 4) s = 4;
    return s;
}

template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <> __device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x), b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
    v[0] = __bfloat162float(a.x); v[1] = __bfloat162float(a.y); v[2] = __bfloat162float(b.x); v[3] = __bfloat162float(b.y);
}
template <typename T>
__device__ __forceinline__ void store4(T* p, const float (&v)[4]);
template <> __device__ __forceinline__ void store4<float>(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a); t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
}

// Statistics can be computed per GROUP of samples (groups G >= 1, N % G == 0: group g = samples
// [g*N/G, (g+1)*N/G)): the reference builds its generator once per device split, so batch-norm
// statistics are per 32- / 64-sample split (TG/CT_gan_cifar_resnet.py:196-199, 314-321); running
// the splits as ONE batch with G = 2 halves the launch count and gives identical results.

// ---- forward stage 1: per-(group, row block, channel) Welford partials ----------------
// blockIdx.x = g * nbg + rb;  ws[(blockIdx.x*C + c)*2 + {0,1}] = {mean, M2} over that block's rows.
template <typename T>
__global__ void __launch_bounds__(BN_LANES * BN_ROWS)
bn_stats_kernel(const T* __restrict__ x, float* __restrict__ ws, int64_t Rg, int C, int rows_per_block, int nbg) {
    ctgan::pdl_entry();
    const int lane = threadIdx.x, ry = threadIdx.y;
    const int c = blockIdx.y * BN_CCH + lane * 4;
    const int g = blockIdx.x / nbg, rb = blockIdx.x - g * nbg;
    const int64_t r0 = (int64_t)g * Rg + (int64_t)rb * rows_per_block;
    const int64_t r1 = min((int64_t)(g + 1) * Rg, r0 + rows_per_block);
    float cnt = 0.f, mean[4] = {0, 0, 0, 0}, m2[4] = {0, 0, 0, 0};
    if (c < C) {
        for (int64_t r = r0 + ry; r < r1; r += BN_ROWS) {
            float v[4];
            load4<T>(x + r * C + c, v);
            cnt += 1.f;
            float inv = 1.f / cnt;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float d = v[j] - mean[j];
                mean[j] += d * inv;
                m2[j] += d * (v[j] - mean[j]);
            }
        }
    }
    __shared__ float s_cnt[BN_ROWS][BN_LANES];
    __shared__ float s_mean[BN_ROWS][BN_LANES][4];
    __shared__ float s_m2[BN_ROWS][BN_LANES][4];
    s_cnt[ry][lane] = cnt;
#pragma unroll
    for (int j = 0; j < 4; ++j) { s_mean[ry][lane][j] = mean[j]; s_m2[ry][lane][j] = m2[j]; }
    __syncthreads();
    if (ry == 0 && c < C) {
        for (int k = 1; k < BN_ROWS; ++k) {
            float nb = s_cnt[k][lane];
            if (nb == 0.f) continue;
            float n = cnt + nb;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float d = s_mean[k][lane][j] - mean[j];
                mean[j] += d * (nb / n);
                m2[j] += s_m2[k][lane][j] + d * d * (cnt * nb / n);
            }
            cnt = n;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            ws[((int64_t)blockIdx.x * C + c + j) * 2 + 0] = mean[j];
            ws[((int64_t)blockIdx.x * C + c + j) * 2 + 1] = m2[j];
        }
    }
}

// ---- forward stage 2: merge the row-block partials (Chan), biased variance --------
// One warp per (group, channel): lane l folds partials l, l+32, ... sequentially, then a 5-step
// butterfly merges the 32 (count, mean, M2) triples.  save_mean / save_invstd are [G][C].
__device__ __forceinline__ void chan_merge(float& cnt, float& mean, float& m2, float cb, float mb, float m2b) {
    if (cb > 0.f) {
        float n = cnt + cb;
        float d = mb - mean;
        mean += d * (cb / n);
        m2 += m2b + d * d * (cnt * cb / n);
        cnt = n;
    }
}
__global__ void __launch_bounds__(128)
bn_finalize_kernel(const float* __restrict__ ws, float* __restrict__ save_mean,
                   float* __restrict__ save_invstd, int64_t Rg, int C, int nbg,
                   int rows_per_block, float eps) {
    ctgan::pdl_entry();
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int g = blockIdx.y;
    if (c >= C) return;
    float cnt = 0.f, mean = 0.f, m2 = 0.f;
    for (int b = lane; b < nbg; b += 32) {
        int64_t r0 = (int64_t)b * rows_per_block;
        float nbc = (float)(min(Rg, r0 + rows_per_block) - r0);
        const float* p = ws + ((int64_t)(g * nbg + b) * C + c) * 2;
        chan_merge(cnt, mean, m2, nbc, p[0], p[1]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float cb = __shfl_xor_sync(0xffffffffu, cnt, o);
        float mb = __shfl_xor_sync(0xffffffffu, mean, o);
        float m2b = __shfl_xor_sync(0xffffffffu, m2, o);
        chan_merge(cnt, mean, m2, cb, mb, m2b);
    }
    if (lane == 0) {
        save_mean[(int64_t)g * C + c] = mean;
        save_invstd[(int64_t)g * C + c] = rsqrtf(m2 / (float)Rg + eps);
    }
}

// ---- forward stage 3: normalise, affine (per-label rows), optional ReLU -------------
template <typename T>
__global__ void bn_apply_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                const int32_t* __restrict__ labels, const float* __restrict__ mean,
                                const float* __restrict__ invstd, T* __restrict__ y,
                                int64_t R, int HW, int C, int relu, int n_per_group) {
    ctgan::pdl_entry();
    const int C4 = C / 4;
    int64_t total = R * C4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % C4) * 4;
        int64_t r = i / C4;
        int n = (int)(r / HW);
        int l = labels ? labels[n] : 0;
        const float* mu = mean + (int64_t)(n / n_per_group) * C;
        const float* is = invstd + (int64_t)(n / n_per_group) * C;
        float v[4], o[4];
        load4<T>(x + r * C + c, v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float xh = (v[j] - mu[c + j]) * is[c + j];
            float t = xh * gamma[(int64_t)l * C + c + j] + beta[(int64_t)l * C + c + j];
            o[j] = relu ? fmaxf(t, 0.f) : t;
        }
        store4<T>(y + r * C + c, o);
    }
}

// ---- BF16 fast paths (C/8 a power of two <= 256): 16-byte accesses, C/8 threads per row --------------------
// Same workspace layout and finalize kernel as above.  Per thread plain sum / sum of squares over its <= ~64 rows
// (converted to (mean, M2) before any merge, so the cancellation stays at the 1e-7 level), then Chan merges.
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int e = 0; e < 4; ++e) { f[2 * e] = __bfloat162float(h[e].x); f[2 * e + 1] = __bfloat162float(h[e].y); }
}
__global__ void __launch_bounds__(256)
bn_stats_bf16_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ ws, int64_t Rg, int C, int rows_per_block, int nbg) {
    ctgan::pdl_entry();
    __shared__ float sh_mean[2048], sh_m2[2048], sh_cnt[256];
    const int tpr = C >> 3;
    const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr, rstep = 256 / tpr;
    const int g = blockIdx.x / nbg, rb = blockIdx.x - g * nbg;
    const int64_t r0 = (int64_t)g * Rg + (int64_t)rb * rows_per_block;
    const int64_t r1 = min((int64_t)(g + 1) * Rg, r0 + rows_per_block);
    float s[8], ss[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] = 0.f; ss[e] = 0.f; }
    const __nv_bfloat16* base = x + cg * 8;
    int64_t r = r0 + rl;
    float cnt = 0.f;
    for (; r + 3 * (int64_t)rstep < r1; r += 4 * (int64_t)rstep) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(base + (r + u * (int64_t)rstep) * C));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float f[8];
            unpack8(v[u], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) { s[e] += f[e]; ss[e] = fmaf(f[e], f[e], ss[e]); }
        }
        cnt += 4.f;
    }
    for (; r < r1; r += rstep) {
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(base + r * C)), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) { s[e] += f[e]; ss[e] = fmaf(f[e], f[e], ss[e]); }
        cnt += 1.f;
    }
    const float inv = cnt > 0.f ? 1.f / cnt : 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float m = s[e] * inv;
        sh_mean[rl * C + cg * 8 + e] = m;
        sh_m2[rl * C + cg * 8 + e] = fmaxf(ss[e] - s[e] * m, 0.f);
    }
    if (cg == 0) sh_cnt[rl] = cnt;
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float n = 0.f, mean = 0.f, m2 = 0.f;
        for (int k = 0; k < rstep; ++k) chan_merge(n, mean, m2, sh_cnt[k], sh_mean[k * C + c], sh_m2[k * C + c]);
        ws[((int64_t)blockIdx.x * C + c) * 2 + 0] = mean;
        ws[((int64_t)blockIdx.x * C + c) * 2 + 1] = m2;
    }
}

// blockIdx.y = sample, blockIdx.x = chunk of its HW rows: the per-(sample, channel) scale / shift are computed once
// per thread, then y = relu(x * scale + shift) with 16-byte loads and stores.
__global__ void __launch_bounds__(256)
bn_apply_bf16_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     const int32_t* __restrict__ labels, const float* __restrict__ mean, const float* __restrict__ invstd,
                     __nv_bfloat16* __restrict__ y, int HW, int C, int relu, int n_per_group, int rows_per_chunk) {
    ctgan::pdl_entry();
    const int tpr = C >> 3;
    const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr, rstep = 256 / tpr;
    const int n = blockIdx.y;
    const int l = labels ? labels[n] : 0;
    const int64_t go = (int64_t)(n / n_per_group) * C + cg * 8, lo = (int64_t)l * C + cg * 8;
    float sc[8], sf[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        sc[e] = gamma[lo + e] * invstd[go + e];
        sf[e] = beta[lo + e] - mean[go + e] * sc[e];
    }
    const int h0 = blockIdx.x * rows_per_chunk, h1 = min(HW, h0 + rows_per_chunk);
    const int64_t base = (int64_t)n * HW * C + cg * 8;
    for (int h = h0 + rl; h < h1; h += rstep) {
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + base + (int64_t)h * C)), f);
        uint4 ov;
        __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float a = fmaf(f[2 * e], sc[2 * e], sf[2 * e]), b = fmaf(f[2 * e + 1], sc[2 * e + 1], sf[2 * e + 1]);
            if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
            op[e] = __floats2bfloat162_rn(a, b);
        }
        *reinterpret_cast<uint4*>(y + base + (int64_t)h * C) = ov;
    }
}

// ---- backward stage 1: per-(sample, hw split, channel) sums of dy and dy*xhat ---------
// ws[((n*S + s)*C + c)*2 + {0,1}]
template <typename T>
__global__ void __launch_bounds__(BN_LANES * BN_ROWS)
bn_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ y,
                     const float* __restrict__ mean, const float* __restrict__ invstd, float* __restrict__ ws,
                     int HW, int C, int S, int relu, int n_per_group) {
    ctgan::pdl_entry();
    const int lane = threadIdx.x, ry = threadIdx.y;
    const int c = blockIdx.y * BN_CCH + lane * 4;
    const int n = blockIdx.x / S, s = blockIdx.x % S;
    const int per = (HW + S - 1) / S;
    const int h0 = s * per, h1 = min(HW, h0 + per);
    float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    if (c < C) {
        float mu[4], is[4];
        const int64_t go = (int64_t)(n / n_per_group) * C;
#pragma unroll
        for (int j = 0; j < 4; ++j) { mu[j] = mean[go + c + j]; is[j] = invstd[go + c + j]; }
        for (int h = h0 + ry; h < h1; h += BN_ROWS) {
            int64_t off = ((int64_t)n * HW + h) * C + c;
            float g[4], v[4];
            load4<T>(dy + off, g);
            load4<T>(x + off, v);
            if (relu) {
                float o[4];
                load4<T>(y + off, o);
#pragma unroll
                for (int j = 0; j < 4; ++j) g[j] = o[j] > 0.f ? g[j] : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) { s1[j] += g[j]; s2[j] += g[j] * (v[j] - mu[j]) * is[j]; }
        }
    }
    __shared__ float sh1[BN_ROWS][BN_LANES][4];
    __shared__ float sh2[BN_ROWS][BN_LANES][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { sh1[ry][lane][j] = s1[j]; sh2[ry][lane][j] = s2[j]; }
    __syncthreads();
    if (ry == 0 && c < C) {
        for (int k = 1; k < BN_ROWS; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) { s1[j] += sh1[k][lane][j]; s2[j] += sh2[k][lane][j]; }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            ws[((int64_t)blockIdx.x * C + c + j) * 2 + 0] = s1[j];
            ws[((int64_t)blockIdx.x * C + c + j) * 2 + 1] = s2[j];
        }
    }
}

// BF16 fast path of stage 1 (same workspace layout): blockIdx.y = sample, blockIdx.x = hw split.
__global__ void __launch_bounds__(256)
bn_bwd_reduce_bf16_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                          const __nv_bfloat16* __restrict__ y, const float* __restrict__ mean,
                          const float* __restrict__ invstd, float* __restrict__ ws, int HW, int C, int S, int relu,
                          int n_per_group) {
    ctgan::pdl_entry();
    __shared__ float sh1[2048], sh2[2048];
    const int tpr = C >> 3;
    const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr, rstep = 256 / tpr;
    const int n = blockIdx.y, sp = blockIdx.x;
    const int per = (HW + S - 1) / S;
    const int h0 = sp * per, h1 = min(HW, h0 + per);
    const int64_t go = (int64_t)(n / n_per_group) * C + cg * 8;
    float mu[8], is[8], s1[8], s2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { mu[e] = mean[go + e]; is[e] = invstd[go + e]; s1[e] = 0.f; s2[e] = 0.f; }
    const int64_t base = (int64_t)n * HW * C + cg * 8;
    for (int h = h0 + rl; h < h1; h += rstep) {
        const int64_t off = base + (int64_t)h * C;
        float g[8], v[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(dy + off)), g);
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + off)), v);
        if (relu) {
            float o[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(y + off)), o);
#pragma unroll
            for (int e = 0; e < 8; ++e) g[e] = o[e] > 0.f ? g[e] : 0.f;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) { s1[e] += g[e]; s2[e] = fmaf(g[e], (v[e] - mu[e]) * is[e], s2[e]); }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) { sh1[rl * C + cg * 8 + e] = s1[e]; sh2[rl * C + cg * 8 + e] = s2[e]; }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float a = 0.f, b = 0.f;
        for (int k = 0; k < rstep; ++k) { a += sh1[k * C + c]; b += sh2[k * C + c]; }
        ws[(((int64_t)n * S + sp) * C + c) * 2 + 0] = a;
        ws[(((int64_t)n * S + sp) * C + c) * 2 + 1] = b;
    }
}

// BF16 fast path of stage 3: blockIdx.y = sample, blockIdx.x = chunk of its HW rows.
__global__ void __launch_bounds__(256)
bn_bwd_apply_bf16_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                         const __nv_bfloat16* __restrict__ y, const float* __restrict__ gamma,
                         const int32_t* __restrict__ labels, const float* __restrict__ mean,
                         const float* __restrict__ invstd, const float* __restrict__ coef, __nv_bfloat16* __restrict__ dx,
                         int HW, int C, int relu, int n_per_group, int groups, int rows_per_chunk) {
    ctgan::pdl_entry();
    const int tpr = C >> 3;
    const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr, rstep = 256 / tpr;
    const int n = blockIdx.y;
    const int l = labels ? labels[n] : 0;
    const int64_t go = (int64_t)(n / n_per_group) * C + cg * 8, lo = (int64_t)l * C + cg * 8;
    float mu[8], is[8], gm[8], c1[8], c2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        mu[e] = mean[go + e]; is[e] = invstd[go + e]; gm[e] = gamma[lo + e];
        c1[e] = coef[go + e]; c2[e] = coef[(int64_t)groups * C + go + e];
    }
    const int h0 = blockIdx.x * rows_per_chunk, h1 = min(HW, h0 + rows_per_chunk);
    const int64_t base = (int64_t)n * HW * C + cg * 8;
    for (int h = h0 + rl; h < h1; h += rstep) {
        const int64_t off = base + (int64_t)h * C;
        float g[8], v[8], o[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(dy + off)), g);
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + off)), v);
        if (relu) {
            float yo[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(y + off)), yo);
#pragma unroll
            for (int e = 0; e < 8; ++e) g[e] = yo[e] > 0.f ? g[e] : 0.f;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float xh = (v[e] - mu[e]) * is[e];
            o[e] = is[e] * (gm[e] * g[e] - c1[e] - xh * c2[e]);
        }
        uint4 ov;
        __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
        for (int e = 0; e < 4; ++e) op[e] = __floats2bfloat162_rn(o[2 * e], o[2 * e + 1]);
        *reinterpret_cast<uint4*>(dx + off) = ov;
    }
}

// ---- backward stage 2: per channel: table gradients + the two means BN needs ----------
// coef[g*C + c] = mean_Rg(gamma_l * dy), coef[G*C + g*C + c] = mean_Rg(gamma_l * dy * xhat).
// One warp per channel; the per-label table sums (shared by all groups) go through shared-memory atomics.
__global__ void __launch_bounds__(128)
bn_bwd_finalize_kernel(const float* __restrict__ ws, const float* __restrict__ gamma,
                       const int32_t* __restrict__ labels, float* __restrict__ dgamma,
                       float* __restrict__ dbeta, float* __restrict__ coef,
                       int N, int S, int C, int n_labels, float inv_Rg, int groups) {
    ctgan::pdl_entry();
    extern __shared__ float tab[];                    // [4 warps][2][n_labels]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 4 + warp;
    float* t1 = tab + (size_t)warp * 2 * n_labels;
    float* t2 = t1 + n_labels;
    for (int l = lane; l < 2 * n_labels; l += 32) t1[l] = 0.f;
    __syncwarp();
    if (c >= C) return;
    const int npg = N / groups;
    for (int g = 0; g < groups; ++g) {
        float a1 = 0.f, a2 = 0.f;
        for (int n = g * npg + lane; n < (g + 1) * npg; n += 32) {
            int l = labels ? labels[n] : 0;
            float s1 = 0.f, s2 = 0.f;
            for (int s = 0; s < S; ++s) {
                s1 += ws[(((int64_t)n * S + s) * C + c) * 2];
                s2 += ws[(((int64_t)n * S + s) * C + c) * 2 + 1];
            }
            float gm = gamma[(int64_t)l * C + c];
            a1 += gm * s1; a2 += gm * s2;
            atomicAdd(t1 + l, s1);
            atomicAdd(t2 + l, s2);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a1 += __shfl_xor_sync(0xffffffffu, a1, o);
            a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        }
        if (lane == 0) { coef[(int64_t)g * C + c] = a1 * inv_Rg; coef[(int64_t)(groups + g) * C + c] = a2 * inv_Rg; }
    }
    __syncwarp();
    for (int l = lane; l < n_labels; l += 32) {
        dbeta[(int64_t)l * C + c] = t1[l];
        dgamma[(int64_t)l * C + c] = t2[l];
    }
}

// ---- backward stage 3: dx = invstd * (gamma_l*dy - mean(.) - xhat*mean(. xhat)) ----------
template <typename T>
__global__ void bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ y,
                                    const float* __restrict__ gamma, const int32_t* __restrict__ labels,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ coef, T* __restrict__ dx,
                                    int64_t R, int HW, int C, int relu, int n_per_group, int groups) {
    ctgan::pdl_entry();
    const int C4 = C / 4;
    int64_t total = R * C4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % C4) * 4;
        int64_t r = i / C4;
        int n = (int)(r / HW);
        int l = labels ? labels[n] : 0;
        const int64_t go = (int64_t)(n / n_per_group) * C;
        float g[4], v[4], o[4];
        load4<T>(dy + r * C + c, g);
        load4<T>(x + r * C + c, v);
        if (relu) {
            float yo[4];
            load4<T>(y + r * C + c, yo);
#pragma unroll
            for (int j = 0; j < 4; ++j) g[j] = yo[j] > 0.f ? g[j] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float is = invstd[go + c + j];
            float xh = (v[j] - mean[go + c + j]) * is;
            o[j] = is * (gamma[(int64_t)l * C + c + j] * g[j] - coef[go + c + j] - xh * coef[(int64_t)groups * C + go + c + j]);
        }
        store4<T>(dx + r * C + c, o);
    }
}

// ---- BF16 "fused" path (ctgan_bn_fwd_fused / ctgan_bn_bwd_fused): two kernels per direction instead of three --------
// The generator step of the round-1 timeline spent 36 % of its kernel time in batch norm: stats / finalize / apply (and
// reduce / finalize / apply backward) are three dependent launches, the finalize kernels are latency chains of
// uncoalesced loads on 32 blocks, and one 16-byte load is in flight per thread.  Here
//   * the statistics kernel adds per-block sums with red.global into a [G][C][2] accumulator (zeroed by a memset node);
//     the sums are of d = x - shift, shift = the group's first row, so that E[d^2] - E[d]^2 does not cancel;
//   * the apply kernel derives mean / invstd from the accumulator itself (no finalize launch), keeps 4 rows in flight
//     per thread and can write its output nearest-neighbour upsampled 2x (CTGAN_BN_UP2: the tf.concat x4 +
//     depth_to_space of UpsampleConv, TG/CT_gan_cifar_resnet.py:100-107, fused into the normalisation that precedes it);
//   * backward: the ReLU pattern is recomputed from x ( y > 0  <=>  fma(x, scale, shift) > 0, the very expression the
//     forward evaluated ), so y is neither saved nor read; per-(sample, split) sums go straight into the per-label
//     dgamma / dbeta tables (CTGAN_BN_ACCUM: the flat gradient bucket) and the two per-group means with red.global;
//     with CTGAN_BN_UP2 the incoming gradient has the upsampled shape and is summed 2x2 on load (the adjoint).
__device__ __forceinline__ void bn_load8(const __nv_bfloat16* p, float (&f)[8]) {
    unpack8(__ldg(reinterpret_cast<const uint4*>(p)), f);
}

__global__ void __launch_bounds__(256)
bn_sums_bf16_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ acc, int64_t Rg, int C, int rows_per_block, int nbg) {
    ctgan::pdl_entry();
    __shared__ float sh1[2048], sh2[2048];
    const int tpr = C >> 3;
    const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr, rstep = 256 / tpr;
    const int g = blockIdx.x / nbg, rb = blockIdx.x - g * nbg;
    const int64_t r0 = (int64_t)g * Rg + (int64_t)rb * rows_per_block;
    const int64_t r1 = min((int64_t)(g + 1) * Rg, r0 + rows_per_block);
    const __nv_bfloat16* base = x + cg * 8;
    float shift[8], s[8], ss[8];
    bn_load8(base + (int64_t)g * Rg * C, shift);
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] = 0.f; ss[e] = 0.f; }
    // 8 rows in flight per thread in EVERY iteration (predicated tail: no single-load remainder loop)
    for (int64_t r = r0 + rl; r < r1; r += 8 * (int64_t)rstep) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (r + u * (int64_t)rstep < r1) v[u] = __ldg(reinterpret_cast<const uint4*>(base + (r + u * (int64_t)rstep) * C));
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (r + u * (int64_t)rstep < r1) {
                float f[8];
                unpack8(v[u], f);
#pragma unroll
                for (int e = 0; e < 8; ++e) { const float d = f[e] - shift[e]; s[e] += d; ss[e] = fmaf(d, d, ss[e]); }
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) { sh1[rl * C + cg * 8 + e] = s[e]; sh2[rl * C + cg * 8 + e] = ss[e]; }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float a = 0.f, b = 0.f;
        for (int k = 0; k < rstep; ++k) { a += sh1[k * C + c]; b += sh2[k * C + c]; }
        atomicAdd(acc + ((int64_t)g * C + c) * 2 + 0, a);
        atomicAdd(acc + ((int64_t)g * C + c) * 2 + 1, b);
    }
}

// blockIdx.y = sample, blockIdx.x = chunk of its H*W pixels.
template <int UP2>
__global__ void __launch_bounds__(256)
bn_apply_fused_bf16_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                           const int32_t* __restrict__ labels, const float* __restrict__ acc, float* __restrict__ save_mean,
                           float* __restrict__ save_invstd, __nv_bfloat16* __restrict__ y, int H, int W, int C, int relu,
                           int n_per_group, int rows_per_chunk, float inv_Rg, float eps) {
    ctgan::pdl_entry();
    const int tpr = C >> 3, HW = H * W;
    const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr, rstep = 256 / tpr;
    const int n = blockIdx.y, g = n / n_per_group;
    const int l = labels ? labels[n] : 0;
    const int64_t go = (int64_t)g * C + cg * 8, lo = (int64_t)l * C + cg * 8;
    float sc[8], sf[8];
    {
        float shift[8];
        bn_load8(x + (int64_t)g * n_per_group * HW * C + cg * 8, shift);
        const bool writer = blockIdx.x == 0 && n == g * n_per_group && rl == 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float m1 = acc[(go + e) * 2] * inv_Rg, m2 = acc[(go + e) * 2 + 1] * inv_Rg;
            const float mean = shift[e] + m1;
            const float is = rsqrtf(fmaxf(m2 - m1 * m1, 0.f) + eps);
            if (writer) { save_mean[go + e] = mean; save_invstd[go + e] = is; }
            sc[e] = gamma[lo + e] * is;
            sf[e] = beta[lo + e] - mean * sc[e];
        }
    }
    const int h0 = blockIdx.x * rows_per_chunk, h1 = min(HW, h0 + rows_per_chunk);
    const int64_t base = (int64_t)n * HW * C + cg * 8;
    const int64_t obase = (UP2 ? 4 : 1) * (int64_t)n * HW * C + cg * 8;
    for (int h = h0 + rl; h < h1; h += 4 * rstep) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (h + u * rstep < h1) v[u] = __ldg(reinterpret_cast<const uint4*>(x + base + (int64_t)(h + u * rstep) * C));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int hh = h + u * rstep;
            if (hh < h1) {
                float f[8];
                unpack8(v[u], f);
                uint4 ov;
                __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float a = fmaf(f[2 * e], sc[2 * e], sf[2 * e]), b = fmaf(f[2 * e + 1], sc[2 * e + 1], sf[2 * e + 1]);
                    if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                    op[e] = __floats2bfloat162_rn(a, b);
                }
                if (UP2) {
                    const int ph = hh / W, pw = hh - ph * W;
                    __nv_bfloat16* o = y + obase + ((int64_t)(2 * ph) * (2 * W) + 2 * pw) * C;
                    *reinterpret_cast<uint4*>(o) = ov;
                    *reinterpret_cast<uint4*>(o + C) = ov;
                    *reinterpret_cast<uint4*>(o + (int64_t)2 * W * C) = ov;
                    *reinterpret_cast<uint4*>(o + (int64_t)2 * W * C + C) = ov;
                } else {
                    *reinterpret_cast<uint4*>(y + obase + (int64_t)hh * C) = ov;
                }
            }
        }
    }
}

// the gradient that reaches the normalisation output at pixel hh of sample n: dy itself, or (UP2) the sum over the 2x2
// block of the upsampled gradient.  Held as raw 16-byte loads so that several rows are in flight before the first use.
template <int UP2> struct BnGradRaw;
template <> struct BnGradRaw<0> {
    uint4 a;
    __device__ __forceinline__ void load(const __nv_bfloat16* __restrict__ dy, int64_t n, int hh, int H, int W, int C, int cg) {
        a = __ldg(reinterpret_cast<const uint4*>(dy + (n * H * W + hh) * C + cg * 8));
    }
    __device__ __forceinline__ void get(float (&g)[8]) const { unpack8(a, g); }
};
template <> struct BnGradRaw<1> {
    uint4 a, b, c, d;
    __device__ __forceinline__ void load(const __nv_bfloat16* __restrict__ dy, int64_t n, int hh, int H, int W, int C, int cg) {
        const int ph = hh / W, pw = hh - ph * W;
        const __nv_bfloat16* p = dy + 4 * n * H * W * C + ((int64_t)(2 * ph) * (2 * W) + 2 * pw) * C + cg * 8;
        a = __ldg(reinterpret_cast<const uint4*>(p)); b = __ldg(reinterpret_cast<const uint4*>(p + C));
        c = __ldg(reinterpret_cast<const uint4*>(p + (int64_t)2 * W * C)); d = __ldg(reinterpret_cast<const uint4*>(p + (int64_t)2 * W * C + C));
    }
    __device__ __forceinline__ void get(float (&g)[8]) const {
        float fa[8], fb[8], fc[8], fd[8];
        unpack8(a, fa); unpack8(b, fb); unpack8(c, fc); unpack8(d, fd);
#pragma unroll
        for (int e = 0; e < 8; ++e) g[e] = (fa[e] + fb[e]) + (fc[e] + fd[e]);
    }
};

// blockIdx.y = sample, blockIdx.x = split of its H*W pixels
template <int UP2>
__global__ void __launch_bounds__(256)
bn_bwd_sums_bf16_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                        const float* __restrict__ beta, const int32_t* __restrict__ labels, const float* __restrict__ mean,
                        const float* __restrict__ invstd, float* __restrict__ dgamma, float* __restrict__ dbeta,
                        float* __restrict__ coef, int H, int W, int C, int S, int relu, int n_per_group) {
    ctgan::pdl_entry();
    __shared__ float sh1[2048], sh2[2048];
    const int tpr = C >> 3, HW = H * W;
    const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr, rstep = 256 / tpr;
    const int n = blockIdx.y, sp = blockIdx.x, g = n / n_per_group;
    const int l = labels ? labels[n] : 0;
    const int per = (HW + S - 1) / S;
    const int h0 = sp * per, h1 = min(HW, h0 + per);
    const int64_t go = (int64_t)g * C + cg * 8, lo = (int64_t)l * C + cg * 8;
    float mu[8], is[8], sc[8], sf[8], s1[8], s2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        mu[e] = mean[go + e]; is[e] = invstd[go + e];
        sc[e] = gamma[lo + e] * is[e];
        sf[e] = beta[lo + e] - mu[e] * sc[e];
        s1[e] = 0.f; s2[e] = 0.f;
    }
    const int64_t base = (int64_t)n * HW * C + cg * 8;
    constexpr int ROWS = UP2 ? 2 : 4;                    // rows in flight per thread (UP2: four gradient loads per row)
    for (int h = h0 + rl; h < h1; h += ROWS * rstep) {
        BnGradRaw<UP2> graw[ROWS];
        uint4 vraw[ROWS];
#pragma unroll
        for (int u = 0; u < ROWS; ++u) {
            const int hh = h + u * rstep;
            if (hh < h1) {
                graw[u].load(dy, n, hh, H, W, C, cg);
                vraw[u] = __ldg(reinterpret_cast<const uint4*>(x + base + (int64_t)hh * C));
            }
        }
#pragma unroll
        for (int u = 0; u < ROWS; ++u) {
            if (h + u * rstep < h1) {
                float gq[8], vq[8];
                graw[u].get(gq);
                unpack8(vraw[u], vq);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float gg = gq[e];
                    if (relu) gg = fmaf(vq[e], sc[e], sf[e]) > 0.f ? gg : 0.f;
                    s1[e] += gg;
                    s2[e] = fmaf(gg, (vq[e] - mu[e]) * is[e], s2[e]);
                }
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) { sh1[rl * C + cg * 8 + e] = s1[e]; sh2[rl * C + cg * 8 + e] = s2[e]; }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float a = 0.f, b = 0.f;
        for (int k = 0; k < rstep; ++k) { a += sh1[k * C + c]; b += sh2[k * C + c]; }
        const float gm = gamma[(int64_t)l * C + c];
        atomicAdd(dbeta + (int64_t)l * C + c, a);
        atomicAdd(dgamma + (int64_t)l * C + c, b);
        atomicAdd(coef + ((int64_t)g * C + c) * 2 + 0, gm * a);
        atomicAdd(coef + ((int64_t)g * C + c) * 2 + 1, gm * b);
    }
}

// blockIdx.y = sample (walked in REVERSE order of the sums kernel: its last reads are the hottest lines of the L2),
// blockIdx.x = chunk of its H*W pixels.
template <int UP2>
__global__ void __launch_bounds__(256)
bn_bwd_apply_fused_bf16_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                               const float* __restrict__ gamma, const float* __restrict__ beta,
                               const int32_t* __restrict__ labels, const float* __restrict__ mean,
                               const float* __restrict__ invstd, const float* __restrict__ coef,
                               __nv_bfloat16* __restrict__ dx, int H, int W, int C, int relu, int n_per_group,
                               int rows_per_chunk, float inv_Rg) {
    ctgan::pdl_entry();
    const int tpr = C >> 3, HW = H * W;
    const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr, rstep = 256 / tpr;
    const int n = gridDim.y - 1 - blockIdx.y, g = n / n_per_group;
    const int l = labels ? labels[n] : 0;
    const int64_t go = (int64_t)g * C + cg * 8, lo = (int64_t)l * C + cg * 8;
    float mu[8], is[8], gm[8], sc[8], sf[8], c1[8], c2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        mu[e] = mean[go + e]; is[e] = invstd[go + e]; gm[e] = gamma[lo + e];
        sc[e] = gm[e] * is[e];
        sf[e] = beta[lo + e] - mu[e] * sc[e];
        c1[e] = coef[(go + e) * 2] * inv_Rg; c2[e] = coef[(go + e) * 2 + 1] * inv_Rg;
    }
    const int chunk = gridDim.x - 1 - blockIdx.x;
    const int h0 = chunk * rows_per_chunk, h1 = min(HW, h0 + rows_per_chunk);
    const int64_t base = (int64_t)n * HW * C + cg * 8;
    constexpr int ROWS = UP2 ? 2 : 4;
    for (int h = h0 + rl; h < h1; h += ROWS * rstep) {
        BnGradRaw<UP2> graw[ROWS];
        uint4 vraw[ROWS];
#pragma unroll
        for (int u = 0; u < ROWS; ++u) {
            const int hh = h + u * rstep;
            if (hh < h1) {
                graw[u].load(dy, n, hh, H, W, C, cg);
                vraw[u] = __ldg(reinterpret_cast<const uint4*>(x + base + (int64_t)hh * C));
            }
        }
#pragma unroll
        for (int u = 0; u < ROWS; ++u) {
            const int hh = h + u * rstep;
            if (hh < h1) {
                float gq[8], vq[8], o[8];
                graw[u].get(gq);
                unpack8(vraw[u], vq);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float gg = gq[e];
                    if (relu) gg = fmaf(vq[e], sc[e], sf[e]) > 0.f ? gg : 0.f;
                    const float xh = (vq[e] - mu[e]) * is[e];
                    o[e] = is[e] * (gm[e] * gg - c1[e] - xh * c2[e]);
                }
                uint4 ov;
                __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
                for (int e = 0; e < 4; ++e) op[e] = __floats2bfloat162_rn(o[2 * e], o[2 * e + 1]);
                *reinterpret_cast<uint4*>(dx + base + (int64_t)hh * C) = ov;
            }
        }
    }
}

// ---- bias gradient: column sums -------------------------------------------------
__global__ void __launch_bounds__(256)
bias_grad_kernel(const void* __restrict__ dy, float* __restrict__ db, int64_t rows, int C, int dt, int rows_per_block) {
    ctgan::pdl_entry();
    // blockDim = (32 channel lanes, 8 row lanes); one channel per lane
    const int lane = threadIdx.x, ry = threadIdx.y;
    const int c = blockIdx.y * 32 + lane;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    float s = 0.f;
    if (c < C)
        for (int64_t r = r0 + ry; r < r1; r += 8) s += ld_act(dy, r * C + c, dt);
    __shared__ float sh[8][33];
    sh[ry][lane] = s;
    __syncthreads();
    if (ry == 0 && c < C) {
        for (int k = 1; k < 8; ++k) s += sh[k][lane];
        atomicAdd(db + c, s);
    }
}

// BF16, C/8 a power of two <= 256: 16-byte loads, C/8 threads per row, 4 rows in flight per thread
__global__ void __launch_bounds__(256)
bias_grad_bf16_kernel(const __nv_bfloat16* __restrict__ dy, float* __restrict__ db, int64_t rows, int C, int rows_per_block) {
    ctgan::pdl_entry();
    __shared__ float sh[256 * 8];
    const int tpr = C >> 3;                                  // threads per row
    const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr, rstep = 256 / tpr;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    const __nv_bfloat16* base = dy + cg * 8;
    int64_t r = r0 + rl;
    for (; r + 3 * (int64_t)rstep < r1; r += 4 * (int64_t)rstep) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(base + (r + u * (int64_t)rstep) * C));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[u]);
#pragma unroll
            for (int e = 0; e < 4; ++e) { acc[2 * e] += __bfloat162float(h[e].x); acc[2 * e + 1] += __bfloat162float(h[e].y); }
        }
    }
    for (; r < r1; r += rstep) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + r * C));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int e = 0; e < 4; ++e) { acc[2 * e] += __bfloat162float(h[e].x); acc[2 * e + 1] += __bfloat162float(h[e].y); }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) sh[rl * C + cg * 8 + e] = acc[e];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float s = 0.f;
        for (int k = 0; k < rstep; ++k) s += sh[k * C + c];
        atomicAdd(db + c, s);
    }
}

}  // namespace ctgan

using namespace ctgan;

extern "C" int64_t ctgan_bn_workspace_floats(int N, int HW, int C, int groups) {
    if (groups < 1) groups = 1;
    int64_t Rg = (int64_t)(N / groups) * HW;
    int64_t fwd = (int64_t)bn_fwd_rowblocks(Rg) * groups * C * 2;
    int64_t bwd = (int64_t)N * bn_bwd_splits(HW) * C * 2 + 2 * (int64_t)groups * C;
    return fwd > bwd ? fwd : bwd;
}

extern "C" int ctgan_bn_fwd(const void* x, const float* gamma, const float* beta, const int32_t* labels,
                            void* y, float* save_mean, float* save_invstd, float* ws,
                            int N, int HW, int C, float eps, int relu, int groups, int dtype, void* stream) {
    CTGAN_REQUIRE(N > 0 && HW > 0 && C > 0 && C % 4 == 0, CTGAN_ERR_UNSUPPORTED, "bn_fwd: C must be a positive multiple of 4");
    CTGAN_REQUIRE(groups >= 1 && N % groups == 0, CTGAN_ERR_BAD_DESC, "bn_fwd: groups must divide N");
    CTGAN_REQUIRE(x && gamma && beta && y && save_mean && save_invstd && ws, CTGAN_ERR_BAD_DESC, "bn_fwd: null pointer");
    CTGAN_REQUIRE(dtype_ok(dtype), CTGAN_ERR_BAD_DESC, "bn_fwd: bad dtype");
    cudaStream_t st = as_stream(stream);
    const int64_t R = (int64_t)N * HW, Rg = R / groups;
    int nbg = bn_fwd_rowblocks(Rg);
    int rpb = (int)((Rg + nbg - 1) / nbg);
    nbg = (int)((Rg + rpb - 1) / rpb);
    const int tpr = C / 8;
    const bool fast = dtype == CTGAN_BF16 && C % 8 == 0 && tpr <= 256 && (tpr & (tpr - 1)) == 0 &&
                      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    dim3 blk(BN_LANES, BN_ROWS), grid(nbg * groups, ceil_div(C, BN_CCH));
    if (fast) CTGAN_LAUNCH((bn_stats_bf16_kernel), nbg * groups, 256, 0, st, (const __nv_bfloat16*)x, ws, Rg, C, rpb, nbg);
    else if (dtype == CTGAN_F32) CTGAN_LAUNCH((bn_stats_kernel<float>), grid, blk, 0, st, (const float*)x, ws, Rg, C, rpb, nbg);
    else CTGAN_LAUNCH((bn_stats_kernel<__nv_bfloat16>), grid, blk, 0, st, (const __nv_bfloat16*)x, ws, Rg, C, rpb, nbg);
    CTGAN_CHECK_LAUNCH("bn_stats");
    CTGAN_LAUNCH((bn_finalize_kernel), dim3(ceil_div(C, 4), groups), 128, 0, st, ws, save_mean, save_invstd, Rg, C, nbg, rpb, eps);
    CTGAN_CHECK_LAUNCH("bn_finalize");
    if (fast && N <= 65535) {
        const int rstep = 256 / tpr;
        int rpc = 4 * rstep;                                   // >= 4 rows per thread, more when there are plenty of blocks
        while ((int64_t)N * ceil_div(HW, rpc) > 16 * (int64_t)sm_count() && rpc < HW) rpc *= 2;
        CTGAN_LAUNCH((bn_apply_bf16_kernel), dim3(ceil_div(HW, rpc), N), 256, 0, st, (const __nv_bfloat16*)x, gamma, beta, labels, save_mean,
                                                                        save_invstd, (__nv_bfloat16*)y, HW, C, relu, N / groups, rpc);
        CTGAN_CHECK_LAUNCH("bn_apply");
        return 0;
    }
    int g2 = elementwise_grid(R * (C / 4), 256);
    if (dtype == CTGAN_F32)
        CTGAN_LAUNCH((bn_apply_kernel<float>), g2, 256, 0, st, (const float*)x, gamma, beta, labels, save_mean, save_invstd, (float*)y, R, HW, C, relu, N / groups);
    else
        CTGAN_LAUNCH((bn_apply_kernel<__nv_bfloat16>), g2, 256, 0, st, (const __nv_bfloat16*)x, gamma, beta, labels, save_mean, save_invstd, (__nv_bfloat16*)y, R, HW, C, relu, N / groups);
    CTGAN_CHECK_LAUNCH("bn_apply");
    return 0;
}

extern "C" int ctgan_bn_bwd(const void* dy, const void* x, const void* y, const float* gamma,
                            const int32_t* labels, const float* save_mean, const float* save_invstd,
                            void* dx, float* dgamma, float* dbeta, float* ws,
                            int N, int HW, int C, int n_labels, int relu, int groups, int dtype, void* stream) {
    CTGAN_REQUIRE(N > 0 && HW > 0 && C > 0 && C % 4 == 0 && n_labels > 0, CTGAN_ERR_UNSUPPORTED, "bn_bwd: C must be a positive multiple of 4");
    CTGAN_REQUIRE(groups >= 1 && N % groups == 0, CTGAN_ERR_BAD_DESC, "bn_bwd: groups must divide N");
    CTGAN_REQUIRE(dy && x && gamma && save_mean && save_invstd && dx && dgamma && dbeta && ws && (!relu || y),
                  CTGAN_ERR_BAD_DESC, "bn_bwd: null pointer");
    CTGAN_REQUIRE(dtype_ok(dtype), CTGAN_ERR_BAD_DESC, "bn_bwd: bad dtype");
    cudaStream_t st = as_stream(stream);
    const int64_t R = (int64_t)N * HW, Rg = R / groups;
    int S = bn_bwd_splits(HW);
    float* coef = ws + (int64_t)N * S * C * 2;
    const int tpr = C / 8;
    const bool fast = dtype == CTGAN_BF16 && C % 8 == 0 && tpr <= 256 && (tpr & (tpr - 1)) == 0 && N <= 65535 &&
                      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx) |
                        reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    dim3 blk(BN_LANES, BN_ROWS), grid(N * S, ceil_div(C, BN_CCH));
    if (fast)
        CTGAN_LAUNCH((bn_bwd_reduce_bf16_kernel), dim3(S, N), 256, 0, st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (const __nv_bfloat16*)y,
                                                             save_mean, save_invstd, ws, HW, C, S, relu, N / groups);
    else if (dtype == CTGAN_F32)
        CTGAN_LAUNCH((bn_bwd_reduce_kernel<float>), grid, blk, 0, st, (const float*)dy, (const float*)x, (const float*)y, save_mean, save_invstd, ws, HW, C, S, relu, N / groups);
    else
        CTGAN_LAUNCH((bn_bwd_reduce_kernel<__nv_bfloat16>), grid, blk, 0, st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (const __nv_bfloat16*)y, save_mean, save_invstd, ws, HW, C, S, relu, N / groups);
    CTGAN_CHECK_LAUNCH("bn_bwd_reduce");
    CTGAN_LAUNCH((bn_bwd_finalize_kernel), ceil_div(C, 4), 128, sizeof(float) * 8 * n_labels, st, ws, gamma, labels, dgamma, dbeta, coef, N, S, C, n_labels, 1.f / (float)Rg, groups);
    CTGAN_CHECK_LAUNCH("bn_bwd_finalize");
    if (fast) {
        const int rstep = 256 / tpr;
        int rpc = 4 * rstep;
        while ((int64_t)N * ceil_div(HW, rpc) > 16 * (int64_t)sm_count() && rpc < HW) rpc *= 2;
        CTGAN_LAUNCH((bn_bwd_apply_bf16_kernel), dim3(ceil_div(HW, rpc), N), 256, 0, st, 
            (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (const __nv_bfloat16*)y, gamma, labels, save_mean, save_invstd, coef,
            (__nv_bfloat16*)dx, HW, C, relu, N / groups, groups, rpc);
        CTGAN_CHECK_LAUNCH("bn_bwd_apply");
        return 0;
    }
    int g2 = elementwise_grid(R * (C / 4), 256);
    if (dtype == CTGAN_F32)
        CTGAN_LAUNCH((bn_bwd_apply_kernel<float>), g2, 256, 0, st, (const float*)dy, (const float*)x, (const float*)y, gamma, labels, save_mean, save_invstd, coef, (float*)dx, R, HW, C, relu, N / groups, groups);
    else
        CTGAN_LAUNCH((bn_bwd_apply_kernel<__nv_bfloat16>), g2, 256, 0, st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (const __nv_bfloat16*)y, gamma, labels, save_mean, save_invstd, coef, (__nv_bfloat16*)dx, R, HW, C, relu, N / groups, groups);
    CTGAN_CHECK_LAUNCH("bn_bwd_apply");
    return 0;
}

static bool bn_fused_ok(int N, int H, int W, int C, int groups, int dtype) {
    const int tpr = C / 8;
    return dtype == CTGAN_BF16 && N > 0 && N <= 65535 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && tpr <= 256 &&
           (tpr & (tpr - 1)) == 0 && groups >= 1 && N % groups == 0;
}
static int bn_rows_per_chunk(int N, int HW, int tpr, int rows_in_flight) {
    // ~4 blocks per SM: every block first derives its 8 channels' scale / shift from the accumulators (a chain of dependent
    // loads); with one 4-row batch per block (the first version: 2048 blocks for 128x32x32) that prologue was the kernel
    const int rstep = 256 / tpr;
    int rpc = rows_in_flight * rstep;
    while ((int64_t)N * ceil_div(HW, rpc) > 4 * (int64_t)sm_count() && rpc < HW) rpc *= 2;
    return rpc;
}

extern "C" int ctgan_bn_fused_ok(int N, int H, int W, int C, int groups, int dtype) { return bn_fused_ok(N, H, W, C, groups, dtype) ? 1 : 0; }

extern "C" int ctgan_bn_fwd_fused(const void* x, const float* gamma, const float* beta, const int32_t* labels, void* y,
                                  float* save_mean, float* save_invstd, float* ws, int N, int H, int W, int C, float eps,
                                  int flags, int groups, void* stream) {
    CTGAN_REQUIRE(bn_fused_ok(N, H, W, C, groups, CTGAN_BF16), CTGAN_ERR_UNSUPPORTED, "bn_fwd_fused: needs BF16, C/8 a power of two <= 256, groups | N");
    CTGAN_REQUIRE(x && gamma && beta && y && save_mean && save_invstd && ws, CTGAN_ERR_BAD_DESC, "bn_fwd_fused: null pointer");
    CTGAN_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0, CTGAN_ERR_BAD_DESC, "bn_fwd_fused: x, y must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    const int HW = H * W, tpr = C / 8, rstep = 256 / tpr;
    const int64_t Rg = (int64_t)(N / groups) * HW;
    cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(float) * 2 * (size_t)groups * C, st);
    if (e != cudaSuccess) return cuda_status(e, "bn_fwd_fused memset");
    // ~4 blocks per SM, at least 8 row steps each (8 loads in flight per thread)
    int64_t nbg = (4 * (int64_t)sm_count() + groups - 1) / groups;
    const int64_t max_nbg = (Rg + 8 * rstep - 1) / (8 * rstep);
    if (nbg > max_nbg) nbg = max_nbg;
    if (nbg < 1) nbg = 1;
    const int rpb = (int)((Rg + nbg - 1) / nbg);
    nbg = (Rg + rpb - 1) / rpb;
    CTGAN_LAUNCH((bn_sums_bf16_kernel), (unsigned)(nbg * groups), 256, 0, st, (const __nv_bfloat16*)x, ws, Rg, C, rpb, (int)nbg);
    CTGAN_CHECK_LAUNCH("bn_sums");
    const int rpc = bn_rows_per_chunk(N, HW, tpr, 4);
    const int relu = (flags & CTGAN_BN_RELU) ? 1 : 0;
    if (flags & CTGAN_BN_UP2)
        CTGAN_LAUNCH((bn_apply_fused_bf16_kernel<1>), dim3(ceil_div(HW, rpc), N), 256, 0, st, (const __nv_bfloat16*)x, gamma, beta, labels,
                     (const float*)ws, save_mean, save_invstd, (__nv_bfloat16*)y, H, W, C, relu, N / groups, rpc, 1.f / (float)Rg, eps);
    else
        CTGAN_LAUNCH((bn_apply_fused_bf16_kernel<0>), dim3(ceil_div(HW, rpc), N), 256, 0, st, (const __nv_bfloat16*)x, gamma, beta, labels,
                     (const float*)ws, save_mean, save_invstd, (__nv_bfloat16*)y, H, W, C, relu, N / groups, rpc, 1.f / (float)Rg, eps);
    CTGAN_CHECK_LAUNCH("bn_apply_fused");
    return 0;
}

extern "C" int ctgan_bn_bwd_fused(const void* dy, const void* x, const float* gamma, const float* beta, const int32_t* labels,
                                  const float* save_mean, const float* save_invstd, void* dx, float* dgamma, float* dbeta,
                                  float* ws, int N, int H, int W, int C, int n_labels, int flags, int groups, void* stream) {
    CTGAN_REQUIRE(bn_fused_ok(N, H, W, C, groups, CTGAN_BF16) && n_labels > 0, CTGAN_ERR_UNSUPPORTED, "bn_bwd_fused: needs BF16, C/8 a power of two <= 256, groups | N");
    CTGAN_REQUIRE(dy && x && gamma && beta && save_mean && save_invstd && dx && dgamma && dbeta && ws, CTGAN_ERR_BAD_DESC, "bn_bwd_fused: null pointer");
    CTGAN_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0,
                  CTGAN_ERR_BAD_DESC, "bn_bwd_fused: dy, x, dx must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    const int HW = H * W, tpr = C / 8, rstep = 256 / tpr;
    const int64_t Rg = (int64_t)(N / groups) * HW;
    cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(float) * 2 * (size_t)groups * C, st);
    if (e == cudaSuccess && !(flags & CTGAN_BN_ACCUM)) {
        e = cudaMemsetAsync(dgamma, 0, sizeof(float) * (size_t)n_labels * C, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(dbeta, 0, sizeof(float) * (size_t)n_labels * C, st);
    }
    if (e != cudaSuccess) return cuda_status(e, "bn_bwd_fused memset");
    // splits of a sample's pixels: ~4 blocks per SM, at least 4 row steps per block
    int S = (int)((4 * (int64_t)sm_count() + N - 1) / N);
    const int maxS = ceil_div(HW, 4 * rstep);
    if (S > maxS) S = maxS;
    if (S < 1) S = 1;
    const int relu = (flags & CTGAN_BN_RELU) ? 1 : 0;
    const int rpc = bn_rows_per_chunk(N, HW, tpr, 4);
    const float inv_Rg = 1.f / (float)Rg;
    if (flags & CTGAN_BN_UP2) {
        CTGAN_LAUNCH((bn_bwd_sums_bf16_kernel<1>), dim3(S, N), 256, 0, st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, gamma, beta, labels,
                     save_mean, save_invstd, dgamma, dbeta, ws, H, W, C, S, relu, N / groups);
        CTGAN_CHECK_LAUNCH("bn_bwd_sums");
        CTGAN_LAUNCH((bn_bwd_apply_fused_bf16_kernel<1>), dim3(ceil_div(HW, rpc), N), 256, 0, st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x,
                     gamma, beta, labels, save_mean, save_invstd, (const float*)ws, (__nv_bfloat16*)dx, H, W, C, relu, N / groups, rpc, inv_Rg);
    } else {
        CTGAN_LAUNCH((bn_bwd_sums_bf16_kernel<0>), dim3(S, N), 256, 0, st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, gamma, beta, labels,
                     save_mean, save_invstd, dgamma, dbeta, ws, H, W, C, S, relu, N / groups);
        CTGAN_CHECK_LAUNCH("bn_bwd_sums");
        CTGAN_LAUNCH((bn_bwd_apply_fused_bf16_kernel<0>), dim3(ceil_div(HW, rpc), N), 256, 0, st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x,
                     gamma, beta, labels, save_mean, save_invstd, (const float*)ws, (__nv_bfloat16*)dx, H, W, C, relu, N / groups, rpc, inv_Rg);
    }
    CTGAN_CHECK_LAUNCH("bn_bwd_apply_fused");
    return 0;
}

extern "C" int ctgan_bias_grad(const void* dy, float* db, int64_t rows, int C, int dtype, int accumulate, void* stream) {
    CTGAN_REQUIRE(rows > 0 && C > 0 && dtype_ok(dtype) && dy && db, CTGAN_ERR_BAD_DESC, "bias_grad: bad args");
    cudaStream_t st = as_stream(stream);
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(db, 0, sizeof(float) * (size_t)C, st);
        if (e != cudaSuccess) return cuda_status(e, "bias_grad memset");
    }
    const int tpr = C / 8;
    if (dtype == CTGAN_BF16 && C % 8 == 0 && tpr <= 256 && (tpr & (tpr - 1)) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0) {
        const int rstep = 256 / tpr;
        int64_t nb = (rows + 4 * rstep - 1) / (4 * rstep);              // >= 4 rows per thread
        const int64_t cap = 4 * (int64_t)sm_count();
        if (nb > cap) nb = cap;
        int rpb = (int)((rows + nb - 1) / nb);
        nb = (rows + rpb - 1) / rpb;
        CTGAN_LAUNCH((bias_grad_bf16_kernel), (unsigned)nb, 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(dy), db, rows, C, rpb);
        CTGAN_CHECK_LAUNCH("bias_grad");
        return 0;
    }
    int cblocks = ceil_div(C, 32);
    int64_t want = (2 * (int64_t)sm_count() + cblocks - 1) / cblocks;
    int64_t maxb = (rows + 63) / 64;
    int64_t nb = want < maxb ? want : maxb;
    if (nb < 1) nb = 1;
    int rpb = (int)((rows + nb - 1) / nb);
    nb = (rows + rpb - 1) / rpb;
    dim3 blk(32, 8), grid((unsigned)nb, cblocks);
    CTGAN_LAUNCH((bias_grad_kernel), grid, blk, 0, st, dy, db, rows, C, dtype, rpb);
    CTGAN_CHECK_LAUNCH("bias_grad");
    return 0;
}
