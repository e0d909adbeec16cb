// layernorm.cu -- layer normalisation over (C,H,W) per sample with per-channel scale / offset: the critic's Normalize
// of TG/CT_gan_64x64.py:87-93 (op: TG/tflib/ops/layernorm.py:6-21, eps 1e-5).  SURVEY.md 8(f) row N4; validated on a B200 by
// tests/test_64x64_gpu.py (kernel family against the PyTorch-CPU definitions, and the CT_gan_64x64.py step against the oracle).
//
// Unlike every other critic op, layer norm is not piecewise linear, so the gradient penalty (a derivative of a derivative)
// needs its genuine second-order terms.  With, per sample, M = C*H*W, xh = (x - mean) * r, r = rsqrt(var + eps),
// mean_s(.) the mean over the sample's M elements and
//     core(u) = r * (u - mean_s(u) - xh * mean_s(u * xh))                      (symmetric: <c, core(u)> = <core(c), u>)
// the family is
//     forward            y   = xh * gamma_c + beta_c
//     backward           dx  = core(gamma * gy),   dgamma_c = sum gy * xh,   dbeta_c = sum gy
//     backward of dx     ggy = gamma * core(c),    ggamma_c = sum gy * core(c),
//                        gx  = -r^2 * [ xh * Q + B * (a - mean a) + A * (b - mean b) - 2 * xh * A * B ]
//                              a = c, b = gamma * gy, A = mean_s(a*xh), B = mean_s(b*xh), Q = mean_s(a*b) - mean a * mean b - A*B
// (derivation and its check against autograd: tests/test_layernorm_host.py).  All kernels are HBM-bound streams over the
// activation; per-sample sums are produced by P partial blocks per sample and folded by the consumer.
#include "common.cuh"

namespace ctgan {
namespace {

constexpr int LN_THREADS = 256;
constexpr int LN_MAX_P = 32;            // partial blocks per sample

template <int K>
__device__ __forceinline__ void block_sum(float (&v)[K], float* out /* K floats, global */) {
    __shared__ float sm[K][LN_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float s = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sm[k][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < LN_THREADS / 32; ++w) s += sm[threadIdx.x][w];
        out[threadIdx.x] = s;
    }
}

// sums of the P partials of sample n (K values each), by every thread (P <= 32: a short loop over L2-resident floats)
template <int K>
__device__ __forceinline__ void fold(const float* __restrict__ ws, int n, int P, float (&s)[K]) {
#pragma unroll
    for (int k = 0; k < K; ++k) s[k] = 0.f;
    for (int p = 0; p < P; ++p) {
#pragma unroll
        for (int k = 0; k < K; ++k) s[k] += ws[((int64_t)n * P + p) * K + k];
    }
}

// Vector width V: 1 (any dtype / shape) or 8 (BF16, C % 8 == 0, M % 8 == 0, 16-byte aligned tensors: one 16-byte access per
// tensor and thread -- round 2; the scalar 2-byte accesses of the first version ran at a fraction of the HBM rate).
template <int V> __device__ __forceinline__ void ld_vec(const void* p, int64_t i, int dt, float (&o)[V]);
template <> __device__ __forceinline__ void ld_vec<1>(const void* p, int64_t i, int dt, float (&o)[1]) { o[0] = ld_act(p, i, dt); }
template <> __device__ __forceinline__ void ld_vec<8>(const void* p, int64_t i, int, float (&o)[8]) {
    const uint4 q = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p) + i);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[e]);
        o[2 * e] = __bfloat162float(h.x); o[2 * e + 1] = __bfloat162float(h.y);
    }
}
template <int V> __device__ __forceinline__ void st_vec(void* p, int64_t i, int dt, const float (&o)[V]);
template <> __device__ __forceinline__ void st_vec<1>(void* p, int64_t i, int dt, const float (&o)[1]) { st_act(p, i, dt, o[0]); }
template <> __device__ __forceinline__ void st_vec<8>(void* p, int64_t i, int, const float (&o)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(o[2 * e], o[2 * e + 1]);
        w[e] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p) + i) = make_uint4(w[0], w[1], w[2], w[3]);
}
// V consecutive per-channel parameters starting at channel c (c % V == 0)
template <int V> __device__ __forceinline__ void ld_par(const float* __restrict__ g, int c, float (&o)[V]) {
    if (V == 8) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(g + c)), b = __ldg(reinterpret_cast<const float4*>(g + c + 4));
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[V > 4 ? 4 : 0] = b.x; o[V > 5 ? 5 : 0] = b.y; o[V > 6 ? 6 : 0] = b.z; o[V > 7 ? 7 : 0] = b.w;
    } else {
        o[0] = __ldg(g + c);
    }
}

template <int V>
__device__ __forceinline__ void chunk_range(int64_t M, int P, int p, int64_t* lo, int64_t* hi) {
    int64_t per = (M + P - 1) / P;
    per = (per + V - 1) / V * V;
    *lo = min(M, per * p);
    *hi = min(M, *lo + per);
}

// ---- statistics: shifted sums (shift = the sample's first element) so that var = E[d^2] - E[d]^2 does not cancel
template <int V>
__global__ void __launch_bounds__(LN_THREADS)
ln_stats_partial_kernel(const void* __restrict__ x, float* __restrict__ ws, int64_t M, int P, int dt) {
    pdl_entry();
    const int n = blockIdx.y, p = blockIdx.x;
    const int64_t base = (int64_t)n * M;
    const float shift = ld_act(x, base, dt);
    int64_t lo, hi;
    chunk_range<V>(M, P, p, &lo, &hi);
    float v[2] = {0.f, 0.f};
    for (int64_t i = lo + (int64_t)threadIdx.x * V; i < hi; i += LN_THREADS * V) {
        float xv[V];
        ld_vec<V>(x, base + i, dt, xv);
#pragma unroll
        for (int e = 0; e < V; ++e) { const float d = xv[e] - shift; v[0] += d; v[1] += d * d; }
    }
    block_sum<2>(v, ws + ((int64_t)n * P + p) * 2);
}

__global__ void ln_stats_finalize_kernel(const void* __restrict__ x, const float* __restrict__ ws, float* __restrict__ mean,
                                         float* __restrict__ rstd, int N, int64_t M, int P, float eps, int dt) {
    pdl_entry();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float s[2];
    fold<2>(ws, n, P, s);
    const float shift = ld_act(x, (int64_t)n * M, dt);
    const float m1 = s[0] / (float)M, m2 = s[1] / (float)M;
    mean[n] = shift + m1;
    rstd[n] = rsqrtf(fmaxf(m2 - m1 * m1, 0.f) + eps);
}

// ---- forward apply
template <int V>
__global__ void __launch_bounds__(LN_THREADS)
ln_apply_kernel(const void* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ mean, const float* __restrict__ rstd, void* __restrict__ y, int64_t total, int64_t M,
                int C, int dt) {
    pdl_entry();
    for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * V; i < total; i += (int64_t)gridDim.x * blockDim.x * V) {
        const int n = (int)(i / M), c = (int)(i % C);
        const float mu = mean[n], r = rstd[n];
        float xv[V], g[V], b[V], o[V];
        ld_vec<V>(x, i, dt, xv); ld_par<V>(gamma, c, g); ld_par<V>(beta, c, b);
#pragma unroll
        for (int e = 0; e < V; ++e) o[e] = (xv[e] - mu) * r * g[e] + b[e];
        st_vec<V>(y, i, dt, o);
    }
}

// ---- core(u): partial sums of u and u*xh, then the apply
template <int V>
__global__ void __launch_bounds__(LN_THREADS)
ln_core_partial_kernel(const void* __restrict__ v, const void* __restrict__ x, const float* __restrict__ gamma,
                       const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ ws, int64_t M, int C,
                       int P, int pre_scale, int dt) {
    pdl_entry();
    const int n = blockIdx.y, p = blockIdx.x;
    const int64_t base = (int64_t)n * M;
    const float mu = mean[n], r = rstd[n];
    int64_t lo, hi;
    chunk_range<V>(M, P, p, &lo, &hi);
    float s[2] = {0.f, 0.f};
    for (int64_t i = lo + (int64_t)threadIdx.x * V; i < hi; i += LN_THREADS * V) {
        float u[V], xv[V], g[V];
        ld_vec<V>(v, base + i, dt, u); ld_vec<V>(x, base + i, dt, xv);
        if (pre_scale) ld_par<V>(gamma, (int)((base + i) % C), g);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const float ue = pre_scale ? u[e] * g[e] : u[e];
            s[0] += ue; s[1] += ue * (xv[e] - mu) * r;
        }
    }
    block_sum<2>(s, ws + ((int64_t)n * P + p) * 2);
}

template <int V>
__global__ void __launch_bounds__(LN_THREADS)
ln_core_apply_kernel(const void* __restrict__ v, const void* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ ws,
                     void* __restrict__ out, int64_t M, int C, int P, int pre_scale, int post_scale, int dt) {
    pdl_entry();
    const int n = blockIdx.y;
    const int64_t base = (int64_t)n * M;
    float s[2];
    fold<2>(ws, n, P, s);
    const float m1 = s[0] / (float)M, m2 = s[1] / (float)M, mu = mean[n], r = rstd[n];
    for (int64_t i = (blockIdx.x * (int64_t)LN_THREADS + threadIdx.x) * V; i < M; i += (int64_t)gridDim.x * LN_THREADS * V) {
        float u[V], xv[V], g[V], o[V];
        ld_vec<V>(v, base + i, dt, u); ld_vec<V>(x, base + i, dt, xv); ld_par<V>(gamma, (int)((base + i) % C), g);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const float ue = pre_scale ? u[e] * g[e] : u[e];
            const float xh = (xv[e] - mu) * r;
            const float oe = r * (ue - m1 - xh * m2);
            o[e] = post_scale ? oe * g[e] : oe;
        }
        st_vec<V>(out, base + i, dt, o);
    }
}

// ---- dgamma_c += sum v * xh,  dbeta_c += sum v   (rows = N*H*W pixels; thread = channel, block = a slab of rows)
__global__ void __launch_bounds__(LN_THREADS)
ln_param_grad_kernel(const void* __restrict__ v, const void* __restrict__ x, const float* __restrict__ mean,
                     const float* __restrict__ rstd, float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows,
                     int64_t rows_per_sample, int C, int rows_per_block, int dt) {
    pdl_entry();
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    for (int c = threadIdx.x; c < C; c += LN_THREADS) {
        float sg = 0.f, sb = 0.f;
        for (int64_t row = r0; row < r1; ++row) {
            const int n = (int)(row / rows_per_sample);
            const float val = ld_act(v, row * C + c, dt);
            sb += val;
            sg += val * (ld_act(x, row * C + c, dt) - mean[n]) * rstd[n];
        }
        atomicAdd(dgamma + c, sg);
        if (dbeta) atomicAdd(dbeta + c, sb);
    }
}

// the same with 16-byte accesses (BF16): thread = 8 channels of one row, 256 / (C/8) rows per pass, reduced through shared memory
__global__ void __launch_bounds__(LN_THREADS)
ln_param_grad_vec8_kernel(const __nv_bfloat16* __restrict__ v, const __nv_bfloat16* __restrict__ x, const float* __restrict__ mean,
                          const float* __restrict__ rstd, float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows,
                          int64_t rows_per_sample, int C, int rows_per_block) {
    pdl_entry();
    __shared__ float red[LN_THREADS][17];
    const int tpr = C >> 3, cg = threadIdx.x % tpr, rsub = threadIdx.x / tpr, rstep = LN_THREADS / tpr;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    float sg[8], sb[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) sg[e] = sb[e] = 0.f;
    for (int64_t row = r0 + rsub; row < r1; row += rstep) {
        const int n = (int)(row / rows_per_sample);
        const float mu = mean[n], r = rstd[n];
        float vv[8], xv[8];
        ld_vec<8>(v, row * C + cg * 8, CTGAN_BF16, vv); ld_vec<8>(x, row * C + cg * 8, CTGAN_BF16, xv);
#pragma unroll
        for (int e = 0; e < 8; ++e) { sb[e] += vv[e]; sg[e] += vv[e] * (xv[e] - mu) * r; }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) { red[threadIdx.x][e] = sg[e]; red[threadIdx.x][8 + e] = sb[e]; }
    __syncthreads();
    // thread t < 2*C: column t % C of (dgamma | dbeta), summed over the rstep row groups
    for (int t = threadIdx.x; t < 2 * C; t += LN_THREADS) {
        const int which = t / C, c = t % C;
        float s = 0.f;
        for (int g = 0; g < rstep; ++g) s += red[g * tpr + (c >> 3)][which * 8 + (c & 7)];
        if (which == 0) atomicAdd(dgamma + c, s);
        else if (dbeta) atomicAdd(dbeta + c, s);
    }
}

// ---- the x-derivative of the backward: five per-sample sums, then the element-wise formula
template <int V>
__global__ void __launch_bounds__(LN_THREADS)
ln_bwd2_partial_kernel(const void* __restrict__ cin, const void* __restrict__ gy, const void* __restrict__ x,
                       const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                       float* __restrict__ ws, int64_t M, int C, int P, int dt) {
    pdl_entry();
    const int n = blockIdx.y, p = blockIdx.x;
    const int64_t base = (int64_t)n * M;
    const float mu = mean[n], r = rstd[n];
    int64_t lo, hi;
    chunk_range<V>(M, P, p, &lo, &hi);
    float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int64_t i = lo + (int64_t)threadIdx.x * V; i < hi; i += LN_THREADS * V) {
        float av[V], gv[V], xv[V], g[V];
        ld_vec<V>(cin, base + i, dt, av); ld_vec<V>(gy, base + i, dt, gv); ld_vec<V>(x, base + i, dt, xv);
        ld_par<V>(gamma, (int)((base + i) % C), g);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const float a = av[e], b = gv[e] * g[e], xh = (xv[e] - mu) * r;
            s[0] += a; s[1] += b; s[2] += a * xh; s[3] += b * xh; s[4] += a * b;
        }
    }
    block_sum<5>(s, ws + ((int64_t)n * P + p) * 5);
}

template <int V>
__global__ void __launch_bounds__(LN_THREADS)
ln_bwd2_apply_kernel(const void* __restrict__ cin, const void* __restrict__ gy, const void* __restrict__ x,
                     const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ ws, void* __restrict__ gx, int64_t M, int C, int P, int dt) {
    pdl_entry();
    const int n = blockIdx.y;
    const int64_t base = (int64_t)n * M;
    float s[5];
    fold<5>(ws, n, P, s);
    const float inv = 1.f / (float)M;
    const float abar = s[0] * inv, bbar = s[1] * inv, A = s[2] * inv, B = s[3] * inv;
    const float Q = s[4] * inv - abar * bbar - A * B;
    const float mu = mean[n], r = rstd[n], nr2 = -r * r;
    for (int64_t i = (blockIdx.x * (int64_t)LN_THREADS + threadIdx.x) * V; i < M; i += (int64_t)gridDim.x * LN_THREADS * V) {
        float av[V], gv[V], xv[V], g[V], o[V];
        ld_vec<V>(cin, base + i, dt, av); ld_vec<V>(gy, base + i, dt, gv); ld_vec<V>(x, base + i, dt, xv);
        ld_par<V>(gamma, (int)((base + i) % C), g);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const float a = av[e], b = gv[e] * g[e], xh = (xv[e] - mu) * r;
            o[e] = nr2 * (xh * Q + B * (a - abar) + A * (b - bbar) - 2.f * xh * A * B);
        }
        st_vec<V>(gx, base + i, dt, o);
    }
}

// 16-byte path: BF16, whole 8-channel groups, 16-byte aligned tensors
static bool ln_vec8(int64_t M, int C, int dt, const void* a, const void* b = nullptr, const void* c = nullptr, const void* d = nullptr) {
    const uintptr_t bits = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
                           reinterpret_cast<uintptr_t>(d);
    return dt == CTGAN_BF16 && C % 8 == 0 && M % 8 == 0 && (bits & 15) == 0;
}

static int ln_partials(int64_t M) {
    int64_t p = (M + 4095) / 4096;
    if (p < 1) p = 1;
    if (p > LN_MAX_P) p = LN_MAX_P;
    return (int)p;
}

static int check_ln(int N, int64_t M, int C, int dt, const char* who) {
    CTGAN_REQUIRE(N > 0 && N <= 65535 && M > 0 && C > 0 && M % C == 0 && dtype_ok(dt), CTGAN_ERR_BAD_DESC, "%s: bad shape", who);
    return 0;
}

}  // namespace
}  // namespace ctgan

using namespace ctgan;

extern "C" int64_t ctgan_ln_workspace_floats(int N, int64_t M) { return (int64_t)N * LN_MAX_P * 5; }

extern "C" int ctgan_ln_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, float* ws,
                            int N, int64_t M, int C, float eps, int dtype, void* stream) {
    if (int r = check_ln(N, M, C, dtype, "ln_fwd")) return r;
    CTGAN_REQUIRE(x && gamma && beta && y && mean && rstd && ws, CTGAN_ERR_BAD_DESC, "ln_fwd: null pointer");
    cudaStream_t st = as_stream(stream);
    const int P = ln_partials(M);
    const bool v8 = ln_vec8(M, C, dtype, x, y, gamma, beta);
    if (v8) CTGAN_LAUNCH((ln_stats_partial_kernel<8>), dim3(P, N), LN_THREADS, 0, st, x, ws, M, P, dtype);
    else    CTGAN_LAUNCH((ln_stats_partial_kernel<1>), dim3(P, N), LN_THREADS, 0, st, x, ws, M, P, dtype);
    CTGAN_CHECK_LAUNCH("ln_stats_partial");
    CTGAN_LAUNCH((ln_stats_finalize_kernel), (N + 127) / 128, 128, 0, st, x, (const float*)ws, mean, rstd, N, M, P, eps, dtype);
    CTGAN_CHECK_LAUNCH("ln_stats_finalize");
    const int64_t total = (int64_t)N * M;
    if (v8) CTGAN_LAUNCH((ln_apply_kernel<8>), elementwise_grid(total / 8, LN_THREADS), LN_THREADS, 0, st, x, gamma, beta, (const float*)mean,
                         (const float*)rstd, y, total, M, C, dtype);
    else    CTGAN_LAUNCH((ln_apply_kernel<1>), elementwise_grid(total, LN_THREADS), LN_THREADS, 0, st, x, gamma, beta, (const float*)mean,
                         (const float*)rstd, y, total, M, C, dtype);
    CTGAN_CHECK_LAUNCH("ln_apply");
    return 0;
}

extern "C" int ctgan_ln_core(const void* v, const void* x, const float* gamma, const float* mean, const float* rstd, void* out,
                             float* ws, int N, int64_t M, int C, int pre_scale, int post_scale, int dtype, void* stream) {
    if (int r = check_ln(N, M, C, dtype, "ln_core")) return r;
    CTGAN_REQUIRE(v && x && gamma && mean && rstd && out && ws, CTGAN_ERR_BAD_DESC, "ln_core: null pointer");
    cudaStream_t st = as_stream(stream);
    const int P = ln_partials(M);
    if (ln_vec8(M, C, dtype, v, x, out, gamma)) {
        CTGAN_LAUNCH((ln_core_partial_kernel<8>), dim3(P, N), LN_THREADS, 0, st, v, x, gamma, mean, rstd, ws, M, C, P, pre_scale, dtype);
        CTGAN_CHECK_LAUNCH("ln_core_partial");
        CTGAN_LAUNCH((ln_core_apply_kernel<8>), dim3(P, N), LN_THREADS, 0, st, v, x, gamma, mean, rstd, (const float*)ws, out, M, C, P,
                     pre_scale, post_scale, dtype);
    } else {
        CTGAN_LAUNCH((ln_core_partial_kernel<1>), dim3(P, N), LN_THREADS, 0, st, v, x, gamma, mean, rstd, ws, M, C, P, pre_scale, dtype);
        CTGAN_CHECK_LAUNCH("ln_core_partial");
        CTGAN_LAUNCH((ln_core_apply_kernel<1>), dim3(P, N), LN_THREADS, 0, st, v, x, gamma, mean, rstd, (const float*)ws, out, M, C, P,
                     pre_scale, post_scale, dtype);
    }
    CTGAN_CHECK_LAUNCH("ln_core_apply");
    return 0;
}

extern "C" int ctgan_ln_param_grad(const void* v, const void* x, const float* mean, const float* rstd, float* dgamma,
                                   float* dbeta, int N, int64_t M, int C, int dtype, void* stream) {
    if (int r = check_ln(N, M, C, dtype, "ln_param_grad")) return r;
    CTGAN_REQUIRE(v && x && mean && rstd && dgamma, CTGAN_ERR_BAD_DESC, "ln_param_grad: null pointer");
    const int64_t rows = (int64_t)N * (M / C);
    int rows_per_block = (int)((rows + (int64_t)sm_count() * 4 - 1) / ((int64_t)sm_count() * 4));
    if (rows_per_block < 8) rows_per_block = 8;
    const int blocks = (int)((rows + rows_per_block - 1) / rows_per_block);
    const int tpr = C / 8;
    if (ln_vec8(M, C, dtype, v, x) && tpr <= LN_THREADS && LN_THREADS % tpr == 0 && 2 * C <= 16 * LN_THREADS)
        CTGAN_LAUNCH((ln_param_grad_vec8_kernel), blocks, LN_THREADS, 0, as_stream(stream), (const __nv_bfloat16*)v, (const __nv_bfloat16*)x,
                     mean, rstd, dgamma, dbeta, rows, M / C, C, rows_per_block);
    else
        CTGAN_LAUNCH((ln_param_grad_kernel), blocks, LN_THREADS, 0, as_stream(stream), v, x, mean, rstd, dgamma, dbeta, rows, M / C, C,
                     rows_per_block, dtype);
    CTGAN_CHECK_LAUNCH("ln_param_grad");
    return 0;
}

extern "C" int ctgan_ln_bwd2_x(const void* c, const void* gy, const void* x, const float* gamma, const float* mean,
                               const float* rstd, void* gx, float* ws, int N, int64_t M, int C, int dtype, void* stream) {
    if (int r = check_ln(N, M, C, dtype, "ln_bwd2_x")) return r;
    CTGAN_REQUIRE(c && gy && x && gamma && mean && rstd && gx && ws, CTGAN_ERR_BAD_DESC, "ln_bwd2_x: null pointer");
    cudaStream_t st = as_stream(stream);
    const int P = ln_partials(M);
    if (ln_vec8(M, C, dtype, c, gy, x, gx) && (reinterpret_cast<uintptr_t>(gamma) & 15) == 0) {
        CTGAN_LAUNCH((ln_bwd2_partial_kernel<8>), dim3(P, N), LN_THREADS, 0, st, c, gy, x, gamma, mean, rstd, ws, M, C, P, dtype);
        CTGAN_CHECK_LAUNCH("ln_bwd2_partial");
        CTGAN_LAUNCH((ln_bwd2_apply_kernel<8>), dim3(P, N), LN_THREADS, 0, st, c, gy, x, gamma, mean, rstd, (const float*)ws, gx, M, C, P, dtype);
    } else {
        CTGAN_LAUNCH((ln_bwd2_partial_kernel<1>), dim3(P, N), LN_THREADS, 0, st, c, gy, x, gamma, mean, rstd, ws, M, C, P, dtype);
        CTGAN_CHECK_LAUNCH("ln_bwd2_partial");
        CTGAN_LAUNCH((ln_bwd2_apply_kernel<1>), dim3(P, N), LN_THREADS, 0, st, c, gy, x, gamma, mean, rstd, (const float*)ws, gx, M, C, P, dtype);
    }
    CTGAN_CHECK_LAUNCH("ln_bwd2_apply");
    return 0;
}
