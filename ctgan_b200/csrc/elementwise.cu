// elementwise.cu -- HBM-bound element-wise, data-movement and RNG kernels.
//
// All are grid-stride kernels moving 16 bytes per thread per access when the buffers
// allow it (8 x bf16 / 4 x float), falling back to scalar accesses otherwise.  Grids are
// capped at 8 CTAs per SM (common.cuh: elementwise_grid).
#include "common.cuh"

namespace ctgan {

template <typename T, int V> struct alignas(sizeof(T) * V) Pack { T v[V]; };

template <typename T> constexpr int vec_width() { return 16 / sizeof(T); }

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- generic map kernels: out[i] = f(a[i]) / f(a[i], b[i]) -----------------
template <typename T, int V, typename F>
__global__ void map1_kernel(const T* __restrict__ a, T* __restrict__ out, int64_t nvec, F f) {
    ctgan::pdl_entry();
    using P = Pack<T, V>;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        P pa = reinterpret_cast<const P*>(a)[i], po;
#pragma unroll
        for (int j = 0; j < V; ++j) po.v[j] = from_f<T>(f(to_f<T>(pa.v[j])));
        reinterpret_cast<P*>(out)[i] = po;
    }
}
template <typename T, int V, typename F>
__global__ void map2_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int64_t nvec, F f) {
    ctgan::pdl_entry();
    using P = Pack<T, V>;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        P pa = reinterpret_cast<const P*>(a)[i], pb = reinterpret_cast<const P*>(b)[i], po;
#pragma unroll
        for (int j = 0; j < V; ++j) po.v[j] = from_f<T>(f(to_f<T>(pa.v[j]), to_f<T>(pb.v[j])));
        reinterpret_cast<P*>(out)[i] = po;
    }
}

template <typename T, typename F>
static int launch_map1(const void* a, void* out, int64_t n, F f, cudaStream_t st, const char* what) {
    if (n <= 0) return 0;
    constexpr int V = vec_width<T>();
    if (aligned16(a) && aligned16(out) && n % V == 0) {
        int64_t nv = n / V;
        CTGAN_LAUNCH((map1_kernel<T, V, F>), elementwise_grid(nv, 256), 256, 0, st, (const T*)a, (T*)out, nv, f);
    } else {
        CTGAN_LAUNCH((map1_kernel<T, 1, F>), elementwise_grid(n, 256), 256, 0, st, (const T*)a, (T*)out, n, f);
    }
    CTGAN_CHECK_LAUNCH(what);
    return 0;
}
template <typename T, typename F>
static int launch_map2(const void* a, const void* b, void* out, int64_t n, F f, cudaStream_t st, const char* what) {
    if (n <= 0) return 0;
    constexpr int V = vec_width<T>();
    if (aligned16(a) && aligned16(b) && aligned16(out) && n % V == 0) {
        int64_t nv = n / V;
        CTGAN_LAUNCH((map2_kernel<T, V, F>), elementwise_grid(nv, 256), 256, 0, st, (const T*)a, (const T*)b, (T*)out, nv, f);
    } else {
        CTGAN_LAUNCH((map2_kernel<T, 1, F>), elementwise_grid(n, 256), 256, 0, st, (const T*)a, (const T*)b, (T*)out, n, f);
    }
    CTGAN_CHECK_LAUNCH(what);
    return 0;
}

struct AddOp  { __device__ float operator()(float a, float b) const { return a + b; } };
struct MulOp  { __device__ float operator()(float a, float b) const { return a * b; } };
struct ReluMaskOp { __device__ float operator()(float g, float y) const { return y > 0.f ? g : 0.f; } };   // g * [relu output > 0]
struct ScaleOp { float s; __device__ float operator()(float a) const { return a * s; } };
struct TanhOp { __device__ float operator()(float a) const { return tanhf(a); } };
struct SigmOp { __device__ float operator()(float a) const { return 1.f / (1.f + __expf(-a)); } };
// backward given forward OUTPUT y and upstream dy
struct TanhBwd { __device__ float operator()(float y, float dy) const { return dy * (1.f - y * y); } };
struct SigmBwd { __device__ float operator()(float y, float dy) const { return dy * y * (1.f - y); } };

// ---- cast --------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ x, TO* __restrict__ y, int64_t n) {
    ctgan::pdl_entry();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = from_f<TO>(to_f<TI>(x[i]));
}

// ---- fused activation + dropout ---------------------------------------------
template <typename T, int V>
__global__ void act_dropout_kernel(const T* __restrict__ x, const float* __restrict__ u, T* __restrict__ y,
                                   T* __restrict__ m, int64_t nvec, float slope, float keep,
                                   uint64_t seed, uint64_t offset, const uint64_t* __restrict__ dyn) {
    ctgan::pdl_entry();
    using P = Pack<T, V>;
    if (dyn) offset += dyn[0];
    const bool drop = keep < 1.f;
    const float inv_keep = 1.f / keep;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        P px = reinterpret_cast<const P*>(x)[i], py, pm;
        uint32_t r[4];
#pragma unroll
        for (int j = 0; j < V; ++j) {
            float xv = to_f<T>(px.v[j]);
            float mult = xv > 0.f ? 1.f : slope;
            if (drop) {
                int64_t e = i * V + j;
                float uu;
                if (u) {
                    uu = u[e];
                } else {
                    uint64_t se = offset + (uint64_t)e;
                    // offset is required to be a multiple of 4 when V > 1, so lanes j&3 share a block
                    if (V == 1 || (j & 3) == 0) Philox::block(seed, se >> 2, r);
                    uu = Philox::to_uniform(r[se & 3]);
                }
                mult *= floorf(keep + uu) * inv_keep;
            }
            pm.v[j] = from_f<T>(mult);
            // multiply by the ROUNDED multiplier so that y == x * m holds exactly for the saved m
            py.v[j] = from_f<T>(xv * to_f<T>(pm.v[j]));
        }
        reinterpret_cast<P*>(y)[i] = py;
        if (m) reinterpret_cast<P*>(m)[i] = pm;
    }
}

// ---- fused dropout -> (skip connection, ReLU) fork and the mask algebra of its backward ---------------------
// A residual block input after dropout feeds the block's shortcut (d) AND the block's first ReLU (r = relu(d)):
//   d = x * md,  r = x * mdr,   md = floor(keep+u)/keep,  mdr = md * [x > 0]        (one kernel instead of two)
// backward:        gx = gd * md + gr * mdr                                           (mask_sum2: one instead of three)
// double backward: c -> (c * md, c * mdr)                                            (mask_fork2: one instead of two)
// `ma == nullptr` means an identity mask on that branch (a fork without dropout).
template <typename T, int V>
__global__ void fork_dropout_relu_kernel(const T* __restrict__ x, const float* __restrict__ u, T* __restrict__ d,
                                         T* __restrict__ r, T* __restrict__ md, T* __restrict__ mdr, int64_t nvec,
                                         float keep, uint64_t seed, uint64_t offset, const uint64_t* __restrict__ dyn) {
    ctgan::pdl_entry();
    using P = Pack<T, V>;
    if (dyn) offset += dyn[0];
    const float inv_keep = 1.f / keep;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        P px = reinterpret_cast<const P*>(x)[i], pd, pr, pmd, pmdr;
        uint32_t rb[4];
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float xv = to_f<T>(px.v[j]);
            const int64_t e = i * V + j;
            float uu;
            if (u) {
                uu = u[e];
            } else {
                const uint64_t se = offset + (uint64_t)e;
                if (V == 1 || (j & 3) == 0) Philox::block(seed, se >> 2, rb);
                uu = Philox::to_uniform(rb[se & 3]);
            }
            pmd.v[j] = from_f<T>(floorf(keep + uu) * inv_keep);
            const float mdf = to_f<T>(pmd.v[j]);                       // the ROUNDED multiplier: d == x * md exactly
            pmdr.v[j] = from_f<T>(xv > 0.f ? mdf : 0.f);
            pd.v[j] = from_f<T>(xv * mdf);
            pr.v[j] = from_f<T>(xv > 0.f ? xv * mdf : 0.f);
        }
        reinterpret_cast<P*>(d)[i] = pd;
        reinterpret_cast<P*>(r)[i] = pr;
        reinterpret_cast<P*>(md)[i] = pmd;
        reinterpret_cast<P*>(mdr)[i] = pmdr;
    }
}
template <typename T, int V>
__global__ void mask_sum2_kernel(const T* __restrict__ a, const T* __restrict__ ma, const T* __restrict__ b,
                                 const T* __restrict__ mb, T* __restrict__ out, int64_t nvec) {
    ctgan::pdl_entry();
    using P = Pack<T, V>;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        P pa = reinterpret_cast<const P*>(a)[i], pb = reinterpret_cast<const P*>(b)[i], pmb = reinterpret_cast<const P*>(mb)[i], po, pma;
        if (ma) pma = reinterpret_cast<const P*>(ma)[i];
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float av = ma ? to_f<T>(pa.v[j]) * to_f<T>(pma.v[j]) : to_f<T>(pa.v[j]);
            po.v[j] = from_f<T>(fmaf(to_f<T>(pb.v[j]), to_f<T>(pmb.v[j]), av));
        }
        reinterpret_cast<P*>(out)[i] = po;
    }
}
template <typename T, int V>
__global__ void mask_fork2_kernel(const T* __restrict__ c, const T* __restrict__ ma, const T* __restrict__ mb,
                                  T* __restrict__ o1, T* __restrict__ o2, int64_t nvec) {
    ctgan::pdl_entry();
    using P = Pack<T, V>;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        P pc = reinterpret_cast<const P*>(c)[i], pma = reinterpret_cast<const P*>(ma)[i], pmb = reinterpret_cast<const P*>(mb)[i], p1, p2;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float cv = to_f<T>(pc.v[j]);
            p1.v[j] = from_f<T>(cv * to_f<T>(pma.v[j]));
            p2.v[j] = from_f<T>(cv * to_f<T>(pmb.v[j]));
        }
        reinterpret_cast<P*>(o1)[i] = p1;
        reinterpret_cast<P*>(o2)[i] = p2;
    }
}

// ---- 2x2 mean pool + skip add + fork, and its adjoint -----------------------------------------------------
// End of a down-sampling critic block: x = meanpool2x2(y) + s, then the next block's input fork (see above).
//   pool_add_fork (COMPUTE = 1): masks from the data: m1 = dropout multiplier (keep < 1; else identity, m1 not written),
//                                m2 = m1 * [x > 0];  o1 = x * m1, o2 = x * m2           (replaces pool, add, dropout, relu)
//   pool_add_fork (COMPUTE = 0): the same linear map with GIVEN masks (m1 may be null = identity)
//   mask_sum2_up: gx = a*m1 + b*m2 (low resolution) and gy = 0.25 * gx replicated 2x2      (the adjoint of the above)
template <typename T, int V, int COMPUTE>
__global__ void pool_add_fork_kernel(const T* __restrict__ y, const T* __restrict__ s, const float* __restrict__ u,
                                     T* __restrict__ m1, T* __restrict__ m2, T* __restrict__ o1, T* __restrict__ o2,
                                     int N, int H, int W, int C, float keep, uint64_t seed, uint64_t offset,
                                     const uint64_t* __restrict__ dyn) {
    ctgan::pdl_entry();
    using P = Pack<T, V>;
    const int Ho = H / 2, Wo = W / 2, CV = C / V;
    if (COMPUTE && dyn) offset += dyn[0];
    const bool drop = COMPUTE && keep < 1.f;
    const float inv_keep = 1.f / keep;
    const int64_t total = (int64_t)N * Ho * Wo * CV;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int cv = i % CV; int64_t t = i / CV;
        const int wo = t % Wo; t /= Wo;                                 // t = n * Ho + ho
        const P* r0 = reinterpret_cast<const P*>(y + ((2 * t) * W + 2 * wo) * (int64_t)C) + cv;
        const P* r1 = reinterpret_cast<const P*>(y + ((2 * t + 1) * W + 2 * wo) * (int64_t)C) + cv;
        const P a = r0[0], b = r0[CV], c = r1[0], d = r1[CV], ps = reinterpret_cast<const P*>(s)[i];
        P pm1, pm2, p1, p2;
        if (!COMPUTE) { pm2 = reinterpret_cast<const P*>(m2)[i]; if (m1) pm1 = reinterpret_cast<const P*>(m1)[i]; }
        uint32_t rb[4];
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float xf = 0.25f * (to_f<T>(a.v[j]) + to_f<T>(b.v[j]) + to_f<T>(c.v[j]) + to_f<T>(d.v[j])) + to_f<T>(ps.v[j]);
            const float x = to_f<T>(from_f<T>(xf));                     // the block output as it would be stored
            float f1 = 1.f, f2;
            if (COMPUTE) {
                if (drop) {
                    const int64_t e = i * V + j;
                    float uu;
                    if (u) uu = u[e];
                    else {
                        const uint64_t se = offset + (uint64_t)e;
                        if (V == 1 || (j & 3) == 0) Philox::block(seed, se >> 2, rb);
                        uu = Philox::to_uniform(rb[se & 3]);
                    }
                    pm1.v[j] = from_f<T>(floorf(keep + uu) * inv_keep);
                    f1 = to_f<T>(pm1.v[j]);
                }
                f2 = x > 0.f ? f1 : 0.f;
                pm2.v[j] = from_f<T>(f2);
            } else {
                if (m1) f1 = to_f<T>(pm1.v[j]);
                f2 = to_f<T>(pm2.v[j]);
            }
            p1.v[j] = from_f<T>(x * f1);
            p2.v[j] = from_f<T>(x * f2);
        }
        reinterpret_cast<P*>(o1)[i] = p1;
        reinterpret_cast<P*>(o2)[i] = p2;
        if (COMPUTE) {
            reinterpret_cast<P*>(m2)[i] = pm2;
            if (drop) reinterpret_cast<P*>(m1)[i] = pm1;
        }
    }
}
template <typename T, int V>
__global__ void mask_sum2_up_kernel(const T* __restrict__ a, const T* __restrict__ m1, const T* __restrict__ b,
                                    const T* __restrict__ m2, T* __restrict__ gx, T* __restrict__ gy,
                                    int N, int Ho, int Wo, int C) {
    ctgan::pdl_entry();
    using P = Pack<T, V>;
    const int W = 2 * Wo, CV = C / V;
    const int64_t total = (int64_t)N * Ho * Wo * CV;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int cv = i % CV; int64_t t = i / CV;
        const int wo = t % Wo; t /= Wo;
        const P pa = reinterpret_cast<const P*>(a)[i], pb = reinterpret_cast<const P*>(b)[i], pm2 = reinterpret_cast<const P*>(m2)[i];
        P pm1, px, py;
        if (m1) pm1 = reinterpret_cast<const P*>(m1)[i];
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float av = m1 ? to_f<T>(pa.v[j]) * to_f<T>(pm1.v[j]) : to_f<T>(pa.v[j]);
            px.v[j] = from_f<T>(fmaf(to_f<T>(pb.v[j]), to_f<T>(pm2.v[j]), av));
            py.v[j] = from_f<T>(0.25f * to_f<T>(px.v[j]));
        }
        reinterpret_cast<P*>(gx)[i] = px;
        P* r0 = reinterpret_cast<P*>(gy + ((2 * t) * W + 2 * wo) * (int64_t)C) + cv;
        P* r1 = reinterpret_cast<P*>(gy + ((2 * t + 1) * W + 2 * wo) * (int64_t)C) + cv;
        r0[0] = py; r0[CV] = py; r1[0] = py; r1[CV] = py;
    }
}

template <typename T>
__global__ void bias_add_kernel(const T* __restrict__ x, const float* __restrict__ b, T* __restrict__ y, int64_t total, int C) {
    ctgan::pdl_entry();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = from_f<T>(to_f<T>(x[i]) + b[i % C]);
}

// ---- pooling / upsampling on NHWC --------------------------------------------
// V channels (16 bytes when the channel count allows) per thread; index math once per vector.
template <typename T, int V>
__global__ void pool2x2_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, float scale) {
    ctgan::pdl_entry();
    using P = Pack<T, V>;
    const int Ho = H / 2, Wo = W / 2, CV = C / V;
    int64_t total = (int64_t)N * Ho * Wo * CV;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int cv = i % CV; int64_t t = i / CV;
        int wo = t % Wo; t /= Wo;
        int ho = t % Ho; int n = t / Ho;
        const P* p = reinterpret_cast<const P*>(x + (((int64_t)n * H + 2 * ho) * W + 2 * wo) * C) + cv;
        P a = p[0], b = p[CV], c = p[(int64_t)W * CV], d = p[(int64_t)W * CV + CV], o;
#pragma unroll
        for (int k = 0; k < V; ++k)
            o.v[k] = from_f<T>((to_f<T>(a.v[k]) + to_f<T>(c.v[k]) + to_f<T>(b.v[k]) + to_f<T>(d.v[k])) * scale);
        reinterpret_cast<P*>(y)[i] = o;
    }
}
template <typename T, int V>
__global__ void upsample2x_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, float scale) {
    ctgan::pdl_entry();
    using P = Pack<T, V>;
    // one thread per INPUT vector: one load, four stores (the 2x2 replicas)
    const int Wo = 2 * W, CV = C / V;
    int64_t total = (int64_t)N * H * W * CV;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int cv = i % CV; int64_t t = i / CV;
        int w = t % W; t /= W;                       // t = n * H + h
        P a = reinterpret_cast<const P*>(x)[i], o;
#pragma unroll
        for (int k = 0; k < V; ++k) o.v[k] = from_f<T>(to_f<T>(a.v[k]) * scale);
        P* row0 = reinterpret_cast<P*>(y + ((2 * t) * Wo + 2 * w) * (int64_t)C) + cv;
        P* row1 = reinterpret_cast<P*>(y + ((2 * t + 1) * Wo + 2 * w) * (int64_t)C) + cv;
        row0[0] = o; row0[CV] = o; row1[0] = o; row1[CV] = o;
    }
}
template <typename T>
__global__ void spatial_sum_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int HW, int C, float scale) {
    ctgan::pdl_entry();
    int64_t total = (int64_t)N * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = i % C; int n = i / C;
        const T* p = x + (int64_t)n * HW * C + c;
        float s = 0.f;
        for (int j = 0; j < HW; ++j) s += to_f<T>(p[(int64_t)j * C]);
        y[i] = from_f<T>(s * scale);
    }
}
template <typename T>
__global__ void spatial_bcast_kernel(const T* __restrict__ y, T* __restrict__ x, int N, int HW, int C, float scale) {
    ctgan::pdl_entry();
    int64_t total = (int64_t)N * HW * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = i % C; int n = i / ((int64_t)HW * C);
        x[i] = from_f<T>(to_f<T>(y[(int64_t)n * C + c]) * scale);
    }
}

// ---- layout ----------------------------------------------------------------
// Both directions index by the OUTPUT element; the 1/3-channel image tensors they are
// used on are tiny, so no smem transpose.
__global__ void nchw_to_nhwc_kernel(const void* __restrict__ x, int xdt, void* __restrict__ y, int ydt,
                                    int N, int C, int H, int W) {
    ctgan::pdl_entry();
    int64_t total = (int64_t)N * C * H * W;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = i % C; int64_t t = i / C;
        int w = t % W; t /= W;
        int h = t % H; int n = t / H;
        st_act(y, i, ydt, ld_act(x, (((int64_t)n * C + c) * H + h) * W + w, xdt));
    }
}
__global__ void nhwc_to_nchw_kernel(const void* __restrict__ x, int xdt, void* __restrict__ y, int ydt,
                                    int N, int C, int H, int W) {
    ctgan::pdl_entry();
    int64_t total = (int64_t)N * C * H * W;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int w = i % W; int64_t t = i / W;
        int h = t % H; t /= H;
        int c = t % C; int n = t / C;
        st_act(y, i, ydt, ld_act(x, (((int64_t)n * H + h) * W + w) * C + c, xdt));
    }
}
template <typename T>
__global__ void crop_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int h, int w, int fwd) {
    ctgan::pdl_entry();
    // fwd: y[N,h,w,C] = x[N,:h,:w,C];  !fwd: y is [N,H,W,C] zero-padded copy of x[N,h,w,C]
    int64_t total = fwd ? (int64_t)N * h * w * C : (int64_t)N * H * W * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = i % C; int64_t t = i / C;
        if (fwd) {
            int ww = t % w; t /= w;
            int hh = t % h; int n = t / h;
            y[i] = x[(((int64_t)n * H + hh) * W + ww) * C + c];
        } else {
            int ww = t % W; t /= W;
            int hh = t % H; int n = t / H;
            y[i] = (hh < h && ww < w) ? x[(((int64_t)n * h + hh) * w + ww) * C + c] : from_f<T>(0.f);
        }
    }
}

// y2 (nullable): a second destination for the same values -- the real batch occupies two row ranges of the stacked critic
// input (passes ' and ''), written here instead of by a concatenation kernel
template <typename TIn>
__global__ void prep_real_div_kernel(const TIn* __restrict__ x, float* __restrict__ y, float* __restrict__ y2, int64_t n, float denom,
                                     float noise_hi, uint64_t seed, uint64_t offset, const uint64_t* __restrict__ dyn) {
    ctgan::pdl_entry();
    if (dyn) offset += dyn[0];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v = 2.f * (__fdiv_rn((float)x[i], denom) - 0.5f);       // true division: /255 is not exact as a multiply
        if (noise_hi > 0.f) v += noise_hi * Philox::uniform_at(seed, offset + (uint64_t)i);
        y[i] = v;
        if (y2) y2[i] = v;
    }
}
__global__ void interpolate_kernel(const float* __restrict__ real, const float* __restrict__ fake,
                                   const float* __restrict__ alpha, float* __restrict__ out, int B, int P) {
    ctgan::pdl_entry();
    int64_t total = (int64_t)B * P;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int b = i / P;
        float r = real[i];
        out[i] = r + alpha[b] * (fake[i] - r);
    }
}

// ---- RNG ---------------------------------------------------------------------
__global__ void philox_uniform_kernel(float* __restrict__ out, int64_t n, float lo, float hi, uint64_t seed, uint64_t offset,
                                      const uint64_t* __restrict__ dyn) {
    ctgan::pdl_entry();
    if (dyn) offset += dyn[0];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = lo + (hi - lo) * Philox::uniform_at(seed, offset + (uint64_t)i);
}
__global__ void philox_normal_kernel(float* __restrict__ out, int64_t n, uint64_t seed, uint64_t offset,
                                     const uint64_t* __restrict__ dyn) {
    ctgan::pdl_entry();
    if (dyn) offset += dyn[0];
    // element i uses stream elements (2i, 2i+1): Box-Muller, cosine branch only
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t e = offset + 2ull * (uint64_t)i;
        float u1 = Philox::uniform_at(seed, e), u2 = Philox::uniform_at(seed, e + 1);
        u1 = fmaxf(u1, 5.9604645e-8f);                                   // avoid log(0)
        out[i] = sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
    }
}
__global__ void philox_labels_kernel(int32_t* __restrict__ out, int64_t n, int n_labels, uint64_t seed, uint64_t offset,
                                     const uint64_t* __restrict__ dyn) {
    ctgan::pdl_entry();
    if (dyn) offset += dyn[0];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (int32_t)(Philox::uniform_at(seed, offset + (uint64_t)i) * (float)n_labels);
}

__global__ void counter_add_kernel(uint64_t* ctr, uint64_t delta) {
    ctgan::pdl_entry(); ctr[0] += delta; }

}  // namespace ctgan

using namespace ctgan;

#define DISPATCH_T(dtype, CALL_F32, CALL_BF16)                                  \
    do {                                                                        \
        if ((dtype) == CTGAN_F32) { CALL_F32; }                                 \
        else if ((dtype) == CTGAN_BF16) { CALL_BF16; }                          \
        else { set_error("bad dtype %d", (int)(dtype)); return CTGAN_ERR_BAD_DESC; } \
    } while (0)

extern "C" int ctgan_add(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream) {
    DISPATCH_T(dtype, return launch_map2<float>(a, b, out, n, AddOp{}, as_stream(stream), "add"),
                      return launch_map2<__nv_bfloat16>(a, b, out, n, AddOp{}, as_stream(stream), "add"));
}
extern "C" int ctgan_mul(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream) {
    DISPATCH_T(dtype, return launch_map2<float>(a, b, out, n, MulOp{}, as_stream(stream), "mul"),
                      return launch_map2<__nv_bfloat16>(a, b, out, n, MulOp{}, as_stream(stream), "mul"));
}
extern "C" int ctgan_scale(const void* a, float s, void* out, int64_t n, int dtype, void* stream) {
    DISPATCH_T(dtype, return launch_map1<float>(a, out, n, ScaleOp{s}, as_stream(stream), "scale"),
                      return launch_map1<__nv_bfloat16>(a, out, n, ScaleOp{s}, as_stream(stream), "scale"));
}
extern "C" int ctgan_unary_fwd(const void* x, void* y, int64_t n, int dtype, int kind, void* stream) {
    CTGAN_REQUIRE(kind == 0 || kind == 1, CTGAN_ERR_BAD_DESC, "unary_fwd: kind must be 0 (tanh) or 1 (sigmoid)");
    cudaStream_t st = as_stream(stream);
    if (kind == 0)
        DISPATCH_T(dtype, return launch_map1<float>(x, y, n, TanhOp{}, st, "tanh"),
                          return launch_map1<__nv_bfloat16>(x, y, n, TanhOp{}, st, "tanh"));
    DISPATCH_T(dtype, return launch_map1<float>(x, y, n, SigmOp{}, st, "sigmoid"),
                      return launch_map1<__nv_bfloat16>(x, y, n, SigmOp{}, st, "sigmoid"));
}
extern "C" int ctgan_unary_bwd(const void* y, const void* dy, void* dx, int64_t n, int dtype, int kind, void* stream) {
    CTGAN_REQUIRE(kind == 0 || kind == 1, CTGAN_ERR_BAD_DESC, "unary_bwd: kind must be 0 (tanh) or 1 (sigmoid)");
    cudaStream_t st = as_stream(stream);
    if (kind == 0)
        DISPATCH_T(dtype, return launch_map2<float>(y, dy, dx, n, TanhBwd{}, st, "tanh_bwd"),
                          return launch_map2<__nv_bfloat16>(y, dy, dx, n, TanhBwd{}, st, "tanh_bwd"));
    DISPATCH_T(dtype, return launch_map2<float>(y, dy, dx, n, SigmBwd{}, st, "sigmoid_bwd"),
                      return launch_map2<__nv_bfloat16>(y, dy, dx, n, SigmBwd{}, st, "sigmoid_bwd"));
}

extern "C" int ctgan_cast(const void* x, int xdt, void* y, int ydt, int64_t n, void* stream) {
    CTGAN_REQUIRE(dtype_ok(xdt) && dtype_ok(ydt), CTGAN_ERR_BAD_DESC, "cast: bad dtype");
    if (n <= 0) return 0;
    cudaStream_t st = as_stream(stream);
    int grid = elementwise_grid(n, 256);
    if (xdt == CTGAN_F32 && ydt == CTGAN_BF16) CTGAN_LAUNCH((cast_kernel<float, __nv_bfloat16>), grid, 256, 0, st, (const float*)x, (__nv_bfloat16*)y, n);
    else if (xdt == CTGAN_BF16 && ydt == CTGAN_F32) CTGAN_LAUNCH((cast_kernel<__nv_bfloat16, float>), grid, 256, 0, st, (const __nv_bfloat16*)x, (float*)y, n);
    else if (xdt == CTGAN_F32) CTGAN_LAUNCH((cast_kernel<float, float>), grid, 256, 0, st, (const float*)x, (float*)y, n);
    else CTGAN_LAUNCH((cast_kernel<__nv_bfloat16, __nv_bfloat16>), grid, 256, 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, n);
    CTGAN_CHECK_LAUNCH("cast");
    return 0;
}

template <typename T>
static int launch_act_dropout(const void* x, const float* u, void* y, void* m, int64_t n, float slope, float keep,
                              uint64_t seed, uint64_t offset, const uint64_t* dyn, cudaStream_t st) {
    constexpr int V = vec_width<T>();
    bool vec = aligned16(x) && aligned16(y) && (m == nullptr || aligned16(m)) && n % V == 0 && (offset & 3) == 0;
    if (vec) {
        int64_t nv = n / V;
        CTGAN_LAUNCH((act_dropout_kernel<T, V>), elementwise_grid(nv, 256), 256, 0, st, (const T*)x, u, (T*)y, (T*)m, nv, slope, keep, seed, offset, dyn);
    } else {
        CTGAN_LAUNCH((act_dropout_kernel<T, 1>), elementwise_grid(n, 256), 256, 0, st, (const T*)x, u, (T*)y, (T*)m, n, slope, keep, seed, offset, dyn);
    }
    CTGAN_CHECK_LAUNCH("act_dropout_fwd");
    return 0;
}
extern "C" int ctgan_act_dropout_fwd(const void* x, const float* u, void* y, void* m, int64_t n, int dtype,
                                     float slope, float keep, uint64_t seed, uint64_t offset,
                                     const uint64_t* dyn_offset, void* stream) {
    CTGAN_REQUIRE(keep > 0.f && keep <= 1.f, CTGAN_ERR_BAD_DESC, "act_dropout_fwd: keep must be in (0,1]");
    if (n <= 0) return 0;
    DISPATCH_T(dtype, return launch_act_dropout<float>(x, u, y, m, n, slope, keep, seed, offset, dyn_offset, as_stream(stream)),
                      return launch_act_dropout<__nv_bfloat16>(x, u, y, m, n, slope, keep, seed, offset, dyn_offset, as_stream(stream)));
}

template <typename T>
static int launch_fork(const void* x, const float* u, void* d, void* r, void* md, void* mdr, int64_t n, float keep,
                       uint64_t seed, uint64_t offset, const uint64_t* dyn, cudaStream_t st) {
    constexpr int V = vec_width<T>();
    const bool vec = aligned16(x) && aligned16(d) && aligned16(r) && aligned16(md) && aligned16(mdr) && n % V == 0 && (offset & 3) == 0;
    if (vec) CTGAN_LAUNCH((fork_dropout_relu_kernel<T, V>), elementwise_grid(n / V, 256), 256, 0, st, (const T*)x, u, (T*)d, (T*)r, (T*)md, (T*)mdr, n / V, keep, seed, offset, dyn);
    else     CTGAN_LAUNCH((fork_dropout_relu_kernel<T, 1>), elementwise_grid(n, 256), 256, 0, st, (const T*)x, u, (T*)d, (T*)r, (T*)md, (T*)mdr, n, keep, seed, offset, dyn);
    CTGAN_CHECK_LAUNCH("fork_dropout_relu");
    return 0;
}
extern "C" int ctgan_fork_dropout_relu(const void* x, const float* u, void* d, void* r, void* md, void* mdr, int64_t n, int dtype,
                                       float keep, uint64_t seed, uint64_t offset, const uint64_t* dyn_offset, void* stream) {
    CTGAN_REQUIRE(keep > 0.f && keep <= 1.f && x && d && r && md && mdr, CTGAN_ERR_BAD_DESC, "fork_dropout_relu: bad args");
    if (n <= 0) return 0;
    DISPATCH_T(dtype, return launch_fork<float>(x, u, d, r, md, mdr, n, keep, seed, offset, dyn_offset, as_stream(stream)),
                      return launch_fork<__nv_bfloat16>(x, u, d, r, md, mdr, n, keep, seed, offset, dyn_offset, as_stream(stream)));
}
template <typename T>
static int launch_mask_sum2(const void* a, const void* ma, const void* b, const void* mb, void* out, int64_t n, cudaStream_t st) {
    constexpr int V = vec_width<T>();
    const bool vec = aligned16(a) && (!ma || aligned16(ma)) && aligned16(b) && aligned16(mb) && aligned16(out) && n % V == 0;
    if (vec) CTGAN_LAUNCH((mask_sum2_kernel<T, V>), elementwise_grid(n / V, 256), 256, 0, st, (const T*)a, (const T*)ma, (const T*)b, (const T*)mb, (T*)out, n / V);
    else     CTGAN_LAUNCH((mask_sum2_kernel<T, 1>), elementwise_grid(n, 256), 256, 0, st, (const T*)a, (const T*)ma, (const T*)b, (const T*)mb, (T*)out, n);
    CTGAN_CHECK_LAUNCH("mask_sum2");
    return 0;
}
extern "C" int ctgan_mask_sum2(const void* a, const void* ma, const void* b, const void* mb, void* out, int64_t n, int dtype, void* stream) {
    CTGAN_REQUIRE(a && b && mb && out, CTGAN_ERR_BAD_DESC, "mask_sum2: null pointer");
    if (n <= 0) return 0;
    DISPATCH_T(dtype, return launch_mask_sum2<float>(a, ma, b, mb, out, n, as_stream(stream)),
                      return launch_mask_sum2<__nv_bfloat16>(a, ma, b, mb, out, n, as_stream(stream)));
}
template <typename T>
static int launch_mask_fork2(const void* c, const void* ma, const void* mb, void* o1, void* o2, int64_t n, cudaStream_t st) {
    constexpr int V = vec_width<T>();
    const bool vec = aligned16(c) && aligned16(ma) && aligned16(mb) && aligned16(o1) && aligned16(o2) && n % V == 0;
    if (vec) CTGAN_LAUNCH((mask_fork2_kernel<T, V>), elementwise_grid(n / V, 256), 256, 0, st, (const T*)c, (const T*)ma, (const T*)mb, (T*)o1, (T*)o2, n / V);
    else     CTGAN_LAUNCH((mask_fork2_kernel<T, 1>), elementwise_grid(n, 256), 256, 0, st, (const T*)c, (const T*)ma, (const T*)mb, (T*)o1, (T*)o2, n);
    CTGAN_CHECK_LAUNCH("mask_fork2");
    return 0;
}
extern "C" int ctgan_mask_fork2(const void* c, const void* ma, const void* mb, void* o1, void* o2, int64_t n, int dtype, void* stream) {
    CTGAN_REQUIRE(c && ma && mb && o1 && o2, CTGAN_ERR_BAD_DESC, "mask_fork2: null pointer");
    if (n <= 0) return 0;
    DISPATCH_T(dtype, return launch_mask_fork2<float>(c, ma, mb, o1, o2, n, as_stream(stream)),
                      return launch_mask_fork2<__nv_bfloat16>(c, ma, mb, o1, o2, n, as_stream(stream)));
}
extern "C" int ctgan_mul_relu_mask(const void* g, const void* y, void* out, int64_t n, int dtype, void* stream) {
    DISPATCH_T(dtype, return launch_map2<float>(g, y, out, n, ReluMaskOp{}, as_stream(stream), "mul_relu_mask"),
                      return launch_map2<__nv_bfloat16>(g, y, out, n, ReluMaskOp{}, as_stream(stream), "mul_relu_mask"));
}

template <typename T>
static int launch_pool_add_fork(int compute, const void* y, const void* s, const float* u, void* m1, void* m2, void* o1, void* o2,
                                int N, int H, int W, int C, float keep, uint64_t seed, uint64_t offset, const uint64_t* dyn,
                                cudaStream_t st) {
    constexpr int V = vec_width<T>();
    const bool vec = C % V == 0 && aligned16(y) && aligned16(s) && (!m1 || aligned16(m1)) && aligned16(m2) && aligned16(o1) &&
                     aligned16(o2) && (offset & 3) == 0;
    const int64_t n = (int64_t)N * (H / 2) * (W / 2) * C;
#define CTGAN_PAF(VV, CC) CTGAN_LAUNCH((pool_add_fork_kernel<T, VV, CC>), elementwise_grid(n / VV, 256), 256, 0, st,  \
        (const T*)y, (const T*)s, u, (T*)m1, (T*)m2, (T*)o1, (T*)o2, N, H, W, C, keep, seed, offset, dyn)
    if (compute) { if (vec) CTGAN_PAF(V, 1); else CTGAN_PAF(1, 1); }
    else         { if (vec) CTGAN_PAF(V, 0); else CTGAN_PAF(1, 0); }
#undef CTGAN_PAF
    CTGAN_CHECK_LAUNCH("pool_add_fork");
    return 0;
}
extern "C" int ctgan_pool_add_fork(int compute_masks, const void* y, const void* s, const float* u, void* m1, void* m2, void* o1,
                                   void* o2, int N, int H, int W, int C, int dtype, float keep, uint64_t seed, uint64_t offset,
                                   const uint64_t* dyn_offset, void* stream) {
    CTGAN_REQUIRE(y && s && m2 && o1 && o2 && N > 0 && H > 0 && W > 0 && C > 0 && H % 2 == 0 && W % 2 == 0 && keep > 0.f && keep <= 1.f,
                  CTGAN_ERR_BAD_DESC, "pool_add_fork: bad args");
    CTGAN_REQUIRE(!(compute_masks && keep < 1.f) || m1, CTGAN_ERR_BAD_DESC, "pool_add_fork: dropout needs the m1 output");
    DISPATCH_T(dtype, return launch_pool_add_fork<float>(compute_masks, y, s, u, m1, m2, o1, o2, N, H, W, C, keep, seed, offset, dyn_offset, as_stream(stream)),
                      return launch_pool_add_fork<__nv_bfloat16>(compute_masks, y, s, u, m1, m2, o1, o2, N, H, W, C, keep, seed, offset, dyn_offset, as_stream(stream)));
}
template <typename T>
static int launch_mask_sum2_up(const void* a, const void* m1, const void* b, const void* m2, void* gx, void* gy, int N, int Ho, int Wo,
                               int C, cudaStream_t st) {
    constexpr int V = vec_width<T>();
    const bool vec = C % V == 0 && aligned16(a) && (!m1 || aligned16(m1)) && aligned16(b) && aligned16(m2) && aligned16(gx) && aligned16(gy);
    const int64_t n = (int64_t)N * Ho * Wo * C;
    if (vec) CTGAN_LAUNCH((mask_sum2_up_kernel<T, V>), elementwise_grid(n / V, 256), 256, 0, st, (const T*)a, (const T*)m1, (const T*)b, (const T*)m2, (T*)gx, (T*)gy, N, Ho, Wo, C);
    else     CTGAN_LAUNCH((mask_sum2_up_kernel<T, 1>), elementwise_grid(n, 256), 256, 0, st, (const T*)a, (const T*)m1, (const T*)b, (const T*)m2, (T*)gx, (T*)gy, N, Ho, Wo, C);
    CTGAN_CHECK_LAUNCH("mask_sum2_up");
    return 0;
}
extern "C" int ctgan_mask_sum2_up(const void* a, const void* m1, const void* b, const void* m2, void* gx, void* gy, int N, int Ho, int Wo,
                                  int C, int dtype, void* stream) {
    CTGAN_REQUIRE(a && b && m2 && gx && gy && N > 0 && Ho > 0 && Wo > 0 && C > 0, CTGAN_ERR_BAD_DESC, "mask_sum2_up: bad args");
    DISPATCH_T(dtype, return launch_mask_sum2_up<float>(a, m1, b, m2, gx, gy, N, Ho, Wo, C, as_stream(stream)),
                      return launch_mask_sum2_up<__nv_bfloat16>(a, m1, b, m2, gx, gy, N, Ho, Wo, C, as_stream(stream)));
}

extern "C" int ctgan_bias_add(const void* x, const float* b, void* y, int64_t rows, int C, int dtype, void* stream) {
    CTGAN_REQUIRE(x && b && y && rows > 0 && C > 0, CTGAN_ERR_BAD_DESC, "bias_add: bad args");
    int64_t total = rows * C;
    int grid = elementwise_grid(total, 256);
    cudaStream_t st = as_stream(stream);
    DISPATCH_T(dtype, (CTGAN_LAUNCH((bias_add_kernel<float>), grid, 256, 0, st, (const float*)x, b, (float*)y, total, C)),
                      (CTGAN_LAUNCH((bias_add_kernel<__nv_bfloat16>), grid, 256, 0, st, (const __nv_bfloat16*)x, b, (__nv_bfloat16*)y, total, C)));
    CTGAN_CHECK_LAUNCH("bias_add");
    return 0;
}
template <typename T>
static int launch_resample(bool pool, const void* x, void* y, int N, int H, int W, int C, float scale, cudaStream_t st) {
    constexpr int V = vec_width<T>();
    const int64_t out_elems = pool ? (int64_t)N * (H / 2) * (W / 2) * C : (int64_t)N * H * W * C;   // threads: pool = outputs, upsample = inputs
    const bool vec = C % V == 0 && aligned16(x) && aligned16(y);
    const int grid = elementwise_grid(vec ? out_elems / V : out_elems, 256);
    if (pool) {
        if (vec) CTGAN_LAUNCH((pool2x2_kernel<T, V>), grid, 256, 0, st, (const T*)x, (T*)y, N, H, W, C, scale);
        else     CTGAN_LAUNCH((pool2x2_kernel<T, 1>), grid, 256, 0, st, (const T*)x, (T*)y, N, H, W, C, scale);
    } else {
        if (vec) CTGAN_LAUNCH((upsample2x_kernel<T, V>), grid, 256, 0, st, (const T*)x, (T*)y, N, H, W, C, scale);
        else     CTGAN_LAUNCH((upsample2x_kernel<T, 1>), grid, 256, 0, st, (const T*)x, (T*)y, N, H, W, C, scale);
    }
    CTGAN_CHECK_LAUNCH(pool ? "pool2x2" : "upsample2x");
    return 0;
}
extern "C" int ctgan_pool2x2(const void* x, void* y, int N, int H, int W, int C, float scale, int dtype, void* stream) {
    CTGAN_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && H % 2 == 0 && W % 2 == 0, CTGAN_ERR_BAD_DESC, "pool2x2: H and W must be even and positive");
    DISPATCH_T(dtype, return launch_resample<float>(true, x, y, N, H, W, C, scale, as_stream(stream)),
                      return launch_resample<__nv_bfloat16>(true, x, y, N, H, W, C, scale, as_stream(stream)));
}
extern "C" int ctgan_upsample2x(const void* x, void* y, int N, int H, int W, int C, float scale, int dtype, void* stream) {
    CTGAN_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, CTGAN_ERR_BAD_DESC, "upsample2x: bad shape");
    DISPATCH_T(dtype, return launch_resample<float>(false, x, y, N, H, W, C, scale, as_stream(stream)),
                      return launch_resample<__nv_bfloat16>(false, x, y, N, H, W, C, scale, as_stream(stream)));
}
extern "C" int ctgan_spatial_sum(const void* x, void* y, int N, int HW, int C, float scale, int dtype, void* stream) {
    CTGAN_REQUIRE(N > 0 && HW > 0 && C > 0, CTGAN_ERR_BAD_DESC, "spatial_sum: bad shape");
    int grid = elementwise_grid((int64_t)N * C, 128);
    cudaStream_t st = as_stream(stream);
    DISPATCH_T(dtype, (CTGAN_LAUNCH((spatial_sum_kernel<float>), grid, 128, 0, st, (const float*)x, (float*)y, N, HW, C, scale)),
                      (CTGAN_LAUNCH((spatial_sum_kernel<__nv_bfloat16>), grid, 128, 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, N, HW, C, scale)));
    CTGAN_CHECK_LAUNCH("spatial_sum");
    return 0;
}
extern "C" int ctgan_spatial_bcast(const void* y, void* x, int N, int HW, int C, float scale, int dtype, void* stream) {
    CTGAN_REQUIRE(N > 0 && HW > 0 && C > 0, CTGAN_ERR_BAD_DESC, "spatial_bcast: bad shape");
    int grid = elementwise_grid((int64_t)N * HW * C, 256);
    cudaStream_t st = as_stream(stream);
    DISPATCH_T(dtype, (CTGAN_LAUNCH((spatial_bcast_kernel<float>), grid, 256, 0, st, (const float*)y, (float*)x, N, HW, C, scale)),
                      (CTGAN_LAUNCH((spatial_bcast_kernel<__nv_bfloat16>), grid, 256, 0, st, (const __nv_bfloat16*)y, (__nv_bfloat16*)x, N, HW, C, scale)));
    CTGAN_CHECK_LAUNCH("spatial_bcast");
    return 0;
}

extern "C" int ctgan_nchw_to_nhwc(const void* x, int xdt, void* y, int ydt, int N, int C, int H, int W, void* stream) {
    CTGAN_REQUIRE(dtype_ok(xdt) && dtype_ok(ydt) && N > 0 && C > 0 && H > 0 && W > 0, CTGAN_ERR_BAD_DESC, "nchw_to_nhwc: bad args");
    CTGAN_LAUNCH((nchw_to_nhwc_kernel), elementwise_grid((int64_t)N * C * H * W, 256), 256, 0, as_stream(stream), x, xdt, y, ydt, N, C, H, W);
    CTGAN_CHECK_LAUNCH("nchw_to_nhwc");
    return 0;
}
extern "C" int ctgan_nhwc_to_nchw(const void* x, int xdt, void* y, int ydt, int N, int C, int H, int W, void* stream) {
    CTGAN_REQUIRE(dtype_ok(xdt) && dtype_ok(ydt) && N > 0 && C > 0 && H > 0 && W > 0, CTGAN_ERR_BAD_DESC, "nhwc_to_nchw: bad args");
    CTGAN_LAUNCH((nhwc_to_nchw_kernel), elementwise_grid((int64_t)N * C * H * W, 256), 256, 0, as_stream(stream), x, xdt, y, ydt, N, C, H, W);
    CTGAN_CHECK_LAUNCH("nhwc_to_nchw");
    return 0;
}
static int crop_impl(const void* x, void* y, int N, int H, int W, int C, int h, int w, int dtype, int fwd, void* stream) {
    CTGAN_REQUIRE(N > 0 && C > 0 && h > 0 && w > 0 && h <= H && w <= W, CTGAN_ERR_BAD_DESC, "crop: bad shape");
    int64_t total = fwd ? (int64_t)N * h * w * C : (int64_t)N * H * W * C;
    int grid = elementwise_grid(total, 256);
    cudaStream_t st = as_stream(stream);
    DISPATCH_T(dtype, (CTGAN_LAUNCH((crop_kernel<float>), grid, 256, 0, st, (const float*)x, (float*)y, N, H, W, C, h, w, fwd)),
                      (CTGAN_LAUNCH((crop_kernel<__nv_bfloat16>), grid, 256, 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, N, H, W, C, h, w, fwd)));
    CTGAN_CHECK_LAUNCH("crop");
    return 0;
}
extern "C" int ctgan_crop(const void* x, void* y, int N, int H, int W, int C, int h, int w, int dtype, void* stream) {
    return crop_impl(x, y, N, H, W, C, h, w, dtype, 1, stream);
}
extern "C" int ctgan_crop_bwd(const void* dy, void* dx, int N, int H, int W, int C, int h, int w, int dtype, void* stream) {
    return crop_impl(dy, dx, N, H, W, C, h, w, dtype, 0, stream);
}

extern "C" int ctgan_prep_real(const int32_t* x, float* y, int64_t n, float denom, float noise_hi,
                               uint64_t seed, uint64_t offset, const uint64_t* dyn_offset, void* stream) {
    return ctgan_prep_real_dup(x, CTGAN_F32 /* int32 */, y, nullptr, n, denom, noise_hi, seed, offset, dyn_offset, stream);
}
extern "C" int ctgan_prep_real_u8(const uint8_t* x, float* y, int64_t n, float denom, float noise_hi,
                                  uint64_t seed, uint64_t offset, const uint64_t* dyn_offset, void* stream) {
    return ctgan_prep_real_dup(x, 1 /* uint8 */, y, nullptr, n, denom, noise_hi, seed, offset, dyn_offset, stream);
}
extern "C" int ctgan_prep_real_dup(const void* x, int x_is_u8, float* y, float* y2, int64_t n, float denom, float noise_hi,
                                   uint64_t seed, uint64_t offset, const uint64_t* dyn_offset, void* stream) {
    CTGAN_REQUIRE(denom > 0.f && x && y, CTGAN_ERR_BAD_DESC, "prep_real: denom must be positive, x / y non-null");
    if (n <= 0) return 0;
    if (x_is_u8)
        CTGAN_LAUNCH((prep_real_div_kernel<uint8_t>), elementwise_grid(n, 256), 256, 0, as_stream(stream), (const uint8_t*)x, y, y2, n, denom, noise_hi, seed, offset, dyn_offset);
    else
        CTGAN_LAUNCH((prep_real_div_kernel<int32_t>), elementwise_grid(n, 256), 256, 0, as_stream(stream), (const int32_t*)x, y, y2, n, denom, noise_hi, seed, offset, dyn_offset);
    CTGAN_CHECK_LAUNCH("prep_real");
    return 0;
}
extern "C" int ctgan_memset_zero(void* p, int64_t bytes, void* stream) {
    CTGAN_REQUIRE(p && bytes >= 0, CTGAN_ERR_BAD_DESC, "memset_zero: bad args");
    return cuda_status(cudaMemsetAsync(p, 0, (size_t)bytes, as_stream(stream)), "memset_zero");
}
extern "C" int ctgan_interpolate(const float* real, const float* fake, const float* alpha, float* out,
                                 int B, int P, void* stream) {
    CTGAN_REQUIRE(B > 0 && P > 0, CTGAN_ERR_BAD_DESC, "interpolate: bad shape");
    CTGAN_LAUNCH((interpolate_kernel), elementwise_grid((int64_t)B * P, 256), 256, 0, as_stream(stream), real, fake, alpha, out, B, P);
    CTGAN_CHECK_LAUNCH("interpolate");
    return 0;
}

extern "C" int ctgan_counter_add(uint64_t* counter, uint64_t delta, void* stream) {
    CTGAN_REQUIRE(counter != nullptr, CTGAN_ERR_BAD_DESC, "counter_add: null pointer");
    CTGAN_LAUNCH((counter_add_kernel), 1, 1, 0, as_stream(stream), counter, delta);
    CTGAN_CHECK_LAUNCH("counter_add");
    return 0;
}
extern "C" int ctgan_philox_uniform(float* out, int64_t n, float lo, float hi, uint64_t seed, uint64_t offset,
                                    const uint64_t* dyn_offset, void* stream) {
    if (n <= 0) return 0;
    CTGAN_LAUNCH((philox_uniform_kernel), elementwise_grid(n, 256), 256, 0, as_stream(stream), out, n, lo, hi, seed, offset, dyn_offset);
    CTGAN_CHECK_LAUNCH("philox_uniform");
    return 0;
}
extern "C" int ctgan_philox_normal(float* out, int64_t n, uint64_t seed, uint64_t offset, const uint64_t* dyn_offset, void* stream) {
    if (n <= 0) return 0;
    CTGAN_LAUNCH((philox_normal_kernel), elementwise_grid(n, 256), 256, 0, as_stream(stream), out, n, seed, offset, dyn_offset);
    CTGAN_CHECK_LAUNCH("philox_normal");
    return 0;
}
extern "C" int ctgan_philox_labels(int32_t* out, int64_t n, int n_labels, uint64_t seed, uint64_t offset,
                                   const uint64_t* dyn_offset, void* stream) {
    CTGAN_REQUIRE(n_labels > 0, CTGAN_ERR_BAD_DESC, "philox_labels: n_labels must be positive");
    if (n <= 0) return 0;
    CTGAN_LAUNCH((philox_labels_kernel), elementwise_grid(n, 256), 256, 0, as_stream(stream), out, n, n_labels, seed, offset, dyn_offset);
    CTGAN_CHECK_LAUNCH("philox_labels");
    return 0;
}
