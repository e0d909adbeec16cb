// conv_thin.cu -- convolution kernels for layers with <= 4 channels on one side.
//
// The image-facing layers of the three scripts (Discriminator.1 / Discriminator.1.Conv1 /
// Discriminator.1.Shortcut read 1-3 channels; Generator.Output / Generator.5 write 1-3) have
// K = 27..75 or N = 1..3: GEMM tiles are >90 % padding there and the work is L2/HBM-bound (the
// 128-channel side is read or written once).  Warp-per-pixel kernels replace the tiled implicit
// GEMM for them.  In all of them a lane owns 4 consecutive channels of the WIDE side (one
// 8/16-byte load), all taps of a pixel are loaded before any arithmetic (memory-level
// parallelism instead of a dependent chain), and the filter tap geometry is a template
// parameter so no division appears in an inner loop.
//   thin_contract      wide -> thin : fprop with Cout <= 4, dgrad with Cin <= 4
//                      filters resident in shared memory, warp-shuffle reduction per pixel
//   thin_wgrad_wide_x  Cout <= 3    : dW[tap][ci][o]  = sum_pix x[pix+tap][ci] * dy[pix][o]
//   thin_wgrad_wide_dy Cin  <= 4    : dW[(tap,c)][co] = sum_pix x[pix+tap][c]  * dy[pix][co]
//                      lane k gathers patch element k once per pixel and the warp broadcasts it
//                      with shuffles; accumulators live in registers over the warp's pixel range,
//                      then shared-memory and global fp32 atomics.
// Selected by ctgan_conv_fprop/dgrad/wgrad (conv_simt.cu) when the shape qualifies.
#include "common.cuh"

namespace ctgan {
namespace thin {

struct Geom {
    int N, H, W, Cin, Ho, Wo, Cout, kh, kw, stride, pt, pl;
};

template <typename T> __device__ __forceinline__ void ld4(const T* p, float (&v)[4]);
template <> __device__ __forceinline__ void ld4<float>(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void ld4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x), b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
    v[0] = __bfloat162float(a.x); v[1] = __bfloat162float(a.y); v[2] = __bfloat162float(b.x); v[3] = __bfloat162float(b.y);
}

// ------------------------------------------------------------------ wide -> thin
// FPROP (DG == false): out[n,p,q,o<T] = bias[o] + sum_{r,s,c} x[n, p*S+r-pt, q*S+s-pl, c] * w[r,s,c,o]
// DGRAD (DG == true):  out[n,h,w,c<T] = sum_{r,s,o} dy[n, (h+pt-r)/S, (w+pl-s)/S, o] * w[r,s,c,o]
// `wide` = Cin (fprop) / Cout (dgrad).  Shared memory holds the filter as float4 ws[tap][j][wide/4]
// = (tap, wide channel 4*g+j) -> the <=4 thin-channel weights.  KS = filter size, S = stride.
template <typename TA, bool DG, int KS, int S>
__global__ void __launch_bounds__(256)
thin_contract_kernel(Geom g, const TA* __restrict__ src, const float* __restrict__ w, const float* __restrict__ bias,
                     TA* __restrict__ out, int relu)
{
    ctgan::pdl_entry();
    constexpr int TAPS = KS * KS;
    extern __shared__ float4 ws4[];
    const int wide = DG ? g.Cout : g.Cin;
    const int T = DG ? g.Cin : g.Cout;
    const int wg = wide / 4;
    for (int e = threadIdx.x; e < TAPS * wide; e += blockDim.x) {
        int tap = e / wide, wc = e - tap * wide;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        for (int t = 0; t < T; ++t) {
            int c = DG ? t : wc, o = DG ? wc : t;
            v[t] = w[((int64_t)tap * g.Cin + c) * g.Cout + o];
        }
        ws4[(tap * 4 + (wc & 3)) * wg + (wc >> 2)] = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int OH = DG ? g.H : g.Ho, OW = DG ? g.W : g.Wo;      // output spatial extent
    const int SH = DG ? g.Ho : g.H, SW = DG ? g.Wo : g.W;      // source spatial extent
    const int64_t npix = (int64_t)g.N * OH * OW;
    for (int64_t pix = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); pix < npix;
         pix += (int64_t)gridDim.x * warps_per_block) {
        const int q = (int)(pix % OW); const int64_t t_ = pix / OW;
        const int p = (int)(t_ % OH); const int n = (int)(t_ / OH);
        const TA* img = src + (int64_t)n * SH * SW * wide;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c4 = lane; c4 < wg; c4 += 32) {
            float v[TAPS][4];
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap) {
                const int r = tap / KS, s = tap % KS;
                int hh, ww; bool ok;
                if (DG) {
                    const int a = p + g.pt - r, b = q + g.pl - s;
                    ok = a >= 0 && b >= 0 && (a % S) == 0 && (b % S) == 0;
                    hh = a / S; ww = b / S;
                } else {
                    hh = p * S + r - g.pt; ww = q * S + s - g.pl;
                    ok = hh >= 0 && ww >= 0;
                }
                ok = ok && hh < SH && ww < SW;
                if (ok) ld4<TA>(img + ((int64_t)hh * SW + ww) * wide + c4 * 4, v[tap]);
                else { v[tap][0] = v[tap][1] = v[tap][2] = v[tap][3] = 0.f; }
            }
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap) {
                const float4* wt = ws4 + (size_t)tap * 4 * wg + c4;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 f = wt[j * wg];
                    acc[0] = fmaf(v[tap][j], f.x, acc[0]); acc[1] = fmaf(v[tap][j], f.y, acc[1]);
                    acc[2] = fmaf(v[tap][j], f.z, acc[2]); acc[3] = fmaf(v[tap][j], f.w, acc[3]);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int t = 0; t < 4; ++t) acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], o);
        }
        if (lane < T) {
            float v = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
            if (!DG && bias) v += bias[lane];
            if (relu) v = fmaxf(v, 0.f);
            out[pix * T + lane] = from_f<TA>(v);
        }
    }
}

// ------------------------------------------------------------------ wgrad, Cout <= 3 (wide = Cin)
// dW[tap][ci][o] += sum_pix x[n, p*S+r-pt, q*S+s-pl, ci] * dy[n,p,q,o]
template <typename TA, int KS>
__global__ void __launch_bounds__(256)
thin_wgrad_wide_x_kernel(Geom g, const TA* __restrict__ x, const TA* __restrict__ dy, float* __restrict__ dw)
{
    ctgan::pdl_entry();
    constexpr int TAPS = KS * KS;
    extern __shared__ float sdw[];                  // [TAPS][Cin][4]
    const int T = g.Cout, wide = g.Cin, wg = wide / 4, S = g.stride;
    for (int e = threadIdx.x; e < TAPS * wide * 4; e += blockDim.x) sdw[e] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int64_t npix = (int64_t)g.N * g.Ho * g.Wo;
    for (int c4 = lane; c4 < wg; c4 += 32) {
        float acc[TAPS][4][3];
#pragma unroll
        for (int t = 0; t < TAPS; ++t)
#pragma unroll
            for (int j = 0; j < 4; ++j) { acc[t][j][0] = acc[t][j][1] = acc[t][j][2] = 0.f; }
        for (int64_t pix = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); pix < npix;
             pix += (int64_t)gridDim.x * warps_per_block) {
            const int q = (int)(pix % g.Wo); const int64_t t_ = pix / g.Wo;
            const int p = (int)(t_ % g.Ho); const int n = (int)(t_ / g.Ho);
            const TA* img = x + (int64_t)n * g.H * g.W * wide + c4 * 4;
            float d[3] = {0.f, 0.f, 0.f};
            for (int t = 0; t < T; ++t) d[t] = to_f<TA>(dy[pix * T + t]);
            float v[TAPS][4];
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap) {
                const int hh = p * S + tap / KS - g.pt, ww = q * S + tap % KS - g.pl;
                if (hh >= 0 && hh < g.H && ww >= 0 && ww < g.W) ld4<TA>(img + ((int64_t)hh * g.W + ww) * wide, v[tap]);
                else { v[tap][0] = v[tap][1] = v[tap][2] = v[tap][3] = 0.f; }
            }
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[tap][j][0] = fmaf(v[tap][j], d[0], acc[tap][j][0]);
                    acc[tap][j][1] = fmaf(v[tap][j], d[1], acc[tap][j][1]);
                    acc[tap][j][2] = fmaf(v[tap][j], d[2], acc[tap][j][2]);
                }
        }
#pragma unroll
        for (int tap = 0; tap < TAPS; ++tap)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float* dst = sdw + ((size_t)tap * wide + c4 * 4 + j) * 4;
                atomicAdd(dst + 0, acc[tap][j][0]); atomicAdd(dst + 1, acc[tap][j][1]); atomicAdd(dst + 2, acc[tap][j][2]);
            }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < TAPS * wide; e += blockDim.x)
        for (int t = 0; t < T; ++t) atomicAdd(dw + (int64_t)e * T + t, sdw[e * 4 + t]);
}

// ------------------------------------------------------------------ wgrad, Cin <= 4 (wide = Cout)
// dW[row=(tap,c)][co] += sum_pix x[n, p*S+r-pt, q*S+s-pl, c] * dy[n,p,q,co].
// blockIdx.y selects a group of 32 rows: lane k owns patch row row0+k (its (r,s,c) is decoded once),
// gathers it per pixel, and the warp broadcasts the 32 values with shuffles.
template <typename TA>
__global__ void __launch_bounds__(256)
thin_wgrad_wide_dy_kernel(Geom g, const TA* __restrict__ x, const TA* __restrict__ dy, float* __restrict__ dw)
{
    ctgan::pdl_entry();
    constexpr int ROWS = 32;
    extern __shared__ float sdw[];                  // [ROWS][Cout]
    const int T = g.Cin, wide = g.Cout, wg = wide / 4, S = g.stride;
    const int KT = g.kh * g.kw * T;
    const int row0 = blockIdx.y * ROWS;
    const int nrows = min(ROWS, KT - row0);
    for (int e = threadIdx.x; e < ROWS * wide; e += blockDim.x) sdw[e] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int64_t npix = (int64_t)g.N * g.Ho * g.Wo;
    // this lane's patch row
    const bool row_ok = lane < nrows;
    const int my_row = row_ok ? row0 + lane : 0;
    const int my_c = my_row % T, my_tap = my_row / T;
    const int my_r = my_tap / g.kw - g.pt, my_s = my_tap % g.kw - g.pl;
    for (int c4base = 0; c4base < wg; c4base += 32) {
        const int c4 = c4base + lane;
        const bool col_ok = c4 < wg;
        float acc[ROWS][4];
#pragma unroll
        for (int k = 0; k < ROWS; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f;
        for (int64_t pix = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); pix < npix;
             pix += (int64_t)gridDim.x * warps_per_block) {
            const int q = (int)(pix % g.Wo); const int64_t t_ = pix / g.Wo;
            const int p = (int)(t_ % g.Ho); const int n = (int)(t_ / g.Ho);
            float d[4] = {0.f, 0.f, 0.f, 0.f};
            if (col_ok) ld4<TA>(dy + pix * wide + c4 * 4, d);
            const int hh = p * S + my_r, ww = q * S + my_s;
            float mine = 0.f;
            if (row_ok && hh >= 0 && hh < g.H && ww >= 0 && ww < g.W)
                mine = to_f<TA>(x[(((int64_t)n * g.H + hh) * g.W + ww) * T + my_c]);
#pragma unroll
            for (int k = 0; k < ROWS; ++k) {
                const float v = __shfl_sync(0xffffffffu, mine, k);
                acc[k][0] = fmaf(v, d[0], acc[k][0]); acc[k][1] = fmaf(v, d[1], acc[k][1]);
                acc[k][2] = fmaf(v, d[2], acc[k][2]); acc[k][3] = fmaf(v, d[3], acc[k][3]);
            }
        }
        if (col_ok) {
#pragma unroll
            for (int k = 0; k < ROWS; ++k) {
                if (k < nrows) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) atomicAdd(sdw + (size_t)k * wide + c4 * 4 + j, acc[k][j]);
                }
            }
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nrows * wide; e += blockDim.x)
        atomicAdd(dw + (int64_t)row0 * wide + e, sdw[e]);
}

static Geom make_geom(const ctgan_conv_desc* d) {
    Geom g;
    g.N = d->N; g.H = d->H; g.W = d->W; g.Cin = d->Cin; g.Ho = d->Ho; g.Wo = d->Wo; g.Cout = d->Cout;
    g.kh = d->kh; g.kw = d->kw; g.stride = d->stride; g.pt = d->pad_t; g.pl = d->pad_l;
    return g;
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// grid for warp-per-pixel kernels: 8 warps per CTA, `ctas_per_sm` CTAs per SM, never more warps than pixels
static int pixel_grid(int64_t npix, int ctas_per_sm) {
    int64_t want = (npix + 7) / 8;
    int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

static bool tap_geometry_ok(const ctgan_conv_desc* d) {
    if (d->kh != d->kw) return false;
    return (d->kh == 1 && d->stride == 1) || (d->kh == 3 && d->stride == 1) || (d->kh == 5 && d->stride == 2);
}

template <typename TA, bool DG, int KS, int S>
static int launch_contract_t(const Geom& g, int wide, int64_t npix, const void* src, const float* w, const float* bias,
                             void* out, int relu, cudaStream_t st) {
    size_t smem = (size_t)KS * KS * wide * 16;
    static bool set = false;
    if (!set) {
        cudaError_t e = cudaFuncSetAttribute(thin_contract_kernel<TA, DG, KS, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
        if (e != cudaSuccess) return cuda_status(e, "thin_contract smem attribute");
        set = true;
    }
    CTGAN_LAUNCH((thin_contract_kernel<TA, DG, KS, S>), pixel_grid(npix, KS == 5 ? 2 : 4), 256, smem, st, g, (const TA*)src, w, bias, (TA*)out, relu);
    CTGAN_CHECK_LAUNCH("thin_contract");
    return 0;
}

template <typename TA, bool DG>
static int launch_contract(const ctgan_conv_desc* d, const void* src, const float* w, const float* bias,
                           void* out, int relu, cudaStream_t st) {
    Geom g = make_geom(d);
    const int wide = DG ? d->Cout : d->Cin;
    const int64_t npix = (int64_t)d->N * (DG ? d->H : d->Ho) * (DG ? d->W : d->Wo);
    if (d->kh == 1) return launch_contract_t<TA, DG, 1, 1>(g, wide, npix, src, w, bias, out, relu, st);
    if (d->kh == 3) return launch_contract_t<TA, DG, 3, 1>(g, wide, npix, src, w, bias, out, relu, st);
    return launch_contract_t<TA, DG, 5, 2>(g, wide, npix, src, w, bias, out, relu, st);
}

// each try_* returns 1 when it handled the call (*rc = status), 0 when the shape does not qualify
int try_fprop(const ctgan_conv_desc* d, const void* x, const float* w, const float* bias, void* y, int flags,
              cudaStream_t st, int* rc) {
    if (!(d->Cout <= 4 && d->Cin % 4 == 0 && d->Cin >= 16 && d->x_dtype == d->y_dtype && al16(x) && tap_geometry_ok(d) &&
          (size_t)d->kh * d->kw * d->Cin * 16 <= 98304))
        return 0;
    *rc = d->x_dtype == CTGAN_F32 ? launch_contract<float, false>(d, x, w, bias, y, flags & CTGAN_EPI_RELU, st)
                                  : launch_contract<__nv_bfloat16, false>(d, x, w, bias, y, flags & CTGAN_EPI_RELU, st);
    return 1;
}

int try_dgrad(const ctgan_conv_desc* d, const void* dy, const float* w, void* dx, cudaStream_t st, int* rc) {
    if (!(d->Cin <= 4 && d->Cout % 4 == 0 && d->Cout >= 16 && d->x_dtype == d->y_dtype && al16(dy) && tap_geometry_ok(d) &&
          (size_t)d->kh * d->kw * d->Cout * 16 <= 98304))
        return 0;
    *rc = d->x_dtype == CTGAN_F32 ? launch_contract<float, true>(d, dy, w, nullptr, dx, 0, st)
                                  : launch_contract<__nv_bfloat16, true>(d, dy, w, nullptr, dx, 0, st);
    return 1;
}

template <typename TA, int KS>
static int launch_wgrad_wide_x_t(const ctgan_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st) {
    Geom g = make_geom(d);
    size_t smem = (size_t)KS * KS * d->Cin * 16;
    static bool set = false;
    if (!set) {
        cudaError_t e = cudaFuncSetAttribute(thin_wgrad_wide_x_kernel<TA, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
        if (e != cudaSuccess) return cuda_status(e, "thin_wgrad_wide_x smem attribute");
        set = true;
    }
    CTGAN_LAUNCH((thin_wgrad_wide_x_kernel<TA, KS>), pixel_grid((int64_t)d->N * d->Ho * d->Wo, 1), 256, smem, st, g, (const TA*)x, (const TA*)dy, dw);
    CTGAN_CHECK_LAUNCH("thin_wgrad_wide_x");
    return 0;
}

template <typename TA>
static int launch_wgrad_wide_dy(const ctgan_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st) {
    Geom g = make_geom(d);
    const int KT = d->kh * d->kw * d->Cin;
    size_t smem = (size_t)32 * d->Cout * 4;
    dim3 grid(pixel_grid((int64_t)d->N * d->Ho * d->Wo, 2), (KT + 31) / 32);
    static bool set = false;
    if (!set) {
        cudaError_t e = cudaFuncSetAttribute(thin_wgrad_wide_dy_kernel<TA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
        if (e != cudaSuccess) return cuda_status(e, "thin_wgrad_wide_dy smem attribute");
        set = true;
    }
    CTGAN_LAUNCH((thin_wgrad_wide_dy_kernel<TA>), grid, 256, smem, st, g, (const TA*)x, (const TA*)dy, dw);
    CTGAN_CHECK_LAUNCH("thin_wgrad_wide_dy");
    return 0;
}

static bool wide_x_ok(const ctgan_conv_desc* d, const void* x) {
    return d->Cout <= 3 && d->Cin % 4 == 0 && d->Cin >= 16 && d->kh == d->kw && (d->kh == 1 || d->kh == 3) && al16(x) &&
           (size_t)d->kh * d->kw * d->Cin * 16 <= 98304;
}
static bool wide_dy_ok(const ctgan_conv_desc* d, const void* dy) {
    return d->Cin <= 4 && d->Cout % 4 == 0 && d->Cout >= 16 && al16(dy) && (size_t)32 * d->Cout * 4 <= 98304;
}

bool wgrad_ok(const ctgan_conv_desc* d, const void* x, const void* dy) {
    return d->x_dtype == d->y_dtype && (wide_x_ok(d, x) || wide_dy_ok(d, dy));
}

// dw must already hold the value to accumulate onto (the caller zeroes it for accumulate == 0)
int try_wgrad(const ctgan_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st, int* rc) {
    if (d->x_dtype != d->y_dtype) return 0;
    const bool f32 = d->x_dtype == CTGAN_F32;
    if (wide_x_ok(d, x)) {
        if (d->kh == 1) *rc = f32 ? launch_wgrad_wide_x_t<float, 1>(d, x, dy, dw, st) : launch_wgrad_wide_x_t<__nv_bfloat16, 1>(d, x, dy, dw, st);
        else            *rc = f32 ? launch_wgrad_wide_x_t<float, 3>(d, x, dy, dw, st) : launch_wgrad_wide_x_t<__nv_bfloat16, 3>(d, x, dy, dw, st);
        return 1;
    }
    if (wide_dy_ok(d, dy)) {
        *rc = f32 ? launch_wgrad_wide_dy<float>(d, x, dy, dw, st) : launch_wgrad_wide_dy<__nv_bfloat16>(d, x, dy, dw, st);
        return 1;
    }
    return 0;
}

}  // namespace thin
}  // namespace ctgan
