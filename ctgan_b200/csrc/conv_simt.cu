// conv_simt.cu -- generic implicit-GEMM convolution kernels on the FP32 FMA pipes.
//
// These cover every shape the tcgen05 path (conv_tc.cu) does not: the 1/3-channel
// first and last layers, stride-2 5x5 DCGAN layers, GEMV-shaped linears, and the whole
// "fp32 path" (float activations).  One tiled kernel, three gather modes:
//   FPROP  C[pixel][co]      = sum_{(r,s,ci)} x[pixel shifted][ci] * w[(r,s,ci)][co]
//   DGRAD  C[in pixel][ci]   = sum_{(r,s,co)} dy[(pixel+pad-tap)/stride][co] * w[r,s,ci,co]
//   WGRAD  C[(r,s,ci)][co]   = sum_{pixel}    x[pixel shifted][ci] * dy[pixel][co]   (split over pixels)
// Replaces tf.nn.conv2d / Conv2DBackpropInput / Conv2DBackpropFilter / tf.matmul
// (TG/tflib/ops/conv2d.py:106-112, deconv2d.py:97-103, linear.py:132-136).
#include "common.cuh"

namespace ctgan {

namespace thin {   // conv_thin.cu: warp-per-pixel kernels for <= 4 channels on one side
int try_fprop(const ctgan_conv_desc* d, const void* x, const float* w, const float* bias, void* y, int flags,
              cudaStream_t st, int* rc);
int try_dgrad(const ctgan_conv_desc* d, const void* dy, const float* w, void* dx, cudaStream_t st, int* rc);
bool wgrad_ok(const ctgan_conv_desc* d, const void* x, const void* dy);
int try_wgrad(const ctgan_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st, int* rc);
}
namespace head {   // conv_head.cu: Linear(K -> N <= 16) on 2-D activations (the critic heads)
int try_fprop(const ctgan_conv_desc* d, const void* x, const float* w, const float* bias, void* y, int flags, cudaStream_t st, int* rc);
int try_dgrad(const ctgan_conv_desc* d, const void* dy, const float* w, void* dx, cudaStream_t st, int* rc);
int try_wgrad(const ctgan_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st, int* rc);
bool ok(const ctgan_conv_desc* d);
}


enum { MODE_FPROP = 0, MODE_DGRAD = 1, MODE_WGRAD = 2 };

struct Geom {
    int N, H, W, Cin, Ho, Wo, Cout, kh, kw, stride, pt, pl;
    int xdt, ydt;
};

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

template <int MODE>
__global__ void __launch_bounds__(NT)
igemm_simt_kernel(Geom g, const void* __restrict__ Asrc, const void* __restrict__ Bsrc,
                  void* __restrict__ out, const float* __restrict__ bias, int flags,
                  int M, int Ncol, int K, int k_per_split, int use_atomic)
{
    ctgan::pdl_entry();
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int kbeg = blockIdx.z * k_per_split;
    const int kend = min(K, kbeg + k_per_split);

    const int tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    // ---- per-thread fixed row decomposition for the A gather
    // FPROP/DGRAD: thread loads rows am[i] = tid/16 + 16*i at column tid%16
    // WGRAD:       thread loads row  tid%64 at columns tid/64 + 4*i
    int a_n[4], a_h[4], a_w[4];
    bool a_ok[4];
    if (MODE == MODE_FPROP || MODE == MODE_DGRAD) {
        const int SH = (MODE == MODE_FPROP) ? g.Ho : g.H;
        const int SW = (MODE == MODE_FPROP) ? g.Wo : g.W;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int m = m0 + (tid >> 4) + 16 * i;
            a_ok[i] = m < M;
            int mm = a_ok[i] ? m : 0;
            int w_ = mm % SW; int t = mm / SW;
            int h_ = t % SH;  int n_ = t / SH;
            a_n[i] = n_;
            if (MODE == MODE_FPROP) { a_h[i] = h_ * g.stride - g.pt; a_w[i] = w_ * g.stride - g.pl; }
            else                    { a_h[i] = h_ + g.pt;            a_w[i] = w_ + g.pl; }
        }
    } else {
        int m = m0 + (tid & 63);
        a_ok[0] = m < M;
        int mm = a_ok[0] ? m : 0;
        int ci = mm % g.Cin; int t = mm / g.Cin;
        int s = t % g.kw;    int r = t / g.kw;
        a_n[0] = ci; a_h[0] = r - g.pt; a_w[0] = s - g.pl;    // reuse slots: (ci, r-pt, s-pl)
    }

    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        // ------------------------------------------------ A tile -> As[k][m]
        if (MODE == MODE_FPROP) {
            int k = k0 + (tid & 15);
            bool kok = k < kend;
            int kk = kok ? k : 0;
            int ci = kk % g.Cin; int t = kk / g.Cin;
            int s = t % g.kw;    int r = t / g.kw;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float v = 0.f;
                int hh = a_h[i] + r, ww = a_w[i] + s;
                if (kok && a_ok[i] && hh >= 0 && hh < g.H && ww >= 0 && ww < g.W)
                    v = ld_act(Asrc, (((int64_t)a_n[i] * g.H + hh) * g.W + ww) * g.Cin + ci, g.xdt);
                As[tid & 15][(tid >> 4) + 16 * i] = v;
            }
        } else if (MODE == MODE_DGRAD) {
            int k = k0 + (tid & 15);
            bool kok = k < kend;
            int kk = kok ? k : 0;
            int co = kk % g.Cout; int t = kk / g.Cout;
            int s = t % g.kw;     int r = t / g.kw;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float v = 0.f;
                int hh = a_h[i] - r, ww = a_w[i] - s;
                if (kok && a_ok[i] && hh >= 0 && ww >= 0) {
                    int ho = hh / g.stride, wo = ww / g.stride;
                    if (ho * g.stride == hh && wo * g.stride == ww && ho < g.Ho && wo < g.Wo)
                        v = ld_act(Asrc, (((int64_t)a_n[i] * g.Ho + ho) * g.Wo + wo) * g.Cout + co, g.ydt);
                }
                As[tid & 15][(tid >> 4) + 16 * i] = v;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int kl = (tid >> 6) + 4 * i;
                int p = k0 + kl;
                float v = 0.f;
                if (p < kend && a_ok[0]) {
                    int wo = p % g.Wo; int t = p / g.Wo;
                    int ho = t % g.Ho; int n_ = t / g.Ho;
                    int hh = ho * g.stride + a_h[0], ww = wo * g.stride + a_w[0];
                    if (hh >= 0 && hh < g.H && ww >= 0 && ww < g.W)
                        v = ld_act(Asrc, (((int64_t)n_ * g.H + hh) * g.W + ww) * g.Cin + a_n[0], g.xdt);
                }
                As[kl][tid & 63] = v;
            }
        }
        // ------------------------------------------------ B tile -> Bs[k][n]
        if (MODE == MODE_FPROP) {
            const float* wsrc = reinterpret_cast<const float*>(Bsrc);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int kl = (tid >> 6) + 4 * i;
                int k = k0 + kl, n = n0 + (tid & 63);
                Bs[kl][tid & 63] = (k < kend && n < Ncol) ? wsrc[(int64_t)k * g.Cout + n] : 0.f;
            }
        } else if (MODE == MODE_DGRAD) {
            const float* wsrc = reinterpret_cast<const float*>(Bsrc);
            int k = k0 + (tid & 15);
            bool kok = k < kend;
            int kk = kok ? k : 0;
            int co = kk % g.Cout; int t = kk / g.Cout;        // t = r*kw + s
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int nl = (tid >> 4) + 16 * i;
                int n = n0 + nl;                               // ci
                Bs[tid & 15][nl] = (kok && n < Ncol) ? wsrc[((int64_t)t * g.Cin + n) * g.Cout + co] : 0.f;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int kl = (tid >> 6) + 4 * i;
                int p = k0 + kl, n = n0 + (tid & 63);
                Bs[kl][tid & 63] = (p < kend && n < Ncol) ? ld_act(Bsrc, (int64_t)p * g.Cout + n, g.ydt) : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ------------------------------------------------ epilogue
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= Ncol) continue;
            float v = acc[i][j];
            int64_t o = (int64_t)m * Ncol + n;
            if (MODE == MODE_FPROP) {
                if (bias) v += bias[n];
                if (flags & CTGAN_EPI_RELU) v = fmaxf(v, 0.f);
                st_act(out, o, g.ydt, v);
            } else if (MODE == MODE_DGRAD) {
                st_act(out, o, g.xdt, v);
            } else {
                float* dw = reinterpret_cast<float*>(out);
                if (use_atomic) atomicAdd(dw + o, v);
                else            dw[o] = v;
            }
        }
    }
}

static int check_desc(const ctgan_conv_desc* d) {
    CTGAN_REQUIRE(d != nullptr, CTGAN_ERR_BAD_DESC, "conv: null descriptor");
    CTGAN_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Ho > 0 && d->Wo > 0 && d->Cout > 0 &&
                  d->kh > 0 && d->kw > 0 && d->stride > 0 && d->pad_t >= 0 && d->pad_l >= 0,
                  CTGAN_ERR_BAD_DESC, "conv: non-positive dimension in descriptor");
    CTGAN_REQUIRE(dtype_ok(d->x_dtype) && dtype_ok(d->y_dtype), CTGAN_ERR_BAD_DESC, "conv: bad dtype");
    // every output position must read at least partly defined input rows: (Ho-1)*s - pad < H
    CTGAN_REQUIRE((d->Ho - 1) * d->stride - d->pad_t < d->H && (d->Wo - 1) * d->stride - d->pad_l < d->W,
                  CTGAN_ERR_BAD_DESC, "conv: output extent inconsistent with input/stride/pad");
    return 0;
}

static Geom make_geom(const ctgan_conv_desc* d) {
    Geom g;
    g.N = d->N; g.H = d->H; g.W = d->W; g.Cin = d->Cin; g.Ho = d->Ho; g.Wo = d->Wo; g.Cout = d->Cout;
    g.kh = d->kh; g.kw = d->kw; g.stride = d->stride; g.pt = d->pad_t; g.pl = d->pad_l;
    g.xdt = d->x_dtype; g.ydt = d->y_dtype;
    return g;
}

}  // namespace ctgan

using namespace ctgan;

extern "C" int ctgan_conv_fprop(const ctgan_conv_desc* d, const void* x, const float* w,
                                const float* bias, void* y, int flags, void* stream) {
    if (int r = check_desc(d)) return r;
    CTGAN_REQUIRE(x && w && y, CTGAN_ERR_BAD_DESC, "conv_fprop: null pointer");
    { int rc = 0; if (head::try_fprop(d, x, w, bias, y, flags, as_stream(stream), &rc)) return rc; }
    { int rc = 0; if (thin::try_fprop(d, x, w, bias, y, flags, as_stream(stream), &rc)) return rc; }
    Geom g = make_geom(d);
    int64_t M64 = (int64_t)d->N * d->Ho * d->Wo;
    CTGAN_REQUIRE(M64 < (1ll << 31), CTGAN_ERR_UNSUPPORTED, "conv_fprop: too many output pixels");
    int M = (int)M64, Ncol = d->Cout, K = d->kh * d->kw * d->Cin;
    dim3 grid(ceil_div(M, BM), ceil_div(Ncol, BN), 1);
    CTGAN_LAUNCH((igemm_simt_kernel<MODE_FPROP>), grid, NT, 0, as_stream(stream), g, x, w, y, bias, flags, M, Ncol, K, K, 0);
    CTGAN_CHECK_LAUNCH("conv_fprop");
    return 0;
}

extern "C" int ctgan_conv_dgrad(const ctgan_conv_desc* d, const void* dy, const float* w,
                                void* dx, void* stream) {
    if (int r = check_desc(d)) return r;
    CTGAN_REQUIRE(dy && w && dx, CTGAN_ERR_BAD_DESC, "conv_dgrad: null pointer");
    { int rc = 0; if (head::try_dgrad(d, dy, w, dx, as_stream(stream), &rc)) return rc; }
    { int rc = 0; if (thin::try_dgrad(d, dy, w, dx, as_stream(stream), &rc)) return rc; }
    Geom g = make_geom(d);
    int64_t M64 = (int64_t)d->N * d->H * d->W;
    CTGAN_REQUIRE(M64 < (1ll << 31), CTGAN_ERR_UNSUPPORTED, "conv_dgrad: too many input pixels");
    int M = (int)M64, Ncol = d->Cin, K = d->kh * d->kw * d->Cout;
    dim3 grid(ceil_div(M, BM), ceil_div(Ncol, BN), 1);
    CTGAN_LAUNCH((igemm_simt_kernel<MODE_DGRAD>), grid, NT, 0, as_stream(stream), g, dy, w, dx, nullptr, 0, M, Ncol, K, K, 0);
    CTGAN_CHECK_LAUNCH("conv_dgrad");
    return 0;
}

extern "C" int ctgan_conv_wgrad(const ctgan_conv_desc* d, const void* x, const void* dy,
                                float* dw, int accumulate, void* stream) {
    if (int r = check_desc(d)) return r;
    CTGAN_REQUIRE(x && dy && dw, CTGAN_ERR_BAD_DESC, "conv_wgrad: null pointer");
    if (head::ok(d)) {
        cudaStream_t ts = as_stream(stream);
        if (!accumulate) {
            cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d->Cin * d->Cout, ts);
            if (e != cudaSuccess) return cuda_status(e, "conv_wgrad memset");
        }
        int rc = 0;
        if (head::try_wgrad(d, x, dy, dw, ts, &rc)) return rc;
    }
    if (thin::wgrad_ok(d, x, dy)) {
        cudaStream_t ts = as_stream(stream);
        if (!accumulate) {
            cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d->kh * d->kw * d->Cin * d->Cout, ts);
            if (e != cudaSuccess) return cuda_status(e, "conv_wgrad memset");
        }
        int rc = 0;
        if (thin::try_wgrad(d, x, dy, dw, ts, &rc)) return rc;
    }
    Geom g = make_geom(d);
    int64_t P64 = (int64_t)d->N * d->Ho * d->Wo;
    CTGAN_REQUIRE(P64 < (1ll << 31), CTGAN_ERR_UNSUPPORTED, "conv_wgrad: too many output pixels");
    int M = d->kh * d->kw * d->Cin, Ncol = d->Cout, K = (int)P64;
    int tiles = ceil_div(M, BM) * ceil_div(Ncol, BN);
    // split the pixel reduction so that ~2 waves of CTAs exist, each with >= 8 k-steps
    int want = (2 * sm_count() + tiles - 1) / tiles;
    int max_split = K / (8 * BK); if (max_split < 1) max_split = 1;
    int split = want < max_split ? want : max_split;
    if (split < 1) split = 1;
    int k_per_split = ceil_div(ceil_div(K, split), BK) * BK;
    split = ceil_div(K, k_per_split);
    int use_atomic = (split > 1 || accumulate) ? 1 : 0;
    cudaStream_t st = as_stream(stream);
    if (use_atomic && !accumulate) {
        cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)M * Ncol, st);
        if (e != cudaSuccess) return cuda_status(e, "conv_wgrad memset");
    }
    dim3 grid(ceil_div(M, BM), ceil_div(Ncol, BN), split);
    CTGAN_LAUNCH((igemm_simt_kernel<MODE_WGRAD>), grid, NT, 0, st, g, x, dy, dw, nullptr, 0, M, Ncol, K, k_per_split, use_atomic);
    CTGAN_CHECK_LAUNCH("conv_wgrad");
    return 0;
}
