// conv_s2d.cu -- the stride-2 5x5 'SAME' convolutions of the DCGAN-style critics / generators
// (TG/CT_gan_cifar.py:84-94, TG/CT_gan_mnist.py:92-102, and as their dgrad the Deconv2D layers
// TG/tflib/ops/deconv2d.py:97-103) re-expressed so that they run on the stride-1 tcgen05 kernels of conv_tc.cu.
//
// A stride-2 correlation reads x at rows 2p + r - pad_t.  Writing the row as 2*(p + b) + dy (b = block offset,
// dy = parity) turns it into a stride-1 correlation over the SPACE-TO-DEPTH image
//     xs[n, i, j, (dy*2 + dx)*C + c] = x[n, 2i + dy, 2j + dx, c]                  (zero where 2i+dy >= H or 2j+dx >= W)
// with the embedded filter
//     W3[R, S, (dy*2 + dx)*C + c, o] = w[2(R-1) + dy + pad_t, 2(S-1) + dx + pad_l, c, o]   (zero outside the k x k taps)
// For k = 5 and TF-SAME padding (pad 1 on an even extent, 2 on an odd one) the block offsets are exactly {-1, 0, +1}:
// a 3x3, pad-1, stride-1 conv with 4C input channels -- 36/25 of the FLOPs, on tensor cores instead of FP32 FMAs.
// dgrad = depth_to_space(dgrad3x3(dy, W3)); wgrad = the gather below of wgrad3x3(xs, dy).
//
// The kernels here are pure data movement (HBM-bound, 16-byte accesses on the activations).
#include "common.cuh"

namespace ctgan {
namespace {

// x [N,H,W,C] -> xs [N,Hs,Ws,4C]; VEC elements of T per thread (C % VEC == 0)
// mul (nullable): a multiplier in the SPACE-TO-DEPTH layout applied to every element moved (the dropout / LeakyReLU
// multiplier a fused conv epilogue stored in that layout: backward = depth_to_space(g * m), its adjoint = space_to_depth(c) * m)
template <typename T, int VEC, bool FWD>
__global__ void s2d_kernel(const T* __restrict__ src, const T* __restrict__ mul, T* __restrict__ dst, int N, int H, int W, int C,
                           int Hs, int Ws, int relu_pattern) {
    pdl_entry();
    struct alignas(sizeof(T) * VEC) Vec { T v[VEC]; };
    const int cv = C / VEC;
    if (FWD) {
        // one thread per (n, i, j, q, c-vector) of xs
        const int64_t total = (int64_t)N * Hs * Ws * 4 * cv;
        for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
            int c = (int)(idx % cv); int64_t t = idx / cv;
            int q = (int)(t & 3); t >>= 2;
            int j = (int)(t % Ws); t /= Ws;
            int i = (int)(t % Hs); int n = (int)(t / Hs);
            const int h = 2 * i + (q >> 1), w = 2 * j + (q & 1);
            Vec val;
            if (h < H && w < W) {
                val = reinterpret_cast<const Vec*>(src)[(((int64_t)n * H + h) * W + w) * cv + c];
            } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) val.v[e] = from_f<T>(0.f);
            }
            if (mul) {
                const Vec m = reinterpret_cast<const Vec*>(mul)[idx];
#pragma unroll
                for (int e = 0; e < VEC; ++e)
                    val.v[e] = relu_pattern ? (to_f<T>(m.v[e]) > 0.f ? val.v[e] : from_f<T>(0.f))
                                            : from_f<T>(to_f<T>(val.v[e]) * to_f<T>(m.v[e]));
            }
            reinterpret_cast<Vec*>(dst)[idx] = val;
        }
    } else {
        // depth_to_space (+ crop to H x W): one thread per (n, h, w, c-vector) of x
        const int64_t total = (int64_t)N * H * W * cv;
        for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
            int c = (int)(idx % cv); int64_t t = idx / cv;
            int w = (int)(t % W); t /= W;
            int h = (int)(t % H); int n = (int)(t / H);
            const int q = ((h & 1) << 1) | (w & 1);
            const int64_t sidx = ((((int64_t)n * Hs + (h >> 1)) * Ws + (w >> 1)) * 4 + q) * cv + c;
            Vec val = reinterpret_cast<const Vec*>(src)[sidx];
            if (mul) {
                const Vec m = reinterpret_cast<const Vec*>(mul)[sidx];
#pragma unroll
                for (int e = 0; e < VEC; ++e)
                    val.v[e] = relu_pattern ? (to_f<T>(m.v[e]) > 0.f ? val.v[e] : from_f<T>(0.f))
                                            : from_f<T>(to_f<T>(val.v[e]) * to_f<T>(m.v[e]));
            }
            reinterpret_cast<Vec*>(dst)[idx] = val;
        }
    }
}

// source tap of embedded tap R (0..2) and parity d: r = 2(R-1) + d + pad; valid when 0 <= r < k
__device__ __forceinline__ int s2d_src_tap(int R, int d, int pad) { return 2 * (R - 1) + d + pad; }

// w HWIO [k][k][C][O] float -> bf16 operands of the 3x3 conv with 4C input channels:
//   wp_f [T][O][4C]           (fprop: rows = output channels, K = input channels)         T = R*3 + S
//   wp_d [8-T][4C][O]         (dgrad: tap-flipped, rows = input channels, K = output channels)
// One CTA = one 32 (kc) x 32 (o) tile of one embedded tap: w is read and wp_d written along o, wp_f written along kc
// through a shared-memory transpose, so that all three streams are coalesced.
__global__ void __launch_bounds__(256)
pack_filter_s2d_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp_f,
                       __nv_bfloat16* __restrict__ wp_d, int k, int C, int O, int pad_t, int pad_l) {
    pdl_entry();
    __shared__ __nv_bfloat16 tile[32][33];
    const int C4 = 4 * C;
    const int T = blockIdx.z, R = T / 3, S = T - 3 * R;
    const int kc0 = blockIdx.y * 32, o0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           // 32 x 8
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int kc = kc0 + ty + 8 * j, o = o0 + tx;
        float v = 0.f;
        if (kc < C4 && o < O) {
            const int q = kc / C, c = kc - q * C;
            const int r = s2d_src_tap(R, q >> 1, pad_t), s = s2d_src_tap(S, q & 1, pad_l);
            if (r >= 0 && r < k && s >= 0 && s < k) v = w[(((int64_t)r * k + s) * C + c) * O + o];
        }
        const __nv_bfloat16 b = __float2bfloat16_rn(v);
        tile[ty + 8 * j][tx] = b;
        if (wp_d && kc < C4 && o < O) wp_d[((int64_t)(8 - T) * C4 + kc) * O + o] = b;
    }
    __syncthreads();
    if (wp_f) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = o0 + ty + 8 * j, kc = kc0 + tx;
            if (kc < C4 && o < O) wp_f[((int64_t)T * O + o) * C4 + kc] = tile[tx][ty + 8 * j];
        }
    }
}

// dw HWIO [k][k][C][O] (+)= the entries of dW3 HWIO [3][3][4C][O] that carry filter taps (the adjoint of the embedding)
__global__ void s2d_filter_grad_kernel(const float* __restrict__ dw3, float* __restrict__ dw, int k, int C, int O,
                                       int pad_t, int pad_l, int accumulate) {
    pdl_entry();
    const int64_t total = (int64_t)k * k * C * O;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int o = (int)(i % O); int64_t t = i / O;
        int c = (int)(t % C); t /= C;
        int s = (int)(t % k); int r = (int)(t / k);
        // r - pad = 2*(R-1) + dy with dy in {0,1}:  R = floor((r - pad) / 2) + 1
        const int ar = r - pad_t + 2, as = s - pad_l + 2;          // >= 0 for pad <= 2
        const int R = ar >> 1, dy = ar & 1, S = as >> 1, dx = as & 1;
        const float v = dw3[(((int64_t)(R * 3 + S) * 4 + (dy * 2 + dx)) * C + c) * O + o];
        if (accumulate) atomicAdd(dw + i, v); else dw[i] = v;
    }
}


// ------------------------------------------------------------------ stride-2 convs with a thin (<= 8 channel) input
// Discriminator.1 (3 -> DIM, 5x5/2: TG/CT_gan_cifar.py:84, 1 -> DIM: TG/CT_gan_mnist.py:92) and -- as the dgrad of that
// geometry -- the generators' last Deconv2D (DIM -> 3 / 1: TG/CT_gan_cifar.py:75, TG/CT_gan_mnist.py:83).  K = taps*C =
// 75 (25) fits one 128-column im2col row per OUTPUT pixel, so the family becomes 1x1 tensor-core GEMMs:
//   fprop: y    = col x W128,           col[p][(r*kw+s)*C + c] = x[n, st*ho + r - pad_t, st*wo + s - pad_l, c]   (zero outside / k >= taps*C)
//   dgrad: dx   = col2im(dy x W128^T)   dx[n,h,w,c] = sum over taps with (h + pad_t - r) = st*ho of dcol[(n,ho,wo)][(r*kw+s)*C + c]
//   wgrad: dW   = first taps*C rows of col^T x dy  (HWIO order == the column order)
__global__ void __launch_bounds__(256)
im2col_strided_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ col,
                      int N, int H, int W, int C, int Ho, int Wo, int kh, int kw, int stride, int pad_t, int pad_l) {
    pdl_entry();
    __shared__ uint32_t tab[128];                 // per column k: (r, s, c, valid) as bytes
    if (threadIdx.x < 128) {
        const int k = threadIdx.x;
        uint32_t e = 0;
        if (k < kh * kw * C) {
            const int t = k / C, c = k - t * C;
            const int r = t / kw, s = t - r * kw;
            e = (uint32_t)r | ((uint32_t)s << 8) | ((uint32_t)c << 16) | (1u << 24);
        }
        tab[k] = e;
    }
    __syncthreads();
    const int64_t total = (int64_t)N * Ho * Wo * 16;         // 16 threads per output pixel, one 16-byte store each
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i >> 4;
        const int g = (int)(i & 15);
        const int wo = (int)(p % Wo); const int64_t q = p / Wo;
        const int ho = (int)(q % Ho); const int64_t n = q / Ho;
        const int h0 = ho * stride - pad_t, w0 = wo * stride - pad_l;
        const __nv_bfloat16* img = src + n * (int64_t)H * W * C;
        __align__(16) __nv_bfloat16 v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const uint32_t te = tab[g * 8 + e];
            const int hh = h0 + (int)(te & 0xff), ww = w0 + (int)((te >> 8) & 0xff), c = (int)((te >> 16) & 0xff);
            const bool ok = (te >> 24) && hh >= 0 && hh < H && ww >= 0 && ww < W;
            v[e] = ok ? img[((int64_t)hh * W + ww) * C + c] : __float2bfloat16_rn(0.f);
        }
        *reinterpret_cast<uint4*>(col + p * 128 + g * 8) = *reinterpret_cast<const uint4*>(v);
    }
}

__global__ void __launch_bounds__(128)
col2im_strided_kernel(const __nv_bfloat16* __restrict__ col, const float* __restrict__ bias, __nv_bfloat16* __restrict__ dst,
                      int N, int H, int W, int C, int Ho, int Wo, int kh, int kw, int stride, int pad_t, int pad_l) {
    pdl_entry();
    const int64_t total = (int64_t)N * H * W;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int w = (int)(p % W); const int64_t q = p / W;
        const int h = (int)(q % H); const int64_t n = q / H;
        float acc[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = (bias && c < C) ? bias[c] : 0.f;
        for (int r = 0; r < kh; ++r) {
            const int th = h + pad_t - r;
            if (th < 0 || th % stride) continue;
            const int ho = th / stride;
            if (ho >= Ho) continue;
            for (int s = 0; s < kw; ++s) {
                const int tw = w + pad_l - s;
                if (tw < 0 || tw % stride) continue;
                const int wo = tw / stride;
                if (wo >= Wo) continue;
                const __nv_bfloat16* row = col + ((n * Ho + ho) * Wo + wo) * 128 + (r * kw + s) * C;
#pragma unroll
                for (int c = 0; c < 8; ++c) if (c < C) acc[c] += __bfloat162float(row[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) if (c < C) dst[p * C + c] = __float2bfloat16_rn(acc[c]);
    }
}

// w [Kreal][O] float (HWIO flattened) -> wp_f [O][128] and wp_d [128][O] bf16, rows/columns k >= Kreal zero
__global__ void __launch_bounds__(256)
pack_filter_padk_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp_f, __nv_bfloat16* __restrict__ wp_d,
                        int Kreal, int O) {
    pdl_entry();
    __shared__ __nv_bfloat16 tile[32][33];
    const int k0 = blockIdx.y * 32, o0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int k = k0 + ty + 8 * j, o = o0 + tx;
        const float v = (k < Kreal && o < O) ? w[(int64_t)k * O + o] : 0.f;
        const __nv_bfloat16 b = __float2bfloat16_rn(v);
        tile[ty + 8 * j][tx] = b;
        if (wp_d && o < O) wp_d[(int64_t)k * O + o] = b;
    }
    __syncthreads();
    if (wp_f) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = o0 + ty + 8 * j, k = k0 + tx;
            if (o < O) wp_f[(int64_t)o * 128 + k] = tile[tx][ty + 8 * j];
        }
    }
}

__global__ void add_prefix_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n, int accumulate) {
    pdl_entry();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (accumulate) atomicAdd(dst + i, src[i]); else dst[i] = src[i];
    }
}

static int check_strided_thin(const ctgan_conv_desc* d, int C, const char* who) {
    CTGAN_REQUIRE(d != nullptr, CTGAN_ERR_BAD_DESC, "%s: null descriptor", who);
    CTGAN_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->Ho > 0 && d->Wo > 0 && d->kh > 0 && d->kw > 0 && d->stride >= 1 &&
                  d->pad_t >= 0 && d->pad_l >= 0 && d->kh < 256 && d->kw < 256, CTGAN_ERR_BAD_DESC, "%s: bad geometry", who);
    CTGAN_REQUIRE(C > 0 && C <= 8 && d->kh * d->kw * C <= 128, CTGAN_ERR_UNSUPPORTED, "%s: needs C <= 8 and taps*C <= 128", who);
    return 0;
}

static int check_s2d_filter(int k, int C, int O, int pad_t, int pad_l, const char* who) {
    CTGAN_REQUIRE(k > 0 && C > 0 && O > 0 && pad_t >= 0 && pad_l >= 0 && pad_t <= 2 && pad_l <= 2, CTGAN_ERR_BAD_DESC, "%s: bad args", who);
    // block offsets floor((r - pad)/2) of the taps r = 0..k-1 must be exactly {-1, 0, 1}
    CTGAN_REQUIRE(k - 1 - pad_t <= 3 && k - 1 - pad_l <= 3, CTGAN_ERR_UNSUPPORTED,
                  "%s: k=%d pad=(%d,%d) does not embed into a 3x3 space-to-depth filter", who, k, pad_t, pad_l);
    return 0;
}

// ------------------------------------------------------------------ ConvMeanPool(3x3) as one stride-2 conv
// Mean-pooling the output of a 'SAME' 3x3 conv over 2x2 windows (TG/CT_gan_cifar_resnet.py:89-92) is ONE 'SAME' 4x4 /
// stride-2 conv (pad 1) with   W4[u][v] = 1/4 * sum_{a,b in {0,1}} w[u-a][v-b]   (terms outside the 3x3 filter dropped):
// 16 taps at a quarter of the positions instead of 9 at all of them, 2.25x fewer multiply-adds.
// w3 HWIO [3][3][CO] -> w4 [4][4][CO], CO = Cin*Cout.
__global__ void box_filter_kernel(const float* __restrict__ w3, float* __restrict__ w4, int64_t CO) {
    pdl_entry();
    const int64_t total = 16 * CO;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(i / CO); const int64_t e = i - (int64_t)t * CO;
        const int u = t >> 2, v = t & 3;
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int r = u - a, s = v - b;
                if (r >= 0 && r < 3 && s >= 0 && s < 3) acc += w3[(int64_t)(r * 3 + s) * CO + e];
            }
        w4[i] = 0.25f * acc;
    }
}
// the adjoint, folding the gradient of W4 into the 3x3 filter's and clearing it for the next step:
// dw3[r][s] += 1/4 * sum_{a,b} dW4[r+a][s+b];  dW4 = 0.   One thread per (cin, cout) element.
__global__ void box_filter_grad_kernel(float* __restrict__ dw4, float* __restrict__ dw3, int64_t CO) {
    pdl_entry();
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < CO; e += (int64_t)gridDim.x * blockDim.x) {
        float g[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) { g[t] = dw4[(int64_t)t * CO + e]; dw4[(int64_t)t * CO + e] = 0.f; }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s)
                dw3[(int64_t)(r * 3 + s) * CO + e] += 0.25f * (g[r * 4 + s] + g[r * 4 + s + 1] + g[(r + 1) * 4 + s] + g[(r + 1) * 4 + s + 1]);
    }
}

template <bool FWD>
static int s2d_impl(const void* src, const void* mul, void* dst, int N, int H, int W, int C, int dtype, void* stream, const char* who,
                    int relu_pattern = 0) {
    CTGAN_REQUIRE(src && dst && N > 0 && H > 0 && W > 0 && C > 0 && dtype_ok(dtype), CTGAN_ERR_BAD_DESC, "%s: bad args", who);
    const int Hs = (H + 1) / 2, Ws = (W + 1) / 2;
    cudaStream_t st = as_stream(stream);
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(mul)) & 15) == 0;
    if (dtype == CTGAN_BF16 && C % 8 == 0 && aligned) {
        const int64_t total = FWD ? (int64_t)N * Hs * Ws * 4 * (C / 8) : (int64_t)N * H * W * (C / 8);
        CTGAN_LAUNCH((s2d_kernel<__nv_bfloat16, 8, FWD>), elementwise_grid(total, 256), 256, 0, st,
                     (const __nv_bfloat16*)src, (const __nv_bfloat16*)mul, (__nv_bfloat16*)dst, N, H, W, C, Hs, Ws, relu_pattern);
    } else if (dtype == CTGAN_BF16) {
        const int64_t total = FWD ? (int64_t)N * Hs * Ws * 4 * C : (int64_t)N * H * W * C;
        CTGAN_LAUNCH((s2d_kernel<__nv_bfloat16, 1, FWD>), elementwise_grid(total, 256), 256, 0, st,
                     (const __nv_bfloat16*)src, (const __nv_bfloat16*)mul, (__nv_bfloat16*)dst, N, H, W, C, Hs, Ws, relu_pattern);
    } else if (C % 4 == 0 && aligned) {
        const int64_t total = FWD ? (int64_t)N * Hs * Ws * 4 * (C / 4) : (int64_t)N * H * W * (C / 4);
        CTGAN_LAUNCH((s2d_kernel<float, 4, FWD>), elementwise_grid(total, 256), 256, 0, st,
                     (const float*)src, (const float*)mul, (float*)dst, N, H, W, C, Hs, Ws, relu_pattern);
    } else {
        const int64_t total = FWD ? (int64_t)N * Hs * Ws * 4 * C : (int64_t)N * H * W * C;
        CTGAN_LAUNCH((s2d_kernel<float, 1, FWD>), elementwise_grid(total, 256), 256, 0, st,
                     (const float*)src, (const float*)mul, (float*)dst, N, H, W, C, Hs, Ws, relu_pattern);
    }
    CTGAN_CHECK_LAUNCH(who);
    return 0;
}

}  // namespace
}  // namespace ctgan

using namespace ctgan;

extern "C" int ctgan_space_to_depth(const void* x, void* xs, int N, int H, int W, int C, int dtype, void* stream) {
    return s2d_impl<true>(x, nullptr, xs, N, H, W, C, dtype, stream, "space_to_depth");
}

extern "C" int ctgan_depth_to_space(const void* xs, void* x, int N, int H, int W, int C, int dtype, void* stream) {
    return s2d_impl<false>(xs, nullptr, x, N, H, W, C, dtype, stream, "depth_to_space");
}

extern "C" int ctgan_space_to_depth_mul(const void* x, const void* ms, void* xs, int N, int H, int W, int C, int dtype, void* stream) {
    CTGAN_REQUIRE(ms != nullptr, CTGAN_ERR_BAD_DESC, "space_to_depth_mul: null multiplier");
    return s2d_impl<true>(x, ms, xs, N, H, W, C, dtype, stream, "space_to_depth_mul");
}

/* xs = space_to_depth(x) where the tensor `pattern` (space-to-depth layout, a ReLU output) is positive, 0 elsewhere */
extern "C" int ctgan_space_to_depth_mask(const void* x, const void* pattern, void* xs, int N, int H, int W, int C, int dtype, void* stream) {
    CTGAN_REQUIRE(pattern != nullptr, CTGAN_ERR_BAD_DESC, "space_to_depth_mask: null pattern");
    return s2d_impl<true>(x, pattern, xs, N, H, W, C, dtype, stream, "space_to_depth_mask", 1);
}

extern "C" int ctgan_depth_to_space_mul(const void* xs, const void* ms, void* x, int N, int H, int W, int C, int dtype, void* stream) {
    CTGAN_REQUIRE(ms != nullptr, CTGAN_ERR_BAD_DESC, "depth_to_space_mul: null multiplier");
    return s2d_impl<false>(xs, ms, x, N, H, W, C, dtype, stream, "depth_to_space_mul");
}

extern "C" int ctgan_pack_filter_s2d(const float* w, void* wp_f, void* wp_d, int k, int Cin, int Cout, int pad_t, int pad_l,
                                     void* stream) {
    if (int r = check_s2d_filter(k, Cin, Cout, pad_t, pad_l, "pack_filter_s2d")) return r;
    CTGAN_REQUIRE(w && (wp_f || wp_d), CTGAN_ERR_BAD_DESC, "pack_filter_s2d: null pointer");
    CTGAN_REQUIRE((4 * Cin + 31) / 32 <= 65535, CTGAN_ERR_UNSUPPORTED, "pack_filter_s2d: Cin too large");
    CTGAN_LAUNCH((pack_filter_s2d_kernel), dim3((Cout + 31) / 32, (4 * Cin + 31) / 32, 9), 256, 0, as_stream(stream), w,
                 reinterpret_cast<__nv_bfloat16*>(wp_f), reinterpret_cast<__nv_bfloat16*>(wp_d), k, Cin, Cout, pad_t, pad_l);
    CTGAN_CHECK_LAUNCH("pack_filter_s2d");
    return 0;
}

extern "C" int ctgan_s2d_filter_grad(const float* dw3, float* dw, int k, int Cin, int Cout, int pad_t, int pad_l,
                                     int accumulate, void* stream) {
    if (int r = check_s2d_filter(k, Cin, Cout, pad_t, pad_l, "s2d_filter_grad")) return r;
    CTGAN_REQUIRE(dw3 && dw, CTGAN_ERR_BAD_DESC, "s2d_filter_grad: null pointer");
    const int64_t total = (int64_t)k * k * Cin * Cout;
    CTGAN_LAUNCH((s2d_filter_grad_kernel), elementwise_grid(total, 256), 256, 0, as_stream(stream), dw3, dw, k, Cin, Cout,
                 pad_t, pad_l, accumulate);
    CTGAN_CHECK_LAUNCH("s2d_filter_grad");
    return 0;
}

extern "C" int ctgan_im2col_strided(const ctgan_conv_desc* d, int C, const void* src, void* col, void* stream) {
    if (int r = check_strided_thin(d, C, "im2col_strided")) return r;
    CTGAN_REQUIRE(src && col && (reinterpret_cast<uintptr_t>(col) & 15) == 0, CTGAN_ERR_BAD_DESC, "im2col_strided: bad pointers");
    const int64_t total = (int64_t)d->N * d->Ho * d->Wo * 16;
    CTGAN_LAUNCH((im2col_strided_kernel), elementwise_grid(total, 256), 256, 0, as_stream(stream),
                 reinterpret_cast<const __nv_bfloat16*>(src), reinterpret_cast<__nv_bfloat16*>(col), d->N, d->H, d->W, C, d->Ho, d->Wo,
                 d->kh, d->kw, d->stride, d->pad_t, d->pad_l);
    CTGAN_CHECK_LAUNCH("im2col_strided");
    return 0;
}

extern "C" int ctgan_col2im_strided(const ctgan_conv_desc* d, int C, const void* col, const float* bias, void* dst, void* stream) {
    if (int r = check_strided_thin(d, C, "col2im_strided")) return r;
    CTGAN_REQUIRE(col && dst, CTGAN_ERR_BAD_DESC, "col2im_strided: null pointer");
    const int64_t total = (int64_t)d->N * d->H * d->W;
    CTGAN_LAUNCH((col2im_strided_kernel), elementwise_grid(total, 128), 128, 0, as_stream(stream),
                 reinterpret_cast<const __nv_bfloat16*>(col), bias, reinterpret_cast<__nv_bfloat16*>(dst), d->N, d->H, d->W, C, d->Ho, d->Wo,
                 d->kh, d->kw, d->stride, d->pad_t, d->pad_l);
    CTGAN_CHECK_LAUNCH("col2im_strided");
    return 0;
}

extern "C" int ctgan_pack_filter_padk(const float* w, void* wp_f, void* wp_d, int Kreal, int Cout, void* stream) {
    CTGAN_REQUIRE(w && (wp_f || wp_d) && Kreal > 0 && Kreal <= 128 && Cout > 0, CTGAN_ERR_BAD_DESC, "pack_filter_padk: bad args");
    CTGAN_LAUNCH((pack_filter_padk_kernel), dim3((Cout + 31) / 32, 4), 256, 0, as_stream(stream), w,
                 reinterpret_cast<__nv_bfloat16*>(wp_f), reinterpret_cast<__nv_bfloat16*>(wp_d), Kreal, Cout);
    CTGAN_CHECK_LAUNCH("pack_filter_padk");
    return 0;
}

extern "C" int ctgan_add_prefix(const float* src, float* dst, int64_t n, int accumulate, void* stream) {
    CTGAN_REQUIRE(src && dst && n >= 0, CTGAN_ERR_BAD_DESC, "add_prefix: bad args");
    if (n == 0) return 0;
    CTGAN_LAUNCH((add_prefix_kernel), elementwise_grid(n, 256), 256, 0, as_stream(stream), src, dst, n, accumulate);
    CTGAN_CHECK_LAUNCH("add_prefix");
    return 0;
}

extern "C" int ctgan_box_filter(const float* w3, float* w4, int Cin, int Cout, void* stream) {
    CTGAN_REQUIRE(w3 && w4 && Cin > 0 && Cout > 0, CTGAN_ERR_BAD_DESC, "box_filter: bad args");
    const int64_t CO = (int64_t)Cin * Cout;
    CTGAN_LAUNCH((box_filter_kernel), elementwise_grid(16 * CO, 256), 256, 0, as_stream(stream), w3, w4, CO);
    CTGAN_CHECK_LAUNCH("box_filter");
    return 0;
}

extern "C" int ctgan_box_filter_grad(float* dw4, float* dw3, int Cin, int Cout, void* stream) {
    CTGAN_REQUIRE(dw4 && dw3 && Cin > 0 && Cout > 0, CTGAN_ERR_BAD_DESC, "box_filter_grad: bad args");
    const int64_t CO = (int64_t)Cin * Cout;
    CTGAN_LAUNCH((box_filter_grad_kernel), elementwise_grid(CO, 128), 128, 0, as_stream(stream), dw4, dw3, CO);
    CTGAN_CHECK_LAUNCH("box_filter_grad");
    return 0;
}
