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
template <typename T, int VEC, bool FWD>
__global__ void s2d_kernel(const T* __restrict__ src, T* __restrict__ dst, int N, int H, int W, int C, int Hs, int Ws) {
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
            reinterpret_cast<Vec*>(dst)[idx] =
                reinterpret_cast<const Vec*>(src)[((((int64_t)n * Hs + (h >> 1)) * Ws + (w >> 1)) * 4 + q) * cv + c];
        }
    }
}

// source tap of embedded tap R (0..2) and parity d: r = 2(R-1) + d + pad; valid when 0 <= r < k
__device__ __forceinline__ int s2d_src_tap(int R, int d, int pad) { return 2 * (R - 1) + d + pad; }

// w HWIO [k][k][C][O] float -> bf16 operands of the 3x3 conv with 4C input channels:
//   wp_f [T][O][4C]           (fprop: rows = output channels, K = input channels)         T = R*3 + S
//   wp_d [8-T][4C][O]         (dgrad: tap-flipped, rows = input channels, K = output channels)
__global__ void pack_filter_s2d_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp_f,
                                       __nv_bfloat16* __restrict__ wp_d, int k, int C, int O, int pad_t, int pad_l) {
    pdl_entry();
    const int C4 = 4 * C;
    const int64_t total = (int64_t)9 * O * C4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int kc = (int)(i % C4); int64_t t = i / C4;
        int o = (int)(t % O); int T = (int)(t / O);
        const int R = T / 3, S = T - 3 * R;
        const int q = kc / C, c = kc - q * C;
        const int r = s2d_src_tap(R, q >> 1, pad_t), s = s2d_src_tap(S, q & 1, pad_l);
        float v = 0.f;
        if (r >= 0 && r < k && s >= 0 && s < k) v = w[(((int64_t)r * k + s) * C + c) * O + o];
        const __nv_bfloat16 b = __float2bfloat16_rn(v);
        if (wp_f) wp_f[i] = b;
        if (wp_d) wp_d[((int64_t)(8 - T) * C4 + kc) * O + o] = b;
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

static int check_s2d_filter(int k, int C, int O, int pad_t, int pad_l, const char* who) {
    CTGAN_REQUIRE(k > 0 && C > 0 && O > 0 && pad_t >= 0 && pad_l >= 0 && pad_t <= 2 && pad_l <= 2, CTGAN_ERR_BAD_DESC, "%s: bad args", who);
    // block offsets floor((r - pad)/2) of the taps r = 0..k-1 must be exactly {-1, 0, 1}
    CTGAN_REQUIRE(k - 1 - pad_t <= 3 && k - 1 - pad_l <= 3, CTGAN_ERR_UNSUPPORTED,
                  "%s: k=%d pad=(%d,%d) does not embed into a 3x3 space-to-depth filter", who, k, pad_t, pad_l);
    return 0;
}

template <bool FWD>
static int s2d_impl(const void* src, void* dst, int N, int H, int W, int C, int dtype, void* stream, const char* who) {
    CTGAN_REQUIRE(src && dst && N > 0 && H > 0 && W > 0 && C > 0 && dtype_ok(dtype), CTGAN_ERR_BAD_DESC, "%s: bad args", who);
    const int Hs = (H + 1) / 2, Ws = (W + 1) / 2;
    cudaStream_t st = as_stream(stream);
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
    if (dtype == CTGAN_BF16 && C % 8 == 0 && aligned) {
        const int64_t total = FWD ? (int64_t)N * Hs * Ws * 4 * (C / 8) : (int64_t)N * H * W * (C / 8);
        CTGAN_LAUNCH((s2d_kernel<__nv_bfloat16, 8, FWD>), elementwise_grid(total, 256), 256, 0, st,
                     (const __nv_bfloat16*)src, (__nv_bfloat16*)dst, N, H, W, C, Hs, Ws);
    } else if (dtype == CTGAN_BF16) {
        const int64_t total = FWD ? (int64_t)N * Hs * Ws * 4 * C : (int64_t)N * H * W * C;
        CTGAN_LAUNCH((s2d_kernel<__nv_bfloat16, 1, FWD>), elementwise_grid(total, 256), 256, 0, st,
                     (const __nv_bfloat16*)src, (__nv_bfloat16*)dst, N, H, W, C, Hs, Ws);
    } else if (C % 4 == 0 && aligned) {
        const int64_t total = FWD ? (int64_t)N * Hs * Ws * 4 * (C / 4) : (int64_t)N * H * W * (C / 4);
        CTGAN_LAUNCH((s2d_kernel<float, 4, FWD>), elementwise_grid(total, 256), 256, 0, st,
                     (const float*)src, (float*)dst, N, H, W, C, Hs, Ws);
    } else {
        const int64_t total = FWD ? (int64_t)N * Hs * Ws * 4 * C : (int64_t)N * H * W * C;
        CTGAN_LAUNCH((s2d_kernel<float, 1, FWD>), elementwise_grid(total, 256), 256, 0, st,
                     (const float*)src, (float*)dst, N, H, W, C, Hs, Ws);
    }
    CTGAN_CHECK_LAUNCH(who);
    return 0;
}

}  // namespace
}  // namespace ctgan

using namespace ctgan;

extern "C" int ctgan_space_to_depth(const void* x, void* xs, int N, int H, int W, int C, int dtype, void* stream) {
    return s2d_impl<true>(x, xs, N, H, W, C, dtype, stream, "space_to_depth");
}

extern "C" int ctgan_depth_to_space(const void* xs, void* x, int N, int H, int W, int C, int dtype, void* stream) {
    return s2d_impl<false>(xs, x, N, H, W, C, dtype, stream, "depth_to_space");
}

extern "C" int ctgan_pack_filter_s2d(const float* w, void* wp_f, void* wp_d, int k, int Cin, int Cout, int pad_t, int pad_l,
                                     void* stream) {
    if (int r = check_s2d_filter(k, Cin, Cout, pad_t, pad_l, "pack_filter_s2d")) return r;
    CTGAN_REQUIRE(w && (wp_f || wp_d), CTGAN_ERR_BAD_DESC, "pack_filter_s2d: null pointer");
    const int64_t total = (int64_t)36 * Cin * Cout;
    CTGAN_LAUNCH((pack_filter_s2d_kernel), elementwise_grid(total, 256), 256, 0, as_stream(stream), w,
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
