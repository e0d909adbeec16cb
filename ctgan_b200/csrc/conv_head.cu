// conv_head.cu -- the critic heads: Linear(DIM_D -> 1) 'Discriminator.Output' and Linear(DIM_D -> 10)
// 'Discriminator.ACGANOutput' on the [B, DIM_D] pooled features (TG/CT_gan_cifar_resnet.py:180-187; op =
// TG/tflib/ops/linear.py:132-136), and their dgrad / wgrad.  M = batch rows (64-192), K = DIM_D, N <= 16: a tiled GEMM
// spends its time in pipeline latency, so each member is one short kernel:
//   fprop: one warp per row, lanes stride K, N partial sums reduced with shuffles
//   dgrad: one thread per (row, k)
//   wgrad: one thread per (k, n), rows split over blockIdx.y, atomics into dw
#include "common.cuh"

namespace ctgan {
namespace head {

constexpr int MAX_N = 16;

__global__ void __launch_bounds__(256)
head_fprop_kernel(const void* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, void* __restrict__ y,
                  int M, int K, int N, int xdt, int ydt, int relu) {
    ctgan::pdl_entry();
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (m >= M) return;
    float acc[MAX_N];
#pragma unroll
    for (int n = 0; n < MAX_N; ++n) acc[n] = 0.f;
    for (int k = lane; k < K; k += 32) {
        const float xv = ld_act(x, (int64_t)m * K + k, xdt);
        const float* wr = w + (int64_t)k * N;
#pragma unroll
        for (int n = 0; n < MAX_N; ++n) if (n < N) acc[n] = fmaf(xv, wr[n], acc[n]);
    }
#pragma unroll
    for (int n = 0; n < MAX_N; ++n) {
        if (n < N) {
            float v = acc[n];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) {
                v += bias ? bias[n] : 0.f;
                st_act(y, (int64_t)m * N + n, ydt, relu ? fmaxf(v, 0.f) : v);
            }
        }
    }
}

// N == 1 with a long K (the DCGAN critics' 'Discriminator.Output' Linear(4*4*4*DIM -> 1) on the flattened features,
// TG/CT_gan_cifar.py:98, TG/CT_gan_mnist.py:106): one 128-thread CTA per row, 16-byte loads of x, float4 loads of w,
// 4 independent loads in flight per lane -- the warp-per-row kernel above spends one DRAM latency per 32 elements.
__global__ void __launch_bounds__(128)
head_fprop_n1_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                     void* __restrict__ y, int M, int K, int ydt, int relu) {
    ctgan::pdl_entry();
    __shared__ float part[4];
    const int m = blockIdx.x;
    const __nv_bfloat16* xr = x + (int64_t)m * K;
    float acc = 0.f;
    for (int k = threadIdx.x * 8; k < K; k += 128 * 8) {
        const uint4 raw = *reinterpret_cast<const uint4*>(xr + k);
        const float4 w0 = *reinterpret_cast<const float4*>(w + k), w1 = *reinterpret_cast<const float4*>(w + k + 4);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
        const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]), c = __bfloat1622float2(h[2]), d = __bfloat1622float2(h[3]);
        acc = fmaf(a.x, w0.x, acc); acc = fmaf(a.y, w0.y, acc); acc = fmaf(b.x, w0.z, acc); acc = fmaf(b.y, w0.w, acc);
        acc = fmaf(c.x, w1.x, acc); acc = fmaf(c.y, w1.y, acc); acc = fmaf(d.x, w1.z, acc); acc = fmaf(d.y, w1.w, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float v = part[0] + part[1] + part[2] + part[3] + (bias ? bias[0] : 0.f);
        st_act(y, m, ydt, relu ? fmaxf(v, 0.f) : v);
    }
}

__global__ void __launch_bounds__(256)
head_dgrad_kernel(const void* __restrict__ dy, const float* __restrict__ w, void* __restrict__ dx, int M, int K, int N, int xdt, int ydt) {
    ctgan::pdl_entry();
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)M * K) return;
    const int k = (int)(i % K), m = (int)(i / K);
    const float* wr = w + (int64_t)k * N;
    float v = 0.f;
    for (int n = 0; n < N; ++n) v = fmaf(ld_act(dy, (int64_t)m * N + n, ydt), wr[n], v);
    st_act(dx, i, xdt, v);
}

__global__ void __launch_bounds__(256)
head_wgrad_kernel(const void* __restrict__ x, const void* __restrict__ dy, float* __restrict__ dw, int M, int K, int N, int xdt, int ydt,
                  int rows_per_split) {
    ctgan::pdl_entry();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;                 // i = n * K + k: consecutive threads read consecutive x
    if (i >= K * N) return;
    const int k = i % K, n = i / K;
    const int m0 = blockIdx.y * rows_per_split, m1 = min(M, m0 + rows_per_split);
    float v = 0.f;
#pragma unroll 4
    for (int m = m0; m < m1; ++m) v = fmaf(ld_act(x, (int64_t)m * K + k, xdt), ld_act(dy, (int64_t)m * N + n, ydt), v);
    atomicAdd(dw + (int64_t)k * N + n, v);
}

static bool is_head(const ctgan_conv_desc* d) {
    return d->H == 1 && d->W == 1 && d->Ho == 1 && d->Wo == 1 && d->kh == 1 && d->kw == 1 && d->stride == 1 &&
           d->Cout <= MAX_N && d->Cin >= 32 && d->Cin <= 16384 && d->N <= (1 << 20);
}

int try_fprop(const ctgan_conv_desc* d, const void* x, const float* w, const float* bias, void* y, int flags, cudaStream_t st, int* rc) {
    if (!is_head(d)) return 0;
    if (d->Cout == 1 && d->x_dtype == CTGAN_BF16 && d->Cin % 8 == 0 && d->Cin >= 1024 &&
        ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w)) & 15) == 0) {
        CTGAN_LAUNCH((head_fprop_n1_kernel), (unsigned)d->N, 128, 0, st, reinterpret_cast<const __nv_bfloat16*>(x), w, bias, y, d->N, d->Cin,
                                                                  d->y_dtype, (flags & CTGAN_EPI_RELU) ? 1 : 0);
        *rc = 0;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) *rc = cuda_status(e, "head_fprop_n1"); else ++g_kernel_launches;
        return 1;
    }
    const int64_t threads = (int64_t)d->N * 32;
    CTGAN_LAUNCH((head_fprop_kernel), (unsigned)((threads + 255) / 256), 256, 0, st, x, w, bias, y, d->N, d->Cin, d->Cout, d->x_dtype, d->y_dtype,
                                                                        (flags & CTGAN_EPI_RELU) ? 1 : 0);
    *rc = 0;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) *rc = cuda_status(e, "head_fprop"); else ++g_kernel_launches;
    return 1;
}

int try_dgrad(const ctgan_conv_desc* d, const void* dy, const float* w, void* dx, cudaStream_t st, int* rc) {
    if (!is_head(d)) return 0;
    const int64_t threads = (int64_t)d->N * d->Cin;
    CTGAN_LAUNCH((head_dgrad_kernel), (unsigned)((threads + 255) / 256), 256, 0, st, dy, w, dx, d->N, d->Cin, d->Cout, d->x_dtype, d->y_dtype);
    *rc = 0;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) *rc = cuda_status(e, "head_dgrad"); else ++g_kernel_launches;
    return 1;
}

// dw must already hold the value to accumulate onto (the caller zero-fills it when not accumulating)
int try_wgrad(const ctgan_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st, int* rc) {
    if (!is_head(d)) return 0;
    const int threads = d->Cin * d->Cout;
    int splits = d->N / 32; if (splits < 1) splits = 1; if (splits > 16) splits = 16;
    const int rps = (d->N + splits - 1) / splits;
    splits = (d->N + rps - 1) / rps;
    CTGAN_LAUNCH((head_wgrad_kernel), dim3((threads + 255) / 256, splits), 256, 0, st, x, dy, dw, d->N, d->Cin, d->Cout, d->x_dtype, d->y_dtype, rps);
    *rc = 0;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) *rc = cuda_status(e, "head_wgrad"); else ++g_kernel_launches;
    return 1;
}
bool ok(const ctgan_conv_desc* d) { return is_head(d); }

}  // namespace head
}  // namespace ctgan
