// Micro-benchmark: raw tcgen05.mma issue/execute rate from shared memory (no TMA traffic), one CTA per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tests/micro/mma_rate tests/micro/mma_rate.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int N>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, int stages, long long* cycles, int mode, const uint8_t* gsrc) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint64_t scratch[8];
    __shared__ uint64_t done_bar;
    __shared__ uint64_t cp_bar[4];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < stages * (16384 + N * 128) / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&scratch[i])));
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&cp_bar[i])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done_bar)));
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&done_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_slot;
    if (warp == 1 && lane == 0) {
        const uint32_t idesc = make_idesc(128, N);
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (mode & 2) {
                uint32_t ok = 0;
                while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&done_bar)));
                asm volatile("tcgen05.fence::after_thread_sync;");
            }
            const uint32_t base = smem_u32(smem) + (uint32_t)(it % stages) * (16384 + N * 128);
            const uint64_t a = make_desc(base, 16, 1024), b = make_desc(base + 16384, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t acc = (it | k) != 0;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(a + 2 * k), "l"(b + 2 * k), "r"(idesc), "r"(acc));
            }
            if (mode & 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&scratch[it & 7])));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)));
        }
        long long t1 = clock64();
        if (blockIdx.x == 0) cycles[0] = t1 - t0;
    }
    if ((mode & 4) && warp == 2 && lane == 0) {
        // stream 32 KB per iteration from global into the LAST stage region (never read by the MMAs when stages>1)
        const uint32_t dst = smem_u32(smem) + (uint32_t)(stages - 1) * (16384 + N * 128);
        uint32_t ph[4] = {0, 0, 0, 0};
        for (int it = 0; it < iters; ++it) {
            const int b = it & 3;
            if (it >= 4) { uint32_t ok = 0; while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&cp_bar[b])), "r"(ph[b])); ph[b] ^= 1; }
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&cp_bar[b])), "r"(32768));
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(gsrc + ((size_t)(blockIdx.x * 64 + (it & 63)) * 32768)), "r"(32768), "r"(smem_u32(&cp_bar[b])) : "memory");
        }
        for (int b = 0; b < 4; ++b) { uint32_t ok = 0; while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&cp_bar[b])), "r"(ph[b])); }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

template <int N>
void run(int stages, int mode = 0) {
    static uint8_t* gsrc = nullptr;
    if (!gsrc) { cudaMalloc(&gsrc, (size_t)148 * 64 * 32768); cudaMemset(gsrc, 0x3c, (size_t)148 * 64 * 32768); }
    long long* d; cudaMalloc(&d, 8);
    size_t smem = (size_t)stages * (16384 + N * 128) + 2048;
    cudaFuncSetAttribute(mma_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int iters = 4000;
    mma_rate_kernel<N><<<148, 128, smem>>>(iters, stages, d, mode, gsrc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    mma_rate_kernel<N><<<148, 128, smem>>>(iters, stages, d, mode, gsrc);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    double flops = 2.0 * 148 * iters * 4 * 128.0 * N * 16;
    printf("mode=%d N=%3d stages=%d: %.1f cycles per MMA (M=128,N=%d,K=16), %.3f ms, %.0f TFLOP/s, err=%s\n", mode, N, stages,
           (double)c / (iters * 4), N, ms, flops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}

int main() {
    run<128>(6, 0); run<128>(6, 1); run<128>(6, 2); run<128>(6, 3); run<128>(6, 4); run<128>(6, 7); run<256>(4, 0);
    return 0;
}
