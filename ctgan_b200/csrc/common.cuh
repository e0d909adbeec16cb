// common.cuh -- shared helpers for libctgan_sm100 (dtype access, error plumbing, Philox).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/ctgan_sm100.h"

namespace ctgan {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
int  cuda_status(cudaError_t e, const char* what);

extern unsigned long long g_kernel_launches;   // kernels launched by this library (api.cu)

#define CTGAN_CHECK_LAUNCH(what)                                   \
    do {                                                           \
        ++::ctgan::g_kernel_launches;                              \
        cudaError_t _e = cudaGetLastError();                       \
        if (_e != cudaSuccess) return ::ctgan::cuda_status(_e, what); \
    } while (0)

// ---------------------------------------------------------------- programmatic dependent launch
// A step is a chain of ~200 short kernels; with plain stream order every boundary costs a full drain + launch.  All
// kernels of the library are launched with the programmatic-stream-serialization attribute and begin with
//     griddepcontrol.launch_dependents;   (the next kernel's CTAs may become resident once all of ours have started)
//     griddepcontrol.wait;                (block until every prerequisite grid has completed and its writes are visible)
// so the dependent kernel's launch latency, CTA scheduling and (for the tcgen05 kernels) barrier / TMEM / tensor-map
// set-up overlap the tail of its predecessor.  No global memory is touched before the wait: ordering semantics are
// exactly those of the stream.  CUDA-graph capture turns the attribute into programmatic dependency edges.
extern int g_pdl;                                   // ctgan_set_pdl(): 0 = plain launches (api.cu)
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_entry() { pdl_launch_dependents(); pdl_wait(); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define CTGAN_LAUNCH(kernel, grid, block, smem, stream, ...) \
    (void)::ctgan::launch_pdl(kernel, grid, block, smem, stream, __VA_ARGS__)

#define CTGAN_REQUIRE(cond, code, ...)                             \
    do {                                                           \
        if (!(cond)) { ::ctgan::set_error(__VA_ARGS__); return (code); } \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
static inline size_t dtype_size(int dt) { return dt == CTGAN_BF16 ? 2 : 4; }
static inline bool dtype_ok(int dt) { return dt == CTGAN_F32 || dt == CTGAN_BF16; }

// grid size for a grid-stride element-wise kernel: enough CTAs for ~8 waves max, multiple of SMs
int elementwise_grid(int64_t work_items, int threads);
int sm_count();

// ---------------------------------------------------------------- dtype access
// Runtime-dtype loads/stores (uniform branch) for the generic kernels.
__device__ __forceinline__ float ld_act(const void* p, int64_t i, int dt) {
    return dt == CTGAN_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i])
                            : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void st_act(void* p, int64_t i, int dt, float v) {
    if (dt == CTGAN_BF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
    else                  reinterpret_cast<float*>(p)[i] = v;
}

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---------------------------------------------------------------- Philox4x32-10
struct Philox {
    static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    static constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    __host__ __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
        uint64_t p0 = (uint64_t)M0 * c[0];
        uint64_t p1 = (uint64_t)M1 * c[2];
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    // 4 x 32 random bits for 64-bit counter `ctr` under 64-bit key `seed`
    __host__ __device__ static inline void block(uint64_t seed, uint64_t ctr, uint32_t (&out)[4]) {
        uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
        uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
        for (int i = 0; i < 10; ++i) { round(c, k0, k1); k0 += W0; k1 += W1; }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
    __host__ __device__ static inline float to_uniform(uint32_t bits) {
        return (float)(bits >> 8) * (1.0f / 16777216.0f);          // [0,1), 24 bits
    }
    // uniform for absolute stream element index e = offset + i
    __host__ __device__ static inline float uniform_at(uint64_t seed, uint64_t e) {
        uint32_t r[4];
        block(seed, e >> 2, r);
        return to_uniform(r[e & 3]);
    }
};

}  // namespace ctgan
