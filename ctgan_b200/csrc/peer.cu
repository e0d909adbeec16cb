// peer.cu -- the data-parallel optimizer update as ONE kernel over NVLink peer memory:
//     reduce-scatter of the flat gradient buckets + TF-semantics Adam on the owned slice + all-gather of the new parameters.
//
// The step shards by batch: every rank holds replicas of the model and a flat fp32 gradient bucket
// (4.2 MB critic / 4.9 MB generator).  With NCCL the update is all-reduce (latency-bound, ~80 us at 8 GPUs and launched from the
// host BETWEEN two CUDA graphs) followed by a full Adam pass on every rank.  Here the bucket and the parameter buffer of every
// rank live in cudaMalloc memory opened on all peers through CUDA IPC (NVSwitch: every GPU reads every peer at NVLink speed), and
// one kernel per rank does, for its slice [rank*n/W, (rank+1)*n/W) of the flat buffers:
//   A. barrier over the peers' flag words (st.release.sys / ld.acquire.sys): "every rank's gradients are complete";
//   B. g = sum over ranks r = 0..W-1 of peer_g[r][i]            (W loads in flight per thread, fixed order: all ranks that
//      update element i -- only its owner -- see the same sum, so replicas stay bit-identical);
//   C. Adam on (p, m, v)[i] with the owner's moments (the moments of the other slices are never touched on this rank);
//   D. the new p[i] is stored into EVERY rank's parameter buffer;
//   E. barrier: "every rank has written its slice everywhere" -- only then may the kernel finish (the next kernel re-packs the
//      bf16 filter operands from the local parameter buffer) and may a peer zero the gradients it let us read.
// No host involvement, no NCCL call, graph-capturable: a data-parallel step is ONE CUDA graph, like the single-GPU step.
// Flags are monotonic epoch counters kept in device memory, so a captured launch works on every replay.
#include "common.cuh"

namespace ctgan {

constexpr int PEER_MAX = 16;

struct PeerTable {
    float* g[PEER_MAX];            // every rank's gradient bucket (peer-mapped), g[rank] = local
    float* p[PEER_MAX];            // every rank's parameter buffer
    unsigned int* flags[PEER_MAX]; // every rank's flag block: [2][PEER_MAX] arrival words + [0..3] local words after them
    int world, rank;
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned int* p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// bounded spin: a dead peer must surface as a trap (launch failure), never as a hung GPU
#define PEER_SPIN(cond)                                                        \
    do {                                                                       \
        long long _t0 = clock64();                                             \
        while (!(cond)) {                                                      \
            if (clock64() - _t0 > 20000000000ll) { __trap(); }                 \
        }                                                                      \
    } while (0)

// flag block layout (unsigned int words): arrive[phase][PEER_MAX] at [phase*PEER_MAX + r], then
//   LOCAL_EPOCH = 2*PEER_MAX      (the epoch of the last completed launch; read at start, bumped by the last CTA out)
//   LOCAL_GO    = 2*PEER_MAX + 1  (CTA 0 -> all CTAs: the peers' gradients are ready)
//   LOCAL_DONE  = 2*PEER_MAX + 2  (count of CTAs that have stored their part of the slice everywhere)
//   LOCAL_EXIT  = 2*PEER_MAX + 3  (count of CTAs that have read LOCAL_EPOCH: the last one out may bump it)
constexpr int F_EPOCH = 2 * PEER_MAX, F_GO = F_EPOCH + 1, F_DONE = F_EPOCH + 2, F_EXIT = F_EPOCH + 3;
constexpr int PEER_FLAG_WORDS = F_EPOCH + 4;

__global__ void __launch_bounds__(256)
peer_reduce_adam_kernel(const __grid_constant__ PeerTable T, float* __restrict__ m, float* __restrict__ v, int64_t n4,
                        float lr_t, float b1, float b2, float eps, float gscale, const float* __restrict__ lr_t_dev) {
    ctgan::pdl_entry();
    const int W = T.world, me = T.rank;
    unsigned int* my = T.flags[me];
    const unsigned int epoch = ld_acquire_gpu(my + F_EPOCH) + 1u;           // same value in every CTA: bumped only after all have read it
    if (lr_t_dev) lr_t = lr_t_dev[0];

    // ---- A: all ranks' gradient buckets are complete
    if (blockIdx.x == 0) {
        if ((int)threadIdx.x < W) {
            __threadfence_system();
            st_release_sys(T.flags[threadIdx.x] + 0 * PEER_MAX + me, epoch);            // tell rank threadIdx.x: I have arrived
            PEER_SPIN(ld_acquire_sys(my + 0 * PEER_MAX + threadIdx.x) >= epoch);        // wait for rank threadIdx.x
        }
        __syncthreads();
        if (threadIdx.x == 0) st_release_gpu(my + F_GO, epoch);
    } else if (threadIdx.x == 0) {
        PEER_SPIN(ld_acquire_gpu(my + F_GO) >= epoch);
    }
    __syncthreads();

    // ---- B..D on the owned slice (float4 granularity)
    const int64_t lo = (n4 * me) / W, hi = (n4 * (me + 1)) / W;
    for (int64_t i = lo + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
        float4 gs[PEER_MAX];
#pragma unroll
        for (int r = 0; r < PEER_MAX; ++r)
            if (r < W) gs[r] = __ldcv(reinterpret_cast<const float4*>(T.g[r]) + i);     // volatile-cached: never a stale line
        float4 g = gs[0];
#pragma unroll
        for (int r = 1; r < PEER_MAX; ++r)
            if (r < W) { g.x += gs[r].x; g.y += gs[r].y; g.z += gs[r].z; g.w += gs[r].w; }
        float4 pp = reinterpret_cast<const float4*>(T.p[me])[i];
        float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        float* ga = &g.x; float* pa = &pp.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gr = ga[j] * gscale;
            ma[j] = b1 * ma[j] + (1.f - b1) * gr;
            va[j] = b2 * va[j] + (1.f - b2) * gr * gr;
            pa[j] -= lr_t * ma[j] / (sqrtf(va[j]) + eps);
        }
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
#pragma unroll
        for (int r = 0; r < PEER_MAX; ++r)
            if (r < W) reinterpret_cast<float4*>(T.p[r])[i] = pp;
    }

    // ---- E: my slice is everywhere; wait until everybody else's is here
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(my + F_DONE, 1u);
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) PEER_SPIN(ld_acquire_gpu(my + F_DONE) >= gridDim.x);      // every CTA of this rank has stored its part
        __syncthreads();
        if ((int)threadIdx.x < W) {
            __threadfence_system();
            st_release_sys(T.flags[threadIdx.x] + 1 * PEER_MAX + me, epoch);
            PEER_SPIN(ld_acquire_sys(my + 1 * PEER_MAX + threadIdx.x) >= epoch);
        }
        __syncthreads();
    }
    // the last CTA out resets the local counters and publishes the epoch for the next launch
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int k = atomicAdd(my + F_EXIT, 1u);
        if (k == gridDim.x - 1) {
            my[F_DONE] = 0u; my[F_EXIT] = 0u;
            __threadfence();
            st_release_gpu(my + F_EPOCH, epoch);
        }
    }
}

}  // namespace ctgan

using namespace ctgan;

/* ---- peer-visible allocations and their CUDA IPC handles (64 bytes each) */
extern "C" int ctgan_peer_alloc(void** out, int64_t bytes) {
    CTGAN_REQUIRE(out && bytes > 0, CTGAN_ERR_BAD_DESC, "peer_alloc: bad args");
    cudaError_t e = cudaMalloc(out, (size_t)bytes);
    if (e != cudaSuccess) return cuda_status(e, "peer_alloc cudaMalloc");
    e = cudaMemset(*out, 0, (size_t)bytes);
    return cuda_status(e, "peer_alloc cudaMemset");
}
extern "C" int ctgan_peer_free(void* ptr) { return cuda_status(cudaFree(ptr), "peer_free"); }
extern "C" int ctgan_ipc_get_handle(const void* ptr, void* handle64) {
    CTGAN_REQUIRE(ptr && handle64, CTGAN_ERR_BAD_DESC, "ipc_get_handle: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    return cuda_status(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), const_cast<void*>(ptr)), "cudaIpcGetMemHandle");
}
extern "C" int ctgan_ipc_open_handle(const void* handle64, void** out) {
    CTGAN_REQUIRE(handle64 && out, CTGAN_ERR_BAD_DESC, "ipc_open_handle: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    return cuda_status(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
}
extern "C" int ctgan_ipc_close_handle(void* ptr) { return cuda_status(cudaIpcCloseMemHandle(ptr), "cudaIpcCloseMemHandle"); }
extern "C" int ctgan_peer_flag_bytes(void) { return PEER_FLAG_WORDS * 4; }

/* g / p / flags: HOST arrays of `world` device pointers (entry `rank` = this rank's own buffers); m, v: this rank's moments */
extern "C" int ctgan_peer_reduce_adam(int world, int rank, float* const* g, float* const* p, void* const* flags, float* m, float* v,
                                      int64_t n, float lr_t, float beta1, float beta2, float eps, float grad_scale,
                                      const float* lr_t_dev, void* stream) {
    CTGAN_REQUIRE(world >= 1 && world <= PEER_MAX && rank >= 0 && rank < world && g && p && flags && m && v, CTGAN_ERR_BAD_DESC,
                  "peer_reduce_adam: bad args (world <= %d)", PEER_MAX);
    CTGAN_REQUIRE(n > 0 && n % 4 == 0, CTGAN_ERR_BAD_DESC, "peer_reduce_adam: n must be a positive multiple of 4");
    PeerTable T = {};
    for (int r = 0; r < world; ++r) {
        CTGAN_REQUIRE(g[r] && p[r] && flags[r] && ((reinterpret_cast<uintptr_t>(g[r]) | reinterpret_cast<uintptr_t>(p[r])) & 15) == 0,
                      CTGAN_ERR_BAD_DESC, "peer_reduce_adam: null or misaligned peer pointer %d", r);
        T.g[r] = g[r]; T.p[r] = p[r]; T.flags[r] = reinterpret_cast<unsigned int*>(flags[r]);
    }
    T.world = world; T.rank = rank;
    const int64_t n4 = n / 4, slice = (n4 + world - 1) / world;
    int grid = (int)((slice + 255) / 256);
    if (grid > sm_count()) grid = sm_count();          // all CTAs must be co-resident: they synchronise through memory
    if (grid < 1) grid = 1;
    CTGAN_LAUNCH((peer_reduce_adam_kernel), grid, 256, 0, as_stream(stream), T, m, v, n4, lr_t, beta1, beta2, eps, grad_scale, lr_t_dev);
    CTGAN_CHECK_LAUNCH("peer_reduce_adam");
    return 0;
}
