// tc_common.cuh -- shared pieces of the tcgen05 / TMA kernels (conv_tc.cu, conv_wgrad_multi.cu, ...): PTX wrappers,
// UMMA shared-memory / instruction descriptors, and the host-side tensor-map builders.  sm_100a only.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace ctgan {
namespace tc {

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a descriptor/programming error must surface as a trap (-> cudaErrorLaunchFailure),
// never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) { __trap(); }
    }
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem], BF16 x BF16 -> FP32
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = TMEM lane)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// one elected lane of a converged warp (the surrounding control flow stays warp-uniform, so descriptor and barrier
// address arithmetic can live in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ------------------------------------------------------------------ descriptors
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4   [32,46) stride byte offset >> 4
//   [46,48) version = 1 (Blackwell)   [49,52) base offset = 0   [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) @4, a/b format BF16 (1) @7/@10,
// a_major @15, b_major @16 (0 = K-major, 1 = MN-major), N>>3 @17, M>>4 @24.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// tcgen05.mma from pre-built 32-bit low descriptor words (the issuing thread stays lean: see conv_fprop_tc_lean_kernel)
__device__ __forceinline__ void umma_bf16_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    // descriptor = {lo: start address (>>4) | LBO(16 B) << 16,  hi: SBO(1024 B) | version 1 | SWIZZLE_128B}
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(0x40004040u) : "memory");
}

// the same for kind::tf32 (a/b format TF32 = 2): fp32 operands in shared memory, 10-bit mantissa products, fp32 accumulate
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(0x40004040u) : "memory");
}

// MN-major TF32 operands exist in ONE shared-memory layout only (CUTLASS: "for mn-major tf32 operands, SW128_32B is the only
// available smem layout"): 128-byte rows whose 32-byte chunks are swizzled by (row mod 4) -- TMA mode SWIZZLE_128B_ATOM_32B,
// descriptor layout type 1 (SWIZZLE_128B_BASE32B), K atoms of 4 rows: SBO = 512 B.
__device__ __forceinline__ void umma_tf32_mn_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    // hi word: SBO (512 B >> 4 = 0x20) | version 1 (bit 46) | layout type 1 (bits 61..63)
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(0x20004020u) : "memory");
}

constexpr int BLOCK_M = 128;       // tile rows = TMEM lanes
constexpr int BLOCK_K = 64;        // bf16 elements per 128-byte swizzle row
constexpr int UMMA_K  = 16;

// ------------------------------------------------------------------ forward-family parameters and epilogue helpers
struct FpropParams {
    int N, H, W, Cin, Cout;
    int kh, kw, pad_t, pad_l;
    int BW, BH, BN;                // pixel box of one M tile: BN*BH*BW == 128
    int tilesW, tilesH, tilesN;
    int hn;                        // 1: the pixels of a tile are ordered [h][n][w] (several small images per tile, halo pipeline)
    int flags;
    __nv_bfloat16* y;
    const float* bias;
    const __nv_bfloat16* residual;
    const __nv_bfloat16* relu_mask;    // output is zeroed where this tensor (shape of y) is <= 0: dgrad into a ReLU output
    // EPI_ACTDROP epilogue (lean kernels): y = v * m, m = (v > 0 ? 1 : slope) * (keep < 1 ? floor(keep + u) / keep : 1), v = conv + bias,
    // u = Philox(seed, offset + dyn[0] + NHWC element index); m is stored next to y; out_s2d: y and m are written in the
    // space-to-depth layout [N, H/2, W/2, 4*Cout] the next stride-2 layer consumes (H, W even)
    __nv_bfloat16* mult;
    float slope, keep;
    unsigned long long seed, offset;
    const unsigned long long* dyn;
    int out_s2d;                   // any epilogue (CTGAN_EPI_OUT_S2D / the actdrop entry): y (and m) in the space-to-depth layout
    int out_d2s;                   // CTGAN_EPI_OUT_D2S: y is a space-to-depth image, written as the plain tensor it stands for
    // CTGAN_EPI_S2D_SKIP: the 3x3 filter embeds a stride-2 skip_k x skip_k one (conv_s2d.cu): (tap, phase) blocks that hold no
    // filter element are all-zero and are neither loaded nor multiplied.  skip_mode 0: phases = 64-channel blocks of the
    // INPUT (fprop over the space-to-depth image), 1: phases = 128-channel blocks of the OUTPUT, taps flipped (its dgrad)
    int skip_k, skip_mode, skip_pad_t, skip_pad_l;
};

// live embedded taps (bit R of the result) of phase coordinate d: source tap 2(R-1) + d + pad inside [0, k)
__device__ __forceinline__ uint32_t s2d_live3(int k, int pad, int d, bool flip) {
    uint32_t m = 0;
#pragma unroll
    for (int R = 0; R < 3; ++R) {
        const int r = 2 * (R - 1) + d + pad;
        if (r >= 0 && r < k) m |= 1u << (flip ? 2 - R : R);
    }
    return m;
}
// (row mask, column mask) of the live taps for input-channel block cb / output-channel block co0 of a launch; 7, 7 = all
__device__ __forceinline__ void s2d_live_masks(const FpropParams& p, int cb, int co0, uint32_t& rmask, uint32_t& cmask) {
    rmask = cmask = 7u;
    if (p.skip_k > 0) {
        const int ph = p.skip_mode ? co0 / (p.Cout >> 2) : (cb * 64) / (p.Cin >> 2);
        rmask = s2d_live3(p.skip_k, p.skip_pad_t, ph >> 1, p.skip_mode != 0);
        cmask = s2d_live3(p.skip_k, p.skip_pad_l, ph & 1, p.skip_mode != 0);
    }
}


enum { EPI_PLAIN = 0, EPI_MASK = 1, EPI_ACTDROP = 2 };

__device__ __forceinline__ void ldg16_bf16(const __nv_bfloat16* p, bool wide, uint32_t (&rw)[8]) {
    if (wide) {
        asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(rw[0]), "=r"(rw[1]), "=r"(rw[2]), "=r"(rw[3]), "=r"(rw[4]), "=r"(rw[5]), "=r"(rw[6]), "=r"(rw[7]) : "l"(p));
    } else {
        const uint4 r0 = *reinterpret_cast<const uint4*>(p), r1 = *reinterpret_cast<const uint4*>(p + 8);
        rw[0] = r0.x; rw[1] = r0.y; rw[2] = r0.z; rw[3] = r0.w; rw[4] = r1.x; rw[5] = r1.y; rw[6] = r1.z; rw[7] = r1.w;
    }
}
__device__ __forceinline__ void stg16_bf16(__nv_bfloat16* p, bool wide, const uint32_t (&ow)[8]) {
    if (wide) {
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     ::"l"(p), "r"(ow[0]), "r"(ow[1]), "r"(ow[2]), "r"(ow[3]), "r"(ow[4]), "r"(ow[5]), "r"(ow[6]), "r"(ow[7]) : "memory");
    } else {
        *reinterpret_cast<uint4*>(p) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        *reinterpret_cast<uint4*>(p + 8) = make_uint4(ow[4], ow[5], ow[6], ow[7]);
    }
}


// ------------------------------------------------------------------ host side: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(ptr);
        else (void)cudaGetLastError();
    }
    return fn;
}

static inline int pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// 4-D map over an NHWC bf16 tensor: dims (C, W, H, N), box (64, BW, BH, BN), 128B swizzle, zero OOB fill
static inline int make_act_map(CUtensorMap* map, const void* base, int N, int H, int W, int C, int BW, int BH, int BN) {
    CTGAN_REQUIRE(BW <= 256 && BH <= 256 && BN <= 256, CTGAN_ERR_UNSUPPORTED, "activation box dimension exceeds 256");
    EncodeTiledFn enc = get_encode_fn();
    CTGAN_REQUIRE(enc != nullptr, CTGAN_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)BW, (cuuint32_t)BH, (cuuint32_t)BN};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CTGAN_REQUIRE(r == CUDA_SUCCESS, CTGAN_ERR_DRIVER, "cuTensorMapEncodeTiled(activation) failed: CUresult %d", (int)r);
    return 0;
}
// the same maps over FLOAT tensors: a 128-byte swizzle row holds 32 channels
static inline int make_act_map_f32(CUtensorMap* map, const void* base, int N, int H, int W, int C, int BW, int BH, int BN,
                                   bool mn_major = false) {
    CTGAN_REQUIRE(BW <= 256 && BH <= 256 && BN <= 256, CTGAN_ERR_UNSUPPORTED, "activation box dimension exceeds 256");
    EncodeTiledFn enc = get_encode_fn();
    CTGAN_REQUIRE(enc != nullptr, CTGAN_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)BW, (cuuint32_t)BH, (cuuint32_t)BN};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    // mn_major: the operand is consumed MN-major by kind::tf32 (wgrad) -> 32-byte swizzle atoms (see umma_tf32_mn_lo)
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CTGAN_REQUIRE(r == CUDA_SUCCESS, CTGAN_ERR_DRIVER, "cuTensorMapEncodeTiled(float activation) failed: CUresult %d", (int)r);
    return 0;
}
static inline int make_filter_map_f32(CUtensorMap* map, const void* base, int taps, int rows, int K, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    CTGAN_REQUIRE(enc != nullptr, CTGAN_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)taps};
    cuuint64_t strides[2] = {(cuuint64_t)K * 4, (cuuint64_t)rows * K * 4};
    cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CTGAN_REQUIRE(r == CUDA_SUCCESS, CTGAN_ERR_DRIVER, "cuTensorMapEncodeTiled(float filter) failed: CUresult %d", (int)r);
    return 0;
}
// The same NHWC bf16 tensor with the two outer dimensions exchanged, dims (C, W, N, H), box (64, BW, BN, BH): a box that
// holds SEVERAL small images lands in shared memory as [h][n][w] pixels, so that a filter-row shift is a uniform offset of
// BN*BW pixel rows (a whole number of swizzle atoms when BN*BW % 8 == 0) -- the halo pipeline for 8x8 / 4x4 images.
static inline int make_act_map_hn(CUtensorMap* map, const void* base, int N, int H, int W, int C, int BW, int BH, int BN) {
    CTGAN_REQUIRE(BW <= 256 && BH <= 256 && BN <= 256, CTGAN_ERR_UNSUPPORTED, "activation box dimension exceeds 256");
    EncodeTiledFn enc = get_encode_fn();
    CTGAN_REQUIRE(enc != nullptr, CTGAN_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)N, (cuuint64_t)H};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)BW, (cuuint32_t)BN, (cuuint32_t)BH};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CTGAN_REQUIRE(r == CUDA_SUCCESS, CTGAN_ERR_DRIVER, "cuTensorMapEncodeTiled(activation, [h][n][w]) failed: CUresult %d", (int)r);
    return 0;
}
// 3-D map over a packed filter [taps][rows][K] bf16: dims (K, rows, taps), box (64, box_rows, 1)
static inline int make_filter_map(CUtensorMap* map, const void* base, int taps, int rows, int K, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    CTGAN_REQUIRE(enc != nullptr, CTGAN_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)taps};
    cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)rows * K * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CTGAN_REQUIRE(r == CUDA_SUCCESS, CTGAN_ERR_DRIVER, "cuTensorMapEncodeTiled(filter) failed: CUresult %d", (int)r);
    return 0;
}

static inline void pixel_box(int H, int W, int pixels, int* BW, int* BH, int* BN) {
    int bw = pow2ceil(W); if (bw > pixels) bw = pixels;
    int bh = pow2ceil(H); if (bh > pixels / bw) bh = pixels / bw;
    *BW = bw; *BH = bh; *BN = pixels / (bw * bh);
}

}  // namespace tc
}  // namespace ctgan
