// conv_tf32.cu -- the fp32-storage path of the conv family on the tensor cores: tcgen05.mma kind::tf32.
//
// north_star names "TF32/BF16 inputs and fp32 accumulation".  The BF16 kernels (conv_tc.cu) store activations in bf16; this
// file runs the SAME implicit-GEMM pipelines on float activations and float filter packs: TMA boxes of 32 channels (one
// 128-byte swizzle row), UMMA K = 8 per instruction (four per k-block, like the four K = 16 steps of a 64-channel bf16
// block), products with TF32 operand precision (10-bit mantissa), fp32 accumulation in TMEM, fp32 epilogue and stores.
// It replaces the FP32-FMA SIMT kernels (conv_simt.cu) for the stride-1 layers with Cin, Cout multiples of 128 when the
// caller opts in (kernels.config.tf32): ~10x their speed at half the BF16 tensor peak.  No storage rounding of
// activations or cotangents; what remains relative to fp32 is the operand rounding of every product (2^-11).
//   conv_fprop_tf32_kernel<HALO, EPI>   fprop, stride-1 dgrad (flipped pack) and Linear; persistent, grouped stages, the
//                                       lean issue loop and double-buffered TMEM accumulators of conv_fprop_tc_lean_kernel
//   conv_wgrad_tf32_multi_kernel        filter gradients of many layers in one launch (work list of conv_wgrad_multi.cu);
//                                       both operands MN-major, four 32-channel groups per 128-channel side
//   pack_filters_f32_kernel             [taps][Cout][Cin] (fprop) / tap-flipped [taps][Cin][Cout] (dgrad) float operands
// Replaces tf.nn.conv2d / Conv2DBackpropInput / Conv2DBackpropFilter / tf.matmul (TG/tflib/ops/conv2d.py:106-112,
// linear.py:132-136) on the fp32 path.
#include "tc_common.cuh"

namespace ctgan {
namespace tc {

constexpr int KB32 = 32;           // float elements per 128-byte swizzle row = channels per k-block

__device__ __forceinline__ void ldg8_f32(const float* p, float (&f)[8]) {
    uint32_t r[8];
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(r[e]);
}
__device__ __forceinline__ void stg8_f32(float* p, const float (&f)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "r"(__float_as_uint(f[0])), "r"(__float_as_uint(f[1])), "r"(__float_as_uint(f[2])), "r"(__float_as_uint(f[3])),
                   "r"(__float_as_uint(f[4])), "r"(__float_as_uint(f[5])), "r"(__float_as_uint(f[6])), "r"(__float_as_uint(f[7])) : "memory");
}

// epilogue of one 128-pixel x 128-channel fp32 accumulator: thread = pixel row; + bias (+ residual) (ReLU) (mask) -> float
template <int EPI>
__device__ __forceinline__ void tf32_epilogue_tile(const FpropParams& p, uint32_t tmem_addr, int q, int lane,
                                                   int w0, int h0, int n0, int co0, bool relu) {
    int t = q * 32 + lane;
    const int bw = t % p.BW; t /= p.BW;
    const int bh = t % p.BH; const int bn = t / p.BH;
    const int n = n0 + bn, h = h0 + bh, w = w0 + bw;
    const bool valid = (n < p.N) && (h < p.H) && (w < p.W);
    const int64_t pix = ((int64_t)n * p.H + h) * p.W + w;
    float* yrow = reinterpret_cast<float*>(p.y) + pix * p.Cout + co0;
    const int64_t rpix = (p.flags & CTGAN_EPI_RES_UP2) ? ((int64_t)n * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1) : pix;
    const float* rrow = p.residual ? reinterpret_cast<const float*>(p.residual) + rpix * p.Cout + co0 : nullptr;
    const float* mrow = (EPI == EPI_MASK) ? reinterpret_cast<const float*>(p.relu_mask) + pix * p.Cout + co0 : nullptr;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v32[32];
        tmem_ld32(tmem_addr + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v32);
        if (valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                float v[8];
                if (p.bias) {
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + co0 + c0 + j));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + co0 + c0 + j + 4));
                    v[0] = b0.x; v[1] = b0.y; v[2] = b0.z; v[3] = b0.w; v[4] = b1.x; v[5] = b1.y; v[6] = b1.z; v[7] = b1.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = 0.f;
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] += __uint_as_float(v32[j + e]);
                if (rrow) {
                    float r[8];
                    ldg8_f32(rrow + c0 + j, r);
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] += r[e];
                }
                if (relu) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
                }
                if (EPI == EPI_MASK) {
                    float m[8];
                    ldg8_f32(mrow + c0 + j, m);
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = m[e] > 0.f ? v[e] : 0.f;
                }
                stg8_f32(yrow + c0 + j, v);
            }
        }
    }
}

template <int HALO, int EPI>
__global__ void __launch_bounds__(192, 1)
conv_fprop_tf32_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                       const FpropParams p, const int n_tiles)
{
    ctgan::pdl_launch_dependents();
    constexpr int BLOCK_N = 128;
    constexpr int STAGES = 3;
    constexpr int NB = HALO ? 3 : 2;                                  // filter boxes (k-blocks) per stage
    constexpr uint32_t A_REGION = HALO ? 24576u : 32768u;             // one halo box | two 16 KB boxes
    constexpr uint32_t B_BYTES = BLOCK_N * 128;                       // 128 rows x 128 B = 16 KB
    constexpr uint32_t STAGE_BYTES = A_REGION + NB * B_BYTES;         // 72 KB | 64 KB
    constexpr int TMEM_COLS = 256;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t s_base = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
    const uint32_t tfull = smem_u32(bars + 2 * STAGES), tempty = smem_u32(bars + 2 * STAGES + 2);

    const int cin_blocks = p.Cin / KB32;
    const int n_blocks = p.Cout / BLOCK_N;
    const int taps = p.kh * p.kw;
    const int kblocks = cin_blocks * taps;                            // non-halo: k-block index = tap * cin_blocks + cb
    const int groups = HALO ? cin_blocks * p.kw : (kblocks + 1) / 2;
    const uint32_t a_bytes = (uint32_t)(HALO ? p.BH + 2 : p.BH) * p.BW * p.BN * 128u;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_w);
        for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull + 8 * s, 1); mbar_init(tempty + 8 * s, 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ctgan::pdl_wait();          // nothing above touches global memory written by earlier kernels

    if (warp == 0 && lane == 0) {
        // ================= TMA producer =================
        int st = 0; uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int nb = tile % n_blocks; int mt = tile / n_blocks;
            const int tw = mt % p.tilesW; mt /= p.tilesW;
            const int th = mt % p.tilesH; const int tn = mt / p.tilesH;
            const int w0 = tw * p.BW, h0 = th * p.BH, n0 = tn * p.BN, co0 = nb * BLOCK_N;
            for (int gi = 0; gi < groups; ++gi) {
                const uint32_t sb = s_base + st * STAGE_BYTES, fb = full0 + 8 * st;
                mbar_wait(empty0 + 8 * st, ph ^ 1);
                if (HALO) {
                    const int cb = gi / p.kw, s = gi - cb * p.kw;
                    mbar_expect_tx(fb, a_bytes + NB * B_BYTES);
                    tma_load_4d(sb, &tmap_x, fb, cb * KB32, w0 + s - p.pad_l, h0 - p.pad_t, n0);
#pragma unroll
                    for (int r = 0; r < NB; ++r)
                        tma_load_3d(sb + A_REGION + r * B_BYTES, &tmap_w, fb, cb * KB32, co0, r * p.kw + s);
                } else {
                    const int kb0 = gi * 2, nk = min(2, kblocks - kb0);
                    mbar_expect_tx(fb, (uint32_t)nk * (a_bytes + B_BYTES));
                    for (int j = 0; j < nk; ++j) {
                        const int kb = kb0 + j, tap = kb / cin_blocks, cb = kb - tap * cin_blocks;
                        const int r = tap / p.kw, s = tap - r * p.kw;
                        tma_load_4d(sb + j * 16384u, &tmap_x, fb, cb * KB32, w0 + s - p.pad_l, h0 + r - p.pad_t, n0);
                        tma_load_3d(sb + A_REGION + j * B_BYTES, &tmap_w, fb, cb * KB32, co0, tap);
                    }
                }
                if (++st == STAGES) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (whole warp runs the loop; one elected lane issues) =================
        constexpr uint32_t idesc = make_idesc_tf32(BLOCK_M, BLOCK_N, 0, 0);
        const uint32_t lo0 = ((s_base & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t row_step = HALO ? ((uint32_t)p.BW * 128u) >> 4 : (16384u >> 4);   // A offset between the NB k-blocks
        int st = 0; uint32_t ph = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            mbar_wait(tempty + 8 * acc, ((uint32_t)(it >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
            for (int gi = 0; gi < groups; ++gi) {
                mbar_wait(full0 + 8 * st, ph);
                tc_fence_after();
                const uint32_t a_lo = lo0 + st * (STAGE_BYTES >> 4);
                const uint32_t b_lo = a_lo + (A_REGION >> 4);
                const int nk = HALO ? NB : min(2, kblocks - gi * 2);
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        if (j < nk) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)              // 8 floats = 32 bytes along K per instruction
                                umma_tf32_lo(d_tmem, a_lo + j * row_step + 2 * k, b_lo + j * (B_BYTES >> 4) + 2 * k, idesc,
                                             (j | k) ? 1u : (gi > 0 ? 1u : 0u));
                        }
                    }
                    umma_commit(empty0 + 8 * st);
                    if (gi == groups - 1) umma_commit(tfull + 8 * acc);
                }
                __syncwarp();
                if (++st == STAGES) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp >= 2) {
        // ================= epilogue warps: TMEM -> registers -> global =================
        const int q = warp & 3;
        const bool relu = (p.flags & CTGAN_EPI_RELU) != 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const int nb = tile % n_blocks; int mt = tile / n_blocks;
            const int tw = mt % p.tilesW; mt /= p.tilesW;
            const int th = mt % p.tilesH; const int tn = mt / p.tilesH;
            mbar_wait(tfull + 8 * acc, (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            tf32_epilogue_tile<EPI>(p, tmem_base + (uint32_t)(acc * BLOCK_N), q, lane, tw * p.BW, th * p.BH, tn * p.BN, nb * BLOCK_N, relu);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty + 8 * acc) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------ filter gradients, many layers per launch
constexpr int WG32_MAX_JOBS = 24;

struct alignas(64) Wgrad32Job {
    CUtensorMap mx, mdy;            // x: box (32 ch, BW, BH + kh - 1, BN);  dY: box (32 ch, BW, BH, BN)
    float* dw;
    int Cin, Cout, kh, kw, pad_t, pad_l;
    int BW, BH, BN, chunksW, chunksH;
    int co_blocks;
    int splits, chunks_per_split, total_chunks;
    int item0;
    uint32_t a_bytes;               // bytes of one 32-channel group of the x box
};
struct Wgrad32Table {
    Wgrad32Job job[WG32_MAX_JOBS];
    int n_jobs, n_items;
};
struct Wg32Item { int j, ci0, co0, s_tap, chunk0, nchunks; };

__device__ __forceinline__ Wg32Item wg32_decode(const Wgrad32Table& tab, int item) {
    int j = 0;
#pragma unroll 1
    while (j + 1 < tab.n_jobs && item >= tab.job[j + 1].item0) ++j;
    const Wgrad32Job& J = tab.job[j];
    int local = item - J.item0;
    const int z = local % J.splits; local /= J.splits;
    const int s = local % J.kw; const int tile = local / J.kw;
    Wg32Item it;
    it.j = j;
    it.ci0 = (tile / J.co_blocks) * 128; it.co0 = (tile % J.co_blocks) * 128;
    it.s_tap = s;
    it.chunk0 = z * J.chunks_per_split;
    it.nchunks = min(J.chunks_per_split, J.total_chunks - it.chunk0);
    return it;
}

__global__ void __launch_bounds__(192, 1)
conv_wgrad_tf32_multi_kernel(const __grid_constant__ Wgrad32Table tab)
{
    ctgan::pdl_launch_dependents();
    constexpr int STAGES = 2;
    constexpr uint32_t A_SLOT = 16384, B_SLOT = 8192;                    // one 32-channel group of the x halo box / of the dY chunk
    constexpr uint32_t STAGE_BYTES = 4 * A_SLOT + 4 * B_SLOT;             // 96 KB
    constexpr int TMEM_COLS = 512;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t s_base = smem_u32(smem);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
    const uint32_t tfull = smem_u32(bars + 2 * STAGES), tempty = smem_u32(bars + 2 * STAGES + 1);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tfull, 1); mbar_init(tempty, 4);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ctgan::pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            int st = 0; uint32_t ph = 0;
            for (int item = blockIdx.x; item < tab.n_items; item += gridDim.x) {
                const Wg32Item it = wg32_decode(tab, item);
                const Wgrad32Job& J = tab.job[it.j];
                const CUtensorMap* mx = &J.mx; const CUtensorMap* mdy = &J.mdy;
                const int dx = it.s_tap - J.pad_l;
                for (int c = 0; c < it.nchunks; ++c) {
                    int ch = it.chunk0 + c;
                    const int cw = ch % J.chunksW; ch /= J.chunksW;
                    const int chh = ch % J.chunksH; const int cn = ch / J.chunksH;
                    const int w0 = cw * J.BW, h0 = chh * J.BH, n0 = cn * J.BN;
                    const uint32_t sb = s_base + st * STAGE_BYTES, fb = full0 + 8 * st;
                    mbar_wait(empty0 + 8 * st, ph ^ 1);
                    mbar_expect_tx(fb, 4 * J.a_bytes + 4 * B_SLOT);
#pragma unroll
                    for (int gq = 0; gq < 4; ++gq) {
                        tma_load_4d(sb + gq * A_SLOT, mx, fb, it.ci0 + 32 * gq, w0 + dx, h0 - J.pad_t, n0);
                        tma_load_4d(sb + 4 * A_SLOT + gq * B_SLOT, mdy, fb, it.co0 + 32 * gq, w0, h0, n0);
                    }
                    if (++st == STAGES) { st = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // both operands MN-major: smem rows are K (pixels), 32 channels = 128 B per row; LBO = distance between 32-channel
        // groups.  TF32 MN-major operands use the 32-byte-atom swizzle (tc_common.cuh: umma_tf32_mn_lo).
        constexpr uint32_t idesc = make_idesc_tf32(BLOCK_M, 128, 1, 1);
        const uint32_t a_lo0 = ((s_base & 0x3FFFFu) >> 4) | ((A_SLOT >> 4) << 16);
        const uint32_t b_lo0 = (((s_base + 4 * A_SLOT) & 0x3FFFFu) >> 4) | ((B_SLOT >> 4) << 16);
        int st = 0; uint32_t ph = 0;
        uint32_t n = 0;
        for (int item = blockIdx.x; item < tab.n_items; item += gridDim.x, ++n) {
            const Wg32Item it = wg32_decode(tab, item);
            const Wgrad32Job& J = tab.job[it.j];
            const int kh = J.kh;
            const uint32_t row_step = ((uint32_t)J.BW * 128u) >> 4;
            mbar_wait(tempty, (n & 1u) ^ 1u);
            tc_fence_after();
            for (int c = 0; c < it.nchunks; ++c) {
                mbar_wait(full0 + 8 * st, ph);
                tc_fence_after();
                const uint32_t a_lo = a_lo0 + st * (STAGE_BYTES >> 4), b_lo = b_lo0 + st * (STAGE_BYTES >> 4);
                if (elect_one()) {
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        if (r < kh) {
#pragma unroll
                            for (int k = 0; k < 8; ++k)               // 8 pixel rows = 1024 bytes along K per instruction
                                umma_tf32_mn_lo(tmem_base + (uint32_t)(r * 128), a_lo + r * row_step + 64 * k, b_lo + 64 * k, idesc,
                                                k ? 1u : (c > 0 ? 1u : 0u));
                        }
                    }
                    umma_commit(empty0 + 8 * st);
                    if (c == it.nchunks - 1) umma_commit(tfull);
                }
                __syncwarp();
                if (++st == STAGES) { st = 0; ph ^= 1; }
            }
        }
    } else {
        const int q = warp & 3;
        uint32_t n = 0;
        for (int item = blockIdx.x; item < tab.n_items; item += gridDim.x, ++n) {
            const Wg32Item it = wg32_decode(tab, item);
            const Wgrad32Job& J = tab.job[it.j];
            mbar_wait(tfull, n & 1u);
            tc_fence_after();
            const int ci = it.ci0 + q * 32 + lane;
            for (int r = 0; r < J.kh; ++r) {
                float* dst = J.dw + ((int64_t)(r * J.kw + it.s_tap) * J.Cin + ci) * J.Cout + it.co0;
#pragma unroll 1
                for (int c0 = 0; c0 < 128; c0 += 32) {
                    uint32_t acc[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(r * 128 + c0), acc);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                                     ::"l"(dst + c0 + j), "f"(__uint_as_float(acc[j])), "f"(__uint_as_float(acc[j + 1])),
                                       "f"(__uint_as_float(acc[j + 2])), "f"(__uint_as_float(acc[j + 3])) : "memory");
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------ filter packs (float -> float, re-laid out)
// entry e (blockIdx.y) = {src offset in the flat parameter buffer, dst offset in the pack buffer (floats), taps, Cin, Cout, flip}
//   flip == 0: wp[t][o][c] = w[t][c][o]          flip == 1: wp[t][c][o] = w[taps-1-t][c][o]
// The tensor core TRUNCATES fp32 operands to TF32; the filter operand is rounded to nearest here (cvt.rna.tf32), which
// removes the systematic shrink of that side of every product.
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
struct Pack32Entry { long long src, dst; int taps, cin, cout, flip; };
__global__ void pack_filters_f32_kernel(const float* __restrict__ flat, float* __restrict__ packs, const Pack32Entry* __restrict__ table) {
    ctgan::pdl_entry();
    const Pack32Entry e = table[blockIdx.y];
    const float* w = flat + e.src;
    float* wp = packs + e.dst;
    const int64_t total = (int64_t)e.taps * e.cin * e.cout;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        if (e.flip) {
            const int64_t t = i / ((int64_t)e.cin * e.cout), rem = i - t * (int64_t)e.cin * e.cout;
            wp[i] = round_tf32(w[(int64_t)(e.taps - 1 - t) * e.cin * e.cout + rem]);
        } else {
            const int c = (int)(i % e.cin); const int64_t q = i / e.cin;
            const int o = (int)(q % e.cout); const int t = (int)(q / e.cout);
            wp[i] = round_tf32(w[((int64_t)t * e.cin + c) * e.cout + o]);
        }
    }
}
__global__ void pack_filter_f32_kernel(const float* __restrict__ w, float* __restrict__ wp, int taps, int Cin, int Cout, int flip) {
    ctgan::pdl_entry();
    const int64_t total = (int64_t)taps * Cin * Cout;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        if (flip) {
            const int64_t t = i / ((int64_t)Cin * Cout), rem = i - t * (int64_t)Cin * Cout;
            wp[i] = round_tf32(w[(int64_t)(taps - 1 - t) * Cin * Cout + rem]);
        } else {
            const int c = (int)(i % Cin); const int64_t q = i / Cin;
            const int o = (int)(q % Cout); const int t = (int)(q / Cout);
            wp[i] = round_tf32(w[((int64_t)t * Cin + c) * Cout + o]);
        }
    }
}

template <int HALO, int EPI>
static int launch_fprop_tf32(const CUtensorMap& mx, const CUtensorMap& mw, const FpropParams& p, cudaStream_t st) {
    constexpr size_t stage = HALO ? (24576 + 3 * 16384) : (32768 + 2 * 16384);
    constexpr size_t smem = 3 * stage + 1024 + (2 * 3 + 4) * 8 + 16;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_fprop_tf32_kernel<HALO, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_status(e, "fprop_tf32 smem attribute");
        attr_set = true;
    }
    const int n_tiles = p.tilesW * p.tilesH * p.tilesN * (p.Cout / 128);
    const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
    CTGAN_LAUNCH((conv_fprop_tf32_kernel<HALO, EPI>), grid, 192, smem, st, mx, mw, p, n_tiles);
    CTGAN_CHECK_LAUNCH("conv_fprop_tf32");
    return 0;
}

}  // namespace tc
}  // namespace ctgan

using namespace ctgan;
using namespace ctgan::tc;

static int check_tf32_desc(const ctgan_conv_desc* d, const char* who) {
    CTGAN_REQUIRE(d != nullptr, CTGAN_ERR_BAD_DESC, "%s: null descriptor", who);
    CTGAN_REQUIRE(d->x_dtype == CTGAN_F32 && d->y_dtype == CTGAN_F32, CTGAN_ERR_UNSUPPORTED, "%s: the TF32 path needs float activations", who);
    CTGAN_REQUIRE(d->stride == 1 && d->Ho == d->H && d->Wo == d->W, CTGAN_ERR_UNSUPPORTED, "%s: needs stride 1 and Ho==H, Wo==W", who);
    CTGAN_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->kh > 0 && d->kw > 0 && d->pad_t >= 0 && d->pad_l >= 0 &&
                  d->pad_t < d->kh && d->pad_l < d->kw, CTGAN_ERR_BAD_DESC, "%s: bad geometry", who);
    CTGAN_REQUIRE(d->Cin > 0 && d->Cin % 32 == 0 && d->Cout > 0 && d->Cout % 128 == 0, CTGAN_ERR_UNSUPPORTED,
                  "%s: Cin must be a multiple of 32 and Cout of 128", who);
    CTGAN_REQUIRE(ctgan_tc_available(), CTGAN_ERR_UNSUPPORTED, "%s: device is not sm_100", who);
    return 0;
}

extern "C" int ctgan_conv_tf32_ok(const ctgan_conv_desc* d) {
    return d && d->x_dtype == CTGAN_F32 && d->y_dtype == CTGAN_F32 && d->stride == 1 && d->Ho == d->H && d->Wo == d->W && d->N > 0 &&
           d->Cin > 0 && d->Cin % 32 == 0 && d->Cout > 0 && d->Cout % 128 == 0 && d->pad_t >= 0 && d->pad_l >= 0 &&
           d->pad_t < d->kh && d->pad_l < d->kw && ctgan_tc_available();
}

extern "C" int ctgan_conv_fprop_tf32(const ctgan_conv_desc* d, const void* x, const void* wp, const float* bias, const void* residual,
                                     const void* relu_mask, void* y, int flags, void* stream) {
    if (int r = check_tf32_desc(d, "conv_fprop_tf32")) return r;
    CTGAN_REQUIRE(x && wp && y, CTGAN_ERR_BAD_DESC, "conv_fprop_tf32: null pointer");
    CTGAN_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wp) | reinterpret_cast<uintptr_t>(y) |
                    reinterpret_cast<uintptr_t>(residual) | reinterpret_cast<uintptr_t>(relu_mask)) & 31) == 0, CTGAN_ERR_BAD_DESC,
                  "conv_fprop_tf32: pointers must be 32-byte aligned");
    CTGAN_REQUIRE(!(flags & CTGAN_EPI_RES_UP2) || (residual && d->H % 2 == 0 && d->W % 2 == 0), CTGAN_ERR_BAD_DESC,
                  "conv_fprop_tf32: RES_UP2 needs a residual and even H, W");
    FpropParams p = {};
    p.N = d->N; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout;
    p.kh = d->kh; p.kw = d->kw; p.pad_t = d->pad_t; p.pad_l = d->pad_l;
    pixel_box(d->H, d->W, BLOCK_M, &p.BW, &p.BH, &p.BN);
    p.tilesW = ceil_div(d->W, p.BW); p.tilesH = ceil_div(d->H, p.BH); p.tilesN = ceil_div(d->N, p.BN);
    p.flags = flags;
    p.y = reinterpret_cast<__nv_bfloat16*>(y);                       // float tensors behind the shared parameter block
    p.bias = bias;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
    p.relu_mask = reinterpret_cast<const __nv_bfloat16*>(relu_mask);
    const bool halo = d->kh == 3 && p.BN == 1 && (p.BW % 8) == 0 && (uint32_t)(p.BH + 2) * p.BW * 128u <= 24576u;
    CUtensorMap mx, mw;
    if (int r = make_act_map_f32(&mx, x, d->N, d->H, d->W, d->Cin, p.BW, halo ? p.BH + 2 : p.BH, p.BN)) return r;
    if (int r = make_filter_map_f32(&mw, wp, d->kh * d->kw, d->Cout, d->Cin, 128)) return r;
    cudaStream_t st = as_stream(stream);
    if (halo) return relu_mask ? launch_fprop_tf32<1, EPI_MASK>(mx, mw, p, st) : launch_fprop_tf32<1, EPI_PLAIN>(mx, mw, p, st);
    return relu_mask ? launch_fprop_tf32<0, EPI_MASK>(mx, mw, p, st) : launch_fprop_tf32<0, EPI_PLAIN>(mx, mw, p, st);
}

extern "C" int ctgan_conv_wgrad_tf32_multi_ok(const ctgan_conv_desc* d) {
    if (!ctgan_conv_tf32_ok(d) || d->Cin % 128) return 0;
    if (!((d->kh == 3 && d->kw == 3) || (d->kh == 1 && d->kw == 1))) return 0;
    int BW, BH, BN;
    pixel_box(d->H, d->W, 64, &BW, &BH, &BN);
    if (d->kh == 1) return 1;
    return BN == 1 && BW % 8 == 0 && (uint32_t)(BH + 2) * BW * 128u <= 16384u;
}

extern "C" int ctgan_conv_wgrad_tf32_multi(int n, const ctgan_conv_desc* descs, const void* const* xs, const void* const* dys,
                                           float* const* dws, void* stream) {
    CTGAN_REQUIRE(n > 0 && descs && xs && dys && dws, CTGAN_ERR_BAD_DESC, "conv_wgrad_tf32_multi: bad args");
    constexpr size_t smem = (size_t)2 * 98304 + 1024 + (2 * 2 + 2) * 8 + 16;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tf32_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_status(e, "wgrad_tf32_multi smem attribute");
        attr_set = true;
    }
    for (int base = 0; base < n; base += WG32_MAX_JOBS) {
        const int nj = (n - base) < WG32_MAX_JOBS ? (n - base) : WG32_MAX_JOBS;
        static thread_local Wgrad32Table tab;
        long long total = 0;
        long long col_work[WG32_MAX_JOBS];
        for (int i = 0; i < nj; ++i) {
            const ctgan_conv_desc* d = descs + base + i;
            CTGAN_REQUIRE(ctgan_conv_wgrad_tf32_multi_ok(d), CTGAN_ERR_UNSUPPORTED, "conv_wgrad_tf32_multi: job %d is not eligible", base + i);
            CTGAN_REQUIRE(xs[base + i] && dys[base + i] && dws[base + i] &&
                          ((reinterpret_cast<uintptr_t>(xs[base + i]) | reinterpret_cast<uintptr_t>(dys[base + i]) |
                            reinterpret_cast<uintptr_t>(dws[base + i])) & 15) == 0,
                          CTGAN_ERR_BAD_DESC, "conv_wgrad_tf32_multi: job %d: null or misaligned pointer", base + i);
            Wgrad32Job& J = tab.job[i];
            J.dw = dws[base + i];
            J.Cin = d->Cin; J.Cout = d->Cout; J.kh = d->kh; J.kw = d->kw; J.pad_t = d->pad_t; J.pad_l = d->pad_l;
            pixel_box(d->H, d->W, 64, &J.BW, &J.BH, &J.BN);
            J.chunksW = ceil_div(d->W, J.BW); J.chunksH = ceil_div(d->H, J.BH);
            J.total_chunks = J.chunksW * J.chunksH * ceil_div(d->N, J.BN);
            J.co_blocks = d->Cout / 128;
            J.a_bytes = (uint32_t)(J.BH + d->kh - 1) * J.BW * J.BN * 128u;
            if (int r = make_act_map_f32(&J.mx, xs[base + i], d->N, d->H, d->W, d->Cin, J.BW, J.BH + d->kh - 1, J.BN, true)) return r;
            if (int r = make_act_map_f32(&J.mdy, dys[base + i], d->N, d->H, d->W, d->Cout, J.BW, J.BH, J.BN, true)) return r;
            col_work[i] = (long long)J.total_chunks * d->kh;
            total += col_work[i] * (d->Cin / 128) * J.co_blocks * d->kw;
        }
        long long target = (total + 2ll * sm_count() - 1) / (2ll * sm_count());
        if (target < 24) target = 24;
        int items = 0;
        for (int i = 0; i < nj; ++i) {
            Wgrad32Job& J = tab.job[i];
            long long s = (col_work[i] + target / 2) / target;
            if (s < 1) s = 1;
            if (s > J.total_chunks) s = J.total_chunks;
            J.chunks_per_split = ceil_div(J.total_chunks, s);
            J.splits = ceil_div(J.total_chunks, J.chunks_per_split);
            J.item0 = items;
            items += (J.Cin / 128) * J.co_blocks * J.kw * J.splits;
        }
        tab.n_jobs = nj; tab.n_items = items;
        const int grid = items < sm_count() ? items : sm_count();
        CTGAN_LAUNCH((conv_wgrad_tf32_multi_kernel), grid, 192, smem, as_stream(stream), tab);
        CTGAN_CHECK_LAUNCH("conv_wgrad_tf32_multi");
    }
    return 0;
}

extern "C" int ctgan_pack_filters_multi_f32(const float* flat, float* packs, const void* table, int n_entries, void* stream) {
    CTGAN_REQUIRE(flat && packs && table && n_entries > 0 && n_entries <= 65535, CTGAN_ERR_BAD_DESC, "pack_filters_multi_f32: bad args");
    CTGAN_LAUNCH((pack_filters_f32_kernel), dim3(64, n_entries), 256, 0, as_stream(stream), flat, packs, reinterpret_cast<const Pack32Entry*>(table));
    CTGAN_CHECK_LAUNCH("pack_filters_multi_f32");
    return 0;
}

extern "C" int ctgan_pack_filter_f32(const float* w, float* wp, int taps, int Cin, int Cout, int transpose_flip, void* stream) {
    CTGAN_REQUIRE(w && wp && taps > 0 && Cin > 0 && Cout > 0, CTGAN_ERR_BAD_DESC, "pack_filter_f32: bad args");
    const int64_t total = (int64_t)taps * Cin * Cout;
    CTGAN_LAUNCH((pack_filter_f32_kernel), elementwise_grid(total, 256), 256, 0, as_stream(stream), w, wp, taps, Cin, Cout, transpose_flip);
    CTGAN_CHECK_LAUNCH("pack_filter_f32");
    return 0;
}
