// conv_tc.cu -- implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM),
// operands staged by TMA, BF16 inputs, FP32 accumulation.  sm_100a only.
//
//   fprop_tc:  Y[n,h,w,co] = sum_{r,s,ci} X[n,h+r-pt,w+s-pl,ci] * Wp[(r,s)][co][ci]   (+bias,+residual,ReLU)
//              GEMM view: M = pixels (128 per CTA: a BN x BH x BW box of the NHWC tensor),
//              N = Cout tile, K = taps x Cin in blocks of 64.  The A operand of tap (r,s) is ONE
//              4-D TMA box load at coordinates shifted by (r-pt, s-pl): TMA's out-of-bounds
//              zero fill IS the SAME padding, so no im2col buffer and no bounds code exist.
//              Both operands are K-major, 128B-swizzled: the box [pixels][64 ch] lands in smem
//              exactly in the canonical UMMA layout (8-row x 128 B atoms, SBO = 1024 B).
//              Also serves stride-1 dgrad (flipped/transposed filter pack) and Linear (H=W=1).
//   wgrad_tc:  dW[(r,s)][ci][co] += sum_{pixels} X[pixel+(r,s)-pad][ci] * dY[pixel][co]
//              GEMM view: M = Cin tile (128), N = Cout tile (128), K = pixels in blocks of 64.
//              The same TMA boxes are consumed as MN-major operands (a_major=b_major=1), so
//              no transposed copy of the activations is ever made.  Up to 4 taps accumulate in
//              4 x 128 TMEM columns per CTA; the pixel range is split across CTAs and reduced
//              with vector fp32 reductions (red.global.add.v4.f32).
//
// Pipeline per CTA (128 threads): warp 0 = TMA producer (one elected lane), warp 1 = MMA
// issuer (one elected lane), warp 2 = TMEM allocator; all four warps run the epilogue
// (tcgen05.ld 32x32b: warp w owns TMEM lanes 32w..32w+31 = tile rows).  smem full/empty
// mbarrier ring between TMA and MMA, tcgen05.commit releases slots and publishes the
// accumulator.  fprop uses 3 x 32 KB stages so two CTAs co-reside per SM and one CTA's
// epilogue overlaps the other's main loop.
//
// Replaces (BF16 path) tf.nn.conv2d / conv2d_transpose / matmul and their gradients:
// TG/tflib/ops/conv2d.py:106-112, deconv2d.py:97-103, linear.py:132-136.
#include "tc_common.cuh"

namespace ctgan {
namespace tc {


// 32-byte global accesses (LDG.256 / STG.256): 16 bf16 channels of one pixel row per instruction = one full sector.
__device__ __forceinline__ void ld16_bf16(const __nv_bfloat16* p, bool wide, float (&f)[16]) {
    uint32_t rw[8];
    if (wide) {
        asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(rw[0]), "=r"(rw[1]), "=r"(rw[2]), "=r"(rw[3]), "=r"(rw[4]), "=r"(rw[5]), "=r"(rw[6]), "=r"(rw[7]) : "l"(p));
    } else {
        const uint4 r0 = *reinterpret_cast<const uint4*>(p), r1 = *reinterpret_cast<const uint4*>(p + 8);
        rw[0] = r0.x; rw[1] = r0.y; rw[2] = r0.z; rw[3] = r0.w; rw[4] = r1.x; rw[5] = r1.y; rw[6] = r1.z; rw[7] = r1.w;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const __nv_bfloat162 r2 = *reinterpret_cast<const __nv_bfloat162*>(&rw[e]);
        f[2 * e] = __bfloat162float(r2.x); f[2 * e + 1] = __bfloat162float(r2.y);
    }
}
__device__ __forceinline__ void st16_bf16(__nv_bfloat16* p, bool wide, const float (&v)[16]) {
    uint32_t ow[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const __nv_bfloat162 o2 = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
        ow[e] = *reinterpret_cast<const uint32_t*>(&o2);
    }
    if (wide) {
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     ::"l"(p), "r"(ow[0]), "r"(ow[1]), "r"(ow[2]), "r"(ow[3]), "r"(ow[4]), "r"(ow[5]), "r"(ow[6]), "r"(ow[7]) : "memory");
    } else {
        *reinterpret_cast<uint4*>(p) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        *reinterpret_cast<uint4*>(p + 8) = make_uint4(ow[4], ow[5], ow[6], ow[7]);
    }
}

// Epilogue of one accumulator tile of NCOLS channels (all fprop kernels): thread = pixel row = TMEM lane; per 32 fp32
// columns from tcgen05.ld: + bias (+ residual) (ReLU) (zero where relu_mask <= 0), bf16 pack, 32-byte stores.
// `bias` points at the tile's first channel (shared or global memory) or is null.
template <int NCOLS, bool BIAS_GLOBAL>
__device__ __forceinline__ void tile_epilogue(const FpropParams& p, uint32_t tmem_addr, const float* bias, int q, int lane,
                                              int w0, int h0, int n0, int co0) {
    int t = q * 32 + lane;
    const int bw = t % p.BW; t /= p.BW;
    const int bh = t % p.BH; const int bn = t / p.BH;
    const int n = n0 + bn, h = h0 + bh, w = w0 + bw;
    const bool valid = (n < p.N) && (h < p.H) && (w < p.W);
    const int64_t pix = ((int64_t)n * p.H + h) * p.W + w;
    __nv_bfloat16* yrow = p.y + pix * p.Cout + co0;
    const int64_t rpix = (p.flags & CTGAN_EPI_RES_UP2) ? ((int64_t)n * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1) : pix;
    const __nv_bfloat16* rrow = p.residual ? p.residual + rpix * p.Cout + co0 : nullptr;
    const __nv_bfloat16* mrow = p.relu_mask ? p.relu_mask + pix * p.Cout + co0 : nullptr;
    const bool relu = (p.flags & CTGAN_EPI_RELU) != 0;
    const bool wide = ((reinterpret_cast<uintptr_t>(p.y) | reinterpret_cast<uintptr_t>(p.residual) |
                        reinterpret_cast<uintptr_t>(p.relu_mask)) & 31) == 0;
    const bool bias_vec = bias != nullptr && (reinterpret_cast<uintptr_t>(bias) & 15) == 0;
#pragma unroll 1
    for (int c0 = 0; c0 < NCOLS; c0 += 32) {
        uint32_t v32[32];
        tmem_ld32(tmem_addr + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v32);     // warp-collective
        if (valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 16) {
                float v[16];
                if (bias_vec) {
#pragma unroll
                    for (int e = 0; e < 16; e += 4) {
                        const float4 b4 = BIAS_GLOBAL ? __ldg(reinterpret_cast<const float4*>(bias + c0 + j + e))
                                                      : *reinterpret_cast<const float4*>(bias + c0 + j + e);
                        v[e] = b4.x; v[e + 1] = b4.y; v[e + 2] = b4.z; v[e + 3] = b4.w;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = bias ? bias[c0 + j + e] : 0.f;
                }
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] += __uint_as_float(v32[j + e]);
                if (rrow) {
                    float r[16];
                    ld16_bf16(rrow + c0 + j, wide, r);
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] += r[e];
                }
                if (relu) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], 0.f);
                }
                if (mrow) {
                    float m[16];
                    ld16_bf16(mrow + c0 + j, wide, m);
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = m[e] > 0.f ? v[e] : 0.f;
                }
                st16_bf16(yrow + c0 + j, wide, v);
            }
        }
    }
}
template <int BLOCK_N>
__device__ __forceinline__ void fprop_epilogue(const FpropParams& p, uint32_t tmem_base, const float* s_bias,
                                               int warp, int lane, int w0, int h0, int n0, int co0) {
    tile_epilogue<BLOCK_N, false>(p, tmem_base, s_bias, warp, lane, w0, h0, n0, co0);
}

// ------------------------------------------------------------------ fprop kernel
template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(128)
conv_fprop_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                     const FpropParams p)
{
    ctgan::pdl_launch_dependents();
    constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;       // 16 KB
    constexpr uint32_t B_BYTES = BLOCK_N * BLOCK_K * 2;
    constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;

    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B atoms need 1024-byte alignment
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
    float* s_bias = reinterpret_cast<float*>(tmem_slot + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), accum_bar = smem_u32(bars + 2 * STAGES);

    // ---- tile coordinates
    int mt = blockIdx.x;
    const int tw = mt % p.tilesW; mt /= p.tilesW;
    const int th = mt % p.tilesH; const int tn = mt / p.tilesH;
    const int w0 = tw * p.BW, h0 = th * p.BH, n0 = tn * p.BN;
    const int co0 = blockIdx.y * BLOCK_N;

    const int cin_blocks = p.Cin / BLOCK_K;
    const int num_kb = p.kh * p.kw * cin_blocks;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_w);
        for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
    if (threadIdx.x < BLOCK_N) s_bias[threadIdx.x] = p.bias ? p.bias[co0 + threadIdx.x] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ctgan::pdl_wait();          // nothing above touches global memory written by earlier kernels

    if (warp == 0 && lane == 0) {
        // ================= TMA producer =================
        int stage = 0; uint32_t phase = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
            const int tap = kb / cin_blocks, cb = kb - tap * cin_blocks;
            const int r = tap / p.kw, s = tap - r * p.kw;
            mbar_wait(empty0 + 8 * stage, phase ^ 1);
            const uint32_t a_dst = smem_base + stage * STAGE_BYTES, b_dst = a_dst + A_BYTES;
            const uint32_t fb = full0 + 8 * stage;
            mbar_expect_tx(fb, STAGE_BYTES);
            tma_load_4d(a_dst, &tmap_x, fb, cb * BLOCK_K, w0 + s - p.pad_l, h0 + r - p.pad_t, n0);
            tma_load_3d(b_dst, &tmap_w, fb, cb * BLOCK_K, co0, tap);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1 && lane == 0) {
        // ================= MMA issuer =================
        constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, 0, 0);
        int stage = 0; uint32_t phase = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(full0 + 8 * stage, phase);
            tc_fence_after();
            const uint32_t a_src = smem_base + stage * STAGE_BYTES, b_src = a_src + A_BYTES;
            const uint64_t a_desc = make_smem_desc(a_src, 16, 1024);
            const uint64_t b_desc = make_smem_desc(b_src, 16, 1024);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                // advance 32 bytes (16 bf16) along K inside the 128-byte swizzle row: +2 in 16-byte units
                umma_bf16(tmem_base, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
            }
            umma_commit(empty0 + 8 * stage);                   // frees this smem slot when the MMAs retire
            if (kb == num_kb - 1) umma_commit(accum_bar);      // accumulator complete
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    }
    __syncwarp();

    // ================= epilogue: TMEM -> registers -> global =================
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    fprop_epilogue<BLOCK_N>(p, tmem_base, s_bias, warp, lane, w0, h0, n0, co0);
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------ fprop kernel, persistent, grouped stages
// Epilogue of one 128-pixel x 128-channel accumulator of the lean / pair kernels (thread = pixel row = TMEM lane, 32-byte
// global accesses).  The variant is a TEMPLATE parameter: an untaken runtime branch for an extra operand measured 3-4 %
// slower on the whole step (registers of the four epilogue warps).
//   EPI_PLAIN    y = [relu](conv + bias [+ residual | residual upsampled 2x])
//   EPI_MASK     ... and y = 0 where relu_mask <= 0  (Conv2DBackpropInput followed by ReluGrad: the dgrad into a tensor
//                that is a ReLU output -- no separate mask-multiply kernel in the backward chains)
//   EPI_ACTDROP  v = conv + bias;  m = (v > 0 ? 1 : slope) * floor(keep + u) / keep;  y = v * m, m stored beside y:
//                Conv2D -> LeakyReLU -> tf.nn.dropout of the DCGAN critics (TG/CT_gan_cifar.py:84-96, CT_gan_mnist.py:92-104,
//                bias at TG/tflib/ops/conv2d.py:114-120) in the conv epilogue.  u comes from Philox4x32-10 in registers at
//                the element's NHWC index (the stream act_dropout_kernel draws from), optionally written in the
//                space-to-depth layout of the next stride-2 layer.
//   S2D = true: the space-to-depth variants (CTGAN_EPI_OUT_S2D / OUT_D2S / S2D_SKIP).  A template parameter for the same reason:
//                with the extra address arithmetic and mask tests as runtime branches the DEFAULT launches of the step ran
//                3.5 % slower (critic graph 924 -> 957 us, same box).
//   NORES = true: launches without a residual operand (most of them): no residual pointer, loads or adds in the loop.
template <int EPI, bool S2D, bool NORES>
__device__ __forceinline__ void lean_epilogue_tile(const FpropParams& p, uint32_t tmem_addr, int q, int lane,
                                                   int w0, int h0, int n0, int co0, bool relu) {
    int t = q * 32 + lane;
    const int bw = t % p.BW; t /= p.BW;
    int bh, bn;
    if (p.hn) { bn = t % p.BN; bh = t / p.BN; }          // [h][n][w] pixel order (several small images per tile)
    else { bh = t % p.BH; bn = t / p.BH; }
    const int n = n0 + bn, h = h0 + bh, w = w0 + bw;
    const bool valid = (n < p.N) && (h < p.H) && (w < p.W);
    const int64_t pix = ((int64_t)n * p.H + h) * p.W + w;
    int64_t orow = pix * p.Cout;                                     // element offset of this pixel's row in y (and m)
    if ((S2D || EPI == EPI_ACTDROP) && p.out_s2d)
        orow = ((((int64_t)n * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1)) * 4 + ((h & 1) * 2 + (w & 1))) * p.Cout;
    __nv_bfloat16* yrow = p.y + orow + co0;
    if (S2D && p.out_d2s) {
        // this conv's output is the space-to-depth image [N, H, W, 4C] of a [N, 2H, 2W, C] tensor (the dgrad of a stride-2
        // conv on the 3x3 route): channel block co0 is phase (dy, dx) = co0 / C; write that tensor in its plain layout
        const int C = p.Cout >> 2, ph = co0 / C;
        yrow = p.y + ((((int64_t)n * (2 * p.H) + 2 * h + (ph >> 1)) * (2 * p.W) + 2 * w + (ph & 1))) * C + (co0 - ph * C);
    }
    const int64_t rpix = (p.flags & CTGAN_EPI_RES_UP2) ? ((int64_t)n * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1) : pix;
    const __nv_bfloat16* rrow = (!NORES && EPI != EPI_ACTDROP && p.residual) ? p.residual + rpix * p.Cout + co0 : nullptr;
    const __nv_bfloat16* mrow = (EPI == EPI_MASK) ? p.relu_mask + pix * p.Cout + co0 : nullptr;
    __nv_bfloat16* mult_row = (EPI == EPI_ACTDROP) ? p.mult + orow + co0 : nullptr;
    // NORES launches are dispatched only with 32-byte aligned y / relu_mask (every torch allocation): no 16-byte fallback code
    const bool wide = NORES ? true : ((reinterpret_cast<uintptr_t>(p.y) | reinterpret_cast<uintptr_t>(p.residual) |
                        (EPI == EPI_MASK ? reinterpret_cast<uintptr_t>(p.relu_mask) : 0) |
                        (EPI == EPI_ACTDROP ? reinterpret_cast<uintptr_t>(p.mult) : 0)) & 31) == 0;
    // EPI_ACTDROP: absolute Philox stream index of channel co0 of this pixel (a multiple of 4: offsets are 4-aligned)
    unsigned long long se0 = 0;
    bool drop = false;
    float inv_keep = 1.f;
    if (EPI == EPI_ACTDROP) {
        drop = p.keep < 1.f;
        inv_keep = 1.f / p.keep;
        se0 = p.offset + (p.dyn ? *p.dyn : 0ull) + (unsigned long long)(pix * p.Cout + co0);
    }
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v32[32];
        tmem_ld32(tmem_addr + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v32);
        if (valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 16) {
                float v[16];
                if (p.bias) {
#pragma unroll
                    for (int e = 0; e < 16; e += 4) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + co0 + c0 + j + e));
                        v[e] = b4.x; v[e + 1] = b4.y; v[e + 2] = b4.z; v[e + 3] = b4.w;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = 0.f;
                }
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] += __uint_as_float(v32[j + e]);
                if (!NORES && rrow) {
                    uint32_t rw[8];
                    ldg16_bf16(rrow + c0 + j, wide, rw);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const __nv_bfloat162 r2 = *reinterpret_cast<const __nv_bfloat162*>(&rw[e]);
                        v[2 * e] += __bfloat162float(r2.x); v[2 * e + 1] += __bfloat162float(r2.y);
                    }
                }
                if (relu) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], 0.f);
                }
                if (EPI == EPI_MASK) {
                    uint32_t mw[8];
                    ldg16_bf16(mrow + c0 + j, wide, mw);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const __nv_bfloat162 m2 = *reinterpret_cast<const __nv_bfloat162*>(&mw[e]);
                        v[2 * e] = __bfloat162float(m2.x) > 0.f ? v[2 * e] : 0.f;
                        v[2 * e + 1] = __bfloat162float(m2.y) > 0.f ? v[2 * e + 1] : 0.f;
                    }
                }
                uint32_t ow[8];
                if (EPI == EPI_ACTDROP) {
                    uint32_t mo[8];
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        uint32_t r[4] = {0u, 0u, 0u, 0u};
                        if (drop) Philox::block(p.seed, (se0 + (unsigned long long)(c0 + j + 4 * b)) >> 2, r);
                        float mm[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float mult = v[4 * b + e] > 0.f ? 1.f : p.slope;
                            if (drop) mult *= floorf(p.keep + Philox::to_uniform(r[e])) * inv_keep;
                            mm[e] = mult;
                        }
                        // y = v * (the ROUNDED multiplier): y == v * m holds exactly for the stored m (as act_dropout_kernel)
                        const __nv_bfloat162 ma = __floats2bfloat162_rn(mm[0], mm[1]), mb = __floats2bfloat162_rn(mm[2], mm[3]);
                        mo[2 * b] = *reinterpret_cast<const uint32_t*>(&ma); mo[2 * b + 1] = *reinterpret_cast<const uint32_t*>(&mb);
                        const __nv_bfloat162 ya = __floats2bfloat162_rn(v[4 * b] * __bfloat162float(ma.x), v[4 * b + 1] * __bfloat162float(ma.y));
                        const __nv_bfloat162 yb = __floats2bfloat162_rn(v[4 * b + 2] * __bfloat162float(mb.x), v[4 * b + 3] * __bfloat162float(mb.y));
                        ow[2 * b] = *reinterpret_cast<const uint32_t*>(&ya); ow[2 * b + 1] = *reinterpret_cast<const uint32_t*>(&yb);
                    }
                    stg16_bf16(mult_row + c0 + j, wide, mo);
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const __nv_bfloat162 o2 = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
                        ow[e] = *reinterpret_cast<const uint32_t*>(&o2);
                    }
                }
                stg16_bf16(yrow + c0 + j, wide, ow);
            }
        }
    }
}

// Same roles as the persistent kernel above, but the MMA-issuing thread is kept as lean as possible: with N = 128 one
// tcgen05.mma occupies the tensor pipe for only 64 cycles, so every scalar instruction of the issuing thread between two
// MMAs shows up as idle tensor time (tests/micro/mma_rate.cu: 75 cycles per MMA with a bare loop, 163 with a barrier
// wait + fence + commit per 4 MMAs).  Here ONE mbarrier wait and ONE commit cover a group of 8-12 MMAs:
//   HALO = 1 (3x3, tiles = whole rows of one image): stage = one (BH+2)-row activation box + the 3 filter boxes of the
//            taps (r = 0..2, s) of one 64-channel block                                            -> 12 MMAs, 72 KB
//   HALO = 0 (1x1, 8x8 / 4x4 tiles, linear): stage = 2 consecutive (activation, filter) box pairs   ->  8 MMAs, 64 KB
// Three stages in flight; descriptors are pre-built 32-bit words plus immediate offsets.

template <int HALO, int EPI, bool S2D, bool NORES>
__global__ void __launch_bounds__(192, 1)
conv_fprop_tc_lean_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                          const FpropParams p, const int n_tiles)
{
    ctgan::pdl_launch_dependents();
    constexpr int BLOCK_N = 128;
    constexpr int STAGES = 3;
    constexpr int NB = HALO ? 3 : 2;                                  // filter boxes (k-blocks) per stage
    constexpr uint32_t A_REGION = HALO ? 24576u : 32768u;             // one halo box | two 16 KB boxes
    constexpr uint32_t B_BYTES = BLOCK_N * BLOCK_K * 2;               // 16 KB
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

    const int cin_blocks = p.Cin / BLOCK_K;
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
                uint32_t rmask = 7u, cmask = 7u;
                if (HALO && S2D) {
                    s2d_live_masks(p, gi / p.kw, co0, rmask, cmask);
                    if (!((cmask >> (gi % p.kw)) & 1u)) continue;      // an all-zero filter column of this phase: no stage
                }
                const uint32_t sb = s_base + st * STAGE_BYTES, fb = full0 + 8 * st;
                mbar_wait(empty0 + 8 * st, ph ^ 1);
                if (HALO) {
                    const int cb = gi / p.kw, s = gi - cb * p.kw;
                    mbar_expect_tx(fb, a_bytes + (S2D ? (uint32_t)__popc(rmask) : (uint32_t)NB) * B_BYTES);
                    if (p.hn) tma_load_4d(sb, &tmap_x, fb, cb * BLOCK_K, w0 + s - p.pad_l, n0, h0 - p.pad_t);   // dims (C, W, N, H)
                    else      tma_load_4d(sb, &tmap_x, fb, cb * BLOCK_K, w0 + s - p.pad_l, h0 - p.pad_t, n0);
#pragma unroll
                    for (int r = 0; r < NB; ++r)
                        if (!S2D || ((rmask >> r) & 1u)) tma_load_3d(sb + A_REGION + r * B_BYTES, &tmap_w, fb, cb * BLOCK_K, co0, r * p.kw + s);
                } else {
                    const int kb0 = gi * 2, nk = min(2, kblocks - kb0);
                    mbar_expect_tx(fb, (uint32_t)nk * (a_bytes + B_BYTES));
                    for (int j = 0; j < nk; ++j) {
                        const int kb = kb0 + j, tap = kb / cin_blocks, cb = kb - tap * cin_blocks;
                        const int r = tap / p.kw, s = tap - r * p.kw;
                        tma_load_4d(sb + j * 16384u, &tmap_x, fb, cb * BLOCK_K, w0 + s - p.pad_l, h0 + r - p.pad_t, n0);
                        tma_load_3d(sb + A_REGION + j * B_BYTES, &tmap_w, fb, cb * BLOCK_K, co0, tap);
                    }
                }
                if (++st == STAGES) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (whole warp runs the loop; one elected lane issues) =================
        constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, 0, 0);
        // low descriptor word of byte address a: (a >> 4) | (LBO 16 B >> 4) << 16
        uint32_t lo[STAGES];
#pragma unroll
        for (int s = 0; s < STAGES; ++s) lo[s] = (((s_base + s * STAGE_BYTES) & 0x3FFFFu) >> 4) | (1u << 16);
        // A offset between the NB k-blocks: one image row (HALO; BN images side by side in [h][n][w] order) | one 16 KB box
        const uint32_t row_step = HALO ? ((uint32_t)p.BW * (p.hn ? p.BN : 1) * 128u) >> 4 : (16384u >> 4);
        int st = 0; uint32_t ph = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            mbar_wait(tempty + 8 * acc, ((uint32_t)(it >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
            const int co0 = S2D ? (tile % n_blocks) * BLOCK_N : 0;
            int last_gi = groups - 1;                               // the last group that is not skipped (see the producer)
            if (HALO && S2D && p.skip_k > 0) {
                for (; last_gi > 0; --last_gi) {
                    uint32_t rm, cm;
                    s2d_live_masks(p, last_gi / p.kw, co0, rm, cm);
                    if ((cm >> (last_gi % p.kw)) & 1u) break;
                }
            }
            uint32_t accf = 0u;                                     // 0: the first MMA of the tile overwrites the accumulator
            for (int gi = 0; gi < groups; ++gi) {
                uint32_t rmask = 7u, cmask = 7u;
                if (HALO && S2D) {
                    s2d_live_masks(p, gi / p.kw, co0, rmask, cmask);
                    if (!((cmask >> (gi % p.kw)) & 1u)) continue;
                }
                mbar_wait(full0 + 8 * st, ph);
                tc_fence_after();
                const uint32_t a_lo = st == 0 ? lo[0] : (st == 1 ? lo[1] : lo[2]);
                const uint32_t b_lo = a_lo + (A_REGION >> 4);
                const int nk = HALO ? NB : min(2, kblocks - gi * 2);
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        if (j < nk && (!(HALO && S2D) || ((rmask >> j) & 1u))) {
#pragma unroll
                            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                                umma_bf16_lo(d_tmem, a_lo + j * row_step + 2 * k, b_lo + j * (B_BYTES >> 4) + 2 * k, idesc,
                                             S2D ? accf : ((j | k) ? 1u : (gi > 0 ? 1u : 0u)));
                                if (S2D) accf = 1u;
                            }
                        }
                    }
                    umma_commit(empty0 + 8 * st);
                    if (gi == last_gi) umma_commit(tfull + 8 * acc);
                }
                if (S2D) accf = 1u;
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
            const int co0 = nb * BLOCK_N;
            mbar_wait(tfull + 8 * acc, (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            lean_epilogue_tile<EPI, S2D, NORES>(p, tmem_base + (uint32_t)(acc * BLOCK_N), q, lane, tw * p.BW, th * p.BH, tn * p.BN, co0, relu);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty + 8 * acc) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------ fprop, 256-pixel work items ("pair" kernel)
// The lean kernel at 128x128 tiles is bound by the L2 -> shared-memory fill (ncu: 453 MB at ~80 % of the chip's LTS cap for
// D.1.Conv2).  Here a work item is TWO vertically adjacent 128-pixel tiles of one image whenever the CTA's tile range
// allows it: one (2*BH+2)-row halo box of x (40 KB) and the same three filter boxes (48 KB) feed 24 MMAs into two TMEM
// accumulators -- 44 KB per 12 MMAs instead of 72 KB.  Two stages of 88 KB; TMEM 2 x (2 x 128) columns double-buffered.
// Tiles are enumerated n-block-major (tile = nb * m_tiles + mt) and split into contiguous per-CTA ranges, so every warp
// role derives the same item sequence (pair if the next tile is the next row block of the same image, else single).
template <int EPI, bool S2D, bool NORES>
__global__ void __launch_bounds__(192, 1)
conv_fprop_tc_pair_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_x2,
                          const __grid_constant__ CUtensorMap tmap_w, const FpropParams p, const int m_tiles, const int n_tiles)
{
    ctgan::pdl_launch_dependents();
    constexpr int BLOCK_N = 128;
    constexpr int STAGES = 2;
    constexpr uint32_t A_REGION = 40960u;                             // (2*BH+2) rows x BW pixels x 128 B <= 40 KB
    constexpr uint32_t B_BYTES = BLOCK_N * BLOCK_K * 2;               // 16 KB
    constexpr uint32_t STAGE_BYTES = A_REGION + 3 * B_BYTES;          // 88 KB
    constexpr int TMEM_COLS = 512;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t s_base = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
    const uint32_t tfull = smem_u32(bars + 2 * STAGES), tempty = smem_u32(bars + 2 * STAGES + 2);

    const int cin_blocks = p.Cin / BLOCK_K;
    const int groups = cin_blocks * p.kw;
    const uint32_t a1_bytes = (uint32_t)(p.BH + 2) * p.BW * 128u, a2_bytes = (uint32_t)(2 * p.BH + 2) * p.BW * 128u;
    // contiguous tile range of this CTA
    const int t_begin = (int)(((long long)n_tiles * blockIdx.x) / gridDim.x);
    const int t_end = (int)(((long long)n_tiles * (blockIdx.x + 1)) / gridDim.x);
    // pair(t): tile t+1 is in range, same n-block, and the next row block of the same image (tilesW == 1 here)
    auto is_pair = [&](int t) -> bool {
        if (t + 1 >= t_end) return false;
        const int mt = t % m_tiles;
        return (mt + 1 < m_tiles) && ((mt % p.tilesH) + 1 < p.tilesH);
    };

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_x2);
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
        for (int t = t_begin; t < t_end;) {
            const bool pair = is_pair(t);
            const int nb = t / m_tiles, mt = t - nb * m_tiles;
            const int th = mt % p.tilesH, tn = mt / p.tilesH;
            const int h0 = th * p.BH, n0 = tn * p.BN, co0 = nb * BLOCK_N;
            for (int gi = 0; gi < groups; ++gi) {
                const int cb = gi / p.kw, s = gi - cb * p.kw;
                uint32_t rmask = 7u, cmask = 7u;
                if (S2D) {
                    s2d_live_masks(p, cb, co0, rmask, cmask);
                    if (!((cmask >> s) & 1u)) continue;                // an all-zero filter column of this phase: no stage
                }
                const uint32_t sb = s_base + st * STAGE_BYTES, fb = full0 + 8 * st;
                mbar_wait(empty0 + 8 * st, ph ^ 1);
                mbar_expect_tx(fb, (pair ? a2_bytes : a1_bytes) + (S2D ? (uint32_t)__popc(rmask) : 3u) * B_BYTES);
                tma_load_4d(sb, pair ? &tmap_x2 : &tmap_x, fb, cb * BLOCK_K, s - p.pad_l, h0 - p.pad_t, n0);
#pragma unroll
                for (int r = 0; r < 3; ++r)
                    if (!S2D || ((rmask >> r) & 1u)) tma_load_3d(sb + A_REGION + r * B_BYTES, &tmap_w, fb, cb * BLOCK_K, co0, r * p.kw + s);
                if (++st == STAGES) { st = 0; ph ^= 1; }
            }
            t += pair ? 2 : 1;
        }
    } else if (warp == 1) {
        // ================= MMA issuer (whole warp runs the loop; one elected lane issues) =================
        constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, 0, 0);
        const uint32_t lo0 = ((s_base & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t row_step = ((uint32_t)p.BW * 128u) >> 4;          // one image row
        const uint32_t tile1_off = (uint32_t)p.BH * row_step;            // lower tile inside the halo box
        int st = 0; uint32_t ph = 0;
        int it = 0;
        for (int t = t_begin; t < t_end; ++it) {
            const bool pair = is_pair(t);
            const int acc = it & 1;
            mbar_wait(tempty + 8 * acc, ((uint32_t)(it >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 2 * BLOCK_N);
            const int co0 = S2D ? (t / m_tiles) * BLOCK_N : 0;
            int last_gi = groups - 1;                               // the last group that is not skipped (see the producer)
            if (S2D && p.skip_k > 0) {
                for (; last_gi > 0; --last_gi) {
                    uint32_t rm, cm;
                    s2d_live_masks(p, last_gi / p.kw, co0, rm, cm);
                    if ((cm >> (last_gi % p.kw)) & 1u) break;
                }
            }
            uint32_t accf = 0u;                                     // 0: the first MMA of the item overwrites the accumulators
            for (int gi = 0; gi < groups; ++gi) {
                uint32_t rmask = 7u, cmask = 7u;
                if (S2D) {
                    s2d_live_masks(p, gi / p.kw, co0, rmask, cmask);
                    if (!((cmask >> (gi % p.kw)) & 1u)) continue;
                }
                mbar_wait(full0 + 8 * st, ph);
                tc_fence_after();
                const uint32_t a_lo = lo0 + st * (STAGE_BYTES >> 4);
                const uint32_t b_lo = a_lo + (A_REGION >> 4);
                if (elect_one()) {
                    if (pair) {
#pragma unroll
                        for (int r = 0; r < 3; ++r) {
                            if (S2D && !((rmask >> r) & 1u)) continue;
#pragma unroll
                            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                                const uint32_t af = S2D ? accf : ((r | k) ? 1u : (gi > 0 ? 1u : 0u));
                                umma_bf16_lo(d_tmem, a_lo + r * row_step + 2 * k, b_lo + r * (B_BYTES >> 4) + 2 * k, idesc, af);
                                umma_bf16_lo(d_tmem + BLOCK_N, a_lo + tile1_off + r * row_step + 2 * k, b_lo + r * (B_BYTES >> 4) + 2 * k,
                                             idesc, af);
                                if (S2D) accf = 1u;
                            }
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < 3; ++r) {
                            if (S2D && !((rmask >> r) & 1u)) continue;
#pragma unroll
                            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                                umma_bf16_lo(d_tmem, a_lo + r * row_step + 2 * k, b_lo + r * (B_BYTES >> 4) + 2 * k, idesc,
                                             S2D ? accf : ((r | k) ? 1u : (gi > 0 ? 1u : 0u)));
                                if (S2D) accf = 1u;
                            }
                        }
                    }
                    umma_commit(empty0 + 8 * st);
                    if (gi == last_gi) umma_commit(tfull + 8 * acc);
                }
                if (S2D) accf = 1u;
                __syncwarp();
                if (++st == STAGES) { st = 0; ph ^= 1; }
            }
            t += pair ? 2 : 1;
        }
    } else if (warp >= 2) {
        // ================= epilogue warps: TMEM -> registers -> global =================
        const int q = warp & 3;
        const bool relu = (p.flags & CTGAN_EPI_RELU) != 0;
        int it = 0;
        for (int t = t_begin; t < t_end; ++it) {
            const bool pair = is_pair(t);
            const int acc = it & 1;
            const int nb = t / m_tiles, mt = t - nb * m_tiles;
            const int th = mt % p.tilesH, tn = mt / p.tilesH;
            const int h0 = th * p.BH, n0 = tn * p.BN, co0 = nb * BLOCK_N;
            mbar_wait(tfull + 8 * acc, (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            lean_epilogue_tile<EPI, S2D, NORES>(p, tmem_base + (uint32_t)(acc * 2 * BLOCK_N), q, lane, 0, h0, n0, co0, relu);
            if (pair) lean_epilogue_tile<EPI, S2D, NORES>(p, tmem_base + (uint32_t)(acc * 2 * BLOCK_N + BLOCK_N), q, lane, 0, h0 + p.BH, n0, co0, relu);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty + 8 * acc) : "memory");
            t += pair ? 2 : 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------ wgrad kernel
struct WgradParams {
    int N, H, W, Cin, Cout;
    int kh, kw, pad_t, pad_l;
    int BW, BH, BN;                // pixel box of one K chunk: BN*BH*BW == 64
    int chunksW, chunksH, chunksN; // pixel chunks per dimension
    int taps_per_cta;
    int chunks_per_split;
    float* dw;
};

template <int STAGES>
__global__ void __launch_bounds__(128)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                     const WgradParams p)
{
    ctgan::pdl_launch_dependents();
    constexpr int BLOCK_N = 128;
    constexpr uint32_t HALF_BYTES = 64 * 64 * 2;              // [64 px][64 ch] bf16 = 8 KB
    constexpr uint32_t A_BYTES = 2 * HALF_BYTES, B_BYTES = 2 * HALF_BYTES;
    constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;       // 32 KB
    constexpr int TMEM_COLS = 512;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), accum_bar = smem_u32(bars + 2 * STAGES);

    const int co_blocks = p.Cout / 128;
    const int ci0 = (blockIdx.x / co_blocks) * 128, co0 = (blockIdx.x % co_blocks) * 128;
    const int taps = p.kh * p.kw;
    const int tap0 = blockIdx.y * p.taps_per_cta;
    const int ntaps = min(p.taps_per_cta, taps - tap0);
    const int total_chunks = p.chunksN * p.chunksH * p.chunksW;
    const int chunk0 = blockIdx.z * p.chunks_per_split;
    const int nchunks = min(p.chunks_per_split, total_chunks - chunk0);
    const int num_kb = nchunks * ntaps;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_dy);
        for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ctgan::pdl_wait();          // nothing above touches global memory written by earlier kernels

    if (warp == 0 && lane == 0) {
        // ================= TMA producer =================
        int stage = 0; uint32_t phase = 0;
        for (int c = 0; c < nchunks; ++c) {
            int ch = chunk0 + c;
            const int cw = ch % p.chunksW; ch /= p.chunksW;
            const int chh = ch % p.chunksH; const int cn = ch / p.chunksH;
            const int w0 = cw * p.BW, h0 = chh * p.BH, n0 = cn * p.BN;
            for (int tl = 0; tl < ntaps; ++tl) {
                const int tap = tap0 + tl;
                const int r = tap / p.kw, s = tap - r * p.kw;
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                const uint32_t a_dst = smem_base + stage * STAGE_BYTES, b_dst = a_dst + A_BYTES;
                const uint32_t fb = full0 + 8 * stage;
                mbar_expect_tx(fb, STAGE_BYTES);
                tma_load_4d(a_dst,              &tmap_x,  fb, ci0,      w0 + s - p.pad_l, h0 + r - p.pad_t, n0);
                tma_load_4d(a_dst + HALF_BYTES, &tmap_x,  fb, ci0 + 64, w0 + s - p.pad_l, h0 + r - p.pad_t, n0);
                tma_load_4d(b_dst,              &tmap_dy, fb, co0,      w0, h0, n0);
                tma_load_4d(b_dst + HALF_BYTES, &tmap_dy, fb, co0 + 64, w0, h0, n0);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ================= MMA issuer =================
        // both operands MN-major: rows of the smem tile are K (pixels), 64 channels = 128 B contiguous
        constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, 1, 1);
        int stage = 0; uint32_t phase = 0;
        for (int c = 0; c < nchunks; ++c) {
            for (int tl = 0; tl < ntaps; ++tl) {
                mbar_wait(full0 + 8 * stage, phase);
                tc_fence_after();
                const uint32_t a_src = smem_base + stage * STAGE_BYTES, b_src = a_src + A_BYTES;
                // LBO = byte distance between the two 64-channel halves, SBO = 8 pixel rows x 128 B
                const uint64_t a_desc = make_smem_desc(a_src, HALF_BYTES, 1024);
                const uint64_t b_desc = make_smem_desc(b_src, HALF_BYTES, 1024);
#pragma unroll
                for (int k = 0; k < 64 / UMMA_K; ++k) {
                    // advance 16 pixel rows = 2048 bytes along K: +128 in 16-byte units
                    umma_bf16(tmem_base + (uint32_t)(tl * BLOCK_N), a_desc + (uint64_t)(128 * k), b_desc + (uint64_t)(128 * k),
                              idesc, (c | k) != 0);
                }
                umma_commit(empty0 + 8 * stage);
                if (c == nchunks - 1 && tl == ntaps - 1) umma_commit(accum_bar);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    }
    __syncwarp();

    // ================= epilogue: TMEM -> vector reductions into dW (float, HWIO) =================
    if (num_kb > 0) {
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int ci = ci0 + warp * 32 + lane;
        for (int tl = 0; tl < ntaps; ++tl) {
            float* dst = p.dw + ((int64_t)(tap0 + tl) * p.Cin + ci) * p.Cout + co0;
#pragma unroll 1
            for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
                uint32_t acc[32];
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(tl * BLOCK_N + c0), acc);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                                 ::"l"(dst + c0 + j), "f"(__uint_as_float(acc[j])), "f"(__uint_as_float(acc[j + 1])),
                                   "f"(__uint_as_float(acc[j + 2])), "f"(__uint_as_float(acc[j + 3])) : "memory");
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// 3x3 wgrad with the lean issue loop of conv_fprop_tc_lean_kernel.  One CTA owns the three taps (r = 0..2, s) of one
// filter COLUMN s = blockIdx.y: they read the same (BH+2)-row halo box of x, shifted by r rows, and the same dY chunk.
//   stage = x halo box, two 64-channel halves (2 x 16 KB slots) + dY chunk, two halves (2 x 8 KB) = 48 KB for 12 MMAs
// (the per-tap kernel above moves 96 KB for the same 12 MMAs and waits/commits three times).
template <int STAGES>
__global__ void __launch_bounds__(128, 1)
conv_wgrad_tc_lean_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                          const WgradParams p)
{
    ctgan::pdl_launch_dependents();
    constexpr int BLOCK_N = 128;
    constexpr uint32_t A_SLOT = 16384, B_HALF = 8192;
    constexpr uint32_t STAGE_BYTES = 2 * A_SLOT + 2 * B_HALF;             // 48 KB
    constexpr int TMEM_COLS = 512;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform
    const uint32_t s_base = smem_u32(smem);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), accum_bar = smem_u32(bars + 2 * STAGES);

    const int co_blocks = p.Cout / 128;
    const int ci0 = (blockIdx.x / co_blocks) * 128, co0 = (blockIdx.x % co_blocks) * 128;
    const int s_tap = blockIdx.y;
    const int total_chunks = p.chunksN * p.chunksH * p.chunksW;
    const int chunk0 = blockIdx.z * p.chunks_per_split;
    const int nchunks = min(p.chunks_per_split, total_chunks - chunk0);
    const uint32_t a_bytes = (uint32_t)(p.BH + 2) * p.BW * 128u;          // one 64-channel half of the halo box

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_dy);
        for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ctgan::pdl_wait();          // nothing above touches global memory written by earlier kernels

    if (warp == 0) {
        if (lane == 0) {
            // ================= TMA producer =================
            int st = 0; uint32_t ph = 0;
            for (int c = 0; c < nchunks; ++c) {
                int ch = chunk0 + c;
                const int cw = ch % p.chunksW; ch /= p.chunksW;
                const int chh = ch % p.chunksH; const int cn = ch / p.chunksH;
                const int w0 = cw * p.BW, h0 = chh * p.BH;
                const uint32_t sb = s_base + st * STAGE_BYTES, fb = full0 + 8 * st;
                mbar_wait(empty0 + 8 * st, ph ^ 1);
                mbar_expect_tx(fb, 2 * a_bytes + 2 * B_HALF);
                tma_load_4d(sb,                       &tmap_x,  fb, ci0,      w0 + s_tap - p.pad_l, h0 - p.pad_t, cn);
                tma_load_4d(sb + A_SLOT,              &tmap_x,  fb, ci0 + 64, w0 + s_tap - p.pad_l, h0 - p.pad_t, cn);
                tma_load_4d(sb + 2 * A_SLOT,          &tmap_dy, fb, co0,      w0, h0, cn);
                tma_load_4d(sb + 2 * A_SLOT + B_HALF, &tmap_dy, fb, co0 + 64, w0, h0, cn);
                if (++st == STAGES) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (whole warp runs the loop; one elected lane issues) =================
        // both operands MN-major: smem rows are K (pixels); LBO = distance between the two 64-channel halves
        constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, 1, 1);
        const uint32_t a_lo0 = ((s_base & 0x3FFFFu) >> 4) | ((A_SLOT >> 4) << 16);
        const uint32_t b_lo0 = (((s_base + 2 * A_SLOT) & 0x3FFFFu) >> 4) | ((B_HALF >> 4) << 16);
        const uint32_t row_step = ((uint32_t)p.BW * 128u) >> 4;          // one image row down = filter row r + 1
        int st = 0; uint32_t ph = 0;
        for (int c = 0; c < nchunks; ++c) {
            mbar_wait(full0 + 8 * st, ph);
            tc_fence_after();
            const uint32_t a_lo = a_lo0 + st * (STAGE_BYTES >> 4), b_lo = b_lo0 + st * (STAGE_BYTES >> 4);
            if (elect_one()) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
#pragma unroll
                    for (int k = 0; k < 64 / UMMA_K; ++k)     // 16 pixel rows = 2048 bytes along K
                        umma_bf16_lo(tmem_base + (uint32_t)(r * BLOCK_N), a_lo + r * row_step + 128 * k, b_lo + 128 * k, idesc,
                                     k ? 1u : (c > 0 ? 1u : 0u));
                }
                umma_commit(empty0 + 8 * st);
                if (c == nchunks - 1) umma_commit(accum_bar);
            }
            __syncwarp();
            if (++st == STAGES) { st = 0; ph ^= 1; }
        }
    }
    __syncwarp();

    // ================= epilogue: TMEM -> vector reductions into dW (float, HWIO) =================
    if (nchunks > 0) {
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int ci = ci0 + warp * 32 + lane;
        for (int r = 0; r < 3; ++r) {
            float* dst = p.dw + ((int64_t)(r * p.kw + s_tap) * p.Cin + ci) * p.Cout + co0;
#pragma unroll 1
            for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
                uint32_t acc[32];
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(r * BLOCK_N + c0), acc);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                                 ::"l"(dst + c0 + j), "f"(__uint_as_float(acc[j])), "f"(__uint_as_float(acc[j + 1])),
                                   "f"(__uint_as_float(acc[j + 2])), "f"(__uint_as_float(acc[j + 3])) : "memory");
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------ filter packing
// transpose_flip == 0: wp[t][o][c] = w[t][c][o];  == 1: wp[t][c][o] = w[taps-1-t][c][o]
__global__ void pack_filter_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp,
                                   int taps, int Cin, int Cout, int transpose_flip) {
    ctgan::pdl_entry();
    int64_t total = (int64_t)taps * Cin * Cout;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        if (transpose_flip) {
            int64_t t = i / ((int64_t)Cin * Cout), rem = i - t * (int64_t)Cin * Cout;
            wp[i] = __float2bfloat16_rn(w[(int64_t)(taps - 1 - t) * Cin * Cout + rem]);
        } else {
            int c = i % Cin; int64_t q = i / Cin;
            int o = q % Cout; int t = q / Cout;
            wp[i] = __float2bfloat16_rn(w[((int64_t)t * Cin + c) * Cout + o]);
        }
    }
}

// Thin-channel filters (one side has C <= 8 channels, the other Cw): 64-row / 64-column zero-padded operands of the
// im2col GEMMs below.  k = tap*C + c indexes the thin side.
//   kind 0: w [taps][C][Cw] -> wp[Cw][64], wp[j][k] = w[k*Cw + j]            (fprop of a C -> Cw conv)
//   kind 1: w [taps][Cw][C] -> wp[Cw][64], wp[j][k] = w[(t*Cw + j)*C + o]    (dgrad of a Cw -> C conv)
//   kind 2: w [taps][Cw][C] -> wp[64][Cw], wp[k][j] = w[(t*Cw + j)*C + o]    (fprop of a Cw -> C conv)
//   kind 3: w [taps][C][Cw] -> wp[64][Cw], wp[k][j] = w[k*Cw + j]            (dgrad of a C -> Cw conv)
__device__ __forceinline__ void pack_thin_body(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp,
                                               int taps, int C, int Cw, int kind, int64_t start, int64_t step) {
    const int64_t total = (int64_t)64 * Cw;
    const int kreal = taps * C;
    for (int64_t i = start; i < total; i += step) {
        int j, k;
        if (kind <= 1) { j = (int)(i >> 6); k = (int)(i & 63); } else { k = (int)(i / Cw); j = (int)(i - (int64_t)k * Cw); }
        float v = 0.f;
        if (k < kreal) {
            if (kind == 0 || kind == 3) v = w[(int64_t)k * Cw + j];
            else { const int t = k / C, o = k - t * C; v = w[((int64_t)t * Cw + j) * C + o]; }
        }
        wp[i] = __float2bfloat16_rn(v);
    }
}
__global__ void pack_thin_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int taps, int C, int Cw, int kind) {
    ctgan::pdl_entry();
    pack_thin_body(w, wp, taps, C, Cw, kind, blockIdx.x * (int64_t)blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
}

// One launch packs every filter of an optimizer: entry e (blockIdx.y) = {src offset in the flat float parameter
// buffer, dst offset in the bf16 pack buffer, taps, Cin, Cout, transpose_flip}.
struct PackEntry { long long src, dst; int taps, cin, cout, flip; };
__global__ void pack_filters_multi_kernel(const float* __restrict__ flat, __nv_bfloat16* __restrict__ packs,
                                          const PackEntry* __restrict__ table) {
    ctgan::pdl_entry();
    const PackEntry e = table[blockIdx.y];
    const float* w = flat + e.src;
    __nv_bfloat16* wp = packs + e.dst;
    if (e.flip >= 2) {                                   // thin kinds 0..3 (cin, cout = HWIO dims)
        pack_thin_body(w, wp, e.taps, min(e.cin, e.cout), max(e.cin, e.cout), e.flip - 2,
                       blockIdx.x * (int64_t)blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
        return;
    }
    const int64_t total = (int64_t)e.taps * e.cin * e.cout;
    if (e.flip) {                                        // same element order inside a tap: coalesced reads and writes
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
            int64_t t = i / ((int64_t)e.cin * e.cout), rem = i - t * (int64_t)e.cin * e.cout;
            wp[i] = __float2bfloat16_rn(w[(int64_t)(e.taps - 1 - t) * e.cin * e.cout + rem]);
        }
        return;
    }
    // flip == 0 is a [Cin][Cout] -> [Cout][Cin] transpose per tap: 32 x 32 tiles through shared memory, so that the float
    // reads run along Cout and the bf16 writes along Cin (the element-wise version read with a stride of Cout floats and
    // made this launch -- the tail of every optimizer step -- 16 us)
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           // 32 x 8
    const int ct = (e.cin + 31) / 32, ot = (e.cout + 31) / 32;
    const int ntiles = e.taps * ct * ot;
    for (int tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
        const int t = tl / (ct * ot), r = tl - t * ct * ot;
        const int c0 = (r / ot) * 32, o0 = (r % ot) * 32;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + ty + 8 * j, o = o0 + tx;
            tile[ty + 8 * j][tx] = (c < e.cin && o < e.cout) ? w[((int64_t)t * e.cin + c) * e.cout + o] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = o0 + ty + 8 * j, c = c0 + tx;
            if (c < e.cin && o < e.cout) wp[((int64_t)t * e.cout + o) * e.cin + c] = __float2bfloat16_rn(tile[tx][ty + 8 * j]);
        }
    }
}

// ------------------------------------------------------------------ thin-channel convolutions (3-channel image side)
// Discriminator.1.Conv1 / .Shortcut (3 -> DIM_D) and Generator.Output (DIM_G -> 3) have one side with C = 3 channels:
// K = taps*C = 27 is far below a tensor-core k-block and a SIMT implicit GEMM wastes most of its tile.  The thin side
// is expanded into col[pixel][64] (k = tap*C + c, zero for k >= taps*C) so that every member of the conv family is a
// [P x 64] x [64 x Cw] (or [P x Cw] x [Cw x 64]) GEMM on the 1x1 tcgen05 kernels, plus one of these gather kernels.
//   im2col: col[p][(t,c)] = src[p + sign*off(t)][c]           off(t) = (r - pad_t, s - pad_l)
//   col2im: dst[p][c]     = bias[c] + sum_t col[p + sign*off(t)][(t,c)]
__global__ void __launch_bounds__(256)
im2col_thin_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ col,
                   int N, int H, int W, int C, int kh, int kw, int pad_t, int pad_l, int sign) {
    ctgan::pdl_entry();
    // per column k: (dh, dw, c) packed as bytes (dh+8, dw+8, c, valid); built once per block
    __shared__ uint32_t tab[64];
    if (threadIdx.x < 64) {
        const int k = threadIdx.x;
        uint32_t e = 0;
        if (k < kh * kw * C) {
            const int t = k / C, c = k - t * C;
            const int r = t / kw, s = t - r * kw;
            e = (uint32_t)(sign * (r - pad_t) + 8) | ((uint32_t)(sign * (s - pad_l) + 8) << 8) | ((uint32_t)c << 16) | (1u << 24);
        }
        tab[k] = e;
    }
    __syncthreads();
    const int64_t total = (int64_t)N * H * W * 8;            // 8 threads per pixel, one 16-byte store each
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i >> 3;
        const int g = (int)(i & 7);
        const int w = (int)(p % W); const int64_t q = p / W;
        const int h = (int)(q % H);
        const __nv_bfloat16* base = src + p * C;                              // pixel (n, h, w)
        __align__(16) __nv_bfloat16 v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const uint32_t te = tab[g * 8 + e];
            const int dh = (int)(te & 0xff) - 8, dw = (int)((te >> 8) & 0xff) - 8, c = (int)((te >> 16) & 0xff);
            const int hh = h + dh, ww = w + dw;
            const bool ok = (te >> 24) && hh >= 0 && hh < H && ww >= 0 && ww < W;
            v[e] = ok ? base[(dh * W + dw) * C + c] : __float2bfloat16_rn(0.f);
        }
        *reinterpret_cast<uint4*>(col + p * 64 + g * 8) = *reinterpret_cast<const uint4*>(v);
    }
}

__global__ void col2im_thin_kernel(const __nv_bfloat16* __restrict__ col, const float* __restrict__ bias,
                                   __nv_bfloat16* __restrict__ dst, int N, int H, int W, int C, int kh, int kw,
                                   int pad_t, int pad_l, int sign) {
    ctgan::pdl_entry();
    const int64_t total = (int64_t)N * H * W;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int w = (int)(p % W); const int64_t q = p / W;
        const int h = (int)(q % H); const int64_t n = q / H;
        float acc[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = (bias && c < C) ? bias[c] : 0.f;
        for (int r = 0; r < kh; ++r) {
            const int hh = h + sign * (r - pad_t);
            if (hh < 0 || hh >= H) continue;
            for (int s = 0; s < kw; ++s) {
                const int ww = w + sign * (s - pad_l);
                if (ww < 0 || ww >= W) continue;
                const __nv_bfloat16* row = col + ((n * H + hh) * W + ww) * 64 + (r * kw + s) * C;
#pragma unroll
                for (int c = 0; c < 8; ++c) if (c < C) acc[c] += __bfloat162float(row[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) if (c < C) dst[p * C + c] = __float2bfloat16_rn(acc[c]);
    }
}

// dw += wide^T x col over all pixels: M = 128 channels of the wide tensor, N = 64 im2col columns, K = pixels.
// Both operands MN-major straight out of TMA (rows = pixels).  Stage = 2 chunks of 64 pixels (48 KB, 8 MMAs).
// The accumulator element (row j, column k = (t, c)) is added to dw[t*sA + c*sB + j*sC].
struct WgradThinParams {
    long long total_chunks;
    int chunks_per_split, C, kreal;
    int Cw;                        // channels of the wide tensor (a multiple of 64; rows >= Cw of the last tile are zero-filled, not stored)
    long long sA, sB, sC;
    float* dw;
};

template <int STAGES>
__global__ void __launch_bounds__(128, 1)
wgrad_thin_tc_kernel(const __grid_constant__ CUtensorMap tmap_wide, const __grid_constant__ CUtensorMap tmap_col,
                     const WgradThinParams p)
{
    ctgan::pdl_launch_dependents();
    constexpr uint32_t A_BYTES = 16384, A_HALF = 8192, B_BYTES = 8192;
    constexpr uint32_t STAGE_BYTES = 2 * (A_BYTES + B_BYTES);             // 48 KB
    constexpr int TMEM_COLS = 64;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t s_base = smem_u32(smem);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), accum_bar = smem_u32(bars + 2 * STAGES);

    const int cw0 = blockIdx.x * 128;
    const long long chunk0 = (long long)blockIdx.z * p.chunks_per_split;
    const int nchunks = (int)min((long long)p.chunks_per_split, p.total_chunks - chunk0);
    const int groups = (nchunks + 1) / 2;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_wide);
        prefetch_tmap(&tmap_col);
        for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ctgan::pdl_wait();          // nothing above touches global memory written by earlier kernels

    if (warp == 0) {
        if (lane == 0) {
            int st = 0; uint32_t ph = 0;
            for (int g = 0; g < groups; ++g) {
                const uint32_t sb = s_base + st * STAGE_BYTES, fb = full0 + 8 * st;
                const int nk = min(2, nchunks - 2 * g);
                mbar_wait(empty0 + 8 * st, ph ^ 1);
                mbar_expect_tx(fb, (uint32_t)nk * (A_BYTES + B_BYTES));
                for (int j = 0; j < nk; ++j) {
                    const int px0 = (int)((chunk0 + 2 * g + j) * 64);
                    tma_load_4d(sb + j * A_BYTES,          &tmap_wide, fb, cw0,      px0, 0, 0);
                    tma_load_4d(sb + j * A_BYTES + A_HALF, &tmap_wide, fb, cw0 + 64, px0, 0, 0);
                    tma_load_4d(sb + 2 * A_BYTES + j * B_BYTES, &tmap_col, fb, 0, px0, 0, 0);
                }
                if (++st == STAGES) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(BLOCK_M, 64, 1, 1);
        const uint32_t a_lo0 = ((s_base & 0x3FFFFu) >> 4) | ((A_HALF >> 4) << 16);
        const uint32_t b_lo0 = (((s_base + 2 * A_BYTES) & 0x3FFFFu) >> 4) | ((B_BYTES >> 4) << 16);
        int st = 0; uint32_t ph = 0;
        for (int g = 0; g < groups; ++g) {
            mbar_wait(full0 + 8 * st, ph);
            tc_fence_after();
            const uint32_t a_lo = a_lo0 + st * (STAGE_BYTES >> 4), b_lo = b_lo0 + st * (STAGE_BYTES >> 4);
            const int nk = min(2, nchunks - 2 * g);
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (j < nk) {
#pragma unroll
                        for (int k = 0; k < 64 / UMMA_K; ++k)
                            umma_bf16_lo(tmem_base, a_lo + j * (A_BYTES >> 4) + 128 * k, b_lo + j * (B_BYTES >> 4) + 128 * k, idesc,
                                         (j | k) ? 1u : (g > 0 ? 1u : 0u));
                    }
                }
                umma_commit(empty0 + 8 * st);
                if (g == groups - 1) umma_commit(accum_bar);
            }
            __syncwarp();
            if (++st == STAGES) { st = 0; ph ^= 1; }
        }
    }
    __syncwarp();

    if (nchunks > 0) {
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const bool row_ok = cw0 + warp * 32 + lane < p.Cw;
        float* dst = p.dw + (long long)(cw0 + warp * 32 + lane) * p.sC;
        int t = 0, c = 0;
#pragma unroll 1
        for (int c0 = 0; c0 < 64 && c0 < p.kreal; c0 += 32) {
            uint32_t acc[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, acc);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (c0 + j < p.kreal) {
                    if (row_ok) atomicAdd(dst + t * p.sA + c * p.sB, __uint_as_float(acc[j]));
                    if (++c == p.C) { c = 0; ++t; }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------ host side

static int check_tc_desc(const ctgan_conv_desc* d, const char* who) {
    CTGAN_REQUIRE(d != nullptr, CTGAN_ERR_BAD_DESC, "%s: null descriptor", who);
    CTGAN_REQUIRE(d->x_dtype == CTGAN_BF16 && d->y_dtype == CTGAN_BF16, CTGAN_ERR_UNSUPPORTED, "%s: tensor-core path needs BF16 activations", who);
    CTGAN_REQUIRE(d->stride == 1 && d->Ho == d->H && d->Wo == d->W, CTGAN_ERR_UNSUPPORTED, "%s: tensor-core path needs stride 1 and Ho==H, Wo==W", who);
    CTGAN_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->kh > 0 && d->kw > 0 && d->pad_t >= 0 && d->pad_l >= 0 &&
                  d->pad_t < d->kh && d->pad_l < d->kw, CTGAN_ERR_BAD_DESC, "%s: bad geometry", who);
    CTGAN_REQUIRE(d->Cin % 64 == 0 && d->Cout % 64 == 0 && d->Cin > 0 && d->Cout > 0, CTGAN_ERR_UNSUPPORTED, "%s: Cin and Cout must be multiples of 64", who);
    CTGAN_REQUIRE(ctgan_tc_available(), CTGAN_ERR_UNSUPPORTED, "%s: device is not sm_100", who);
    return 0;
}

template <int BLOCK_N, int STAGES>
static int launch_fprop(const CUtensorMap& mx, const CUtensorMap& mw, const FpropParams& p, cudaStream_t st) {
    constexpr size_t smem = (size_t)STAGES * (BLOCK_M * BLOCK_K * 2 + BLOCK_N * BLOCK_K * 2) + 1024 /*align slack*/ +
                            (2 * STAGES + 1) * 8 + 16 + BLOCK_N * 4;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_fprop_tc_kernel<BLOCK_N, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_status(e, "fprop_tc smem attribute");
        attr_set = true;
    }
    dim3 grid(p.tilesW * p.tilesH * p.tilesN, p.Cout / BLOCK_N);
    CTGAN_LAUNCH((conv_fprop_tc_kernel<BLOCK_N, STAGES>), grid, 128, smem, st, mx, mw, p);
    CTGAN_CHECK_LAUNCH("conv_fprop_tc");
    return 0;
}

template <int HALO, int EPI, bool S2D = false, bool NORES = false>
static int launch_fprop_lean(const CUtensorMap& mx, const CUtensorMap& mw, const FpropParams& p, cudaStream_t st) {
    constexpr size_t stage = HALO ? (24576 + 3 * 16384) : (32768 + 2 * 16384);
    constexpr size_t smem = 3 * stage + 1024 + (2 * 3 + 4) * 8 + 16;
    static_assert(smem <= 227 * 1024, "lean fprop: shared memory budget");
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_fprop_tc_lean_kernel<HALO, EPI, S2D, NORES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_status(e, "fprop_tc_lean smem attribute");
        attr_set = true;
    }
    const int n_tiles = p.tilesW * p.tilesH * p.tilesN * (p.Cout / 128);
    const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
    CTGAN_LAUNCH((conv_fprop_tc_lean_kernel<HALO, EPI, S2D, NORES>), grid, 192, smem, st, mx, mw, p, n_tiles);
    CTGAN_CHECK_LAUNCH("conv_fprop_tc_lean");
    return 0;
}

template <int EPI, bool S2D = false, bool NORES = false>
static int launch_fprop_pair(const CUtensorMap& mx, const CUtensorMap& mx2, const CUtensorMap& mw, const FpropParams& p, cudaStream_t st) {
    constexpr size_t smem = 2 * (40960 + 3 * 16384) + 1024 + (2 * 2 + 4) * 8 + 16;
    static_assert(smem <= 227 * 1024, "pair fprop: shared memory budget");
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_fprop_tc_pair_kernel<EPI, S2D, NORES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_status(e, "fprop_tc_pair smem attribute");
        attr_set = true;
    }
    const int m_tiles = p.tilesW * p.tilesH * p.tilesN;
    const int n_tiles = m_tiles * (p.Cout / 128);
    const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
    CTGAN_LAUNCH((conv_fprop_tc_pair_kernel<EPI, S2D, NORES>), grid, 192, smem, st, mx, mx2, mw, p, m_tiles, n_tiles);
    CTGAN_CHECK_LAUNCH("conv_fprop_tc_pair");
    return 0;
}

}  // namespace tc
}  // namespace ctgan

using namespace ctgan;
using namespace ctgan::tc;

namespace ctgan { namespace tc {
int splitk_factor(int n_tiles, int groups);                              // conv_splitk.cu
int launch_fprop_splitk(int split, int epi, const CUtensorMap& mx, const CUtensorMap& mw, const FpropParams& p, int n_tiles,
                        cudaStream_t st);
} }
static bool g_use_splitk = true;
/* test / A-B hook: 0 = sub-wave layers run on conv_fprop_tc_lean_kernel<0> instead of the cluster split-K kernel */
extern "C" void ctgan_set_splitk(int on) { g_use_splitk = on != 0; }
static bool g_use_halo = true;
static bool g_use_halo_hn = true;
static int g_fprop_variant = 4;   // 4 = 256-pixel work items where possible (else 3), 3 = persistent grouped-stage (lean) kernel, 1 = one tile per CTA
/* test hook: selects the fprop_tc kernel family (the families are compared in tests/) */
extern "C" void ctgan_set_fprop_variant(int v) { g_fprop_variant = v; }
static bool g_fprop_nores = true;
/* A/B hook: 0 = launches without a residual run the generic instantiations (residual pointer tested at run time) */
extern "C" void ctgan_set_fprop_nores(int on) { g_fprop_nores = on != 0; }
static int g_wgrad_variant = 2;   // 2 = filter-column CTAs sharing one x halo box (3x3), 1 = one (x, dY) box pair per tap
extern "C" void ctgan_set_wgrad_variant(int v) { g_wgrad_variant = v; }
/* test hook: 0 disables the halo-reuse A pipeline (both are compared in tests/) */
// 0: per-tap boxes only; 1: halo boxes wherever eligible; 2: halo boxes for one-image tiles only (not the [h][n][w] ones)
extern "C" void ctgan_set_fprop_halo(int on) { g_use_halo = on != 0; g_use_halo_hn = on == 1; }

// common launcher of the stride-1 tcgen05 forward family; epi selects the epilogue of the lean / pair kernels
static int fprop_tc_launch(const ctgan_conv_desc* d, const void* x, const void* wp, FpropParams& p, int epi, void* stream) {
    p.N = d->N; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout;
    p.kh = d->kh; p.kw = d->kw; p.pad_t = d->pad_t; p.pad_l = d->pad_l;
    pixel_box(d->H, d->W, BLOCK_M, &p.BW, &p.BH, &p.BN);
    p.tilesW = ceil_div(d->W, p.BW); p.tilesH = ceil_div(d->H, p.BH); p.tilesN = ceil_div(d->N, p.BN);
    const int block_n = (d->Cout % 128 == 0) ? 128 : 64;
    CTGAN_REQUIRE(epi != EPI_ACTDROP || block_n == 128, CTGAN_ERR_UNSUPPORTED, "conv_fprop_tc_actdrop: Cout must be a multiple of 128");
    // halo-reuse A pipeline: k x k filters (k > 1) on tiles that are whole rows of one image
    bool halo = g_use_halo && block_n == 128 && d->kh == 3 && p.BN == 1 && (p.BW % 8) == 0 &&
                (uint32_t)(p.BH + d->kh - 1) * p.BW * 128u <= 24576u && g_fprop_variant >= 3;
    // several small images per tile (8x8, 4x4): the halo pipeline on a box laid out [h][n][w] (make_act_map_hn) -- one
    // (BH+2)-row box per (cin block, column shift) instead of one box per tap: 2.4x less activation traffic per tile
    // (a sub-wave layer that cluster split-K takes keeps the per-tap boxes that kernel expects)
    const int n_tiles_all = p.tilesW * p.tilesH * p.tilesN * (d->Cout / 128);
    // launches with a layout-changing epilogue stay on the lean / pair kernels; zero-block skipping (an optimisation only)
    // applies where those kernels run anyway: a sub-wave layer keeps the cluster split-K kernel and multiplies the zeros
    const bool special = (epi != EPI_ACTDROP && p.out_s2d) || p.out_d2s;
    const bool s2d = special || (p.skip_k > 0 && epi != EPI_ACTDROP);    // -> the <.., S2D = true> instantiations
    const bool nores = g_fprop_nores && p.residual == nullptr &&         // -> the <.., NORES = true> instantiations
                       ((reinterpret_cast<uintptr_t>(p.y) | reinterpret_cast<uintptr_t>(p.relu_mask)) & 31) == 0;
    CTGAN_REQUIRE(!(special || p.skip_k > 0) || (block_n == 128 && g_fprop_variant >= 3), CTGAN_ERR_UNSUPPORTED,
                  "conv_fprop_tc: OUT_S2D / OUT_D2S / S2D_SKIP need Cout %% 128 == 0 and the lean kernel family");
    const int split_all = (g_use_splitk && block_n == 128 && !special) ? splitk_factor(n_tiles_all, (d->kh * d->kw * (d->Cin / 64) + 1) / 2) : 0;
    const bool halo_hn = !halo && g_use_halo_hn && block_n == 128 && d->kh == 3 && p.BN > 1 && p.tilesW == 1 && p.tilesH == 1 &&
                         (p.BN * p.BW) % 8 == 0 && (uint32_t)(p.BH + 2) * p.BN * p.BW * 128u <= 24576u && g_fprop_variant >= 3 &&
                         !split_all;
    p.hn = halo_hn ? 1 : 0;
    CUtensorMap mx, mw;
    if (halo_hn) {
        if (int r = make_act_map_hn(&mx, x, d->N, d->H, d->W, d->Cin, p.BW, p.BH + 2, p.BN)) return r;
        halo = true;
    } else if (int r = make_act_map(&mx, x, d->N, d->H, d->W, d->Cin, p.BW, halo ? p.BH + d->kh - 1 : p.BH, p.BN)) return r;
    if (int r = make_filter_map(&mw, wp, d->kh * d->kw, d->Cout, d->Cin, block_n)) return r;
    cudaStream_t st = as_stream(stream);
    const int variant = (epi == EPI_ACTDROP && g_fprop_variant < 3) ? 3 : g_fprop_variant;   // only the lean kernels have that epilogue
    if (variant == 4 && epi != EPI_ACTDROP && halo && p.tilesW == 1 && p.tilesH >= 2 &&
        (uint32_t)(2 * p.BH + 2) * p.BW * 128u <= 40960u && p.tilesH * p.tilesN * (d->Cout / 128) >= 2 * sm_count()) {
        CUtensorMap mx2;                                             // 256-pixel work items (two row blocks per halo box)
        if (int r = make_act_map(&mx2, x, d->N, d->H, d->W, d->Cin, p.BW, 2 * p.BH + 2, 1)) return r;
        if (s2d) return epi == EPI_MASK ? launch_fprop_pair<EPI_MASK, true>(mx, mx2, mw, p, st) : launch_fprop_pair<EPI_PLAIN, true>(mx, mx2, mw, p, st);
        if (nores) return epi == EPI_MASK ? launch_fprop_pair<EPI_MASK, false, true>(mx, mx2, mw, p, st)
                                          : launch_fprop_pair<EPI_PLAIN, false, true>(mx, mx2, mw, p, st);
        return epi == EPI_MASK ? launch_fprop_pair<EPI_MASK>(mx, mx2, mw, p, st) : launch_fprop_pair<EPI_PLAIN>(mx, mx2, mw, p, st);
    }
    if (variant >= 3 && block_n == 128) {                            // persistent, grouped stages, lean issue loop
        if (halo && s2d) {
            if (epi == EPI_MASK) return launch_fprop_lean<1, EPI_MASK, true>(mx, mw, p, st);
            return launch_fprop_lean<1, EPI_PLAIN, true>(mx, mw, p, st);
        }
        if (halo && nores && epi != EPI_ACTDROP)
            return epi == EPI_MASK ? launch_fprop_lean<1, EPI_MASK, false, true>(mx, mw, p, st)
                                   : launch_fprop_lean<1, EPI_PLAIN, false, true>(mx, mw, p, st);
        if (halo) {
            if (epi == EPI_MASK) return launch_fprop_lean<1, EPI_MASK>(mx, mw, p, st);
            if (epi == EPI_ACTDROP) return launch_fprop_lean<1, EPI_ACTDROP>(mx, mw, p, st);
            return launch_fprop_lean<1, EPI_PLAIN>(mx, mw, p, st);
        }
        // fewer tiles than half the SMs: split K over a cluster of 2 / 4 CTAs (conv_splitk.cu)
        const int n_tiles = p.tilesW * p.tilesH * p.tilesN * (d->Cout / 128);
        const int split = (g_use_splitk && !special) ? splitk_factor(n_tiles, (d->kh * d->kw * (d->Cin / 64) + 1) / 2) : 0;
        if (split) return launch_fprop_splitk(split, epi, mx, mw, p, n_tiles, st);
        if (special) {                                               // layout-changing epilogue on per-tap boxes (1x1 / non-halo tiles)
            if (epi == EPI_MASK) return launch_fprop_lean<0, EPI_MASK, true>(mx, mw, p, st);
            return launch_fprop_lean<0, EPI_PLAIN, true>(mx, mw, p, st);
        }
        if (epi == EPI_ACTDROP) return launch_fprop_lean<0, EPI_ACTDROP>(mx, mw, p, st);
        if (nores) return epi == EPI_MASK ? launch_fprop_lean<0, EPI_MASK, false, true>(mx, mw, p, st)
                                          : launch_fprop_lean<0, EPI_PLAIN, false, true>(mx, mw, p, st);
        if (epi == EPI_MASK) return launch_fprop_lean<0, EPI_MASK>(mx, mw, p, st);
        return launch_fprop_lean<0, EPI_PLAIN>(mx, mw, p, st);
    }
    if (block_n == 128) {
        // one tile per CTA (cross-check family).  Few tiles: the K loop is latency-bound, 6-deep TMA ring; many tiles:
        // 3 stages x 2 co-resident CTAs per SM.
        const int ctas = p.tilesW * p.tilesH * p.tilesN * (d->Cout / block_n);
        if (ctas <= sm_count()) return launch_fprop<128, 6>(mx, mw, p, st);
        return launch_fprop<128, 3>(mx, mw, p, st);
    }
    return launch_fprop<64, 4>(mx, mw, p, st);
}

extern "C" int ctgan_conv_fprop_tc(const ctgan_conv_desc* d, const void* x, const void* wp,
                                   const float* bias, const void* residual, void* y, int flags, void* stream) {
    return ctgan_conv_fprop_tc_masked(d, x, wp, bias, residual, nullptr, y, flags, stream);
}

extern "C" int ctgan_conv_fprop_tc_masked(const ctgan_conv_desc* d, const void* x, const void* wp, const float* bias,
                                          const void* residual, const void* relu_mask, void* y, int flags, void* stream) {
    if (int r = check_tc_desc(d, "conv_fprop_tc")) return r;
    CTGAN_REQUIRE((reinterpret_cast<uintptr_t>(relu_mask) & 15) == 0, CTGAN_ERR_BAD_DESC, "conv_fprop_tc: relu_mask must be 16-byte aligned");
    CTGAN_REQUIRE(!(flags & CTGAN_EPI_RES_UP2) || (residual && d->H % 2 == 0 && d->W % 2 == 0), CTGAN_ERR_BAD_DESC,
                  "conv_fprop_tc: RES_UP2 needs a residual and even H, W");
    CTGAN_REQUIRE(x && wp && y, CTGAN_ERR_BAD_DESC, "conv_fprop_tc: null pointer");
    CTGAN_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(wp) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0,
                  CTGAN_ERR_BAD_DESC, "conv_fprop_tc: pointers must be 16-byte aligned");
    FpropParams p = {};
    p.flags = flags;
    p.y = reinterpret_cast<__nv_bfloat16*>(y);
    p.bias = bias;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
    p.relu_mask = reinterpret_cast<const __nv_bfloat16*>(relu_mask);
    p.out_s2d = (flags & CTGAN_EPI_OUT_S2D) ? 1 : 0;
    p.out_d2s = (flags & CTGAN_EPI_OUT_D2S) ? 1 : 0;
    CTGAN_REQUIRE(!p.out_s2d || (d->H % 2 == 0 && d->W % 2 == 0 && !p.out_d2s), CTGAN_ERR_UNSUPPORTED,
                  "conv_fprop_tc: OUT_S2D needs even H, W (and excludes OUT_D2S)");
    CTGAN_REQUIRE(!p.out_d2s || (d->Cout % 512 == 0 && !residual), CTGAN_ERR_UNSUPPORTED,
                  "conv_fprop_tc: OUT_D2S needs Cout = 4C with C %% 128 == 0 and no residual");
    if (flags & CTGAN_EPI_S2D_SKIP) {
        p.skip_mode = (flags >> 9) & 1; p.skip_k = (flags >> 10) & 15; p.skip_pad_t = (flags >> 14) & 3; p.skip_pad_l = (flags >> 16) & 3;
        CTGAN_REQUIRE(d->kh == 3 && d->kw == 3 && p.skip_k >= 3 && p.skip_k <= 7 &&
                      (p.skip_mode ? d->Cout % 512 == 0 : d->Cin % 256 == 0), CTGAN_ERR_UNSUPPORTED,
                      "conv_fprop_tc: S2D_SKIP needs a 3x3 filter over 4C channels, C %% 64 == 0 (input phases) / C %% 128 == 0 (output phases)");
    }
    return fprop_tc_launch(d, x, wp, p, relu_mask ? EPI_MASK : EPI_PLAIN, stream);
}

extern "C" int ctgan_conv_fprop_tc_actdrop(const ctgan_conv_desc* d, const void* x, const void* wp, const float* bias, void* y,
                                           void* mult, float slope, float keep, uint64_t seed, uint64_t offset,
                                           const uint64_t* dyn_offset, int out_s2d, void* stream) {
    if (int r = check_tc_desc(d, "conv_fprop_tc_actdrop")) return r;
    CTGAN_REQUIRE(x && wp && y && mult, CTGAN_ERR_BAD_DESC, "conv_fprop_tc_actdrop: null pointer");
    CTGAN_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wp) | reinterpret_cast<uintptr_t>(y) |
                    reinterpret_cast<uintptr_t>(mult)) & 15) == 0, CTGAN_ERR_BAD_DESC, "conv_fprop_tc_actdrop: pointers must be 16-byte aligned");
    CTGAN_REQUIRE(keep > 0.f && keep <= 1.f && (offset & 3) == 0, CTGAN_ERR_BAD_DESC, "conv_fprop_tc_actdrop: keep in (0,1], offset a multiple of 4");
    CTGAN_REQUIRE(!out_s2d || (d->H % 2 == 0 && d->W % 2 == 0), CTGAN_ERR_UNSUPPORTED, "conv_fprop_tc_actdrop: space-to-depth output needs even H, W");
    FpropParams p = {};
    p.y = reinterpret_cast<__nv_bfloat16*>(y);
    p.bias = bias;
    p.mult = reinterpret_cast<__nv_bfloat16*>(mult);
    p.slope = slope; p.keep = keep; p.seed = seed; p.offset = offset;
    p.dyn = reinterpret_cast<const unsigned long long*>(dyn_offset);
    p.out_s2d = out_s2d ? 1 : 0;
    return fprop_tc_launch(d, x, wp, p, EPI_ACTDROP, stream);
}

extern "C" int ctgan_conv_wgrad_tc(const ctgan_conv_desc* d, const void* x, const void* dy, float* dw, void* stream) {
    if (int r = check_tc_desc(d, "conv_wgrad_tc")) return r;
    CTGAN_REQUIRE(d->Cin % 128 == 0 && d->Cout % 128 == 0, CTGAN_ERR_UNSUPPORTED, "conv_wgrad_tc: Cin and Cout must be multiples of 128");
    CTGAN_REQUIRE(x && dy && dw, CTGAN_ERR_BAD_DESC, "conv_wgrad_tc: null pointer");
    CTGAN_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(dw) & 15) == 0, CTGAN_ERR_BAD_DESC, "conv_wgrad_tc: pointers must be 16-byte aligned");
    constexpr int STAGES = 4;
    WgradParams p;
    p.N = d->N; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout;
    p.kh = d->kh; p.kw = d->kw; p.pad_t = d->pad_t; p.pad_l = d->pad_l;
    pixel_box(d->H, d->W, 64, &p.BW, &p.BH, &p.BN);
    p.chunksW = ceil_div(d->W, p.BW); p.chunksH = ceil_div(d->H, p.BH); p.chunksN = ceil_div(d->N, p.BN);
    const int taps = d->kh * d->kw;
    const bool lean = g_wgrad_variant == 2 && d->kh == 3 && p.BN == 1 && p.BW % 8 == 0 &&
                      (uint32_t)(p.BH + 2) * p.BW * 128u <= 16384u;
    p.taps_per_cta = lean ? 3 : ((taps % 3 == 0) ? 3 : (taps < 4 ? taps : 4));
    const int tap_groups = lean ? d->kw : ceil_div(taps, p.taps_per_cta);
    const int tiles = (d->Cin / 128) * (d->Cout / 128);
    const int total_chunks = p.chunksN * p.chunksH * p.chunksW;
    // about one wave of CTAs (TMEM: 512 columns => one CTA per SM), at least 4 pixel chunks each: the small
    // layers are bound by the latency of the per-CTA K loop, not by the 49 K reductions each CTA adds
    int splits = sm_count() / (tiles * tap_groups);
    if (splits < 1) splits = 1;
    int max_splits = total_chunks / 4; if (max_splits < 1) max_splits = 1;     // (2..16 chunks per CTA measure the same)
    if (splits > max_splits) splits = max_splits;
    p.chunks_per_split = ceil_div(total_chunks, splits);
    splits = ceil_div(total_chunks, p.chunks_per_split);
    p.dw = dw;
    CUtensorMap mx, mdy;
    if (int r = make_act_map(&mx, x, d->N, d->H, d->W, d->Cin, p.BW, lean ? p.BH + 2 : p.BH, p.BN)) return r;
    if (int r = make_act_map(&mdy, dy, d->N, d->H, d->W, d->Cout, p.BW, p.BH, p.BN)) return r;
    if (lean) {
        constexpr size_t smem_lean = (size_t)STAGES * 49152 + 1024 + (2 * STAGES + 1) * 8 + 16;
        static bool lean_attr_set = false;
        if (!lean_attr_set) {
            cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_lean_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_lean);
            if (e != cudaSuccess) return cuda_status(e, "wgrad_tc_lean smem attribute");
            lean_attr_set = true;
        }
        CTGAN_LAUNCH((conv_wgrad_tc_lean_kernel<STAGES>), dim3(tiles, tap_groups, splits), 128, smem_lean, as_stream(stream), mx, mdy, p);
        CTGAN_CHECK_LAUNCH("conv_wgrad_tc_lean");
        return 0;
    }
    constexpr size_t smem = (size_t)STAGES * 32768 + 1024 + (2 * STAGES + 1) * 8 + 16;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_status(e, "wgrad_tc smem attribute");
        attr_set = true;
    }
    dim3 grid(tiles, tap_groups, splits);
    CTGAN_LAUNCH((conv_wgrad_tc_kernel<STAGES>), grid, 128, smem, as_stream(stream), mx, mdy, p);
    CTGAN_CHECK_LAUNCH("conv_wgrad_tc");
    return 0;
}

extern "C" int ctgan_pack_filters_multi(const float* flat, void* packs, const void* table, int n_entries, void* stream) {
    CTGAN_REQUIRE(flat && packs && table && n_entries > 0 && n_entries <= 65535, CTGAN_ERR_BAD_DESC, "pack_filters_multi: bad args");
    CTGAN_LAUNCH((pack_filters_multi_kernel), dim3(64, n_entries), 256, 0, as_stream(stream), flat, reinterpret_cast<__nv_bfloat16*>(packs),
                                                                                reinterpret_cast<const PackEntry*>(table));
    CTGAN_CHECK_LAUNCH("pack_filters_multi");
    return 0;
}

extern "C" int ctgan_pack_filter_bf16(const float* w, void* wp, int taps, int Cin, int Cout, int transpose_flip, void* stream) {
    CTGAN_REQUIRE(w && wp && taps > 0 && Cin > 0 && Cout > 0, CTGAN_ERR_BAD_DESC, "pack_filter_bf16: bad args");
    int64_t total = (int64_t)taps * Cin * Cout;
    CTGAN_LAUNCH((pack_filter_kernel), elementwise_grid(total, 256), 256, 0, as_stream(stream), w, reinterpret_cast<__nv_bfloat16*>(wp), taps, Cin, Cout, transpose_flip);
    CTGAN_CHECK_LAUNCH("pack_filter_bf16");
    return 0;
}

// ------------------------------------------------------------------ thin-channel entry points
static int check_thin(const ctgan_conv_desc* d, int C, const char* who) {
    CTGAN_REQUIRE(d != nullptr, CTGAN_ERR_BAD_DESC, "%s: null descriptor", who);
    CTGAN_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->kh > 0 && d->kw > 0 && d->stride == 1 && d->Ho == d->H && d->Wo == d->W &&
                  d->pad_t > -8 && d->pad_t < 8 && d->pad_l >= 0 && d->pad_l < d->kw, CTGAN_ERR_BAD_DESC, "%s: bad geometry", who);
    // (pad_t outside [0, kh): a group of filter rows of a larger filter, see kernels._thin_split; offsets are stored in bytes +-8)
    CTGAN_REQUIRE(C > 0 && C <= 8 && d->kh * d->kw * C <= 64, CTGAN_ERR_UNSUPPORTED, "%s: needs C <= 8 and taps*C <= 64", who);
    return 0;
}

extern "C" int ctgan_im2col_thin(const ctgan_conv_desc* d, int C, int sign, const void* src, void* col, void* stream) {
    if (int r = check_thin(d, C, "im2col_thin")) return r;
    CTGAN_REQUIRE(src && col && (sign == 1 || sign == -1) && (reinterpret_cast<uintptr_t>(col) & 15) == 0, CTGAN_ERR_BAD_DESC, "im2col_thin: bad args");
    const int64_t total = (int64_t)d->N * d->H * d->W * 8;
    CTGAN_LAUNCH((im2col_thin_kernel), elementwise_grid(total, 256), 256, 0, as_stream(stream), 
        reinterpret_cast<const __nv_bfloat16*>(src), reinterpret_cast<__nv_bfloat16*>(col), d->N, d->H, d->W, C, d->kh, d->kw,
        d->pad_t, d->pad_l, sign);
    CTGAN_CHECK_LAUNCH("im2col_thin");
    return 0;
}

extern "C" int ctgan_col2im_thin(const ctgan_conv_desc* d, int C, int sign, const void* col, const float* bias, void* dst, void* stream) {
    if (int r = check_thin(d, C, "col2im_thin")) return r;
    CTGAN_REQUIRE(col && dst && (sign == 1 || sign == -1), CTGAN_ERR_BAD_DESC, "col2im_thin: bad args");
    const int64_t total = (int64_t)d->N * d->H * d->W;
    CTGAN_LAUNCH((col2im_thin_kernel), elementwise_grid(total, 128), 128, 0, as_stream(stream), 
        reinterpret_cast<const __nv_bfloat16*>(col), bias, reinterpret_cast<__nv_bfloat16*>(dst), d->N, d->H, d->W, C, d->kh, d->kw,
        d->pad_t, d->pad_l, sign);
    CTGAN_CHECK_LAUNCH("col2im_thin");
    return 0;
}

extern "C" int ctgan_pack_filter_thin(const float* w, void* wp, int taps, int C, int Cw, int kind, void* stream) {
    CTGAN_REQUIRE(w && wp && taps > 0 && C > 0 && taps * C <= 64 && Cw > 0 && kind >= 0 && kind <= 3, CTGAN_ERR_BAD_DESC, "pack_filter_thin: bad args");
    CTGAN_LAUNCH((pack_thin_kernel), elementwise_grid((int64_t)64 * Cw, 256), 256, 0, as_stream(stream), w, reinterpret_cast<__nv_bfloat16*>(wp), taps, C, Cw, kind);
    CTGAN_CHECK_LAUNCH("pack_filter_thin");
    return 0;
}

extern "C" int ctgan_wgrad_thin_tc(const void* wide, const void* col, long long P, int Cw, int C, int taps, int mode,
                                   float* dw, void* stream) {
    CTGAN_REQUIRE(wide && col && dw && P > 0 && P < (1ll << 31) && Cw > 0 && Cw % 64 == 0 && C > 0 && taps > 0 && taps * C <= 64 &&
                  (mode == 0 || mode == 1), CTGAN_ERR_BAD_DESC, "wgrad_thin_tc: bad args");
    CTGAN_REQUIRE((reinterpret_cast<uintptr_t>(wide) & 15) == 0 && (reinterpret_cast<uintptr_t>(col) & 15) == 0, CTGAN_ERR_BAD_DESC,
                  "wgrad_thin_tc: pointers must be 16-byte aligned");
    CTGAN_REQUIRE(ctgan_tc_available(), CTGAN_ERR_UNSUPPORTED, "wgrad_thin_tc: device is not sm_100");
    constexpr int STAGES = 4;
    WgradThinParams p;
    p.total_chunks = (P + 63) / 64;
    p.C = C; p.kreal = taps * C; p.dw = dw; p.Cw = Cw;
    if (mode == 0) { p.sA = (long long)C * Cw; p.sB = Cw; p.sC = 1; }          // dw [taps][C][Cw]
    else           { p.sA = (long long)Cw * C; p.sB = 1;  p.sC = C; }          // dw [taps][Cw][C]
    const int tiles = (Cw + 127) / 128;
    long long splits = sm_count() / tiles; if (splits < 1) splits = 1;
    long long max_splits = p.total_chunks / 8; if (max_splits < 1) max_splits = 1;
    if (splits > max_splits) splits = max_splits;
    p.chunks_per_split = (int)((p.total_chunks + splits - 1) / splits);
    splits = (p.total_chunks + p.chunks_per_split - 1) / p.chunks_per_split;
    CUtensorMap mw, mc;
    if (int r = make_act_map(&mw, wide, 1, 1, (int)P, Cw, 64, 1, 1)) return r;
    if (int r = make_act_map(&mc, col, 1, 1, (int)P, 64, 64, 1, 1)) return r;
    constexpr size_t smem = (size_t)STAGES * 49152 + 1024 + (2 * STAGES + 1) * 8 + 16;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_thin_tc_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_status(e, "wgrad_thin_tc smem attribute");
        attr_set = true;
    }
    CTGAN_LAUNCH((wgrad_thin_tc_kernel<STAGES>), dim3(tiles, 1, (unsigned)splits), 128, smem, as_stream(stream), mw, mc, p);
    CTGAN_CHECK_LAUNCH("wgrad_thin_tc");
    return 0;
}
