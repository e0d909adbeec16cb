// conv_splitk.cu -- stride-1 tcgen05 convolution for layers with FEWER TILES THAN SMs: split-K over a thread-block cluster.
//
// At batch 64 the 8x8 layers of the critic (Discriminator.3 / .4, TG/CT_gan_cifar_resnet.py:183-190) are 32 output tiles of
// 128 pixels x 128 channels on 148 SMs.  One tile needs the WHOLE 3x3x128x128 filter (288 KB) and 288 KB of shifted
// activation boxes, and an SM ingests ~64 B/clk from L2: ~5 us of the 9.4 us the launch takes alone (127 TFLOP/s = 0.076 of
// the peak; the family is 20 % of the step's kernel time, profiles/r02_kernel_shares.json).  The work is there, the SMs are
// idle -- so the K dimension (taps x input channels) of each tile is split over a cluster of SPLIT = 2 / 4 CTAs:
//   * CTA r of the cluster loads and multiplies only k-groups [r*G/SPLIT, (r+1)*G/SPLIT) of the tile (G groups of two
//     64-channel k-blocks): 1/SPLIT of the filter and of the activation boxes per SM;
//   * the partial accumulators (TMEM, fp32) are reduce-SCATTERED through distributed shared memory: CTA c owns output
//     channels [c*128/SPLIT, (c+1)*128/SPLIT); every other CTA writes its partial of those columns straight into CTA c's
//     shared memory (st.shared::cluster), one cluster barrier, then the owner adds the SPLIT-1 partials to its own registers
//     and runs the usual epilogue (bias, residual, ReLU / ReLU-backward mask, bf16 pack) on its columns only.
// No global-memory workspace, no second kernel, no atomics: the result is bitwise deterministic.
// Roles per CTA as in conv_fprop_tc_lean_kernel<0>: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM), warps 2..5 =
// epilogue.  One tile per CTA (the grid is tiles x SPLIT <= 148).
#include "tc_common.cuh"

namespace ctgan {
namespace tc {

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int SPLIT, int EPI>
__global__ void __launch_bounds__(192, 1)
conv_fprop_tc_splitk_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                            const FpropParams p)
{
    ctgan::pdl_launch_dependents();
    constexpr int BLOCK_N = 128;
    constexpr int STAGES = SPLIT == 2 ? 3 : 2;
    constexpr uint32_t A_REGION = 32768u, B_BYTES = BLOCK_N * BLOCK_K * 2;      // two (activation, filter) box pairs per stage
    constexpr uint32_t STAGE_BYTES = A_REGION + 2 * B_BYTES;                    // 64 KB
    constexpr int TMEM_COLS = 128;
    constexpr int CW = 128 / SPLIT;                                             // output channels finished by one CTA
    constexpr int UNITS = CW / 32;                                              // 32-column units per owner
    constexpr uint32_t RED_BYTES = (uint32_t)(SPLIT - 1) * UNITS * 8 * 128 * 16;   // [sender][unit][16-byte piece][row]

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t s_base = smem_u32(smem);
    const uint32_t red_base = s_base + STAGES * STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + RED_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), tfull = smem_u32(bars + 2 * STAGES);
    const uint32_t rank = cluster_ctarank();

    const int cin_blocks = p.Cin / BLOCK_K;
    const int n_blocks = p.Cout / BLOCK_N;
    const int kblocks = cin_blocks * p.kh * p.kw;                     // k-block index = tap * cin_blocks + cb
    const int groups = (kblocks + 1) / 2;
    const int g0 = (int)(((long long)groups * rank) / SPLIT), g1 = (int)(((long long)groups * (rank + 1)) / SPLIT);
    const uint32_t a_bytes = (uint32_t)p.BH * p.BW * p.BN * 128u;

    const int tile = blockIdx.x / SPLIT;
    const int nb = tile % n_blocks; int mt = tile / n_blocks;
    const int tw = mt % p.tilesW; mt /= p.tilesW;
    const int th = mt % p.tilesH; const int tn = mt / p.tilesH;
    const int w0 = tw * p.BW, h0 = th * p.BH, n0 = tn * p.BN, co0 = nb * BLOCK_N;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_w);
        for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    cluster_arrive();           // phase 1: "this CTA is resident" -- awaited before the first write into a peer's shared memory
    ctgan::pdl_wait();          // nothing above touches global memory written by earlier kernels

    if (warp == 0) {
        if (lane == 0) {
            // ================= TMA producer =================
            int st = 0; uint32_t ph = 0;
            for (int gi = g0; gi < g1; ++gi) {
                const uint32_t sb = s_base + st * STAGE_BYTES, fb = full0 + 8 * st;
                mbar_wait(empty0 + 8 * st, ph ^ 1);
                const int kb0 = gi * 2, nk = min(2, kblocks - kb0);
                mbar_expect_tx(fb, (uint32_t)nk * (a_bytes + B_BYTES));
                for (int j = 0; j < nk; ++j) {
                    const int kb = kb0 + j, tap = kb / cin_blocks, cb = kb - tap * cin_blocks;
                    const int r = tap / p.kw, s = tap - r * p.kw;
                    tma_load_4d(sb + j * 16384u, &tmap_x, fb, cb * BLOCK_K, w0 + s - p.pad_l, h0 + r - p.pad_t, n0);
                    tma_load_3d(sb + A_REGION + j * B_BYTES, &tmap_w, fb, cb * BLOCK_K, co0, tap);
                }
                if (++st == STAGES) { st = 0; ph ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================= MMA issuer (whole warp runs the loop; one elected lane issues) =================
        constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, 0, 0);
        const uint32_t lo0 = ((s_base & 0x3FFFFu) >> 4) | (1u << 16);
        int st = 0; uint32_t ph = 0;
        for (int gi = g0; gi < g1; ++gi) {
            mbar_wait(full0 + 8 * st, ph);
            tc_fence_after();
            const uint32_t a_lo = lo0 + st * (STAGE_BYTES >> 4), b_lo = a_lo + (A_REGION >> 4);
            const int nk = min(2, kblocks - gi * 2);
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (j < nk) {
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                            umma_bf16_lo(tmem_base, a_lo + j * (16384u >> 4) + 2 * k, b_lo + j * (B_BYTES >> 4) + 2 * k, idesc,
                                         (j | k) ? 1u : (gi > g0 ? 1u : 0u));
                    }
                }
                umma_commit(empty0 + 8 * st);
                if (gi == g1 - 1) umma_commit(tfull);
            }
            __syncwarp();
            if (++st == STAGES) { st = 0; ph ^= 1; }
        }
    }

    // ================= reduce-scatter of the partial accumulators through distributed shared memory =================
    float own[UNITS][32];
    const int q = warp & 3;                                          // TMEM lane quadrant of an epilogue warp
    const int row = q * 32 + lane;                                   // tile row = pixel
    cluster_wait();                                                  // phase 1 done: every CTA of the cluster is resident
    if (warp >= 2) {
        if (g1 > g0) { mbar_wait(tfull, 0); tc_fence_after(); }
#pragma unroll
        for (int u = 0; u < 4; ++u) {                                // 32-column unit u belongs to CTA u / UNITS
            const uint32_t owner = (uint32_t)(u / UNITS);
            uint32_t v32[32];
            if (g1 > g0) {
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(u * 32), v32);
            } else {
#pragma unroll
                for (int e = 0; e < 32; ++e) v32[e] = 0u;
            }
            if (owner == rank) {
#pragma unroll
                for (int e = 0; e < 32; ++e) own[u % UNITS][e] = __uint_as_float(v32[e]);
            } else {
                const uint32_t slot = rank < owner ? rank : rank - 1;              // index among the owner's SPLIT-1 senders
                const uint32_t dst = map_to_cta(red_base + ((slot * UNITS + (uint32_t)(u % UNITS)) * 8 * 128 + (uint32_t)row) * 16, owner);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    st_cluster_v4(dst + (uint32_t)k * 128 * 16, __uint_as_float(v32[4 * k]), __uint_as_float(v32[4 * k + 1]),
                                  __uint_as_float(v32[4 * k + 2]), __uint_as_float(v32[4 * k + 3]));
            }
        }
        tc_fence_before();
    }
    cluster_arrive();                                                // phase 2: my partials are in the owners' shared memory
    cluster_wait();
    if (warp >= 2) {
        // ---- owner: add the SPLIT-1 received partials, then the usual epilogue on columns [rank*CW, rank*CW + CW)
        const int n = n0 + row / (p.BW * p.BH), h = h0 + (row / p.BW) % p.BH, w = w0 + row % p.BW;
        const bool valid = (n < p.N) && (h < p.H) && (w < p.W);
        const int64_t pix = ((int64_t)n * p.H + h) * p.W + w;
        const int cbase = co0 + (int)rank * CW;
        int64_t orow = pix * p.Cout;                                 // element offset of this pixel's row in y (and m)
        if (EPI == EPI_ACTDROP && p.out_s2d)
            orow = ((((int64_t)n * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1)) * 4 + ((h & 1) * 2 + (w & 1))) * p.Cout;
        __nv_bfloat16* yrow = p.y + orow + cbase;
        __nv_bfloat16* mult_row = (EPI == EPI_ACTDROP) ? p.mult + orow + cbase : nullptr;
        // EPI_ACTDROP (see lean_epilogue_tile): absolute Philox stream index of channel cbase of this pixel
        unsigned long long se0 = 0;
        bool drop = false;
        float inv_keep = 1.f;
        if (EPI == EPI_ACTDROP) {
            drop = p.keep < 1.f;
            inv_keep = 1.f / p.keep;
            se0 = p.offset + (p.dyn ? *p.dyn : 0ull) + (unsigned long long)(pix * p.Cout + cbase);
        }
        const int64_t rpix = (p.flags & CTGAN_EPI_RES_UP2) ? ((int64_t)n * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1) : pix;
        const __nv_bfloat16* rrow = (EPI != EPI_ACTDROP && p.residual) ? p.residual + rpix * p.Cout + cbase : nullptr;
        const __nv_bfloat16* mrow = (EPI == EPI_MASK) ? p.relu_mask + pix * p.Cout + cbase : nullptr;
        const bool relu = (p.flags & CTGAN_EPI_RELU) != 0;
        const bool wide = ((reinterpret_cast<uintptr_t>(p.y) | reinterpret_cast<uintptr_t>(p.residual) |
                            (EPI == EPI_MASK ? reinterpret_cast<uintptr_t>(p.relu_mask) : 0) |
                            (EPI == EPI_ACTDROP ? reinterpret_cast<uintptr_t>(p.mult) : 0)) & 31) == 0;
        const float4* red = reinterpret_cast<const float4*>(smem + STAGES * STAGE_BYTES);
#pragma unroll
        for (int u = 0; u < UNITS; ++u) {
#pragma unroll
            for (int s = 0; s < SPLIT - 1; ++s) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float4 t = red[((s * UNITS + u) * 8 + k) * 128 + row];
                    own[u][4 * k] += t.x; own[u][4 * k + 1] += t.y; own[u][4 * k + 2] += t.z; own[u][4 * k + 3] += t.w;
                }
            }
            if (valid) {
#pragma unroll
                for (int j = 0; j < 32; j += 16) {
                    const int c = u * 32 + j;                        // column inside this CTA's CW-wide slice
                    float v[16];
                    if (p.bias) {
#pragma unroll
                        for (int e = 0; e < 16; e += 4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + cbase + c + e));
                            v[e] = own[u][j + e] + b4.x; v[e + 1] = own[u][j + e + 1] + b4.y;
                            v[e + 2] = own[u][j + e + 2] + b4.z; v[e + 3] = own[u][j + e + 3] + b4.w;
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; ++e) v[e] = own[u][j + e];
                    }
                    if (rrow) {
                        uint32_t rw[8];
                        ldg16_bf16(rrow + c, wide, rw);
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
                        ldg16_bf16(mrow + c, wide, mw);
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
                            if (drop) Philox::block(p.seed, (se0 + (unsigned long long)(c + 4 * b)) >> 2, r);
                            float mm[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float mult = v[4 * b + e] > 0.f ? 1.f : p.slope;
                                if (drop) mult *= floorf(p.keep + Philox::to_uniform(r[e])) * inv_keep;
                                mm[e] = mult;
                            }
                            const __nv_bfloat162 ma = __floats2bfloat162_rn(mm[0], mm[1]), mb = __floats2bfloat162_rn(mm[2], mm[3]);
                            mo[2 * b] = *reinterpret_cast<const uint32_t*>(&ma); mo[2 * b + 1] = *reinterpret_cast<const uint32_t*>(&mb);
                            const __nv_bfloat162 ya = __floats2bfloat162_rn(v[4 * b] * __bfloat162float(ma.x), v[4 * b + 1] * __bfloat162float(ma.y));
                            const __nv_bfloat162 yb = __floats2bfloat162_rn(v[4 * b + 2] * __bfloat162float(mb.x), v[4 * b + 3] * __bfloat162float(mb.y));
                            ow[2 * b] = *reinterpret_cast<const uint32_t*>(&ya); ow[2 * b + 1] = *reinterpret_cast<const uint32_t*>(&yb);
                        }
                        stg16_bf16(mult_row + c, wide, mo);
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const __nv_bfloat162 o2 = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
                            ow[e] = *reinterpret_cast<const uint32_t*>(&o2);
                        }
                    }
                    stg16_bf16(yrow + c, wide, ow);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

template <int SPLIT, int EPI>
static int launch_splitk(const CUtensorMap& mx, const CUtensorMap& mw, const FpropParams& p, int n_tiles, cudaStream_t st) {
    constexpr int STAGES = SPLIT == 2 ? 3 : 2;
    constexpr size_t smem = (size_t)STAGES * 65536 + (size_t)(SPLIT - 1) * (128 / SPLIT / 32) * 8 * 128 * 16 + 1024 + (2 * STAGES + 1) * 8 + 16;
    static_assert(smem <= 227 * 1024, "split-K fprop: shared memory budget");
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_fprop_tc_splitk_kernel<SPLIT, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_status(e, "fprop_tc_splitk smem attribute");
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_tiles * SPLIT)); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = SPLIT; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 2;
    (void)cudaLaunchKernelEx(&cfg, conv_fprop_tc_splitk_kernel<SPLIT, EPI>, mx, mw, p);
    CTGAN_CHECK_LAUNCH("conv_fprop_tc_splitk");
    return 0;
}

// Split factor for a layer of n_tiles output tiles and `groups` k-groups (0 = do not split): the grid must stay within
// one wave and every CTA keeps at least two groups.
int splitk_factor(int n_tiles, int groups) {
    if (n_tiles * 4 <= sm_count() && groups >= 8) return 4;
    if (n_tiles * 2 <= sm_count() && groups >= 4) return 2;
    return 0;
}

int launch_fprop_splitk(int split, int epi, const CUtensorMap& mx, const CUtensorMap& mw, const FpropParams& p, int n_tiles,
                        cudaStream_t st) {
    if (split == 4) {
        if (epi == EPI_ACTDROP) return launch_splitk<4, EPI_ACTDROP>(mx, mw, p, n_tiles, st);
        return epi == EPI_MASK ? launch_splitk<4, EPI_MASK>(mx, mw, p, n_tiles, st) : launch_splitk<4, EPI_PLAIN>(mx, mw, p, n_tiles, st);
    }
    if (epi == EPI_ACTDROP) return launch_splitk<2, EPI_ACTDROP>(mx, mw, p, n_tiles, st);
    return epi == EPI_MASK ? launch_splitk<2, EPI_MASK>(mx, mw, p, n_tiles, st) : launch_splitk<2, EPI_PLAIN>(mx, mw, p, n_tiles, st);
}

}  // namespace tc
}  // namespace ctgan
