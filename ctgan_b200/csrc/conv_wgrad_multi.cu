// conv_wgrad_multi.cu -- the filter gradients of MANY layers in ONE persistent tcgen05 launch.
//
// At batch 64 a filter gradient (Conv2DBackpropFilter, autodiff of TG/tflib/ops/conv2d.py:106) is a short GEMM: K = pixels
// is 4 096 .. 196 608, the output 9 x 128 x 128 floats.  Launched one layer at a time (conv_wgrad_tc_lean_kernel) every
// launch has to split its pixel range ~49 ways to fill 148 SMs, so each CTA runs a few hundred cycles of MMAs and then adds a
// full 196 KB accumulator into the same 590 KB of dW with red.global -- the round-1 timeline shows 14 such launches per
// ResNet critic step at 25-50 us each (ideal 1-13 us), serialised behind each other because every CTA owns its SM's whole
// TMEM: ~45 % of the step's span.  Nothing reads a filter gradient before the optimizer step, so the launches are deferred
// (kernels.defer_wgrad) and executed HERE as one work list:
//   job   = one layer's (x, dY, dW) with its own geometry and tensor maps (a table in the kernel's parameter space),
//   item  = (job, 128x128 accumulator tile, filter column s, pixel-chunk range); every CTA walks items blockIdx.x,
//           blockIdx.x + gridDim.x, ... -- ~2-3 items per SM over ALL layers, so an item is long (>= ~50 chunks) whatever
//           the layer's size, the red.global volume drops ~5x and barrier / TMEM / tensor-map set-up is paid once per step.
// Per item the pipeline is conv_wgrad_tc_lean_kernel's: one CTA owns the kh row taps (r, s) of filter column s, which
// read ONE (BH+kh-1)-row halo box of x (shifted by r rows through the descriptor start address) and one dY chunk, both as
// MN-major operands (no transposed copy of an activation exists); kh x 128 TMEM columns; 4-stage TMA ring of 48 KB.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (+ TMEM allocator), 2..5 = epilogue (TMEM -> red.global.add.v4.f32).
#include <algorithm>
#include <utility>
#include "tc_common.cuh"

namespace ctgan {
namespace tc {

constexpr int WG_MAX_JOBS = 24;

struct alignas(64) WgradJob {
    CUtensorMap mx, mdy;            // x: box (64 ch, BW, BH + kh - 1, BN);  dY: box (64 ch, BW, BH, BN)
    float* dw;                      // [kh][kw][Cin][Cout] float, accumulated into
    int Cin, Cout, kh, kw, pad_t, pad_l;
    int BW, BH, BN, chunksW, chunksH;
    int co_blocks;                  // Cout / 128
    int hn;                         // 1: chunks of several small images, boxes laid out [h][n][w] (tensor maps with N and H exchanged)
    // space-to-depth embedding (emb_k > 0): the job is the 3x3 correlation over the 4C-channel image that stands for a
    // stride-2 emb_k x emb_k filter (conv_s2d.cu); dw is THAT filter's gradient [emb_k][emb_k][emb_C][Cout], and tap (R, S),
    // channel block (dy, dx) of the embedded filter is tap (2R + dy - 2 + pad_t, 2S + dx - 2 + pad_l) of it -- or nothing
    // (11 of the 36 combinations for 5x5), which is then neither multiplied nor stored
    int emb_k, emb_C, emb_pad_t, emb_pad_l;
    int splits, chunks_per_split, total_chunks;
    int item0;                      // index of this job's first work item
    uint32_t a_bytes;               // bytes of one 64-channel half of the x box
};
constexpr int WG_MAX_SLOTS = 1024;
struct WgradJobTable {
    WgradJob job[WG_MAX_JOBS];
    int n_jobs, n_items;
    // Work items differ in size (a job's chunk range is cut into `splits` nearly equal parts, but jobs differ), and a CTA runs
    // only ~2 of them: round-robin assignment left the SMs busy 68 % of the launch (ncu: tensor pipe 67 % of active, 46 % of
    // elapsed cycles).  The host assigns items longest-first to the least-loaded CTA; CTA b runs order[b], order[b + grid],
    // ... until a negative entry.  n_slots == 0: no table (more than WG_MAX_SLOTS slots), item = slot.
    int n_slots;
    short order[WG_MAX_SLOTS];
};

struct WgItem { int j, ci0, co0, s_tap, chunk0, nchunks, rmask, s5, dyv, c0; };

__device__ __forceinline__ WgItem wg_decode(const WgradJobTable& tab, int item) {
    int j = 0;
#pragma unroll 1
    while (j + 1 < tab.n_jobs && item >= tab.job[j + 1].item0) ++j;
    const WgradJob& J = tab.job[j];
    int local = item - J.item0;
    const int z = local % J.splits; local /= J.splits;
    WgItem it;
    it.j = j;
    it.chunk0 = z * J.chunks_per_split;
    it.nchunks = min(J.chunks_per_split, J.total_chunks - it.chunk0);
    if (J.emb_k > 0) {
        // items enumerate the REAL filter's columns s5 and the row parity dy: every item has 2-3 live row taps
        const int s5 = local % J.emb_k; local /= J.emb_k;
        const int dyv = local & 1; const int tile = local >> 1;
        const int as = s5 - J.emb_pad_l + 2;
        it.s_tap = as >> 1;
        it.c0 = (tile / J.co_blocks) * 128; it.co0 = (tile % J.co_blocks) * 128;
        it.ci0 = (dyv * 2 + (as & 1)) * J.emb_C + it.c0;
        it.s5 = s5; it.dyv = dyv;
        it.rmask = 0;
#pragma unroll
        for (int R = 0; R < 3; ++R) {
            const int r5 = 2 * R + dyv - 2 + J.emb_pad_t;
            if (r5 >= 0 && r5 < J.emb_k) it.rmask |= 1 << R;
        }
        return it;
    }
    const int s = local % J.kw; const int tile = local / J.kw;
    it.ci0 = (tile / J.co_blocks) * 128; it.co0 = (tile % J.co_blocks) * 128;
    it.s_tap = s;
    it.rmask = (1 << J.kh) - 1; it.s5 = 0; it.dyv = 0; it.c0 = it.ci0;
    return it;
}

// CHUNK = pixels per pipeline stage (the K extent of one stage's MMAs).  <4 stages, 16 KB x-halves, 64 pixels>: the original
// geometry; <3, 24 KB, 64> when a 64-pixel job's halo box is larger (64-pixel-wide images).  <2, 32 KB, 128> (default,
// ctgan_set_wgrad_multi_chunk): 128-pixel chunks -- the halo rows are shared by twice as many image rows (32x32 image:
// 6 rows loaded per 4 used instead of 4 per 2; 64-pixel-wide: 4 per 2 instead of 3 per 1), so a stage moves 80 KB per 24 MMAs
// instead of 48 KB per 12: the kernel is bound by its L2 -> shared-memory fill.
template <int STAGES, uint32_t A_SLOT = 16384, int CHUNK = 64>
__global__ void __launch_bounds__(192, 1)
conv_wgrad_tc_multi_kernel(const __grid_constant__ WgradJobTable tab)
{
    ctgan::pdl_launch_dependents();
    constexpr uint32_t B_HALF = CHUNK * 128;                              // [CHUNK px][64 ch] bf16
    constexpr uint32_t STAGE_BYTES = 2 * A_SLOT + 2 * B_HALF;             // 48 KB (64 KB; 96 KB)
    constexpr int TMEM_COLS = 512;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform
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
    ctgan::pdl_wait();          // nothing above touches global memory written by earlier kernels

    if (warp == 0) {
        if (lane == 0) {
            // ================= TMA producer =================
            int st = 0; uint32_t ph = 0;
            for (int slot = blockIdx.x; slot < (tab.n_slots ? tab.n_slots : tab.n_items); slot += gridDim.x) {
                const int item = tab.n_slots ? (int)tab.order[slot] : slot;
                if (item < 0) break;
                const WgItem it = wg_decode(tab, item);
                const WgradJob& J = tab.job[it.j];
                const CUtensorMap* mx = &J.mx; const CUtensorMap* mdy = &J.mdy;
                const int dx = it.s_tap - J.pad_l;
                for (int c = 0; c < it.nchunks; ++c) {
                    int ch = it.chunk0 + c;
                    const int cw = ch % J.chunksW; ch /= J.chunksW;
                    const int chh = ch % J.chunksH; const int cn = ch / J.chunksH;
                    const int w0 = cw * J.BW, h0 = chh * J.BH, n0 = cn * J.BN;
                    const uint32_t sb = s_base + st * STAGE_BYTES, fb = full0 + 8 * st;
                    mbar_wait(empty0 + 8 * st, ph ^ 1);
                    mbar_expect_tx(fb, 2 * J.a_bytes + 2 * B_HALF);
                    if (J.hn) {                            // dims (C, W, N, H)
                        tma_load_4d(sb,                       mx,  fb, it.ci0,      w0 + dx, n0, h0 - J.pad_t);
                        tma_load_4d(sb + A_SLOT,              mx,  fb, it.ci0 + 64, w0 + dx, n0, h0 - J.pad_t);
                        tma_load_4d(sb + 2 * A_SLOT,          mdy, fb, it.co0,      w0, n0, h0);
                        tma_load_4d(sb + 2 * A_SLOT + B_HALF, mdy, fb, it.co0 + 64, w0, n0, h0);
                    } else {
                        tma_load_4d(sb,                       mx,  fb, it.ci0,      w0 + dx, h0 - J.pad_t, n0);
                        tma_load_4d(sb + A_SLOT,              mx,  fb, it.ci0 + 64, w0 + dx, h0 - J.pad_t, n0);
                        tma_load_4d(sb + 2 * A_SLOT,          mdy, fb, it.co0,      w0, h0, n0);
                        tma_load_4d(sb + 2 * A_SLOT + B_HALF, mdy, fb, it.co0 + 64, w0, h0, n0);
                    }
                    if (++st == STAGES) { st = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (whole warp runs the loop; one elected lane issues) =================
        // both operands MN-major: smem rows are K (pixels); LBO = distance between the two 64-channel halves
        constexpr uint32_t idesc = make_idesc(BLOCK_M, 128, 1, 1);
        const uint32_t a_lo0 = ((s_base & 0x3FFFFu) >> 4) | ((A_SLOT >> 4) << 16);
        const uint32_t b_lo0 = (((s_base + 2 * A_SLOT) & 0x3FFFFu) >> 4) | ((B_HALF >> 4) << 16);
        int st = 0; uint32_t ph = 0;
        uint32_t n = 0;
        for (int slot = blockIdx.x; slot < (tab.n_slots ? tab.n_slots : tab.n_items); slot += gridDim.x, ++n) {
            const int item = tab.n_slots ? (int)tab.order[slot] : slot;
            if (item < 0) break;
            const WgItem it = wg_decode(tab, item);
            const WgradJob& J = tab.job[it.j];
            const int rmask = it.rmask;
            const uint32_t row_step = ((uint32_t)J.BW * (J.hn ? J.BN : 1) * 128u) >> 4;   // one image row down = filter row r + 1
            mbar_wait(tempty, (n & 1u) ^ 1u);                            // the epilogue has drained the accumulators
            tc_fence_after();
            for (int c = 0; c < it.nchunks; ++c) {
                mbar_wait(full0 + 8 * st, ph);
                tc_fence_after();
                const uint32_t a_lo = a_lo0 + st * (STAGE_BYTES >> 4), b_lo = b_lo0 + st * (STAGE_BYTES >> 4);
                if (elect_one()) {
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        if ((rmask >> r) & 1) {
#pragma unroll
                            for (int k = 0; k < CHUNK / UMMA_K; ++k)  // 16 pixel rows = 2048 bytes along K
                                umma_bf16_lo(tmem_base + (uint32_t)(r * 128), a_lo + r * row_step + 128 * k, b_lo + 128 * k, idesc,
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
        // ================= epilogue warps: TMEM -> vector reductions into dW (float, HWIO) =================
        const int q = warp & 3;                                          // TMEM lane quadrant this warp may access
        uint32_t n = 0;
        for (int slot = blockIdx.x; slot < (tab.n_slots ? tab.n_slots : tab.n_items); slot += gridDim.x, ++n) {
            const int item = tab.n_slots ? (int)tab.order[slot] : slot;
            if (item < 0) break;
            const WgItem it = wg_decode(tab, item);
            const WgradJob& J = tab.job[it.j];
            mbar_wait(tfull, n & 1u);
            tc_fence_after();
            const int ci = it.c0 + q * 32 + lane;
            // Cin / Cout = 64 (mod 128): the upper 64-channel operand half is out of bounds (TMA zero fill); those rows /
            // columns of the accumulator are not stored
            const bool row_ok = ci < (J.emb_k > 0 ? J.emb_C : J.Cin);
            for (int r = 0; r < 3; ++r) {
                if (!((it.rmask >> r) & 1)) continue;
                float* dst = J.emb_k > 0
                    ? J.dw + ((int64_t)((2 * r + it.dyv - 2 + J.emb_pad_t) * J.emb_k + it.s5) * J.emb_C + ci) * J.Cout + it.co0
                    : J.dw + ((int64_t)(r * J.kw + it.s_tap) * J.Cin + ci) * J.Cout + it.co0;
#pragma unroll 1
                for (int c0 = 0; c0 < 128; c0 += 32) {
                    uint32_t acc[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(r * 128 + c0), acc);
                    if (!row_ok || it.co0 + c0 >= J.Cout) continue;
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

}  // namespace tc
}  // namespace ctgan

using namespace ctgan;
using namespace ctgan::tc;

// Longest-processing-time-first assignment of `items` work items (cost[i] > 0) to `grid` CTAs: order[b + grid * k] = the k-th
// item CTA b runs, -1 = none.  Returns the number of slots (a multiple of grid), or 0 -- keep round robin (item = slot) -- when
// the table would not fit WG_MAX_SLOTS or the predicted makespan is not at least min_gain_pct % shorter than round robin's.
// rotate: CTA b starts with its (b mod count)-th item.  *gain_pct receives the predicted gain.  Host code, no CUDA calls.
static int wg_assign_items(const int* cost, int items, int grid, int min_gain_pct, bool rotate, short* order, int* gain_pct) {
    if (gain_pct) *gain_pct = 0;
    if (items <= grid || items > WG_MAX_SLOTS || grid <= 0) return 0;
    static thread_local int idx[WG_MAX_SLOTS], cnt[WG_MAX_SLOTS], pos[WG_MAX_SLOTS];
    static thread_local long long load[WG_MAX_SLOTS];
    static thread_local short assign[WG_MAX_SLOTS];                   // CTA of the i-th sorted item
    static thread_local std::pair<long long, int> heap[WG_MAX_SLOTS]; // min-heap of (load, CTA)
    for (int i = 0; i < items; ++i) idx[i] = i;
    std::stable_sort(idx, idx + items, [&](int a, int b) { return cost[a] > cost[b]; });
    for (int b = 0; b < grid; ++b) { load[b] = 0; cnt[b] = 0; pos[b] = 0; heap[b] = std::make_pair(0LL, b); }
    const auto gt = [](const std::pair<long long, int>& x, const std::pair<long long, int>& y) { return x > y; };
    int max_cnt = 0;
    for (int i = 0; i < items; ++i) {
        std::pop_heap(heap, heap + grid, gt);                          // the least-loaded CTA moves to the back
        std::pair<long long, int>& top = heap[grid - 1];
        const int best = top.second;
        assign[i] = (short)best; top.first += cost[idx[i]]; load[best] = top.first; ++cnt[best];
        if (cnt[best] > max_cnt) max_cnt = cnt[best];
        std::push_heap(heap, heap + grid, gt);
    }
    // use the table only where round robin is clearly unbalanced under the same cost model (the generator step's 8 jobs:
    // -31 us; the critic step's 16 jobs mix well by themselves and the launch is bound by its shared-memory fill)
    long long lpt_max = 0, rr_max = 0;
    for (int b = 0; b < grid; ++b) {
        if (load[b] > lpt_max) lpt_max = load[b];
        long long rr = 0;
        for (int i = b; i < items; i += grid) rr += cost[i];
        if (rr > rr_max) rr_max = rr;
    }
    if (gain_pct) *gain_pct = rr_max > 0 ? (int)(100 - 100 * lpt_max / rr_max) : 0;
    if ((long long)max_cnt * grid > WG_MAX_SLOTS || 100 * lpt_max >= (long long)(100 - min_gain_pct) * rr_max) return 0;
    const int n_slots = max_cnt * grid;
    for (int i = 0; i < n_slots; ++i) order[i] = -1;
    for (int i = 0; i < items; ++i) {
        const int b = assign[i];
        // rotate (ctgan_set_wgrad_multi_balance(3, .)): equal-cost items of ONE job then do not finish on every SM at the
        // same moment -- measured equal (critic graph 919-920 vs 917-918 us)
        const int at = rotate ? (pos[b] + b) % cnt[b] : pos[b];
        order[b + grid * at] = (short)idx[i];
        ++pos[b];
    }
    return n_slots;
}
/* the assignment alone, for tests (host code: callable without a GPU): order must hold 1024 entries; returns the slot count */
extern "C" int ctgan_wgrad_multi_assign(const int* cost, int items, int grid, int min_gain_pct, int rotate, short* order, int* gain_pct) {
    return wg_assign_items(cost, items, grid, min_gain_pct, rotate != 0, order, gain_pct);
}

static int g_wgrad_multi_chunk = 0;
/* tuning / test hook: pixels per pipeline stage of the multi-job filter-gradient kernel: 64, 128, or 0 (default) = 128 for a
 * launch with a job on images at least 64 pixels wide (measured: CT_gan_64x64.py +0.9 %, LSUN 128x128 +1.6 %; the CIFAR ResNet
 * step, 32 pixels wide at most, is 0.5 % faster with 64: two 96 KB stages pipeline less deeply than four of 48 KB) */
extern "C" void ctgan_set_wgrad_multi_chunk(int px) { g_wgrad_multi_chunk = (px == 64 || px == 128) ? px : 0; }
static int g_wgrad_multi_balance = 1;
static int g_wgrad_multi_item_overhead = 5;
static int g_wgrad_multi_rotate = 0;
static int g_wgrad_multi_min_gain_pct = 10;      // predicted makespan gain (%) below which round robin is kept
static int g_wgrad_multi_last_gain_pct = 0;
/* diagnostic: predicted makespan gain (%) of the longest-first assignment for the most recent launch */
extern "C" int ctgan_wgrad_multi_last_gain_pct(void) { return g_wgrad_multi_last_gain_pct; }
/* tuning / A-B hook: on = 1 assigns the work items longest-first to the least-loaded CTA (default), 0 = round robin;
 * overhead = the fixed cost of an item (accumulator drain) in 64-pixel chunks used by that assignment */
extern "C" void ctgan_set_wgrad_multi_balance(int on, int overhead) {
    g_wgrad_multi_balance = on & 1; g_wgrad_multi_rotate = (on & 2) ? 1 : 0;       // on = 3: balanced + per-CTA rotation of the order
    if (overhead >= 0) g_wgrad_multi_item_overhead = overhead;
    g_wgrad_multi_min_gain_pct = (on & 4) ? 0 : 10;                                // on = 5: use the table whenever it exists
}
static int g_wgrad_multi_items_per_sm = 2;
/* tuning hook: work items per SM the deferred filter-gradient launch aims for */
extern "C" void ctgan_set_wgrad_multi_items_per_sm(int v) { g_wgrad_multi_items_per_sm = v < 1 ? 1 : v; }

// can (d) be cut into `chunk`-pixel K chunks whose x halo box fits one operand slot?
static bool wgrad_chunk_ok(const ctgan_conv_desc* d, int chunk) {
    const uint32_t slot = chunk == 128 ? 32768u : 24576u;
    int BW, BH, BN;
    pixel_box(d->H, d->W, chunk, &BW, &BH, &BN);
    if (BW > 256 || BH > 256 || BN > 256) return false;
    if (d->kh == 1) return true;
    if (BN == 1) return BW % 8 == 0 && (uint32_t)(BH + 2) * BW * 128u <= slot;
    // several whole small images per chunk (8x8 / 4x4): [h][n][w] boxes, filter row shift = BN * BW pixel rows
    return BW == d->W && BH == d->H && (BN * BW) % 8 == 0 && (uint32_t)(BH + 2) * BN * BW * 128u <= (chunk == 128 ? 32768u : 16384u);
}

/* 1 when (d) can be a job of ctgan_conv_wgrad_tc_multi */
extern "C" int ctgan_conv_wgrad_tc_multi_ok(const ctgan_conv_desc* d) {
    if (!d || !ctgan_tc_available()) return 0;
    if (d->x_dtype != CTGAN_BF16 || d->y_dtype != CTGAN_BF16 || d->stride != 1 || d->Ho != d->H || d->Wo != d->W) return 0;
    if (d->N <= 0 || d->H <= 0 || d->W <= 0 || d->Cin <= 0 || d->Cout <= 0 || d->Cin % 64 || d->Cout % 64) return 0;
    if (!((d->kh == 3 && d->kw == 3) || (d->kh == 1 && d->kw == 1))) return 0;
    if (d->pad_t < 0 || d->pad_l < 0 || d->pad_t >= d->kh || d->pad_l >= d->kw) return 0;
    return wgrad_chunk_ok(d, 64) || (g_wgrad_multi_chunk != 64 && wgrad_chunk_ok(d, 128));
}

extern "C" int ctgan_conv_wgrad_tc_multi(int n, const ctgan_conv_desc* descs, const void* const* xs, const void* const* dys,
                                         float* const* dws, void* stream) {
    return ctgan_conv_wgrad_tc_multi_embed(n, descs, xs, dys, dws, nullptr, stream);
}

extern "C" int ctgan_conv_wgrad_tc_multi_embed(int n, const ctgan_conv_desc* descs, const void* const* xs, const void* const* dys,
                                               float* const* dws, const int* embed, void* stream) {
    CTGAN_REQUIRE(n > 0 && descs && xs && dys && dws, CTGAN_ERR_BAD_DESC, "conv_wgrad_tc_multi: bad args");
    constexpr int STAGES = 4;
    constexpr size_t smem = (size_t)STAGES * 49152 + 1024 + (2 * STAGES + 2) * 8 + 16;
    constexpr size_t smem_big = (size_t)3 * 65536 + 1024 + (2 * 3 + 2) * 8 + 16;
    constexpr size_t smem_128 = (size_t)2 * 98304 + 1024 + (2 * 2 + 2) * 8 + 16;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_multi_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(conv_wgrad_tc_multi_kernel<3, 24576>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_big);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(conv_wgrad_tc_multi_kernel<2, 32768, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_128);
        if (e != cudaSuccess) return cuda_status(e, "wgrad_tc_multi smem attribute");
        attr_set = true;
    }
    for (int base = 0; base < n; base += WG_MAX_JOBS) {
        const int nj = (n - base) < WG_MAX_JOBS ? (n - base) : WG_MAX_JOBS;
        static thread_local WgradJobTable tab;
        // MMA groups (4 instructions) per (accumulator tile, filter column) and in total
        long long total = 0;
        long long col_work[WG_MAX_JOBS];
        bool big = false;                      // some x halo box exceeds 16 KB per 64-channel half
        // 128-pixel chunks when every job of this launch can be cut that way, else 64
        int chunk = g_wgrad_multi_chunk;
        if (chunk == 0) {
            chunk = 64;
            for (int i = 0; i < nj; ++i) if (descs[base + i].W >= 64) chunk = 128;
        }
        for (int i = 0; i < nj && chunk == 128; ++i)
            if (!wgrad_chunk_ok(descs + base + i, 128)) chunk = 64;
        for (int i = 0; i < nj; ++i) {
            const ctgan_conv_desc* d = descs + base + i;
            CTGAN_REQUIRE(ctgan_conv_wgrad_tc_multi_ok(d), CTGAN_ERR_UNSUPPORTED, "conv_wgrad_tc_multi: job %d is not eligible", base + i);
            CTGAN_REQUIRE(xs[base + i] && dys[base + i] && dws[base + i] &&
                          ((reinterpret_cast<uintptr_t>(xs[base + i]) | reinterpret_cast<uintptr_t>(dys[base + i]) |
                            reinterpret_cast<uintptr_t>(dws[base + i])) & 15) == 0,
                          CTGAN_ERR_BAD_DESC, "conv_wgrad_tc_multi: job %d: null or misaligned pointer", base + i);
            WgradJob& J = tab.job[i];
            J.dw = dws[base + i];
            J.Cin = d->Cin; J.Cout = d->Cout; J.kh = d->kh; J.kw = d->kw; J.pad_t = d->pad_t; J.pad_l = d->pad_l;
            CTGAN_REQUIRE(wgrad_chunk_ok(d, chunk), CTGAN_ERR_UNSUPPORTED, "conv_wgrad_tc_multi: job %d does not fit %d-pixel chunks", base + i, chunk);
            pixel_box(d->H, d->W, chunk, &J.BW, &J.BH, &J.BN);
            J.chunksW = ceil_div(d->W, J.BW); J.chunksH = ceil_div(d->H, J.BH);
            J.total_chunks = J.chunksW * J.chunksH * ceil_div(d->N, J.BN);
            J.co_blocks = ceil_div(d->Cout, 128);
            J.a_bytes = (uint32_t)(J.BH + d->kh - 1) * J.BW * J.BN * 128u;
            big = big || J.a_bytes > 16384u;
            J.emb_k = J.emb_C = J.emb_pad_t = J.emb_pad_l = 0;
            if (embed && embed[4 * (base + i)] > 0) {
                const int* e = embed + 4 * (base + i);
                CTGAN_REQUIRE(d->kh == 3 && d->kw == 3 && d->pad_t == 1 && d->pad_l == 1 && e[0] <= 5 && e[1] > 0 && e[1] % 128 == 0 &&
                              4 * e[1] == d->Cin && e[2] >= 0 && e[2] <= 2 && e[3] >= 0 && e[3] <= 2,
                              CTGAN_ERR_UNSUPPORTED, "conv_wgrad_tc_multi: job %d: bad space-to-depth embedding", base + i);
                J.emb_k = e[0]; J.emb_C = e[1]; J.emb_pad_t = e[2]; J.emb_pad_l = e[3];
            }
            J.hn = (d->kh == 3 && J.BN > 1) ? 1 : 0;
            if (J.hn) {
                if (int r = make_act_map_hn(&J.mx, xs[base + i], d->N, d->H, d->W, d->Cin, J.BW, J.BH + d->kh - 1, J.BN)) return r;
                if (int r = make_act_map_hn(&J.mdy, dys[base + i], d->N, d->H, d->W, d->Cout, J.BW, J.BH, J.BN)) return r;
            } else {
                if (int r = make_act_map(&J.mx, xs[base + i], d->N, d->H, d->W, d->Cin, J.BW, J.BH + d->kh - 1, J.BN)) return r;
                if (int r = make_act_map(&J.mdy, dys[base + i], d->N, d->H, d->W, d->Cout, J.BW, J.BH, J.BN)) return r;
            }
            col_work[i] = (long long)J.total_chunks * d->kh * (chunk / 64);
            total += J.emb_k > 0 ? col_work[i] * (J.emb_C / 128) * J.co_blocks * 2 * J.emb_k
                                 : col_work[i] * ceil_div(d->Cin, 128) * J.co_blocks * d->kw;
        }
        // one item ~ total / (items_per_sm * SMs) MMA groups, never below 24 (8 chunks of a 3x3 column)
        long long target = (total + (long long)g_wgrad_multi_items_per_sm * sm_count() - 1) / ((long long)g_wgrad_multi_items_per_sm * sm_count());
        if (target < 24) target = 24;
        int items = 0;
        for (int i = 0; i < nj; ++i) {
            WgradJob& J = tab.job[i];
            long long s = (col_work[i] + target / 2) / target;
            if (s < 1) s = 1;
            if (s > J.total_chunks) s = J.total_chunks;
            J.chunks_per_split = ceil_div(J.total_chunks, s);
            J.splits = ceil_div(J.total_chunks, J.chunks_per_split);
            J.item0 = items;
            items += (J.emb_k > 0 ? (J.emb_C / 128) * 2 * J.emb_k : ceil_div(J.Cin, 128) * J.kw) * J.co_blocks * J.splits;
        }
        tab.n_jobs = nj; tab.n_items = items;
        const int grid = items < sm_count() ? items : sm_count();
        // longest-processing-time-first assignment of the items to the CTAs (see WgradJobTable::order)
        tab.n_slots = 0;
        if (g_wgrad_multi_balance && items > grid && items <= WG_MAX_SLOTS) {
            static thread_local int cost[WG_MAX_SLOTS];
            int k = 0;
            for (int i = 0; i < nj; ++i) {
                const WgradJob& J = tab.job[i];
                const int n_local = (i + 1 < nj ? tab.job[i + 1].item0 : items) - J.item0;
                for (int l = 0; l < n_local; ++l, ++k) {
                    const int z = l % J.splits;
                    const int rest = J.total_chunks - z * J.chunks_per_split;
                    const int nch = rest < J.chunks_per_split ? rest : J.chunks_per_split;
                    cost[k] = nch * (chunk / 64) + g_wgrad_multi_item_overhead;     // in 64-pixel chunks; + the accumulator drain
                }
            }
            tab.n_slots = wg_assign_items(cost, items, grid, g_wgrad_multi_min_gain_pct, g_wgrad_multi_rotate != 0, tab.order,
                                          &g_wgrad_multi_last_gain_pct);
        }
        if (chunk == 128) CTGAN_LAUNCH((conv_wgrad_tc_multi_kernel<2, 32768, 128>), grid, 192, smem_128, as_stream(stream), tab);
        else if (big) CTGAN_LAUNCH((conv_wgrad_tc_multi_kernel<3, 24576>), grid, 192, smem_big, as_stream(stream), tab);
        else     CTGAN_LAUNCH((conv_wgrad_tc_multi_kernel<STAGES>), grid, 192, smem, as_stream(stream), tab);
        CTGAN_CHECK_LAUNCH("conv_wgrad_tc_multi");
    }
    return 0;
}
