// spconv_cg.cu -- sparse 3-D convolution (SubM / strided, + folded BN + ReLU) on the tcgen05 tensor cores whose operand traffic is
// proportional to the number of rulebook PAIRS: only the neighbour rows that exist are fetched.
//
// Replaces spconv 1.x's per-offset gather -> sgemm -> scatter-add used by det3d/models/backbones/scn.py:106-149 (SpMiddleFHD), like
// the lab library's spconv_h2.cu (same numerics: fp16 (hi, lo) planes with an exact power-of-two scale, three kind::f16 products per MAC,
// two main + one cross TMEM accumulator summed in RN fp32), but the 128-row A operand of a (tile, kernel offset) is no longer fetched
// with 32-64 TMA gather4 instructions whose cost is per ROW SLOT, present or not (~5.5 clk per 128-byte row: at the 21 % neighbour fill of
// the 32-channel layers 79 % of the row requests fetched zeros and the layer ran at 0.9 % of the tensor peak).  Here
//   * the rulebook is regrouped ONCE per build (sessd_rulebook_tile_lists; a SubM rulebook serves 2-3 layers) into per-tile, per-offset
//     lists of (input row, tile row) pairs + row masks; a tile's record (~3-8 KB) is copied to shared memory (compacting the neighbour
//     table inside this kernel with shared-memory atomics cost 14.7 k clk per tile, a third of the tile's time);
//   * eight producer warps copy the listed rows with 16-byte cp.async (LDGSTS: global/L2 -> shared, no registers, 2 clk per 128-byte row)
//     straight into the K-major SWIZZLE_128B layout the UMMA descriptors address; rows without a neighbour are never touched;
//   * a stage's missing rows must read as zeros: each producer warp remembers (registers) which of its rows of its stage hold data and
//     clears (st.shared) only the rows that were valid for the stage's previous offset and are not for the new one -- the stages are
//     zeroed once per CTA;
//   * the 32-channel layers stage the weights as [b_hi rows ; b_lo rows] of 64 bytes (SWIZZLE_64B) so that one N = 2 Cout product gives
//     the main and the a_hi x b_lo cross term (4 instead of 6 MMAs per offset: an SS-form MMA costs max(math, operand bytes / 128 B/clk));
//   * CTAs are persistent (two per SM: one tile's epilogue overlaps the other's main loop), so barrier / TMEM / zero-fill setup is paid
//     once, not per 128 rows;
//   * the epilogue writes the NEXT layer's operand format directly -- fp16 (hi, lo) planes scaled by a power of two derived from a
//     rigorous bound |out| <= amax_in * G + max|shift| (G from the weights, host; amax_in measured by the producing layer's epilogue) --
//     and raises the output's abs-max: the separate split kernel (one read + one write of every feature tensor) is gone.
// Each stage has its own producer group (8 / kStages warps): the group copies, waits for ITS copies (cp.async.wait_all), makes them and
// the clears visible to the async proxy (fence.proxy.async by the writing threads) and arrives on the stage's mbarrier, while the other
// groups' fills are in flight -- kStages fills per CTA overlap and nobody hands data over on behalf of another thread.
// Warps: 0-7 producers then epilogue (TMEM lane quadrant = warp & 3, column half = warp >> 2), 8 weight-tile TMA, 9 MMA issue.
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace sessd {

constexpr int kCgBM = 128;
constexpr int kCgMaxK = 27;
constexpr int kCgProdWarps = 8;
constexpr int kCgProdThreads = kCgProdWarps * 32;
constexpr int kCgThreads = kCgProdThreads + 64;

// DEEP = 1: twice the stages, one CTA per SM -- for launches whose tiles do not fill the machine twice over (a single frame: ~100 tiles on 148 SMs), where
// a CTA is alone on its SM anyway and the layer's time is the serial chain of its tile's fills (latency-bound): more fills in flight shorten it
template <int CP, int COUT, int DEEP = 0>
struct CgCfg {
    static constexpr bool kWide = (CP == 64);
    static constexpr int kATile = (kWide ? 2 : 1) * kCgBM * 128;              // bytes: [hi tile ; lo tile] (wide) or one [hi | lo] tile
    static constexpr int kBTile = (kWide ? 2 : 1) * COUT * 128;
    static constexpr int kStage = kATile + (kBTile + 1023) / 1024 * 1024;
    static constexpr int kStages = (kWide ? 2 : 4) * (DEEP ? 2 : 1);           // must divide the 8 producer warps (one group per stage)
    static constexpr int kMeta = kCgBM * kCgMaxK * 4 /*lists*/ + kCgMaxK * 16 /*valid*/ + 32 * 4 /*cnt*/ + 33 * 4 /*klist, nact*/ + 32 * 4 /*off*/ +
                                 (3 * kStages + 1) * 8 /*barriers*/ + 24;
    static constexpr int kSmem = kStages * kStage + kMeta + 1024;
    static constexpr int kTmemCols = (3 * COUT <= 128) ? 128 : 256;
    static constexpr int kCPO = COUT > 32 ? 64 : 32;                          // channels per plane row of the OUTPUT
};

struct CgArgs {
    const __half *planes;              // input [rows][2][CP] fp16, x = (hi + lo) / in_info[1]
    const float *in_info;              // {abs-max of the input tensor, its plane scale}
    const unsigned int *tiles;         // per-tile pair lists (sessd_rulebook_tile_lists), tile_stride words per tile
    int tile_stride;
    const int *d_n_out;
    int kvol, max_out, relu;
    const float *scale, *shift;        // folded BN (scale already times the per-channel weight exponent 2^-e)
    float gain, shift_max;             // |out| <= amax_in * gain + shift_max
    float *out_f32;                    // nullable [max_out][COUT]
    __half *out_planes;                // nullable [max_out (+1)][2][kCPO]
    float *out_info;                   // nullable {abs-max of the output (atomicMax), plane scale}
    long long *dbg;                    // SESSD_CG_PROFILE only
};

__host__ __device__ constexpr uint32_t cg_idesc_f16(int M, int N) {
    return (1u << 4) /*C=F32*/ | (0u << 7) /*A=F16*/ | (0u << 10) /*B=F16*/ | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void cg_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// 16-byte global -> shared copy that bypasses L1 (allocating the gathered rows in L1 was measured: no hits worth the footprint)
__device__ __forceinline__ void cg_cp_async16(uint32_t smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cg_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void cg_sts_zero16(uint32_t saddr) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};\n" ::"r"(saddr), "r"(0) : "memory");
}

__device__ __forceinline__ void cg_tmem_ld16(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// -DSESSD_CG_PROFILE: per-CTA cycle counters ([ctas][16] int64 via sessd_set_cg_dbg; lab measurements only)
//   0 mma wait full_b, 1 mma wait full_a, 2 mma loop total, 3 tiles, 4 wload wait empty, 5 producer(w0) wait empty, 6 producer(w0) loop total,
//   7 epilogue wait acc_full (w0), 8 epilogue total (w0), 9 list build (t0), 10 kernel total (t0), 11 stage fills
#ifdef SESSD_CG_PROFILE
#define CG_T(var) const long long var = clock64()
#define CG_ADD(slot, t0) do { cg_acc[slot] += clock64() - (t0); } while (0)
#define CG_WAIT(slot, cond, stmt) do { const long long _t = clock64(); stmt; if (cond) cg_acc[slot] += clock64() - _t; } while (0)
#else
#define CG_T(var)
#define CG_ADD(slot, t0)
#define CG_WAIT(slot, cond, stmt) stmt
#endif

template <int CP, int COUT, int DEEP>
__global__ void __launch_bounds__(kCgThreads, DEEP ? 1 : 2) spconv_cg_kernel(const __grid_constant__ CUtensorMap map_w, const CgArgs a) {
    using C = CgCfg<CP, COUT, DEEP>;
    const int n_out = min(*a.d_n_out, a.max_out);
    const int ntiles = (n_out + kCgBM - 1) / kCgBM;
    if ((int)blockIdx.x >= ntiles) return;                       // whole CTA leaves together (before any barrier / TMEM use)
    const int kvol = a.kvol;
#ifdef SESSD_CG_PROFILE
    long long cg_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};      // per-thread counters, written once at the end (no memory traffic in the loops)
#endif
    CG_T(t_kernel);
    if (threadIdx.x == 8 * 32) prefetch_tensormap(&map_w);       // the weight-TMA warp's first load finds the descriptor cached

    extern __shared__ unsigned char smem_raw[];
    unsigned char *tiles = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint32_t *s_list = (uint32_t *)(tiles + C::kStages * C::kStage);      // [kvol][128]: (input row << 7) | tile row
    uint32_t *s_valid = s_list + kCgBM * kCgMaxK;                        // [kvol][4]: 128-bit row mask per offset
    int *s_cnt = (int *)(s_valid + kCgMaxK * 4);                         // [32]
    int *s_klist = s_cnt + 32;                                           // [32] + nact
    int *s_nact = s_klist + 32;
    int *s_off = s_nact + 1;                                             // [32] first list entry of every offset
    uint64_t *bars = (uint64_t *)(((uintptr_t)(s_off + 32) + 7) & ~(uintptr_t)7);
    uint64_t *full_a = bars, *full_b = bars + C::kStages, *empty = bars + 2 * C::kStages;
    uint64_t *acc_full = bars + 3 * C::kStages;
    uint32_t *tmem_slot = (uint32_t *)(acc_full + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t tiles_u32 = smem_u32(tiles);

    if (tid == 0) {
        for (int s = 0; s < C::kStages; ++s) { mbar_init(&full_a[s], kCgProdWarps / C::kStages); mbar_init(&full_b[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(C::kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    // every A tile starts as zeros (the B halves of the stages are always fully overwritten by the TMA)
    for (int s = 0; s < C::kStages; ++s)
        for (int o = tid * 16; o < C::kATile; o += kCgThreads * 16) cg_sts_zero16(tiles_u32 + (uint32_t)(s * C::kStage + o));
    cg_fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const float amax_in = __ldg(a.in_info), s_in = __ldg(a.in_info + 1);
    const float inv_act = 1.f / s_in;                            // exact: power of two
    const float s_out = pow2_scale_for_bound(amax_in * a.gain + a.shift_max);
    if (blockIdx.x == 0 && tid == 0 && a.out_info) a.out_info[1] = s_out;
    float vmax = 0.f;
    uint32_t dirty[4] = {0u, 0u, 0u, 0u};                        // rows of this warp's share of its stage that hold data (producer warps)
    int st0 = 0, acc_it = 0;                                     // stage of this tile's first fill / accumulator hand-overs so far (all roles
    uint32_t ph0 = 0;                                            // count alike); ph0 = phase bit of stage st0

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row0 = tile * kCgBM;
        const int rows = min(kCgBM, n_out - row0);
        // ---------------------------------------------------------------- the tile's per-offset pair lists (built once per rulebook by
        // sessd_rulebook_tile_lists: counts, row masks, (input row << 7 | tile row) entries grouped by offset) -> shared memory
        CG_T(t_list);
        {
            const unsigned int *rec = a.tiles + (size_t)tile * (size_t)a.tile_stride;
            const int c = (int)__ldg(rec + lane);                        // every warp ranks the 32 counts itself: no extra block-wide sync
            // the first list entries are requested together with the counts (one L2 round trip instead of two for tiles of <= 1280 pairs;
            // the record is 160 + 128 kvol words long, so the speculative reads stay inside it)
            unsigned int spec[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) spec[u] = (tid + u * kCgThreads < 128 * kvol) ? __ldg(rec + 160 + tid + u * kCgThreads) : 0u;
            const int incl = warp_incl_scan(c, lane);
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (warp == 0) {
                s_cnt[lane] = c;
                s_off[lane] = incl - c;
                const unsigned int m = __ballot_sync(0xffffffffu, c > 0);
                if (c > 0) s_klist[__popc(m & ((1u << lane) - 1u))] = lane;
                if (lane == 0) *s_nact = __popc(m);
            }
            if (tid < kvol * 4) s_valid[tid] = __ldg(rec + 32 + tid);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (tid + u * kCgThreads < total) s_list[tid + u * kCgThreads] = spec[u];
            for (int e = tid + 4 * kCgThreads; e < total; e += kCgThreads) s_list[e] = __ldg(rec + 160 + e);
        }
        __syncthreads();
        const int nact = *s_nact;
        if (tid == 0) { CG_ADD(9, t_list); }

        // stage index / phase of this tile's first fill (all roles advance them alike)
        if (warp == 9) {
            // ===================== MMA issue (one elected lane; the warp walks the loop together: operands stay in uniform registers) ===
            const uint32_t idesc = cg_idesc_f16(kCgBM, COUT);
            const uint32_t idesc2 = cg_idesc_f16(kCgBM, 2 * COUT);
            const uint64_t desc_hi = ((uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29))) << 32;      // SBO | version | SWIZZLE_128B
            const uint64_t desc_b64 = ((uint64_t)((512u >> 4) | (1u << 14) | (4u << 29))) << 32;      // narrow weight stages: SWIZZLE_64B rows
            const uint32_t tiles_lo = ((tiles_u32 >> 4) & 0x3FFFu) | (1u << 16);
            const uint32_t acc_main0 = tmem_base, acc_cross = tmem_base + COUT, acc_main1 = tmem_base + 2 * COUT;
            tc_fence_after();                                    // the previous tile's epilogue read the accumulators before the CTA-wide sync
            int s = st0;
            uint32_t ph = ph0;
            CG_T(t_mma);
            for (int j = 0; j < nact; ++j) {
                CG_WAIT(0, lane == 0, mbar_wait(&full_b[s], ph));
                CG_WAIT(1, lane == 0, mbar_wait(&full_a[s], ph));
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t st_lo = tiles_lo + (uint32_t)s * (C::kStage >> 4);
                    const uint64_t dA = desc_hi | st_lo;
                    const uint64_t dB = desc_hi | (st_lo + (C::kATile >> 4));
                    if constexpr (C::kWide) {
                        const uint64_t dAl = dA + ((kCgBM * 128) >> 4);
                        const uint64_t dBh = dB + (((j & 1) ? COUT * 128 : 0) >> 4);       // the b_hi rows inside the [X ; Y] tile
                        // all K steps of one product first, then the next product: consecutive MMAs of one shape into one accumulator pipeline,
                        // a switch to a product whose accumulator columns OVERLAP the previous one's drains the pipe (one switch per offset, not per K step)
                        if ((j & 1) == 0) {
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                cg_mma_f16(acc_main0, dA + 2 * kk, dB + 2 * kk, idesc2, (j != 0 || kk != 0) ? 1u : 0u);      // [main0|cross] (+)= a_hi x [b_hi;b_lo]
                        } else {
                            if (j == 1) {
                                cg_mma_f16(acc_main1, dA, dBh, idesc, 0u);                                                 // main1  = a_hi x b_hi
                                cg_mma_f16(acc_cross, dA, dB, idesc, 1u);                                                  // cross += a_hi x b_lo
                            }
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                if (j != 1 || kk != 0) cg_mma_f16(acc_cross, dA + 2 * kk, dB + 2 * kk, idesc2, 1u);         // [cross|main1] += a_hi x [b_lo;b_hi]
                        }
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) cg_mma_f16(acc_cross, dAl + 2 * kk, dBh + 2 * kk, idesc, 1u);        // cross += a_lo x b_hi
                    } else {
                        // narrow: A line = [hi 32 | lo 32] halves (SWIZZLE_128B); B stage = [b_hi rows ; b_lo rows] of 64 bytes each (SWIZZLE_64B),
                        // [b_lo ; b_hi] on odd offsets: one N = 2 COUT product gives main and the a_hi x b_lo cross term together (4 instead of
                        // 6 MMAs per offset: the A-operand reads from shared memory bound these layers)
                        const uint64_t dBn = desc_b64 | (st_lo + (C::kATile >> 4));
                        const uint64_t dBh = dBn + (((j & 1) ? COUT * 64 : 0) >> 4);
                        if ((j & 1) == 0) {
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk)
                                cg_mma_f16(acc_main0, dA + 2 * kk, dBn + 2 * kk, idesc2, (j != 0 || kk != 0) ? 1u : 0u);   // [main0|cross] (+)= a_hi x [b_hi;b_lo]
                        } else {
                            if (j == 1) {
                                cg_mma_f16(acc_main1, dA, dBh, idesc, 0u);                                                 // main1  = a_hi x b_hi
                                cg_mma_f16(acc_cross, dA, dBn, idesc, 1u);                                                 // cross += a_hi x b_lo
                            }
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk)
                                if (j != 1 || kk != 0) cg_mma_f16(acc_cross, dA + 2 * kk, dBn + 2 * kk, idesc2, 1u);        // [cross|main1] += a_hi x [b_lo;b_hi]
                        }
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) cg_mma_f16(acc_cross, dA + 4 + 2 * kk, dBh + 2 * kk, idesc, 1u);     // cross += a_lo x b_hi
                    }
                    tc_commit(&empty[s]);
                    if (j == nact - 1) tc_commit(acc_full);
                }
                __syncwarp();
                if (++s == C::kStages) { s = 0; ph ^= 1u; }
            }
            if (lane == 0) { CG_ADD(2, t_mma); }
#ifdef SESSD_CG_PROFILE
            if (lane == 0) { cg_acc[3] += 1; cg_acc[11] += nact; }
#endif
        } else if (warp == 8) {
            // ===================== weight tiles (TMA, one elected lane) =====================
            int s = st0;
            uint32_t ph = ph0;
            for (int j = 0; j < nact; ++j) {
                const int k = s_klist[j];
                CG_WAIT(4, lane == 0, mbar_wait(&empty[s], ph ^ 1u));
                if (elect_one()) {
                    mbar_expect_tx(&full_b[s], C::kBTile);
                    unsigned char *b_tile = tiles + s * C::kStage + C::kATile;
                    if constexpr (C::kWide) {
                        // [b_hi ; b_lo] on even steps, [b_lo ; b_hi] on odd steps (see the MMA issuer)
                        tma_load_4d(b_tile + ((j & 1) ? COUT * 128 : 0), &map_w, &full_b[s], 0, 0, 0, k);
                        tma_load_4d(b_tile + ((j & 1) ? 0 : COUT * 128), &map_w, &full_b[s], 0, 0, 1, k);
                    } else {
                        tma_load_4d(b_tile + ((j & 1) ? COUT * 64 : 0), &map_w, &full_b[s], 0, 0, 0, k);
                        tma_load_4d(b_tile + ((j & 1) ? 0 : COUT * 64), &map_w, &full_b[s], 0, 0, 1, k);
                    }
                }
                __syncwarp();
                if (++s == C::kStages) { s = 0; ph ^= 1u; }
            }
        } else {
            // ===================== producers: copy the rows that exist, clear the rows that stopped existing =====================
            // One producer GROUP per stage (kStages groups of 8 / kStages warps): a group fills only "its" stage, waits for its own copies
            // (cp.async.wait_all), fences them into the async proxy and arrives; the other groups' fills are in flight meanwhile.  (With
            // every warp working on every fill and cp.async.mbarrier.arrive.noinc as the hand-over, a warp's next shared-memory access
            // queued behind the arrival, i.e. behind its outstanding copies: ~900 clk = one L2 round trip per fill, serialised.)
            constexpr int kGroupWarps = kCgProdWarps / C::kStages;       // 2 (four stages) or 4 (two stages)
            constexpr int kGroupThreads = kGroupWarps * 32;
            constexpr int kOwnRows = kCgBM / kGroupWarps;                // rows of the stage whose zero state this warp maintains: 64 or 32
            constexpr int kOwnWords = kOwnRows / 32;
            constexpr int kLanesPerRow = C::kWide ? 16 : 8;
            constexpr int kRowsPerPass = kGroupThreads / kLanesPerRow;
            const int grp = warp / kGroupWarps, gw = warp % kGroupWarps;
            const int gtid = tid - grp * kGroupThreads;
            const int slot = gtid / kLanesPerRow;
            const int c = gtid % kLanesPerRow;
            const int half = c >> 3, cc = c & 7;                         // wide: chunk c of the 256-byte row = (hi | lo tile, 16-byte chunk)
            const uint32_t a_base = tiles_u32 + (uint32_t)(grp * C::kStage);
            CG_T(t_prod);
            // this tile's fills that land in stage `grp`: j = j0, j0 + kStages, ...; the phase bit of fill j follows the ring position
            for (int j = (grp - st0 + C::kStages) % C::kStages; j < nact; j += C::kStages) {
                const uint32_t ph = ph0 ^ ((uint32_t)((st0 + j) / C::kStages) & 1u);
                const int k = s_klist[j];
                if (lane == 0) CG_WAIT(5, warp == 0, mbar_wait(&empty[grp], ph ^ 1u));
                __syncwarp();
#pragma unroll
                for (int w = 0; w < kOwnWords; ++w) {
                    const uint32_t vs = s_valid[k * 4 + gw * kOwnWords + w];
                    const uint32_t z = dirty[w] & ~vs;                     // rows that hold data of the stage's previous offset and get none now
                    dirty[w] = vs;
                    if (z) {
#pragma unroll
                        for (int q4 = 0; q4 < 8; ++q4) {
                            const int r32 = q4 * 4 + (lane >> 3);
                            if ((z >> r32) & 1u) {
                                const uint32_t dst = a_base + (uint32_t)((gw * kOwnRows + w * 32 + r32) * 128 + (lane & 7) * 16);
                                cg_sts_zero16(dst);
                                if constexpr (C::kWide) cg_sts_zero16(dst + kCgBM * 128);
                            }
                        }
                    }
                }
                const int n = s_cnt[k];
                const uint32_t *lst = s_list + s_off[k];
                auto copy_row = [&](uint32_t e) {
                    const uint32_t r = e & 127u;
                    const size_t src = (size_t)(e >> 7);
                    if constexpr (C::kWide)
                        cg_cp_async16(a_base + (uint32_t)half * (kCgBM * 128) + r * 128u + (((uint32_t)cc ^ (r & 7u)) << 4),
                                      a.planes + src * 128 + half * 64 + cc * 8);
                    else
                        cg_cp_async16(a_base + r * 128u + (((uint32_t)cc ^ (r & 7u)) << 4), a.planes + src * 64 + cc * 8);
                };
                // four list entries per round: the shared-memory reads of a round are in flight together (the copies are `asm volatile`
                // with a memory clobber, so the compiler keeps every read behind the previous copy otherwise: one LDS latency per row)
                int i = slot;
                for (; i + 3 * kRowsPerPass < n; i += 4 * kRowsPerPass) {
                    const uint32_t e0 = lst[i], e1 = lst[i + kRowsPerPass], e2 = lst[i + 2 * kRowsPerPass], e3 = lst[i + 3 * kRowsPerPass];
                    copy_row(e0); copy_row(e1); copy_row(e2); copy_row(e3);
                }
                for (; i < n; i += kRowsPerPass) copy_row(lst[i]);
                asm volatile("cp.async.wait_all;\n" ::: "memory");
                cg_fence_proxy_async();                                  // copies and clears (generic proxy) -> visible to the tensor core (measured: free)
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_a[grp]);
            }

            if (tid == 0) { CG_ADD(6, t_prod); }
            CG_T(t_epi);
            // ===================== epilogue: TMEM -> registers -> BN / ReLU -> planes and / or fp32 rows =====================
            const int q = warp & 3, hcol = warp >> 2;
            constexpr int kNcol = COUT / 2;                              // 16 or 32 channels per thread
            const int r = q * 32 + lane;
            float v[kNcol];
            if (nact > 0) {
                if (lane == 0) CG_WAIT(7, warp == 0, mbar_wait(acc_full, (uint32_t)acc_it & 1u));
                __syncwarp();
                tc_fence_after();
                const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(hcol * kNcol);
#pragma unroll
                for (int c0 = 0; c0 < kNcol; c0 += 16) {
                    uint32_t m0[16], cr[16], m1[16];
                    cg_tmem_ld16(lane_base + c0, m0);
                    cg_tmem_ld16(lane_base + COUT + c0, cr);
                    if (nact > 1) cg_tmem_ld16(lane_base + 2 * COUT + c0, m1);
                    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float t = __uint_as_float(m0[i]) + __uint_as_float(cr[i]);
                        if (nact > 1) t += __uint_as_float(m1[i]);
                        v[c0 + i] = t;
                    }
                }
                tc_fence_before();                                       // the next tile's MMAs overwrite the accumulators after the CTA-wide sync
            } else {
#pragma unroll
                for (int i = 0; i < kNcol; ++i) v[i] = 0.f;
            }
            if (r < rows) {
                const size_t orow = (size_t)(row0 + r);
                uint32_t hprev[4] = {0u, 0u, 0u, 0u}, lprev[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int i = 0; i < kNcol; i += 8) {
                    const int n = hcol * kNcol + i;
                    float o[8];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float4 sc = *reinterpret_cast<const float4 *>(a.scale + n + 4 * h);
                        float4 sh = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (a.shift) sh = *reinterpret_cast<const float4 *>(a.shift + n + 4 * h);
                        o[4 * h + 0] = fmaf(v[i + 4 * h + 0] * inv_act, sc.x, sh.x); o[4 * h + 1] = fmaf(v[i + 4 * h + 1] * inv_act, sc.y, sh.y);
                        o[4 * h + 2] = fmaf(v[i + 4 * h + 2] * inv_act, sc.z, sh.z); o[4 * h + 3] = fmaf(v[i + 4 * h + 3] * inv_act, sc.w, sh.w);
                    }
                    if (a.relu) {
#pragma unroll
                        for (int t = 0; t < 8; ++t) o[t] = fmaxf(o[t], 0.f);
                    }
#pragma unroll
                    for (int t = 0; t < 8; ++t) vmax = fmaxf(vmax, fabsf(o[t]));
                    if (a.out_f32) stg256(a.out_f32 + orow * COUT + n, reinterpret_cast<const uint32_t *>(o), reinterpret_cast<const uint32_t *>(o) + 4);
                    if (a.out_planes) {
                        __align__(16) __half2 hi[4], lo[4];
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const float x0 = o[2 * t] * s_out, x1 = o[2 * t + 1] * s_out;
                            hi[t] = __floats2half2_rn(x0, x1);
                            const float2 f = __half22float2(hi[t]);
                            lo[t] = __floats2half2_rn(x0 - f.x, x1 - f.y);
                        }
                        // 32-byte stores (one full sector per lane): the first 8 channels of a 16-channel group wait for the second
                        __half *dst = a.out_planes + orow * (2 * C::kCPO) + n;
                        if ((i & 8) == 0) {
#pragma unroll
                            for (int t = 0; t < 4; ++t) { hprev[t] = *reinterpret_cast<const uint32_t *>(&hi[t]); lprev[t] = *reinterpret_cast<const uint32_t *>(&lo[t]); }
                        } else {
                            stg256(dst - 8, hprev, reinterpret_cast<const uint32_t *>(hi));
                            stg256(dst - 8 + C::kCPO, lprev, reinterpret_cast<const uint32_t *>(lo));
                        }
                    }
                }
            }
            if (tid == 0) { CG_ADD(8, t_epi); }
        }
        {   // advance the ring position by this tile's fills
            const int adv = st0 + nact;
            ph0 ^= (uint32_t)(adv / C::kStages) & 1u;
            st0 = adv % C::kStages;
        }
        acc_it += nact > 0 ? 1 : 0;
        __syncthreads();                                         // lists are rebuilt next; every role is done reading them
    }
    if (warp < kCgProdWarps && a.out_info) {
        const unsigned m = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(vmax));     // non-negative floats order like their bits
        if (lane == 0 && m != 0u) atomicMax(reinterpret_cast<unsigned *>(a.out_info), m);
    }
    if (tid == 0) { CG_ADD(10, t_kernel); }
#ifdef SESSD_CG_PROFILE
    if (a.dbg && lane == 0 && (warp == 0 || warp >= 8)) {
        long long *d = a.dbg + (size_t)blockIdx.x * 16;
        if (warp == 9) { d[0] = cg_acc[0]; d[1] = cg_acc[1]; d[2] = cg_acc[2]; d[3] = cg_acc[3]; d[11] = cg_acc[11]; }
        if (warp == 8) d[4] = cg_acc[4];
        if (warp == 0) { d[5] = cg_acc[5]; d[6] = cg_acc[6]; d[7] = cg_acc[7]; d[8] = cg_acc[8]; d[9] = cg_acc[9]; d[10] = cg_acc[10]; }
    }
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(C::kTmemCols) : "memory");
}

static long long *g_cg_dbg = nullptr;
static int g_cg_deep = 0;          // 1: deep pipeline, one CTA per SM (sessd_set_sp_cg_deep)

template <int CP, int COUT, int DEEP>
static int launch_spconv_cg(const CgArgs &a, const void *w_h2, cudaStream_t st) {
    using C = CgCfg<CP, COUT, DEEP>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(spconv_cg_kernel<CP, COUT, DEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem);
        if (e != cudaSuccess) return (int)e;
        attr_done = true;
    }
    CUtensorMap map_w;
    int rc;
    if (C::kWide) {       // [kvol][2 (hi|lo)][Cout][64]
        const cuuint64_t dims[4] = {64, (cuuint64_t)COUT, 2, (cuuint64_t)a.kvol};
        const cuuint32_t box[4] = {64, (cuuint32_t)COUT, 1, 1};
        rc = encode_map_4d(&map_w, w_h2, dims, box, nullptr, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2);
    } else {              // [kvol][2 (hi|lo)][Cout][32]: 64-byte rows, SWIZZLE_64B
        EncodeTiledFn enc = get_tensor_map_encoder();
        if (!enc) return SESSD_EINVAL;
        const cuuint64_t dims[4] = {32, (cuuint64_t)COUT, 2, (cuuint64_t)a.kvol};
        const cuuint64_t strides[3] = {64, (cuuint64_t)COUT * 64, (cuuint64_t)COUT * 128};
        const cuuint32_t box[4] = {32, (cuuint32_t)COUT, 1, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(w_h2), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        rc = r == CUDA_SUCCESS ? 0 : 700 + (int)r;
    }
    if (rc) return rc;
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        SESSD_CUDA_TRY(cudaGetDevice(&dev));
        SESSD_CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int tiles = div_up(a.max_out, kCgBM);
    const int per_sm = DEEP ? 1 : 2;
    const int grid = tiles < per_sm * num_sms ? tiles : per_sm * num_sms;        // persistent
    SESSD_LAUNCH((spconv_cg_kernel<CP, COUT, DEEP>), grid, kCgThreads, C::kSmem, st, map_w, a);
    return last_error();
}

}  // namespace sessd

using namespace sessd;

#ifdef SESSD_CG_PROFILE
extern "C" void sessd_set_cg_dbg(void *d) { sessd::g_cg_dbg = (long long *)d; }
#endif
// 1: twice the stages and one CTA per SM (launches with fewer tiles than SMs, e.g. single frames: the serial fill chain of a tile is latency-bound);
// 0 (default): two CTAs per SM (many tiles: throughput)
extern "C" void sessd_set_sp_cg_deep(int on) { sessd::g_cg_deep = on ? 1 : 0; }

// S4 (scn.py:106-149), pair-proportional tensor-core path.  d_in_planes [plane_rows][2][cp] fp16 with d_in_info = {abs-max, scale};
// weights / d_scale from ops.pack_weight_sp_h2 (cp = 64: [kvol][2][Cout][64], cp = 32: [kvol][2][Cout][32]; d_scale = BN scale *
// 2^-e[c]); gain / shift_max bound the output (see the header).  Outputs (each nullable, at least one): fp32 rows [max_out][cout],
// planes [>= max_out][2][cout <= 32 ? 32 : 64] + d_out_info = {abs-max (zero it once per frame), scale}.
// Supported (cp, cout): (32,32), (32,64), (64,64).
extern "C" int sessd_spconv_forward_cg(const void *d_in_planes, int cp, int plane_rows, const float *d_in_info, const void *d_tiles, int kvol,
                                       const int *d_n_out, int max_out, const void *d_weight_h2, int cout, const float *d_scale,
                                       const float *d_shift, int relu, float gain, float shift_max, float *d_out_f32, void *d_out_planes,
                                       float *d_out_info, void *stream) {
    if (!d_in_planes || !d_in_info || !d_tiles || !d_n_out || !d_weight_h2 || !d_scale || (!d_out_f32 && !d_out_planes) || max_out < 1 ||
        kvol < 1 || kvol > kCgMaxK || plane_rows < 1 || plane_rows > (1 << 25))
        return SESSD_EINVAL;
    if (d_out_planes && !d_out_info) return SESSD_EINVAL;
    CgArgs a;
    a.planes = (const __half *)d_in_planes; a.in_info = d_in_info; a.tiles = (const unsigned int *)d_tiles; a.tile_stride = 160 + 128 * kvol; a.d_n_out = d_n_out; a.kvol = kvol; a.max_out = max_out;
    a.relu = relu; a.scale = d_scale; a.shift = d_shift; a.gain = gain; a.shift_max = shift_max; a.out_f32 = d_out_f32;
    a.out_planes = (__half *)d_out_planes; a.out_info = d_out_info; a.dbg = g_cg_dbg;
    cudaStream_t st = (cudaStream_t)stream;
#define SESSD_CG_CASE(CPV, CO) \
    if (cp == CPV && cout == CO) return g_cg_deep ? launch_spconv_cg<CPV, CO, 1>(a, d_weight_h2, st) : launch_spconv_cg<CPV, CO, 0>(a, d_weight_h2, st);
    SESSD_CG_CASE(32, 32)
    SESSD_CG_CASE(32, 64)
    SESSD_CG_CASE(64, 64)
#undef SESSD_CG_CASE
    return SESSD_EINVAL;
}
