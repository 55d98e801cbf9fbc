// bevconv_h2.cu -- BEV conv / deconv (+BN+ReLU+residual) on the 5th-gen tensor cores with a TWO-TERM FP16 SPLIT of every fp32 operand.
//
// Replaces the cuDNN conv blocks of det3d/models/necks/rpn_v1.py:135-210 and the 1x1 head convs of
// det3d/models/bbox_heads/mg_head_sessd.py:202-230 (fp32 in, fp32 out, fp32 accumulate), like bevconv_tc.cu, but twice as fast:
//
//   * numerics: x = 2^-s (x_hi + x_lo) with x_hi = fp16_rn(2^s x), x_lo = fp16_rn(2^s x - x_hi): 22+ significand bits, exactly the
//     precision of the 3xTF32 split (tf32 and fp16 both carry 11 bits), but kind::f16 MMAs run at twice the tf32 rate and move half the
//     operand bytes.  fp16's narrow exponent range is handled by exact power-of-two scaling: activations by 2^s with s chosen from the
//     tensor's running abs-max (device scalar written by the producing kernel's epilogue: max maps into [2^10, 2^11)), weights per output
//     channel at pack time (folded into the epilogue scale).  Elements more than 2^13 below the maximum lose low bits of x_lo only: the
//     absolute error stays below 2^-35 of the tensor maximum.  Products a_hi*b_hi + a_hi*b_lo + a_lo*b_hi are accumulated in fp32 (TMEM).
//   * A operand: ONE halo patch (tile 8x16 pixels + the taps' reach, 64 channels, raw fp32) is TMA-loaded per channel chunk and serves
//     all taps (9x less L2->SM activation traffic than one box per tap).  Two groups of four warps read the tap-shifted pixel rows from
//     the (swizzled) patch, scale, split into fp16 hi/lo and store them to TENSOR MEMORY ([hi 16 cols | lo 16 cols] per 32 channels,
//     two fp16 per 32-bit column); all MMAs take A from TMEM (TS form) -- no shared-memory bandwidth for A at all.
//   * B operand: fp16 weight planes, K-major SWIZZLE_128B rows of 64 channels; per (chunk, tap) one [b_hi ; b_lo] 2n-row tile
//     ([b_lo ; b_hi] on odd steps) so that one N=2n MMA produces the main product and the a_hi*b_lo cross term in adjacent accumulators
//     (TMEM: main0 | cross | main1, alternating, because tcgen05 accumulation truncates -- see bevconv_tc.cu).
//   * warps: 0 patch TMA, 1 MMA issue (descriptors precomputed, ring unrolled), 2-5 / 6-9 split groups (2-5 also epilogue), 10 weight TMA.
#include <cuda_fp16.h>

#include <type_traits>

#include "../../include/sessd_b200.h"
#include "tc_common.cuh"

namespace sessd {

constexpr int kH2BM = 128, kH2TileH = 8, kH2TileW = 16;
constexpr int kH2Chunk = 64;                                   // channels per weight stage / per patch (two 32-channel fp32 boxes)
constexpr int kH2BStages = 4;
constexpr int kH2BStageBytes = 2 * 128 * 128;                  // [X ; Y] planes, up to 128 rows of 128 B each
constexpr int kH2PatchMaxPix = (kH2TileH + 2) * (kH2TileW + 2);
constexpr int kH2BoxBytes = ((kH2PatchMaxPix * 128 + 1023) / 1024) * 1024;
constexpr int kH2PatchBytes = 2 * kH2BoxBytes;
constexpr int kH2SmemBytes = kH2BStages * kH2BStageBytes + 2 * kH2PatchBytes + 1024 + 256;
constexpr int kH2Threads = 608;     // w0 patch TMA, w1 MMA, w2-9 split, w10 weight TMA, w11-18 epilogue
constexpr uint32_t kH2ACol = 384;                              // TMEM: 3 accumulators (<= 384 columns) + 4 A slots x 32 columns

struct H2Params {
    int batch, in_h, in_w, cin;
    int out_h, out_w, cout;
    int grid_h, grid_w;
    int out_stride;
    int nclass;
    int cls_ntaps[4], cls_off_y[4], cls_off_x[4];
    int tap_dy[4][9], tap_dx[4][9], tap_w[4][9];
    int relu;
    int n_tile;
    int tiles_x, tiles_y;
    int tiles, nblocks, total;                // pixel tiles, N blocks, work items = nclass * nblocks * tiles
    int cls_order[4];                         // classes by descending tap count (heavy items first in the static round-robin schedule)
    int patch_w, patch_h, org_dy, org_dx;     // patch rows = input rows oy0 + org_dy ... (patch_h of them), same for columns
    const float *amax_in;                     // nullable: abs-max of the input tensor (device scalar)
    float *amax_out;                          // nullable: running abs-max of the output tensor (atomicMax on the float bits)
    int ablate;                               // timing experiments only (garbage results): 1 no split work, 2 no MMAs, 4 no weight reloads, 8 no stores, 32 / 64 drop the A / B handshakes
    long long *dbg;                           // optional [ctas][8] globaltimer stamps
};

__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) /*C=F32*/ | (0u << 7) /*A=F16*/ | (0u << 10) /*B=F16*/ | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[tmem] * B[smem desc], kind::f16 (K = 16 per instruction; A: two fp16 per 32-bit TMEM column)
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// exact power-of-two activation scale 2^s that maps abs-max into [2^10, 2^11) (1 when abs-max is 0 / not finite)
__device__ __forceinline__ float h2_act_scale(float amax) {
    const uint32_t e = (__float_as_uint(amax) >> 23) & 0xFFu;
    if (e == 0 || e == 255) return 1.f;
    int bits = 264 - (int)e;                      // biased exponent of 2^(10 - (e - 127))
    bits = bits < 2 ? 2 : (bits > 252 ? 252 : bits);
    return __uint_as_float((uint32_t)bits << 23);
}

// work item w (persistent CTAs take w = blockIdx.x, + gridDim.x, ...): class-major (heaviest class first), then N block, then pixel tile
struct H2Item { int cls, n0, b, oy0, ox0, ntaps; };
__device__ __forceinline__ H2Item h2_decode(const H2Params &p, int w) {
    H2Item it;
    const int per_cls = p.nblocks * p.tiles;
    const int cr = w / per_cls;
    int rem = w - cr * per_cls;
    const int nb = rem / p.tiles;
    int t = rem - nb * p.tiles;
    it.cls = p.cls_order[cr];
    it.n0 = nb * p.n_tile;
    const int tx = t % p.tiles_x; t /= p.tiles_x;
    const int ty = t % p.tiles_y;
    it.b = t / p.tiles_y;
    it.oy0 = ty * kH2TileH; it.ox0 = tx * kH2TileW;
    it.ntaps = p.cls_ntaps[it.cls];
    return it;
}

__device__ __forceinline__ void tmem_ld_32x32b_x16_nowait(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// profiling aid: CTA 0 writes globaltimer stamps into dbg[4096 + idx]
#define H2_STAMP(idx)                                                                      \
    do {                                                                                   \
        if (p.dbg && blockIdx.x == 0) {                                                    \
            long long ts_;                                                                 \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts_));                        \
            p.dbg[4096 + (idx)] = ts_;                                                     \
        }                                                                                  \
    } while (0)

__global__ void __launch_bounds__(kH2Threads, 1) bev_conv_h2_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                   const __grid_constant__ CUtensorMap map_b,
                                                                   const float *__restrict__ scale, const float *__restrict__ shift,
                                                                   const float *__restrict__ resid, float *__restrict__ out, H2Params p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *tiles = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char *patches = tiles + kH2BStages * kH2BStageBytes;
    uint64_t *bars = (uint64_t *)(patches + 2 * kH2PatchBytes);
    // stage_done[S]: ONE tcgen05.commit per weight stage releases the stage's B tile (to the weight producer) and its two A slots (to the
    // split warps); commits are serialised in the tensor pipe at a few hundred cycles each (scripts/latency_probe.py), so fewer is faster
    uint64_t *patch_full = bars, *patch_empty = bars + 2, *b_full = bars + 4, *stage_done = bars + 8, *a_full = bars + 12;
    uint64_t *acc_full = bars + 20, *acc_free = bars + 21;
    uint32_t *tmem_slot = (uint32_t *)(bars + 22);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = p.cin / kH2Chunk;
    const uint32_t b_plane_bytes = (uint32_t)p.n_tile * 128u;

    if (p.dbg && threadIdx.x == 0) {
        long long ts; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts));
        p.dbg[(size_t)blockIdx.x * 8 + 0] = ts;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(&patch_full[s], 1); mbar_init(&patch_empty[s], 8); }
        for (int s = 0; s < 4; ++s) { mbar_init(&b_full[s], 1); mbar_init(&stage_done[s], 1); mbar_init(&a_full[s], 4); }
        mbar_init(acc_full, 1);
        mbar_init(acc_free, 8);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== activation patches: one halo patch per (item, 64-channel chunk) =====================
        if (lane == 0) {
            const uint32_t box_bytes = (uint32_t)(p.patch_w * p.patch_h) * 128u;
            int gcc = 0;
            for (int w = blockIdx.x; w < p.total; w += gridDim.x) {
                const H2Item it = h2_decode(p, w);
                for (int cc = 0; cc < nchunks; ++cc, ++gcc) {
                    const int pb = gcc & 1;
                    mbar_wait(&patch_empty[pb], ((gcc >> 1) & 1) ^ 1);
                    mbar_expect_tx(&patch_full[pb], 2 * box_bytes);
                    unsigned char *dst = patches + pb * kH2PatchBytes;
                    tma_load_4d(dst, &map_a, &patch_full[pb], cc * kH2Chunk, it.ox0 + p.org_dx, it.oy0 + p.org_dy, it.b);
                    tma_load_4d(dst + kH2BoxBytes, &map_a, &patch_full[pb], cc * kH2Chunk + 32, it.ox0 + p.org_dx, it.oy0 + p.org_dy, it.b);
                }
            }
        }
    } else if (warp == 10) {
        // ===================== weight tiles: one [X ; Y] stage per (item, chunk, tap) =====================
        if (lane == 0) {
            int gbj = 0;
            for (int w = blockIdx.x; w < p.total; w += gridDim.x) {
                const H2Item it = h2_decode(p, w);
                for (int cc = 0; cc < nchunks; ++cc)
                    for (int tap = 0; tap < it.ntaps; ++tap, ++gbj) {
                        const int s = gbj & 3;
                        if (!(p.ablate & 64)) mbar_wait(&stage_done[s], ((gbj >> 2) & 1) ^ 1);
                        if ((p.ablate & 4) && gbj >= kH2BStages) { mbar_arrive(&b_full[s]); continue; }
                        mbar_expect_tx(&b_full[s], 2 * b_plane_bytes);
                        unsigned char *st = tiles + s * kH2BStageBytes;
                        const int wtap = p.tap_w[it.cls][tap];
                        // every item has an even number of stages, so the running parity is also the item-local one
                        const uint32_t hi_off = (gbj & 1) ? b_plane_bytes : 0u, lo_off = (gbj & 1) ? 0u : b_plane_bytes;
                        tma_load_4d(st + hi_off, &map_b, &b_full[s], cc * kH2Chunk, it.n0, wtap, 0);
                        tma_load_4d(st + lo_off, &map_b, &b_full[s], cc * kH2Chunk, it.n0, wtap, 1);
                    }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issue (one thread; keep its scalar work per tcgen05.mma minimal) =====================
        const uint32_t idesc1 = make_idesc_f16(kH2BM, p.n_tile), idesc2 = make_idesc_f16(kH2BM, 2 * p.n_tile);
        const uint32_t acc_main0 = tmem_base, acc_cross = tmem_base + (uint32_t)p.n_tile, acc_main1 = tmem_base + 2 * (uint32_t)p.n_tile;
        const uint64_t desc_hi = ((uint64_t)(((1024u >> 4)) | (1u << 14) | (2u << 29))) << 32;          // SBO | version | SWIZZLE_128B
        const uint32_t tiles_lo = ((smem_u32(tiles) >> 4) & 0x3FFFu) | (1u << 16);
        const uint64_t dcat0 = desc_hi | tiles_lo;
        const uint32_t plane_lo = b_plane_bytes >> 4;
        // compact run-time loop on purpose (no per-stage unrolling): this warp's instruction footprint must stay in the instruction cache
        int gbj = 0, iter = 0;
        for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++iter) {
            const int nbj = nchunks * p.cls_ntaps[p.cls_order[w / (p.nblocks * p.tiles)]];
            if (iter > 0) {                          // the epilogue warps must have drained the previous item's accumulators
                mbar_wait(acc_free, (iter - 1) & 1);
                tc_fence_after();
            }
#pragma unroll 1
            for (int lbj = 0; lbj < nbj; ++lbj, ++gbj) {
                // gbj: running stage counter (ring position / barrier parities); lbj: stage index inside the item (accumulator roles);
                // every item has an even stage count, so both have the same parity
                const int S = gbj & 3, par = gbj & 1;
                const uint64_t dcat = dcat0 + (uint32_t)(S * (kH2BStageBytes >> 4));
                const uint64_t dbhi = dcat + (par ? plane_lo : 0u);
                const uint32_t d2 = par ? acc_cross : acc_main0;        // even stages: [main0|cross], odd stages: [cross|main1]
                if (!(p.ablate & 64)) mbar_wait(&b_full[S], (gbj >> 2) & 1);
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    const int slot = 2 * par + h;
                    if (!(p.ablate & 32)) mbar_wait(&a_full[slot], (gbj >> 1) & 1);
                    tc_fence_after();
                    if (lane == 0 && !(p.ablate & 2)) {
                        const uint32_t a_hi = tmem_base + kH2ACol + (uint32_t)slot * 32u, a_lo = a_hi + 16u;
                        const uint32_t k0 = (uint32_t)(h * 4);                      // 32-channel half: +64 B (>>4); K=16 sub-step: +32 B
                        if (lbj == 1 && h == 0) {
                            tc_mma_f16_ts(acc_cross, a_hi, dcat + k0, idesc1, 1);               // cross += a_hi x b_lo
                            tc_mma_f16_ts(acc_main1, a_hi, dbhi + k0, idesc1, 0);               // main1  = a_hi x b_hi (first write)
                        } else {
                            tc_mma_f16_ts(d2, a_hi, dcat + k0, idesc2, (lbj | h) != 0);          // [main|cross] (+)= a_hi x [b_hi;b_lo]
                        }
                        tc_mma_f16_ts(acc_cross, a_lo, dbhi + k0, idesc1, 1);                   // cross += a_lo x b_hi
                        tc_mma_f16_ts(d2, a_hi + 8, dcat + k0 + 2, idesc2, 1);
                        tc_mma_f16_ts(acc_cross, a_lo + 8, dbhi + k0 + 2, idesc1, 1);
                    }
                    __syncwarp();
                }
                if (lane == 0) {
                    tc_commit(&stage_done[S]);
                    if (lbj == nbj - 1) tc_commit(acc_full);
                }
                __syncwarp();
            }
        }
    } else if (warp < 10) {
        // ===================== split warps: patch (smem, fp32) -> scaled fp16 hi/lo -> tensor memory =====================
        const int q = warp & 3;
        const int r = q * 32 + lane;                 // tile pixel = TMEM lane
        const int grp = (warp - 2) >> 2;             // group g converts channel half g of every (chunk, tap)
        const int ly = r / kH2TileW, lx = r % kH2TileW;
        const float sa = p.amax_in ? h2_act_scale(__ldg(p.amax_in)) : 1.f;
        int gbj = 0, gcc = 0;
        for (int w = blockIdx.x; w < p.total; w += gridDim.x) {
            const int cls = p.cls_order[w / (p.nblocks * p.tiles)];
            const int ntaps = p.cls_ntaps[cls];
            for (int cc = 0; cc < nchunks; ++cc, ++gcc) {
                mbar_wait(&patch_full[gcc & 1], (gcc >> 1) & 1);
                const uint32_t box = smem_u32(patches) + (uint32_t)((gcc & 1) * kH2PatchBytes + grp * kH2BoxBytes);
                for (int tap = 0; tap < ntaps; ++tap, ++gbj) {
                    const int j = 2 * gbj + grp, slot = j & 3;
                    if (p.ablate & 1) {
                        if (gbj >= 2) mbar_wait(&stage_done[(gbj - 2) & 3], ((gbj - 2) >> 2) & 1);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&a_full[slot]);
                        continue;
                    }
                    const int prow = (ly + p.tap_dy[cls][tap] - p.org_dy) * p.patch_w + lx + p.tap_dx[cls][tap] - p.org_dx;
                    const uint32_t a = box + (uint32_t)prow * 128u;
                    uint32_t regs[32];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {    // SWIZZLE_128B box: logical 16-byte chunk c of patch row prow sits at chunk c ^ (prow & 7)
                        const float4 v = lds128(a + (uint32_t)((c ^ (prow & 7)) << 4));
                        const float x0 = v.x * sa, x1 = v.y * sa, x2 = v.z * sa, x3 = v.w * sa;
                        const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
                        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                        const __half2 l01 = __floats2half2_rn(x0 - f01.x, x1 - f01.y), l23 = __floats2half2_rn(x2 - f23.x, x3 - f23.y);
                        regs[2 * c] = *reinterpret_cast<const uint32_t *>(&h01);
                        regs[2 * c + 1] = *reinterpret_cast<const uint32_t *>(&h23);
                        regs[16 + 2 * c] = *reinterpret_cast<const uint32_t *>(&l01);
                        regs[16 + 2 * c + 1] = *reinterpret_cast<const uint32_t *>(&l23);
                    }
                    if (gbj >= 2 && !(p.ablate & 32)) mbar_wait(&stage_done[(gbj - 2) & 3], ((gbj - 2) >> 2) & 1);   // slot last read by the MMAs of stage gbj-2
                    tc_fence_after();
                    tmem_st_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + kH2ACol + (uint32_t)slot * 32u, regs);
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_full[slot]);     // one arrival per warp: 128 per-thread arrivals serialise on the barrier
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&patch_empty[gcc & 1]);  // this warp is done reading the chunk's patch
            }
        }
    } else {
        // ===================== epilogue warps (11-18): drain TMEM to registers, release the accumulators, then BN/ReLU/residual/store
        const int q = warp & 3;
        const int half = (warp - 11) >> 2;           // warps 11-14: first half of the N tile's columns, 15-18: second half
        const int r = q * 32 + lane;
        const int ly = r / kH2TileW, lx = r % kH2TileW;
        const int ncol = p.n_tile >> 1;              // 64 (n_tile 128) or 16 (n_tile 32)
        const float inv_sa = p.amax_in ? 1.f / h2_act_scale(__ldg(p.amax_in)) : 1.f;   // exact: power of two
        float vmax = 0.f;
        int iter = 0;
        for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++iter) {
            const H2Item it = h2_decode(p, w);
            float v[64];
            mbar_wait(acc_full, iter & 1);
            tc_fence_after();
            const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * ncol);
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) {
                if (c0 < ncol) {
                    uint32_t m0[16], cr[16], m1[16];
                    tmem_ld_32x32b_x16_nowait(lane_base + c0, m0);
                    tmem_ld_32x32b_x16_nowait(lane_base + (uint32_t)p.n_tile + c0, cr);
                    tmem_ld_32x32b_x16_nowait(lane_base + 2 * (uint32_t)p.n_tile + c0, m1);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[c0 + i] = (__uint_as_float(m0[i]) + __uint_as_float(cr[i])) + __uint_as_float(m1[i]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_free);    // the next item's MMAs may overwrite the accumulators now
            const int gy = it.oy0 + ly, gx = it.ox0 + lx;
            if (gy < p.grid_h && gx < p.grid_w) {
                const size_t opix = (((size_t)it.b * p.out_h + (size_t)gy * p.out_stride + p.cls_off_y[it.cls]) * p.out_w + (size_t)gx * p.out_stride +
                                     p.cls_off_x[it.cls]);
#pragma unroll
                for (int i = 0; i < 64; i += 4) {
                    const int n = it.n0 + half * ncol + i;
                    if (i < ncol && n < p.cout) {
                        const float4 sc = *reinterpret_cast<const float4 *>(scale + n);
                        float4 sh = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (shift) sh = *reinterpret_cast<const float4 *>(shift + n);
                        float4 o;
                        o.x = fmaf(v[i + 0] * inv_sa, sc.x, sh.x); o.y = fmaf(v[i + 1] * inv_sa, sc.y, sh.y);
                        o.z = fmaf(v[i + 2] * inv_sa, sc.z, sh.z); o.w = fmaf(v[i + 3] * inv_sa, sc.w, sh.w);
                        if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                        const size_t off = opix * p.cout + n;
                        if (resid) {
                            const float4 rr = *reinterpret_cast<const float4 *>(resid + off);
                            o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
                        }
                        vmax = fmaxf(vmax, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
                        if (!(p.ablate & 8)) *reinterpret_cast<float4 *>(out + off) = o;
                    }
                }
            }
        }
        if (p.amax_out) {
            const unsigned m = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(vmax));     // non-negative floats order like their bits
            if (lane == 0 && m != 0u) atomicMax(reinterpret_cast<unsigned *>(p.amax_out), m);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512) : "memory");
    if (p.dbg && threadIdx.x == 0) {
        long long ts; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts));
        p.dbg[(size_t)blockIdx.x * 8 + 3] = ts;
    }
}

static int g_h2_ablate = 0;
static long long *g_h2_dbg = nullptr;

}  // namespace sessd

using namespace sessd;

// profiling experiments only (see H2Params::ablate / dbg)
extern "C" void sessd_set_h2_debug(int ablate_mask, void *d_stamps) { sessd::g_h2_ablate = ablate_mask; sessd::g_h2_dbg = (long long *)d_stamps; }


static int launch_h2(const float *d_in, const void *d_w, int w_taps, int cout_pad, const float *d_scale, const float *d_shift,
                     const float *d_residual, float *d_out, H2Params &p, void *stream) {
    const int n_tile = p.cout <= 32 ? 32 : 128;
    if (cout_pad % n_tile || cout_pad < p.cout || !d_scale) return SESSD_EINVAL;
    int dy0 = 1 << 30, dy1 = -(1 << 30), dx0 = 1 << 30, dx1 = -(1 << 30);
    for (int c = 0; c < p.nclass; ++c)
        for (int t = 0; t < p.cls_ntaps[c]; ++t) {
            dy0 = min(dy0, p.tap_dy[c][t]); dy1 = max(dy1, p.tap_dy[c][t]);
            dx0 = min(dx0, p.tap_dx[c][t]); dx1 = max(dx1, p.tap_dx[c][t]);
        }
    p.org_dy = dy0; p.org_dx = dx0;
    p.patch_h = kH2TileH + dy1 - dy0; p.patch_w = kH2TileW + dx1 - dx0;
    if (p.patch_h * p.patch_w > kH2PatchMaxPix) return SESSD_EINVAL;
    CUtensorMap map_a, map_b;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)p.cin, (cuuint64_t)p.in_w, (cuuint64_t)p.in_h, (cuuint64_t)p.batch};
        const cuuint32_t box[4] = {32, (cuuint32_t)p.patch_w, (cuuint32_t)p.patch_h, 1};
        int rc = encode_map_4d(&map_a, d_in, dims, box);
        if (rc) return rc;
    }
    {
        const cuuint64_t dims[4] = {(cuuint64_t)p.cin, (cuuint64_t)cout_pad, (cuuint64_t)w_taps, 2};
        const cuuint32_t box[4] = {kH2Chunk, (cuuint32_t)n_tile, 1, 1};
        int rc = encode_map_4d(&map_b, d_w, dims, box, nullptr, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2);
        if (rc) return rc;
    }
    static bool attr_done = false;
    if (!attr_done) {
        SESSD_CUDA_TRY(cudaFuncSetAttribute(bev_conv_h2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kH2SmemBytes));
        attr_done = true;
    }
    p.n_tile = n_tile;
    p.ablate = g_h2_ablate;
    p.dbg = g_h2_dbg;
    p.tiles_x = div_up(p.grid_w, kH2TileW);
    p.tiles_y = div_up(p.grid_h, kH2TileH);
    p.tiles = p.tiles_x * p.tiles_y * p.batch;
    p.nblocks = cout_pad / n_tile;
    p.total = p.nclass * p.nblocks * p.tiles;
    for (int c = 0; c < p.nclass; ++c) p.cls_order[c] = c;
    for (int i = 1; i < p.nclass; ++i)            // insertion sort by descending tap count
        for (int k = i; k > 0 && p.cls_ntaps[p.cls_order[k]] > p.cls_ntaps[p.cls_order[k - 1]]; --k) {
            const int tmp = p.cls_order[k]; p.cls_order[k] = p.cls_order[k - 1]; p.cls_order[k - 1] = tmp;
        }
    for (int c = 0; c < p.nclass; ++c)
        if ((p.cls_ntaps[c] * (p.cin / kH2Chunk)) & 1) return SESSD_EINVAL;     // the [main0|cross|main1] alternation needs an even stage count
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        SESSD_CUDA_TRY(cudaGetDevice(&dev));
        SESSD_CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int grid = p.total < num_sms ? p.total : num_sms;
    bev_conv_h2_kernel<<<grid, kH2Threads, kH2SmemBytes, (cudaStream_t)stream>>>(map_a, map_b, d_scale, d_shift, d_residual, d_out, p);
    ++g_launches;
    return last_error();
}

// fp16-split tensor-core variant of sessd_bev_conv (stride-1 tap lists whose reach fits a 10x18 halo patch; cin a multiple of 64).
// d_weight_h2: __half [2 (hi|lo)][ntaps][cout_pad][cin] of the per-output-channel scaled weights 2^e[n] * w (pack: ops.pack_weight_h2);
// d_scale must carry the matching 2^-e[n] (times the folded BN scale).
extern "C" int sessd_bev_conv_h2(const float *d_in, const void *d_weight_h2, int cout_pad, const float *d_scale, const float *d_shift,
                                 const float *d_residual, float *d_out, const sessd_conv_desc *desc, const float *d_amax_in, float *d_amax_out,
                                 void *stream) {
    if (!d_in || !d_weight_h2 || !d_out || !desc || !d_scale) return SESSD_EINVAL;
    const sessd_conv_desc &d = *desc;
    if (d.batch < 1 || d.cin < kH2Chunk || d.cin % kH2Chunk || d.cout < 4 || d.cout % 4 || d.ntaps < 1 || d.ntaps > 9 || d.in_stride != 1 ||
        d.out_stride < 1 || d.grid_h < 1 || d.grid_w < 1)
        return SESSD_EINVAL;
    if ((d.grid_h - 1) * d.out_stride + d.out_off_y >= d.out_h || (d.grid_w - 1) * d.out_stride + d.out_off_x >= d.out_w) return SESSD_EINVAL;
    H2Params p = {};
    p.batch = d.batch; p.in_h = d.in_h; p.in_w = d.in_w; p.cin = d.cin;
    p.out_h = d.out_h; p.out_w = d.out_w; p.cout = d.cout;
    p.grid_h = d.grid_h; p.grid_w = d.grid_w;
    p.out_stride = d.out_stride;
    p.relu = d.relu;
    p.nclass = 1;
    p.cls_ntaps[0] = d.ntaps; p.cls_off_y[0] = d.out_off_y; p.cls_off_x[0] = d.out_off_x;
    for (int t = 0; t < d.ntaps; ++t) { p.tap_dy[0][t] = d.tap_dy[t]; p.tap_dx[0][t] = d.tap_dx[t]; p.tap_w[0][t] = t; }
    p.amax_in = d_amax_in; p.amax_out = d_amax_out;
    return launch_h2(d_in, d_weight_h2, d.ntaps, cout_pad, d_scale, d_shift, d_residual, d_out, p, stream);
}

// ConvTranspose2d(k3, s2, p1, op1) + BN + ReLU (+ residual), four output-parity classes in one launch (see sessd_bev_deconv_tc).
extern "C" int sessd_bev_deconv_h2(const float *d_in, const void *d_weight_h2, int cout_pad, const float *d_scale, const float *d_shift,
                                   const float *d_residual, float *d_out, int batch, int in_h, int in_w, int cin, int cout, int relu,
                                   const float *d_amax_in, float *d_amax_out, void *stream) {
    if (!d_in || !d_weight_h2 || !d_out || !d_scale || batch < 1 || in_h < 1 || in_w < 1 || cin < kH2Chunk || cin % kH2Chunk || cout < 4 || cout % 4)
        return SESSD_EINVAL;
    H2Params p = {};
    p.batch = batch; p.in_h = in_h; p.in_w = in_w; p.cin = cin;
    p.out_h = 2 * in_h; p.out_w = 2 * in_w; p.cout = cout;
    p.grid_h = in_h; p.grid_w = in_w;
    p.out_stride = 2;
    p.relu = relu;
    p.nclass = 4;
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
            const int c = py * 2 + px;
            p.cls_off_y[c] = py; p.cls_off_x[c] = px;
            // out[2y+py] receives in[y+dy] * W[ky] with 2y+py = 2(y+dy) - 1 + ky:  py=0 -> (ky=1,dy=0);  py=1 -> (ky=0,dy=1), (ky=2,dy=0)
            const int kys[2] = {py == 0 ? 1 : 0, 2}, dys[2] = {py == 0 ? 0 : 1, 0}, ny = py == 0 ? 1 : 2;
            const int kxs[2] = {px == 0 ? 1 : 0, 2}, dxs[2] = {px == 0 ? 0 : 1, 0}, nx = px == 0 ? 1 : 2;
            int t = 0;
            for (int a = 0; a < ny; ++a)
                for (int bb = 0; bb < nx; ++bb) {
                    p.tap_dy[c][t] = dys[a]; p.tap_dx[c][t] = dxs[bb]; p.tap_w[c][t] = kys[a] * 3 + kxs[bb];
                    ++t;
                }
            p.cls_ntaps[c] = t;
        }
    p.amax_in = d_amax_in; p.amax_out = d_amax_out;
    return launch_h2(d_in, d_weight_h2, 9, cout_pad, d_scale, d_shift, d_residual, d_out, p, stream);
}

