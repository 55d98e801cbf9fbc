// bevconv_tc.cu -- BEV neck/head convolutions on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), fp32-accurate
// through the 3xTF32 split:   a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi,   a_hi = tf32(a), a_lo = a - a_hi  (exact in fp32).
//
// Same contract as bev_conv_kernel (bevconv.cu): tap-list implicit GEMM over NHWC activations with a fused
// BatchNorm(eval)+ReLU(+residual) epilogue; replaces the cuDNN conv / deconv + BN + ReLU triplets of
// det3d/models/necks/rpn_v1.py:135-210 and the head convs of det3d/models/bbox_heads/mg_head_sessd.py:202-230.
// Single-pass TF32 (10-bit mantissa; what cuDNN silently uses on sm_80+) cannot hold the 1e-4 parity bar through 13
// stacked layers -- and truncation is biased on post-ReLU activations -- hence three tensor-core products per tile.
//
// Tiling: one CTA = one 8x16 output-pixel patch (M = 128) x one 128-wide (or 32-wide, head) slice of Cout.
//   K loop over (tap, 32-channel chunk): BK = 32 tf32 = one 128-byte swizzle row.
//   A tile  : the input patch shifted by the tap offset is a plain 4-D TMA box {32 ch, 16 x, 8 y, 1 image} of the NHWC
//             tensor; out-of-range coordinates (conv padding, deconv fringe, partial tiles) are zero-filled by TMA.
//   B tiles : pre-split weights [plane hi|lo][tap][Cout][Cin] (K-major), box {32, N, 1, 1}.
//   All tiles land in the canonical K-major SWIZZLE_128B layout that tcgen05.mma descriptors address directly.
// Warp roles (192 threads): w0 TMA producer | w1 TMEM alloc + MMA issuer | w2..w5 split a -> (a_hi, a_lo) in smem, then
// epilogue (TMEM -> registers -> BN/ReLU/residual -> global).  3-stage mbarrier pipeline:
//   full[s] (TMA bytes landed) -> split[s] (a_hi/a_lo written, fence.proxy.async) -> 12 x tcgen05.mma (M128 N128 K8)
//   -> tcgen05.commit -> empty[s];  last commit -> acc_full -> epilogue.
// Accumulation: tcgen05.mma adds into its fp32 TMEM accumulator with truncation (measured: ~2^-24 of the accumulator magnitude
// per instruction, biased, i.e. growing LINEARLY with the number of accumulation steps -- 1.6e-5 relative after the 432
// instructions of a 3x3x128 tile).  So the K loop is spread round-robin over THREE "main" accumulators (a_hi*b_hi) plus ONE
// accumulator for the small cross terms (a_lo*b_hi + a_hi*b_lo, 2^-10 smaller), 4 x 128 = all 512 TMEM columns, and the
// epilogue adds the four partial sums in round-to-nearest fp32.  Result: ~1.5e-6 relative per layer.
// Every mbarrier wait is bounded (trap on timeout) so a descriptor mistake aborts the kernel instead of hanging the GPU.
#include <type_traits>

#include "tc_common.cuh"

namespace sessd {

constexpr int kTcStages = 3;
constexpr int kTcBM = 128;                 // pixels per tile: 8 rows x 16 cols
constexpr int kTcTileH = 8, kTcTileW = 16;
constexpr int kTcBK = 32;                  // tf32 elements per K chunk (128 bytes)
constexpr int kTcTileBytes = kTcBM * kTcBK * 4;   // 16 KB
constexpr int kTcStageBytes = 4 * kTcTileBytes;   // A_hi (in place), A_lo, B_hi, B_lo
constexpr int kTcThreads = 192;
constexpr int kTcSmemBytes = kTcStages * kTcStageBytes + 1024 /*align*/ + 256 /*barriers*/;

struct TcParams {
    int batch, in_h, in_w, cin;
    int out_h, out_w, cout;
    int grid_h, grid_w;
    int out_stride;
    // up to 4 "classes" per launch (blockIdx.z): a plain conv is one class; ConvTranspose2d(k3,s2,p1,op1) is its 4 output-parity
    // classes with 1/2/2/4 taps each, all reading the same input and the same 9-tap weight tensor
    int nclass;
    int cls_ntaps[4], cls_off_y[4], cls_off_x[4];
    int tap_dy[4][9], tap_dx[4][9], tap_w[4][9];
    int relu;
    int n_tile;          // 128 or 32: UMMA N and rows of each B tile
    int in_stride;       // 1 or 2 (strided TMA box for the stride-2 conv)
    long long *dbg;      // optional [ctas][8] globaltimer stamps (profiling experiments)
    int ablate;          // timing experiments only (results become garbage): 1 skip the hi/lo split, 2 hi*hi product only, 4 no TMA reloads
    int tiles_x, tiles_y;
};

// ---------------------------------------------------------------------------------------------------------------- kernel
template <int CS>   // CTAs per cluster sharing (multicasting) the weight tiles
__global__ void __launch_bounds__(kTcThreads, 1) bev_conv_tc_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                    const __grid_constant__ CUtensorMap map_b,
                                                                    const float *__restrict__ scale, const float *__restrict__ shift,
                                                                    const float *__restrict__ resid, float *__restrict__ out, TcParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *tiles = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);     // SWIZZLE_128B atoms: 1024 B
    uint64_t *bars = (uint64_t *)(tiles + kTcStages * kTcStageBytes);
    uint64_t *full = bars, *split = bars + kTcStages, *empty = bars + 2 * kTcStages, *acc_full = bars + 3 * kTcStages;
    uint32_t *tmem_slot = (uint32_t *)(bars + 3 * kTcStages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto stamp = [&](int slot) {
        if (p.dbg) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            p.dbg[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 + slot] = (long long)t;
        }
    };
    auto trace = [&](int it, int slot) {
        if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && it < 128) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            p.dbg[4096 + it * 8 + slot] = (long long)t;
        }
    };
    if (threadIdx.x == 0) stamp(0);
    // tile coordinates
    int t = blockIdx.x;
    const int tx = t % p.tiles_x; t /= p.tiles_x;
    const int ty = t % p.tiles_y;
    const int b = t / p.tiles_y;
    const int oy0 = ty * kTcTileH, ox0 = tx * kTcTileW;
    const int n0 = blockIdx.y * p.n_tile;
    const int cls = blockIdx.z;
    const int kchunks = p.cin / kTcBK;
    const int steps = p.cls_ntaps[cls] * kchunks;
    const uint32_t b_tile_bytes = (uint32_t)p.n_tile * kTcBK * 4;

    const uint32_t crank = (CS > 1) ? cluster_cta_rank() : 0u;
    constexpr uint16_t kMask = (uint16_t)((1u << CS) - 1u);
    if (threadIdx.x == 0) {
        for (int s = 0; s < kTcStages; ++s) { mbar_init(&full[s], 1); mbar_init(&split[s], 128); mbar_init(&empty[s], CS); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) {   // TMEM: 128 lanes x 512 fp32 columns = 3 main partial accumulators + 1 cross-term accumulator
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CS > 1) cluster_sync_all();        // peers' barriers must be initialised before any multicast / remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) stamp(1);                      // setup done (barriers, TMEM alloc)

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            for (int it = 0; it < steps; ++it) {
                const int s = it % kTcStages;
                const uint32_t ph = (it / kTcStages) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                trace(it, 0);
                const int tap = it / kchunks, c0 = (it - tap * kchunks) * kTcBK;
                unsigned char *st = tiles + s * kTcStageBytes;
                if ((p.ablate & 4) && it >= kTcStages) { mbar_arrive(&full[s]); continue; }
                mbar_expect_tx(&full[s], kTcTileBytes + 2 * b_tile_bytes);
                const int wtap = p.tap_w[cls][tap];
                tma_load_4d(st, &map_a, &full[s], c0, ox0 * p.in_stride + p.tap_dx[cls][tap], oy0 * p.in_stride + p.tap_dy[cls][tap], b);
                if (CS == 1) {
                    tma_load_4d(st + 2 * kTcTileBytes, &map_b, &full[s], c0, n0, wtap, 0);
                    tma_load_4d(st + 3 * kTcTileBytes, &map_b, &full[s], c0, n0, wtap, 1);
                } else {
                    // this CTA fetches rows [crank*n/CS, (crank+1)*n/CS) of both weight planes and multicasts them to the cluster
                    const int rows = p.n_tile / CS;
                    const uint32_t so = crank * (uint32_t)rows * 128u;
                    tma_load_4d_mc(st + 2 * kTcTileBytes + so, &map_b, &full[s], c0, n0 + (int)crank * rows, wtap, 0, kMask);
                    tma_load_4d_mc(st + 3 * kTcTileBytes + so, &map_b, &full[s], c0, n0 + (int)crank * rows, wtap, 1, kMask);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc_tf32(kTcBM, p.n_tile);
        for (int it = 0; it < steps; ++it) {
            const int s = it % kTcStages;
            const uint32_t ph = (it / kTcStages) & 1;
            mbar_wait(&full[s], ph);
            if (it == 0 && lane == 0) stamp(2);            // first TMA stage landed
            if (lane == 0) trace(it, 3);
            mbar_wait(&split[s], ph);
            tc_fence_after();
            if (lane == 0) trace(it, 4);
            if (lane == 0) {
                const uint32_t a_hi = smem_u32(tiles + s * kTcStageBytes);
                const uint32_t a_lo = a_hi + kTcTileBytes, b_hi = a_hi + 2 * kTcTileBytes, b_lo = a_hi + 3 * kTcTileBytes;
#pragma unroll
                for (int k = 0; k < kTcBK / 8; ++k) {          // UMMA_K = 8 tf32 = 32 bytes inside the 128-byte swizzle row
                    const uint32_t ko = k * 32;
                    const uint64_t dah = make_sw128_desc(a_hi + ko), dal = make_sw128_desc(a_lo + ko);
                    const uint64_t dbh = make_sw128_desc(b_hi + ko), dbl = make_sw128_desc(b_lo + ko);
                    const uint32_t acc_small = tmem_base + 3 * (uint32_t)p.n_tile;
                    const uint32_t acc_main = tmem_base + (uint32_t)(it % 3) * (uint32_t)p.n_tile;
                    if (!(p.ablate & 2)) {
                        tc_mma_tf32(acc_small, dal, dbh, idesc, (it | k) != 0);
                        tc_mma_tf32(acc_small, dah, dbl, idesc, 1);
                    }
                    tc_mma_tf32(acc_main, dah, dbh, idesc, (it >= 3 || k != 0) ? 1u : 0u);
                }
                trace(it, 5);
                if (CS == 1) tc_commit(&empty[s]);                // smem stage reusable once these MMAs retire
                else tc_commit_mc(&empty[s], kMask);              // ... in every CTA of the cluster (peers multicast into it)
                if (it == steps - 1) tc_commit(acc_full);         // accumulator complete
            }
            __syncwarp();
        }
    } else {
        // ===================== split warps (then epilogue) =====================
        const int tid = threadIdx.x - 64;   // 0..127
        for (int it = 0; it < steps; ++it) {
            const int s = it % kTcStages;
            const uint32_t ph = (it / kTcStages) & 1;
            mbar_wait(&full[s], ph);
            float4 *a = reinterpret_cast<float4 *>(tiles + s * kTcStageBytes);
            float4 *lo = reinterpret_cast<float4 *>(tiles + s * kTcStageBytes + kTcTileBytes);
            if (tid == 0) trace(it, 1);
            if (p.ablate & 1) { mbar_arrive(&split[s]); continue; }
#pragma unroll
            for (int j = 0; j < kTcTileBytes / 16 / 128; ++j) {    // 8 x 16-byte chunks per thread; layout agnostic
                const int i = tid + j * 128;
                const float4 v = a[i];
                float4 h, l;
                if (p.ablate & 16) {        // experiment: does the tensor core round tf32 operands to nearest?
                    uint32_t u0, u1, u2, u3;
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u0) : "f"(v.x)); asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u1) : "f"(v.y));
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u2) : "f"(v.z)); asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u3) : "f"(v.w));
                    h.x = __uint_as_float(u0); h.y = __uint_as_float(u1); h.z = __uint_as_float(u2); h.w = __uint_as_float(u3);
                } else {
                    h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                    h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                }
                l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
                if (!(p.ablate & 24)) a[i] = h;     // modes 8 / 16 leave the RAW fp32 value as the "hi" operand
                lo[i] = l;
            }
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy writes -> visible to tcgen05 (async proxy)
            if (tid == 0) trace(it, 2);
            mbar_arrive(&split[s]);
        }
        // ---- epilogue: TMEM -> registers -> BN/ReLU/residual -> global ----
        mbar_wait(acc_full, 0);
        tc_fence_after();
        if (threadIdx.x == 64) stamp(3);            // accumulators complete
        const int q = warp & 3;                     // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;                // accumulator row == pixel within the patch
        const int gy = oy0 + r / kTcTileW, gx = ox0 + r % kTcTileW;
        const bool pix_ok = b < p.batch && gy < p.grid_h && gx < p.grid_w;
        const size_t opix = (((size_t)b * p.out_h + (size_t)gy * p.out_stride + p.cls_off_y[cls]) * p.out_w + (size_t)gx * p.out_stride + p.cls_off_x[cls]);
        const int nmain = steps < 3 ? steps : 3;     // main accumulators that were written at least once
        for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
            uint32_t v[32], u[32];
            const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            tmem_ld_32x32b_x32(lane_base, v);                                       // main 0
            tmem_ld_32x32b_x32(lane_base + 3 * (uint32_t)p.n_tile, u);              // cross terms
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
            for (int m = 1; m < nmain; ++m) {
                tmem_ld_32x32b_x32(lane_base + (uint32_t)m * (uint32_t)p.n_tile, u);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
            }
            if (!pix_ok) continue;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const int n = n0 + c0 + j;
                if (n >= p.cout) break;
                float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
                if (scale) sc = *reinterpret_cast<const float4 *>(scale + n);
                if (shift) sh = *reinterpret_cast<const float4 *>(shift + n);
                float4 o;
                o.x = fmaf(__uint_as_float(v[j + 0]), sc.x, sh.x); o.y = fmaf(__uint_as_float(v[j + 1]), sc.y, sh.y);
                o.z = fmaf(__uint_as_float(v[j + 2]), sc.z, sh.z); o.w = fmaf(__uint_as_float(v[j + 3]), sc.w, sh.w);
                if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                const size_t off = opix * p.cout + n;
                if (resid) {
                    const float4 rr = *reinterpret_cast<const float4 *>(resid + off);
                    o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
                }
                *reinterpret_cast<float4 *>(out + off) = o;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) stamp(4);        // epilogue done
    if (CS > 1) cluster_sync_all();        // nobody exits while a peer may still multicast into / arrive on this CTA
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512) : "memory");
    }
}


// ================================================================================================================
// Variant 2: A operand in TENSOR MEMORY.  Timeline measurements of the variant above (profiles/) show the main loop is bound by
// shared-memory bandwidth (128 B/clk/SM): per k-step the tensor core fetches 12 x (4 KB A + 4 KB B) of operands from smem, the
// split warps read 16 KB and write 32 KB, TMA writes 48 KB => 192 KB => ~1500 clk, measured 1540.  Here the split warps write
// a_hi / a_lo straight into TMEM (tcgen05.st) and the MMAs take A from TMEM, so smem only carries the TMA writes (48 KB), one read of
// the raw A tile (16 KB) and the B operand fetches (48 KB): 112 KB per k-step.  Stages shrink to 48 KB => 4 stages.
// Issue-rate measurements (scripts/mma_probe.py, profiles/): a tf32 tcgen05.mma with K=8 costs ~96 clk for any N <= 128, 138 clk for N=256
// (A from TMEM).  So the two products sharing a_hi are issued as ONE N=2n instruction against the concatenated weight tile
// [b_hi ; b_lo] (the planes sit back to back in smem; odd steps load them in the order [b_lo ; b_hi]):
//   even step: D=[main0|cross] += a_hi x [b_hi;b_lo]      odd step: D=[cross|main1] += a_hi x [b_lo;b_hi]      + cross += a_lo x b_hi (N=n)
// i.e. 138 + 96 = 234 clk per k-sub-step instead of 3 x 107.
// TMEM columns: [0,N) main0 | [N,2N) cross terms | [2N,3N) main1 | [384,448) A slot 0 (hi 32 | lo 32) | [448,512) A slot 1.
// ================================================================================================================
constexpr int kV2Stages = 4;
constexpr int kV2StageBytes = 3 * kTcTileBytes;        // A raw, B_hi, B_lo
constexpr int kV2SmemBytes = kV2Stages * kV2StageBytes + 1024 + 256;
constexpr uint32_t kV2ACol = 384;
constexpr int kV2Threads = 320;                         // w0 TMA, w1 MMA, w2-5 split group 0 (+ epilogue), w6-9 split group 1

__global__ void __launch_bounds__(kV2Threads, 1) bev_conv_tc2_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                     const __grid_constant__ CUtensorMap map_b,
                                                                     const float *__restrict__ scale, const float *__restrict__ shift,
                                                                     const float *__restrict__ resid, float *__restrict__ out, TcParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *tiles = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = (uint64_t *)(tiles + kV2Stages * kV2StageBytes);
    uint64_t *full = bars, *split = bars + kV2Stages, *empty = bars + 2 * kV2Stages, *a_free = bars + 3 * kV2Stages;   // a_free[2]
    uint64_t *acc_full = bars + 3 * kV2Stages + 2;
    uint32_t *tmem_slot = (uint32_t *)(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int t = blockIdx.x;
    const int tx = t % p.tiles_x; t /= p.tiles_x;
    const int ty = t % p.tiles_y;
    const int b = t / p.tiles_y;
    const int oy0 = ty * kTcTileH, ox0 = tx * kTcTileW;
    const int n0 = blockIdx.y * p.n_tile;
    const int cls = blockIdx.z;
    const int kchunks = p.cin / kTcBK;
    const int steps = p.cls_ntaps[cls] * kchunks;
    const uint32_t b_tile_bytes = (uint32_t)p.n_tile * kTcBK * 4;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kV2Stages; ++s) { mbar_init(&full[s], 1); mbar_init(&split[s], 128); mbar_init(&empty[s], 1); }
        mbar_init(&a_free[0], 1); mbar_init(&a_free[1], 1);
        mbar_init(acc_full, 1);
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
        if (lane == 0) {
            for (int it = 0; it < steps; ++it) {
                const int s = it % kV2Stages;
                const uint32_t ph = (it / kV2Stages) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                const int tap = it / kchunks, c0 = (it - tap * kchunks) * kTcBK;
                const int wtap = p.tap_w[cls][tap];
                unsigned char *st = tiles + s * kV2StageBytes;
                if ((p.ablate & 4) && it >= kV2Stages) { mbar_arrive(&full[s]); continue; }
                mbar_expect_tx(&full[s], kTcTileBytes + 2 * b_tile_bytes);
                tma_load_4d(st, &map_a, &full[s], c0, ox0 * p.in_stride + p.tap_dx[cls][tap], oy0 * p.in_stride + p.tap_dy[cls][tap], b);
                // weight planes back to back as one 2n-row K-major tile: [b_hi ; b_lo] on even steps, [b_lo ; b_hi] on odd steps
                const uint32_t hi_off = (it & 1) ? b_tile_bytes : 0u, lo_off = (it & 1) ? 0u : b_tile_bytes;
                tma_load_4d(st + kTcTileBytes + hi_off, &map_b, &full[s], c0, n0, wtap, 0);
                tma_load_4d(st + kTcTileBytes + lo_off, &map_b, &full[s], c0, n0, wtap, 1);
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc1 = make_idesc_tf32(kTcBM, p.n_tile), idesc2 = make_idesc_tf32(kTcBM, 2 * p.n_tile);
        const uint32_t acc_main0 = tmem_base, acc_cross = tmem_base + (uint32_t)p.n_tile, acc_main1 = tmem_base + 2 * (uint32_t)p.n_tile;
        for (int it = 0; it < steps; ++it) {
            const int s = it % kV2Stages;
            const uint32_t ph = (it / kV2Stages) & 1;
            mbar_wait(&full[s], ph);
            mbar_wait(&split[s], ph);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t b_cat = smem_u32(tiles + s * kV2StageBytes) + kTcTileBytes;     // [hi;lo] (even) or [lo;hi] (odd)
                const uint32_t b_hi = b_cat + ((it & 1) ? b_tile_bytes : 0u);
                const uint32_t a_hi = tmem_base + kV2ACol + (uint32_t)(it & 1) * 64u, a_lo = a_hi + 32u;
#pragma unroll
                for (int k = 0; k < kTcBK / 8; ++k) {
                    const uint64_t dcat = make_sw128_desc(b_cat + k * 32), dbh = make_sw128_desc(b_hi + k * 32);
                    if (!(it & 1)) {
                        // [main0 | cross] (+)= a_hi x [b_hi ; b_lo]; the very first instruction overwrites (zero-initialises) both halves
                        tc_mma_tf32_ts(acc_main0, a_hi + k * 8, dcat, idesc2, (it | k) != 0);
                    } else if (it == 1 && k == 0) {
                        // main1 is written for the first time here (overwrite) while cross must accumulate: two N=n instructions once
                        tc_mma_tf32_ts(acc_cross, a_hi + k * 8, dcat, idesc1, 1);                                  // x b_lo
                        tc_mma_tf32_ts(acc_main1, a_hi + k * 8, dbh, idesc1, 0);                                   // x b_hi
                    } else {
                        tc_mma_tf32_ts(acc_cross, a_hi + k * 8, dcat, idesc2, 1);                                  // [cross | main1] += a_hi x [b_lo ; b_hi]
                    }
                    tc_mma_tf32_ts(acc_cross, a_lo + k * 8, dbh, idesc1, 1);                                       // cross += a_lo x b_hi
                }
                tc_commit(&empty[s]);
                tc_commit(&a_free[it & 1]);
                if (it == steps - 1) tc_commit(acc_full);
            }
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        const int r = q * 32 + lane;                 // the tile row (TMEM lane) this thread owns
        const int grp = (warp - 2) >> 2;             // two split groups alternate k-steps (group g owns TMEM A slot g)
        for (int it = grp; it < steps; it += 2) {
            const int s = it % kV2Stages;
            const uint32_t ph = (it / kV2Stages) & 1;
            mbar_wait(&full[s], ph);
            const unsigned char *a = tiles + s * kV2StageBytes + r * 128;
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {            // row r of the SWIZZLE_128B tile: logical chunk c sits at chunk c ^ (r & 7)
                const float4 v = *reinterpret_cast<const float4 *>(a + ((c ^ (r & 7)) << 4));
                const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t h = __float_as_uint(x[e]) & 0xFFFFE000u;
                    hi[c * 4 + e] = h;
                    lo[c * 4 + e] = __float_as_uint(x[e] - __uint_as_float(h));
                }
            }
            // the A slot (it & 1) was last read by the MMAs of step it-2
            if (it >= 2) mbar_wait(&a_free[it & 1], ((it >> 1) - 1) & 1);
            tc_fence_after();
            const uint32_t a_hi = tmem_base + ((uint32_t)(q * 32) << 16) + kV2ACol + (uint32_t)(it & 1) * 64u;
            tmem_st_32x32b_x32(a_hi, hi);
            tmem_st_32x32b_x32(a_hi + 32u, lo);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&split[s]);
        }
        if (grp == 0) {
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int gy = oy0 + r / kTcTileW, gx = ox0 + r % kTcTileW;
        const bool pix_ok = b < p.batch && gy < p.grid_h && gx < p.grid_w;
        const size_t opix = (((size_t)b * p.out_h + (size_t)gy * p.out_stride + p.cls_off_y[cls]) * p.out_w + (size_t)gx * p.out_stride + p.cls_off_x[cls]);
        for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
            uint32_t v[32], u[32];
            const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            tmem_ld_32x32b_x32(lane_base, v);                                  // main0
            tmem_ld_32x32b_x32(lane_base + (uint32_t)p.n_tile, u);             // cross terms
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
            if (steps > 1) {
                tmem_ld_32x32b_x32(lane_base + 2 * (uint32_t)p.n_tile, u);     // main1
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
            }
            if (!pix_ok) continue;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const int n = n0 + c0 + j;
                if (n >= p.cout) break;
                float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
                if (scale) sc = *reinterpret_cast<const float4 *>(scale + n);
                if (shift) sh = *reinterpret_cast<const float4 *>(shift + n);
                float4 o;
                o.x = fmaf(__uint_as_float(v[j + 0]), sc.x, sh.x); o.y = fmaf(__uint_as_float(v[j + 1]), sc.y, sh.y);
                o.z = fmaf(__uint_as_float(v[j + 2]), sc.z, sh.z); o.w = fmaf(__uint_as_float(v[j + 3]), sc.w, sh.w);
                if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                const size_t off = opix * p.cout + n;
                if (resid) {
                    const float4 rr = *reinterpret_cast<const float4 *>(resid + off);
                    o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
                }
                *reinterpret_cast<float4 *>(out + off) = o;
            }
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512) : "memory");
}


// ================================================================================================================
// Variant 3 (hybrid): measured facts (scripts/tc_debug.py E1, scripts/mma_probe.py, per-step traces in profiles/):
//   * the tensor core TRUNCATES raw fp32 bits to tf32, so a_hi needs no copy at all: the raw TMA tile IS the a_hi operand;
//   * variants 1/2 run at ~0.8 us per k-step although their MMA issue time is 0.70 / 0.51 us: the ring is too shallow for the round trip
//     MMA retire -> empty -> TMA (~1 us) -> split -> MMA (variant 1: 3 x 64 KB stages; variant 2: only 2 TMEM slots for a_hi|a_lo).
// Here only a_lo goes to TMEM (32 columns per step => FOUR slots next to the three accumulators), a_hi products are SS-form with the
// concatenated [b_hi;b_lo] tile (N = 2n), stages are 48 KB => four stages: 4 steps of work in flight.
// ================================================================================================================
constexpr int kV3Stages = 4;
constexpr int kV3StageBytes = 3 * kTcTileBytes;        // A raw, B_hi, B_lo
constexpr int kV3SmemBytes = kV3Stages * kV3StageBytes + 1024 + 256;
constexpr uint32_t kV3ACol = 384;
constexpr int kV3Threads = 320;                         // w0 TMA, w1 MMA, w2-5 split group 0 (+ epilogue), w6-9 split group 1

__global__ void __launch_bounds__(kV3Threads, 1) bev_conv_tc3_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                     const __grid_constant__ CUtensorMap map_b,
                                                                     const float *__restrict__ scale, const float *__restrict__ shift,
                                                                     const float *__restrict__ resid, float *__restrict__ out, TcParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *tiles = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = (uint64_t *)(tiles + kV3Stages * kV3StageBytes);
    uint64_t *full = bars, *split = bars + kV3Stages, *empty = bars + 2 * kV3Stages, *a_free = bars + 3 * kV3Stages;   // a_free[4]
    uint64_t *acc_full = bars + 3 * kV3Stages + 4;
    uint32_t *tmem_slot = (uint32_t *)(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int t = blockIdx.x;
    const int tx = t % p.tiles_x; t /= p.tiles_x;
    const int ty = t % p.tiles_y;
    const int b = t / p.tiles_y;
    const int oy0 = ty * kTcTileH, ox0 = tx * kTcTileW;
    const int n0 = blockIdx.y * p.n_tile;
    const int cls = blockIdx.z;
    const int kchunks = p.cin / kTcBK;
    const int steps = p.cls_ntaps[cls] * kchunks;
    const uint32_t b_tile_bytes = (uint32_t)p.n_tile * kTcBK * 4;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kV3Stages; ++s) { mbar_init(&full[s], 1); mbar_init(&split[s], 4); mbar_init(&empty[s], 1); }
        for (int q4 = 0; q4 < 4; ++q4) mbar_init(&a_free[q4], 1);
        mbar_init(acc_full, 1);
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
        if (lane == 0) {
            for (int it = 0; it < steps; ++it) {
                const int s = it % kV3Stages;
                const uint32_t ph = (it / kV3Stages) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                const int tap = it / kchunks, c0 = (it - tap * kchunks) * kTcBK;
                const int wtap = p.tap_w[cls][tap];
                unsigned char *st = tiles + s * kV3StageBytes;
                if ((p.ablate & 4) && it >= kV3Stages) { mbar_arrive(&full[s]); continue; }
                mbar_expect_tx(&full[s], kTcTileBytes + 2 * b_tile_bytes);
                tma_load_4d(st, &map_a, &full[s], c0, ox0 * p.in_stride + p.tap_dx[cls][tap], oy0 * p.in_stride + p.tap_dy[cls][tap], b);
                // weight planes back to back as one 2n-row K-major tile: [b_hi ; b_lo] on even steps, [b_lo ; b_hi] on odd steps
                const uint32_t hi_off = (it & 1) ? b_tile_bytes : 0u, lo_off = (it & 1) ? 0u : b_tile_bytes;
                tma_load_4d(st + kTcTileBytes + hi_off, &map_b, &full[s], c0, n0, wtap, 0);
                tma_load_4d(st + kTcTileBytes + lo_off, &map_b, &full[s], c0, n0, wtap, 1);
            }
        }
    } else if (warp == 1) {
        // The MMA stream is issued by ONE thread: its scalar instruction count per tcgen05.mma is the real ceiling (measured: ~100 clk per
        // MMA with per-instruction descriptor construction).  So: 4-stage ring fully unrolled (stage == it & 3, and with four stages
        // the parity of `it` equals the parity of the stage), descriptors precomputed per stage, only "+ 2k" on the low word per k sub-step.
        const uint32_t idesc1 = make_idesc_tf32(kTcBM, p.n_tile), idesc2 = make_idesc_tf32(kTcBM, 2 * p.n_tile);
        const uint32_t acc_main0 = tmem_base, acc_cross = tmem_base + (uint32_t)p.n_tile, acc_main1 = tmem_base + 2 * (uint32_t)p.n_tile;
        const uint64_t desc_hi = ((uint64_t)(((1024u >> 4)) | (1u << 14) | (2u << 29))) << 32;          // SBO | version | SWIZZLE_128B
        const uint32_t tiles_lo = ((smem_u32(tiles) >> 4) & 0x3FFFu) | (1u << 16);                        // start address field + LBO
        const uint32_t b_rows_lo = b_tile_bytes >> 4;
        uint64_t dA[kV3Stages], dCat[kV3Stages], dBhi[kV3Stages];
        uint32_t aLo[kV3Stages];
#pragma unroll
        for (int sgi = 0; sgi < kV3Stages; ++sgi) {
            const uint32_t st_lo = tiles_lo + (uint32_t)sgi * (kV3StageBytes >> 4);
            dA[sgi] = desc_hi | st_lo;
            dCat[sgi] = desc_hi | (st_lo + (kTcTileBytes >> 4));
            dBhi[sgi] = desc_hi | (st_lo + (kTcTileBytes >> 4) + ((sgi & 1) ? b_rows_lo : 0u));
            aLo[sgi] = tmem_base + kV3ACol + (uint32_t)sgi * 32u;
        }
        auto issue = [&](auto stage_c, int it) {
            constexpr int S = decltype(stage_c)::value;
            const uint32_t ph = (it / kV3Stages) & 1;
            mbar_wait(&full[S], ph);
            mbar_wait(&split[S], ph);
            tc_fence_after();
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < kTcBK / 8; ++k) {
                    const uint64_t da = dA[S] + 2 * k, dcat = dCat[S] + 2 * k, dbh = dBhi[S] + 2 * k;
                    if ((S & 1) == 0) {
                        tc_mma_tf32(acc_main0, da, dcat, idesc2, (it | k) != 0);         // [main0 | cross] (+)= a_hi x [b_hi ; b_lo]
                    } else if (it == 1 && k == 0) {
                        tc_mma_tf32(acc_cross, da, dcat, idesc1, 1);                     // cross += a_hi x b_lo
                        tc_mma_tf32(acc_main1, da, dbh, idesc1, 0);                      // main1  = a_hi x b_hi (first write)
                    } else {
                        tc_mma_tf32(acc_cross, da, dcat, idesc2, 1);                     // [cross | main1] += a_hi x [b_lo ; b_hi]
                    }
                    tc_mma_tf32_ts(acc_cross, aLo[S] + k * 8, dbh, idesc1, 1);           // cross += a_lo x b_hi   (a_lo from TMEM)
                }
                tc_commit(&empty[S]);
                tc_commit(&a_free[S]);
                if (it == steps - 1) tc_commit(acc_full);
            }
            __syncwarp();
        };
        for (int it = 0; it < steps; it += kV3Stages) {
            issue(std::integral_constant<int, 0>{}, it);
            if (it + 1 < steps) issue(std::integral_constant<int, 1>{}, it + 1);
            if (it + 2 < steps) issue(std::integral_constant<int, 2>{}, it + 2);
            if (it + 3 < steps) issue(std::integral_constant<int, 3>{}, it + 3);
        }
    } else {
        const int q = warp & 3;
        const int r = q * 32 + lane;                 // the tile row (TMEM lane) this thread owns
        const int grp = (warp - 2) >> 2;             // two split groups alternate k-steps (group g owns TMEM A slot g)
        for (int it = grp; it < steps; it += 2) {
            const int s = it % kV3Stages;
            const uint32_t ph = (it / kV3Stages) & 1;
            mbar_wait(&full[s], ph);
            const uint32_t a = smem_u32(tiles) + (uint32_t)(s * kV3StageBytes + r * 128);
            uint32_t lo[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {            // row r of the SWIZZLE_128B tile: logical chunk c sits at chunk c ^ (r & 7)
                const float4 v = lds128(a + (uint32_t)((c ^ (r & 7)) << 4));
                const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) lo[c * 4 + e] = __float_as_uint(x[e] - __uint_as_float(__float_as_uint(x[e]) & 0xFFFFE000u));
            }
            // the a_lo slot (it & 3) was last read by the MMAs of step it-4
            if (it >= 4) mbar_wait(&a_free[it & 3], ((it >> 2) - 1) & 1);
            tc_fence_after();
            tmem_st_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + kV3ACol + (uint32_t)(it & 3) * 32u, lo);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&split[s]);       // one arrival per warp (128 per-thread arrivals serialise on the barrier)
        }
        if (grp == 0) {
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int gy = oy0 + r / kTcTileW, gx = ox0 + r % kTcTileW;
        const bool pix_ok = b < p.batch && gy < p.grid_h && gx < p.grid_w;
        const size_t opix = (((size_t)b * p.out_h + (size_t)gy * p.out_stride + p.cls_off_y[cls]) * p.out_w + (size_t)gx * p.out_stride + p.cls_off_x[cls]);
        for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
            uint32_t v[32], u[32];
            const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            tmem_ld_32x32b_x32(lane_base, v);                                  // main0
            tmem_ld_32x32b_x32(lane_base + (uint32_t)p.n_tile, u);             // cross terms
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
            if (steps > 1) {
                tmem_ld_32x32b_x32(lane_base + 2 * (uint32_t)p.n_tile, u);     // main1
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
            }
            if (!pix_ok) continue;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const int n = n0 + c0 + j;
                if (n >= p.cout) break;
                float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
                if (scale) sc = *reinterpret_cast<const float4 *>(scale + n);
                if (shift) sh = *reinterpret_cast<const float4 *>(shift + n);
                float4 o;
                o.x = fmaf(__uint_as_float(v[j + 0]), sc.x, sh.x); o.y = fmaf(__uint_as_float(v[j + 1]), sc.y, sh.y);
                o.z = fmaf(__uint_as_float(v[j + 2]), sc.z, sh.z); o.w = fmaf(__uint_as_float(v[j + 3]), sc.w, sh.w);
                if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                const size_t off = opix * p.cout + n;
                if (resid) {
                    const float4 rr = *reinterpret_cast<const float4 *>(resid + off);
                    o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
                }
                *reinterpret_cast<float4 *>(out + off) = o;
            }
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512) : "memory");
}

static int g_conv_variant = 3;     // 1: A operand from shared memory, 2: A operand from tensor memory, 3: hybrid (default, fastest)
static int g_conv_cluster = 1;
static int g_conv_ablate = 0;
static long long *g_conv_dbg = nullptr;

}  // namespace sessd

using namespace sessd;

static int launch_tc(const float *d_in, const float *d_w, int w_taps, int cout_pad, const float *d_scale, const float *d_shift,
                     const float *d_residual, float *d_out, TcParams &p, void *stream) {
    const int n_tile = p.cout <= 32 ? 32 : 128;
    if (cout_pad % n_tile || cout_pad < p.cout) return SESSD_EINVAL;
    const int cs = (g_conv_variant >= 2) ? 1 : (g_conv_cluster == 8 || g_conv_cluster == 4 || g_conv_cluster == 2) ? g_conv_cluster : 1;
    CUtensorMap map_a, map_b;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)p.cin, (cuuint64_t)p.in_w, (cuuint64_t)p.in_h, (cuuint64_t)p.batch};
        // stride-2 conv: the box TRAVERSES 2x the tile extent with element stride 2 => still 16 x 8 pixels land in smem
        const cuuint32_t box[4] = {kTcBK, (cuuint32_t)(kTcTileW * p.in_stride), (cuuint32_t)(kTcTileH * p.in_stride), 1};
        const cuuint32_t estr[4] = {1, (cuuint32_t)p.in_stride, (cuuint32_t)p.in_stride, 1};
        int rc = encode_map_4d(&map_a, d_in, dims, box, estr);
        if (rc) return rc;
    }
    {
        const cuuint64_t dims[4] = {(cuuint64_t)p.cin, (cuuint64_t)cout_pad, (cuuint64_t)w_taps, 2};
        const cuuint32_t box[4] = {kTcBK, (cuuint32_t)(n_tile / cs), 1, 1};
        int rc = encode_map_4d(&map_b, d_w, dims, box);
        if (rc) return rc;
    }
    static bool attr_done = false;
    if (!attr_done) {
        SESSD_CUDA_TRY(cudaFuncSetAttribute(bev_conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
        SESSD_CUDA_TRY(cudaFuncSetAttribute(bev_conv_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
        SESSD_CUDA_TRY(cudaFuncSetAttribute(bev_conv_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
        SESSD_CUDA_TRY(cudaFuncSetAttribute(bev_conv_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
        attr_done = true;
    }
    p.n_tile = n_tile;
    p.ablate = g_conv_ablate;
    p.dbg = g_conv_dbg;
    p.tiles_x = div_up(p.grid_w, kTcTileW);
    p.tiles_y = div_up(p.grid_h, kTcTileH);
    const int tiles = p.tiles_x * p.tiles_y * p.batch;
    dim3 grid(div_up(tiles, cs) * cs, cout_pad / n_tile, p.nclass);   // padded tiles decode to image index >= batch: masked
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = kTcSmemBytes;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e;
    if (g_conv_variant == 3) {
        static bool attr3 = false;
        if (!attr3) { SESSD_CUDA_TRY(cudaFuncSetAttribute(bev_conv_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kV3SmemBytes)); attr3 = true; }
        cfg.gridDim = dim3(tiles, cout_pad / n_tile, p.nclass);
        cfg.dynamicSmemBytes = kV3SmemBytes;
        cfg.blockDim = dim3(kV3Threads);
        attr[0].val.clusterDim.x = 1;
        e = cudaLaunchKernelEx(&cfg, bev_conv_tc3_kernel, map_a, map_b, d_scale, d_shift, d_residual, d_out, p);
    } else if (g_conv_variant == 2) {
        static bool attr2 = false;
        if (!attr2) { SESSD_CUDA_TRY(cudaFuncSetAttribute(bev_conv_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kV2SmemBytes)); attr2 = true; }
        cfg.gridDim = dim3(tiles, cout_pad / n_tile, p.nclass);
        cfg.dynamicSmemBytes = kV2SmemBytes;
        cfg.blockDim = dim3(kV2Threads);
        attr[0].val.clusterDim.x = 1;
        e = cudaLaunchKernelEx(&cfg, bev_conv_tc2_kernel, map_a, map_b, d_scale, d_shift, d_residual, d_out, p);
    } else if (cs == 8) e = cudaLaunchKernelEx(&cfg, bev_conv_tc_kernel<8>, map_a, map_b, d_scale, d_shift, d_residual, d_out, p);
    else if (cs == 4) e = cudaLaunchKernelEx(&cfg, bev_conv_tc_kernel<4>, map_a, map_b, d_scale, d_shift, d_residual, d_out, p);
    else if (cs == 2) e = cudaLaunchKernelEx(&cfg, bev_conv_tc_kernel<2>, map_a, map_b, d_scale, d_shift, d_residual, d_out, p);
    else e = cudaLaunchKernelEx(&cfg, bev_conv_tc_kernel<1>, map_a, map_b, d_scale, d_shift, d_residual, d_out, p);
    ++g_launches;
    if (e != cudaSuccess) return (int)e;
    return last_error();
}

// Tensor-core variant of sessd_bev_conv.  d_weight_split: [2 (hi, lo)][ntaps][cout_pad][cin] with cout_pad a multiple of the N tile
// (128, or 32 when cout <= 32); hi = tf32-truncated weights, lo = w - hi.
extern "C" int sessd_bev_conv_tc(const float *d_in, const float *d_weight_split, int cout_pad, const float *d_scale, const float *d_shift,
                                 const float *d_residual, float *d_out, const sessd_conv_desc *desc, void *stream) {
    if (!d_in || !d_weight_split || !d_out || !desc) return SESSD_EINVAL;
    const sessd_conv_desc &d = *desc;
    if (d.batch < 1 || d.cin < kTcBK || d.cin % kTcBK || d.cout < 4 || d.cout % 4 || d.ntaps < 1 || d.ntaps > 9 || d.in_stride < 1 || d.in_stride > 2 ||
        d.out_stride < 1 || d.grid_h < 1 || d.grid_w < 1)
        return SESSD_EINVAL;
    if ((d.grid_h - 1) * d.out_stride + d.out_off_y >= d.out_h || (d.grid_w - 1) * d.out_stride + d.out_off_x >= d.out_w) return SESSD_EINVAL;
    TcParams p = {};
    p.batch = d.batch; p.in_h = d.in_h; p.in_w = d.in_w; p.cin = d.cin;
    p.out_h = d.out_h; p.out_w = d.out_w; p.cout = d.cout;
    p.grid_h = d.grid_h; p.grid_w = d.grid_w;
    p.out_stride = d.out_stride;
    p.relu = d.relu;
    p.in_stride = d.in_stride;
    p.nclass = 1;
    p.cls_ntaps[0] = d.ntaps; p.cls_off_y[0] = d.out_off_y; p.cls_off_x[0] = d.out_off_x;
    for (int t = 0; t < d.ntaps; ++t) { p.tap_dy[0][t] = d.tap_dy[t]; p.tap_dx[0][t] = d.tap_dx[t]; p.tap_w[0][t] = t; }
    return launch_tc(d_in, d_weight_split, d.ntaps, cout_pad, d_scale, d_shift, d_residual, d_out, p, stream);
}

// ConvTranspose2d(kernel 3, stride 2, padding 1, output_padding 1) + BN + ReLU (+ residual) in ONE launch: the four output-parity
// classes (1/2/2/4 taps, no multiplications by the zero-stuffed grid) run as blockIdx.z over the same input and the same weight tensor.
// d_weight_split: [2][9][cout_pad][cin] with tap index ky*3+kx of the transposed-conv kernel W[cin][cout][ky][kx].
// Output is [batch, 2*in_h, 2*in_w, cout] NHWC.   (det3d/models/necks/rpn_v1.py:183-195)
extern "C" int sessd_bev_deconv_tc(const float *d_in, const float *d_weight_split, int cout_pad, const float *d_scale, const float *d_shift,
                                   const float *d_residual, float *d_out, int batch, int in_h, int in_w, int cin, int cout, int relu,
                                   void *stream) {
    if (!d_in || !d_weight_split || !d_out || batch < 1 || in_h < 1 || in_w < 1 || cin < kTcBK || cin % kTcBK || cout < 4 || cout % 4) return SESSD_EINVAL;
    TcParams p = {};
    p.batch = batch; p.in_h = in_h; p.in_w = in_w; p.cin = cin;
    p.out_h = 2 * in_h; p.out_w = 2 * in_w; p.cout = cout;
    p.grid_h = in_h; p.grid_w = in_w;
    p.out_stride = 2;
    p.relu = relu;
    p.in_stride = 1;
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
                for (int b = 0; b < nx; ++b) {
                    p.tap_dy[c][t] = dys[a]; p.tap_dx[c][t] = dxs[b]; p.tap_w[c][t] = kys[a] * 3 + kxs[b];
                    ++t;
                }
            p.cls_ntaps[c] = t;
        }
    return launch_tc(d_in, d_weight_split, 9, cout_pad, d_scale, d_shift, d_residual, d_out, p, stream);
}

// tunable: CTAs per cluster that share (TMA-multicast) the weight tiles of sessd_bev_conv_tc: 1, 2 or 4
extern "C" void sessd_set_conv_cluster(int cs) { sessd::g_conv_cluster = cs; }
extern "C" int sessd_get_conv_cluster(void) { return sessd::g_conv_cluster; }
// negative values select timing-ablation modes of the conv kernel (profiling experiments only)
extern "C" void sessd_set_conv_variant(int v) { sessd::g_conv_variant = v; }
extern "C" void sessd_set_conv_ablate(int m) { sessd::g_conv_ablate = m; }
extern "C" void sessd_set_conv_debug_buffer(void *d_buf) { sessd::g_conv_dbg = (long long *)d_buf; }
