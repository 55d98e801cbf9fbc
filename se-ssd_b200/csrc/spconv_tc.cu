// spconv_tc.cu -- sparse 3-D convolution on the tcgen05 tensor cores (3xTF32), output-stationary, fused BN+ReLU.
//
// Tensor-core variant of spconv_gemm_kernel (spconv.cu) for the Cin >= 32 layers of SpMiddleFHD (11 of 14 layers, > 97 % of the
// sparse FLOPs; det3d/models/backbones/scn.py:117-146).  Same contract: out[o,:] = relu(bn(sum_k in[nbr[o,k],:] @ W[k])).
//
// One CTA = 128 output voxels x all Cout.  Per active kernel offset k:
//   * gather warps (2 groups x 4 warps, alternating stages so two offsets' loads are always in flight) read the 128 input rows
//     nbr[o,k] with 16-byte loads (8 lanes per row => 128-byte coalesced segments; zero rows where the neighbour is missing),
//     split every value into tf32 hi / lo parts in registers and store both into shared memory directly in the canonical
//     K-major SWIZZLE_128B layout (chunk ^= row & 7) that the UMMA descriptors address;
//   * the TMA warp streams the pre-split weight tiles W_hi[k], W_lo[k] ([Cout][Cin], K-major) for the stage;
//   * the MMA warp issues, per 8-channel sub-step, ONE N=2*Cout tcgen05.mma of a_hi against the concatenated [b_hi ; b_lo] tile
//     (main and a_hi*b_lo cross term in adjacent TMEM columns) plus one N=Cout MMA for a_lo*b_hi; even / odd offsets alternate between
//     [main0 | cross] and [cross | main1] (weights loaded as [b_lo ; b_hi] on odd steps) so that no accumulator sees more than half of
//     the (truncating) TMEM accumulation steps -- see bevconv_tc.cu.
// Offsets with no neighbour inside the tile are skipped (bitmask built while staging the tile's nbr rows).
// Epilogue: 4 warps read the four TMEM accumulators, add them in RN fp32, apply folded BN + ReLU, store the row once.
// No scatter-add, no atomics on features, no intermediate gather/scatter buffers.
#include <type_traits>

#include "tc_common.cuh"

namespace sessd {

constexpr int kStThreads = 320;          // warps 0-7 gather (+ 0-3 epilogue), 8 MMA, 9 TMA(B)
constexpr int kStBM = 128;
constexpr int kStMaxK = 32;

template <int CIN, int COUT>
struct StCfg {
    static constexpr int kKblk = CIN / 32;                        // 128-byte K blocks per row
    static constexpr int kATile = kStBM * CIN * 4;                // one of a_hi / a_lo
    static constexpr int kBTile = COUT * CIN * 4;                 // one of b_hi / b_lo
    static constexpr int kStage = 2 * kATile + 2 * kBTile;
    static constexpr int kStages = (2 * kStage <= 200 * 1024) ? ((4 * kStage <= 200 * 1024) ? 4 : 2) : 2;
    static constexpr int kSmem = kStages * kStage + kStBM * kStMaxK * 4 + 1024 + 512;
    static constexpr int kTmemCols = (4 * COUT <= 128) ? 128 : 256;
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(kStThreads, 1) spconv_tc_kernel(const float *__restrict__ in_feat, const int *__restrict__ nbr, int kvol,
                                                                   const int *__restrict__ d_n_out, int max_out,
                                                                   const __grid_constant__ CUtensorMap map_w,
                                                                   const float *__restrict__ scale, const float *__restrict__ shift, int relu,
                                                                   float *__restrict__ out_feat) {
    using C = StCfg<CIN, COUT>;
    const int n_out = min(*d_n_out, max_out);
    const int row0 = blockIdx.x * kStBM;
    if (row0 >= n_out) return;                       // whole CTA leaves together
    const int rows = min(kStBM, n_out - row0);

    extern __shared__ unsigned char smem_raw[];
    unsigned char *tiles = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    int *s_nbr = (int *)(tiles + C::kStages * C::kStage);               // [128][kStMaxK]
    uint64_t *bars = (uint64_t *)(s_nbr + kStBM * kStMaxK);
    uint64_t *full_a = bars, *full_b = bars + C::kStages, *empty = bars + 2 * C::kStages, *acc_full = bars + 3 * C::kStages;
    uint32_t *tmem_slot = (uint32_t *)(acc_full + 1);
    int *s_klist = (int *)(tmem_slot + 2);
    int *s_nact = s_klist + kStMaxK;
    unsigned int *s_kmask = (unsigned int *)(s_nact + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < C::kStages; ++s) { mbar_init(&full_a[s], 4); mbar_init(&full_b[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        *s_kmask = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(C::kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    __syncthreads();
    // stage this tile's neighbour rows (contiguous in global memory) and find the occupied kernel offsets
    unsigned int mymask = 0;
    for (int e = tid; e < kStBM * kvol; e += kStThreads) {
        const int r = e / kvol, k = e - r * kvol;
        int v = -1;
        if (r < rows) v = nbr[(size_t)row0 * kvol + e];
        s_nbr[r * kStMaxK + k] = v;
        if (v >= 0) mymask |= 1u << k;
    }
    mymask = __reduce_or_sync(0xffffffffu, mymask);
    if (lane == 0 && mymask) atomicOr(s_kmask, mymask);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        unsigned int m = *s_kmask;
        int c = 0;
        while (m) { const int k = __ffs(m) - 1; m &= m - 1; s_klist[c++] = k; }
        *s_nact = c;
    }
    __syncthreads();
    const int nact = *s_nact;
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 8) {
        // ===================== gather + split warps =====================
        const int grp = warp >> 2;                  // group 0 fills even steps, group 1 odd steps
        const int gt = tid & 127;                   // thread within the group
        const int sub = gt & 7;                     // 16-byte chunk lane within a 128-byte segment
        constexpr int kChunksPerRow = CIN / 4;      // float4 chunks per row (8 per K block)
        constexpr int kPasses = kChunksPerRow / 8;  // K blocks
        for (int j = grp; j < nact; j += 2) {
            const int s = j % C::kStages;
            const uint32_t ph = (j / C::kStages) & 1;
            const int k = s_klist[j];
            mbar_wait(&empty[s], ph ^ 1);
            const uint32_t a_hi = smem_u32(tiles) + (uint32_t)(s * C::kStage);     // explicit shared-space addresses (see lds128)
            const uint32_t a_lo = a_hi + (uint32_t)C::kATile;
            // 16 rows per pass of the group (8 lanes per row), 8 row passes, kPasses K blocks
            float4 v[8][kPasses];
#pragma unroll
            for (int rp = 0; rp < 8; ++rp) {
                const int r = rp * 16 + (gt >> 3);
                int src;
                asm volatile("ld.shared.b32 %0, [%1];\n" : "=r"(src) : "r"(smem_u32(s_nbr) + (uint32_t)((r * kStMaxK + k) * 4)));
#pragma unroll
                for (int kb = 0; kb < kPasses; ++kb) {
                    if (src >= 0) v[rp][kb] = __ldg(reinterpret_cast<const float4 *>(in_feat + (size_t)src * CIN + kb * 32 + sub * 4));
                    else v[rp][kb] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int rp = 0; rp < 8; ++rp) {
                const int r = rp * 16 + (gt >> 3);
#pragma unroll
                for (int kb = 0; kb < kPasses; ++kb) {
                    const float4 x = v[rp][kb];
                    float4 h, l;
                    h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); l.x = x.x - h.x;
                    h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); l.y = x.y - h.y;
                    h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); l.z = x.z - h.z;
                    h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); l.w = x.w - h.w;
                    // canonical K-major SWIZZLE_128B: [K block][row][128 B], 16-byte chunk index XOR (row & 7)
                    const int off = kb * (kStBM * 128) + r * 128 + ((sub ^ (r & 7)) << 4);
                    sts128(a_hi + (uint32_t)off, h);
                    sts128(a_lo + (uint32_t)off, l);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_a[s]);      // one arrival per warp (per-thread arrivals serialise on the barrier)
        }
        // ===================== epilogue (warps 0-3) =====================
        if (warp < 4) {
            mbar_wait(acc_full, 0);
            tc_fence_after();
            const int q = warp & 3;
            const int r = q * 32 + lane;
            const int nmain = nact < 2 ? nact : 2;
            for (int c0 = 0; c0 < COUT; c0 += 32) {
                uint32_t acc[32], u[32];
                const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
                if (nact > 0) {
                    tmem_ld_32x32b_x32(lane_base, acc);
                    tmem_ld_32x32b_x32(lane_base + COUT, u);                      // cross terms
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = __float_as_uint(__uint_as_float(acc[i]) + __uint_as_float(u[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = 0u;
                }
                for (int m = 1; m < nmain; ++m) {
                    tmem_ld_32x32b_x32(lane_base + 2 * m * COUT, u);             // main1
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = __float_as_uint(__uint_as_float(acc[i]) + __uint_as_float(u[i]));
                }
                if (r >= rows) continue;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const int n = c0 + i;
                    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (scale) sc = *reinterpret_cast<const float4 *>(scale + n);
                    if (shift) sh = *reinterpret_cast<const float4 *>(shift + n);
                    float4 o;
                    o.x = fmaf(__uint_as_float(acc[i + 0]), sc.x, sh.x); o.y = fmaf(__uint_as_float(acc[i + 1]), sc.y, sh.y);
                    o.z = fmaf(__uint_as_float(acc[i + 2]), sc.z, sh.z); o.w = fmaf(__uint_as_float(acc[i + 3]), sc.w, sh.w);
                    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    *reinterpret_cast<float4 *>(out_feat + (size_t)(row0 + r) * COUT + n) = o;
                }
            }
        }
    } else if (warp == 8) {
        // ===================== MMA issuer =====================
        // single issuing thread: keep its scalar work per tcgen05.mma minimal (descriptors precomputed per stage, ring unrolled)
        const uint32_t idesc = make_idesc_tf32(kStBM, COUT);
        const uint64_t desc_hi = ((uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29))) << 32;      // SBO | version | SWIZZLE_128B
        const uint32_t tiles_lo = ((smem_u32(tiles) >> 4) & 0x3FFFu) | (1u << 16);
        uint64_t dAh[C::kStages], dAl[C::kStages], dBh[C::kStages], dBl[C::kStages];
#pragma unroll
        for (int sgi = 0; sgi < C::kStages; ++sgi) {
            const uint32_t st_lo = tiles_lo + (uint32_t)sgi * (C::kStage >> 4);
            dAh[sgi] = desc_hi | st_lo;
            dAl[sgi] = desc_hi | (st_lo + (C::kATile >> 4));
            dBh[sgi] = desc_hi | (st_lo + ((2 * C::kATile + ((sgi & 1) ? COUT * 128 : 0)) >> 4));   // b_hi rows inside the [X;Y] block
            dBl[sgi] = desc_hi | (st_lo + (2 * C::kATile >> 4));                                       // the concatenated 2*Cout-row tile
        }
        const uint32_t idesc2 = make_idesc_tf32(kStBM, 2 * COUT);
        const uint32_t acc_main0 = tmem_base, acc_cross = tmem_base + COUT, acc_main1 = tmem_base + 2 * COUT;
        auto issue = [&](auto stage_c, int j) {
            constexpr int S = decltype(stage_c)::value;
            const uint32_t ph = (j / C::kStages) & 1;
            mbar_wait(&full_a[S], ph);
            mbar_wait(&full_b[S], ph);
            tc_fence_after();
            if (lane == 0) {
#pragma unroll
                for (int kb = 0; kb < C::kKblk; ++kb) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint32_t ao = (uint32_t)(kb * (kStBM * 128) + kk * 32) >> 4, bo = (uint32_t)(kb * (2 * COUT * 128) + kk * 32) >> 4;
                        const bool first = (kb | kk) == 0;
                        if ((S & 1) == 0) {
                            tc_mma_tf32(acc_main0, dAh[S] + ao, dBl[S] + bo, idesc2, (j != 0 || !first) ? 1u : 0u);   // [main0|cross] (+)= a_hi x [b_hi;b_lo]
                        } else if (j == 1 && first) {
                            tc_mma_tf32(acc_cross, dAh[S] + ao, dBl[S] + bo, idesc, 1u);                             // cross += a_hi x b_lo
                            tc_mma_tf32(acc_main1, dAh[S] + ao, dBh[S] + bo, idesc, 0u);                             // main1  = a_hi x b_hi
                        } else {
                            tc_mma_tf32(acc_cross, dAh[S] + ao, dBl[S] + bo, idesc2, 1u);                            // [cross|main1] += a_hi x [b_lo;b_hi]
                        }
                        tc_mma_tf32(acc_cross, dAl[S] + ao, dBh[S] + bo, idesc, 1u);                                 // cross += a_lo x b_hi
                    }
                }
                tc_commit(&empty[S]);
                if (j == nact - 1) tc_commit(acc_full);
            }
            __syncwarp();
        };
        for (int j = 0; j < nact; j += C::kStages) {
            issue(std::integral_constant<int, 0>{}, j);
            if (j + 1 < nact) issue(std::integral_constant<int, 1>{}, j + 1);
            if constexpr (C::kStages == 4) {
                if (j + 2 < nact) issue(std::integral_constant<int, 2>{}, j + 2);
                if (j + 3 < nact) issue(std::integral_constant<int, 3>{}, j + 3);
            }
        }
        if (nact == 0 && lane == 0) mbar_arrive(acc_full);     // isolated tile: nothing to accumulate
    } else {
        // ===================== TMA producer for the weight tiles =====================
        if (lane == 0) {
            for (int j = 0; j < nact; ++j) {
                const int s = j % C::kStages;
                const uint32_t ph = (j / C::kStages) & 1;
                const int k = s_klist[j];
                mbar_wait(&empty[s], ph ^ 1);
                unsigned char *b_hi = tiles + s * C::kStage + 2 * C::kATile;
                mbar_expect_tx(&full_b[s], 2 * C::kBTile);
#pragma unroll
                for (int kb = 0; kb < C::kKblk; ++kb) {
                    // per K block one 2*Cout-row tile: [b_hi ; b_lo] on even steps, [b_lo ; b_hi] on odd steps
                    unsigned char *blk = b_hi + kb * (2 * COUT * 128);
                    tma_load_4d(blk + ((j & 1) ? COUT * 128 : 0), &map_w, &full_b[s], kb * 32, 0, k, 0);
                    tma_load_4d(blk + ((j & 1) ? 0 : COUT * 128), &map_w, &full_b[s], kb * 32, 0, k, 1);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(C::kTmemCols) : "memory");
}

template <int CIN, int COUT>
static int launch_spconv_tc(const float *in, const int *nbr, int kvol, const int *d_n, int max_out, const float *w_split, const float *sc,
                            const float *sh, int relu, float *out, cudaStream_t st) {
    using C = StCfg<CIN, COUT>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(spconv_tc_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem);
        if (e != cudaSuccess) return (int)e;
        attr_done = true;
    }
    CUtensorMap map_w;
    const cuuint64_t dims[4] = {(cuuint64_t)CIN, (cuuint64_t)COUT, (cuuint64_t)kvol, 2};
    const cuuint32_t box[4] = {32, (cuuint32_t)COUT, 1, 1};
    int rc = encode_map_4d(&map_w, w_split, dims, box);
    if (rc) return rc;
    const int tiles = div_up(max_out, kStBM);
    SESSD_LAUNCH((spconv_tc_kernel<CIN, COUT>), tiles, kStThreads, C::kSmem, st, in, nbr, kvol, d_n, max_out, map_w, sc, sh, relu, out);
    return last_error();
}

}  // namespace sessd

using namespace sessd;

// d_weight_split: [2 (hi|lo)][kvol][Cout][Cin] (K-major), hi = tf32-truncated, lo = w - hi.  Supported (Cin, Cout): (32,32), (32,64), (64,64).
extern "C" int sessd_spconv_forward_tc(const float *d_in_feat, int cin, const int *d_nbr, int kvol, const int *d_n_out, int max_out,
                                       const float *d_weight_split, int cout, const float *d_scale, const float *d_shift, int relu,
                                       float *d_out_feat, void *stream) {
    if (!d_in_feat || !d_nbr || !d_n_out || !d_weight_split || !d_out_feat || max_out < 1 || kvol < 1 || kvol > kStMaxK) return SESSD_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (cin == 32 && cout == 32) return launch_spconv_tc<32, 32>(d_in_feat, d_nbr, kvol, d_n_out, max_out, d_weight_split, d_scale, d_shift, relu, d_out_feat, st);
    if (cin == 32 && cout == 64) return launch_spconv_tc<32, 64>(d_in_feat, d_nbr, kvol, d_n_out, max_out, d_weight_split, d_scale, d_shift, relu, d_out_feat, st);
    if (cin == 64 && cout == 64) return launch_spconv_tc<64, 64>(d_in_feat, d_nbr, kvol, d_n_out, max_out, d_weight_split, d_scale, d_shift, relu, d_out_feat, st);
    return SESSD_EINVAL;
}
