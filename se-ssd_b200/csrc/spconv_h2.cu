// spconv_h2.cu -- sparse 3-D convolution (SubM / strided, + folded BN + ReLU) on the tcgen05 tensor cores with the gather done by the
// TMA engine (cp.async.bulk.tensor ... tile::gather4) and a TWO-TERM FP16 SPLIT of every fp32 operand.
//
// Replaces spconv 1.x's per-offset gather -> sgemm -> scatter-add used by det3d/models/backbones/scn.py:106-149 (SpMiddleFHD), like
// spconv_tc.cu (3xTF32, SIMT gather warps), but:
//   * features travel between layers as fp16 (hi, lo) planes  planes[row][2][CP]  (x = 2^-s (hi + lo), s from the tensor's abs-max: the
//     same exact power-of-two scaling as bevconv_h2.cu, 22+ significand bits): one row = [hi CP | lo CP] halves = 4*CP bytes, the
//     same bytes as the fp32 row.  sessd_split_h2 produces them from the producing layer's fp32 output;
//   * the 128 input rows of a (tile, kernel offset) are fetched by 32 (CP=32) or 64 (CP=64) TMA gather4 instructions -- four 128-byte rows
//     each, written by the copy engine straight into the K-major SWIZZLE_128B layout the UMMA descriptors address.  Missing neighbours
//     are passed as the out-of-bounds row index `zero_row` = -1: the TMA zero-fills them without memory traffic (reading a real
//     all-zero row instead serialises every SM on two L2 lines: 1.46 ms vs 0.34 ms on 300 k rows at 50 % fill).  No SIMT warp touches
//     the operands: no register staging, no st.shared, no split arithmetic in the main loop;
//   * kind::f16 MMAs (twice the tf32 rate, half the operand bytes in shared memory): a_hi*b_hi -> main0 / main1 (alternating per
//     offset), a_hi*b_lo + a_lo*b_hi -> cross; the three TMEM accumulators are summed in RN fp32 by the epilogue (tcgen05
//     accumulation truncates, see bevconv_tc.cu);
//   * 96-114 KB of shared memory and <= 256 TMEM columns per CTA => TWO CTAs per SM: one tile's prologue / epilogue overlaps the
//     other's main loop without a persistent scheduler, and kernels of other streams can share the SM (the one-CTA-per-SM deep variant
//     costs 9 % of the multi-stream frame throughput).
// Bound (measured): the per-SM TMA row rate, ~5-6 clk per gathered 128-byte row whether it is in bounds or not (~25 B/clk/SM, ~7 TB/s
// over the chip = the L2 -> SM fabric).  At 4 bytes per gathered element and Cout = 64 that caps the tensor pipe near 31 % (ncu: 29-30 %).
// One CTA = 128 consecutive output rows x all Cout; 160 threads: warps 0-3 issue the gathers of their 32 rows (a TMA instruction is
// issued by one elected lane at a time, ~40 clk each: four warps issue in parallel), then run the epilogue (TMEM -> registers -> fp32
// rows + running abs-max of the output for the next layer's split); warp 4 loads the weight tiles and issues the MMAs.
//
// CP = 64 (Cin 64): A stage = hi tile + lo tile (2 x 16 KB); B stage = [b_hi ; b_lo] rows (128 x 128 B, [b_lo ; b_hi] on odd steps) so
//   that ONE N=2*Cout MMA yields main and the a_hi*b_lo cross term in adjacent accumulators; a second N=Cout MMA adds a_lo*b_hi.
// CP = 32 (Cin 32, or Cin 16 zero-padded): the whole [hi | lo] row is ONE 128-byte K-major line, the weight row is [b_hi | b_lo];
//   three K=32 MMAs per offset address the half-lines through the descriptor start address.
#include <cuda_fp16.h>

#include <type_traits>

#include "tc_common.cuh"

namespace sessd {

constexpr int kG4Threads = 160;       // warps 0-3: TMA gather producers, then epilogue; warp 4: weight TMA + MMA issue
constexpr int kG4BM = 128;
constexpr int kG4MaxK = 27;

// DEEP = 0: 2 (wide) / 4 (narrow) stages, <= 113 KB => two CTAs per SM (many tiles: throughput); DEEP = 1: 4 / 8 stages, one CTA per SM
// (few tiles, e.g. one frame: the serial chain over the kernel offsets is latency-bound, more gathers in flight shorten it)
template <int CP, int COUT, int DEEP>
struct G4Cfg {
    static constexpr bool kWide = (CP == 64);
    static constexpr int kATile = (kWide ? 2 : 1) * kG4BM * 128;              // bytes
    static constexpr int kBTile = (kWide ? 2 : 1) * COUT * 128;
    static constexpr int kStage = kATile + (kBTile + 1023) / 1024 * 1024;
    static constexpr int kStages = (kWide ? 2 : 4) * (DEEP ? 2 : 1);
    static constexpr int kSmem = kStages * kStage + kG4BM * kG4MaxK * 4 + 1024 + 512;
    static constexpr int kTmemCols = (3 * COUT <= 64) ? 64 : ((3 * COUT <= 128) ? 128 : 256);
};

__host__ __device__ constexpr uint32_t g4_idesc_f16(int M, int N) {
    return (1u << 4) /*C=F32*/ | (0u << 7) /*A=F16*/ | (0u << 10) /*B=F16*/ | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (K = 16 per instruction = 32 bytes of each K-major row)
__device__ __forceinline__ void g4_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// four rows (row indices i0..i3, column c0) of a 2-D tensor -> four consecutive 128-byte lines of shared memory
__device__ __forceinline__ void tma_gather4(uint32_t smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int i0, int i1, int i2, int i3) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n" ::"r"(
            smem_dst),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(i0), "r"(i1), "r"(i2), "r"(i3)
        : "memory");
}

__device__ __forceinline__ void g4_tmem_ld16(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// exact power-of-two activation scale 2^s that maps abs-max into [2^10, 2^11) (1 when abs-max is 0 / not finite); same rule as
// bevconv_h2.cu so that both consumers of a tensor agree
__device__ __forceinline__ float g4_act_scale(float amax) {
    const uint32_t e = (__float_as_uint(amax) >> 23) & 0xFFu;
    if (e == 0 || e == 255) return 1.f;
    int bits = 264 - (int)e;
    bits = bits < 2 ? 2 : (bits > 252 ? 252 : bits);
    return __uint_as_float((uint32_t)bits << 23);
}

template <int CP, int COUT, int DEEP>
__global__ void __launch_bounds__(kG4Threads, DEEP ? 1 : 2) spconv_h2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                                                                   const int *__restrict__ nbr, int kvol, const int *__restrict__ d_n_out,
                                                                   int max_out, int zero_row, const float *__restrict__ amax_in,
                                                                   const float *__restrict__ scale, const float *__restrict__ shift, int relu,
                                                                   float *__restrict__ out_feat, float *__restrict__ amax_out) {
    using C = G4Cfg<CP, COUT, DEEP>;
    const int n_out = min(*d_n_out, max_out);
    const int row0 = blockIdx.x * kG4BM;
    if (row0 >= n_out) return;                       // whole CTA leaves together
    const int rows = min(kG4BM, n_out - row0);

    extern __shared__ unsigned char smem_raw[];
    unsigned char *tiles = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    int *s_nbr = (int *)(tiles + C::kStages * C::kStage);               // [128][kvol]
    uint64_t *bars = (uint64_t *)(s_nbr + kG4BM * kG4MaxK);
    uint64_t *full = bars, *empty = bars + C::kStages, *acc_full = bars + 2 * C::kStages;
    uint32_t *tmem_slot = (uint32_t *)(acc_full + 1);
    int *s_klist = (int *)(tmem_slot + 2);
    int *s_nact = s_klist + 32;
    unsigned int *s_kmask = (unsigned int *)(s_nact + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < C::kStages; ++s) { mbar_init(&full[s], C::kWide ? 1 : 4); mbar_init(&empty[s], 1); }   // narrow: one arrival per producer warp
        mbar_init(acc_full, 1);
        *s_kmask = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(C::kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    __syncthreads();
    // stage this tile's neighbour table (contiguous in global memory) and find the occupied kernel offsets
    unsigned int mymask = 0;
    const int nent = kG4BM * kvol;
    for (int e = tid; e < nent; e += kG4Threads) {
        const int r = e / kvol, k = e - r * kvol;
        int v = -1;
        if (r < rows) v = nbr[(size_t)row0 * kvol + e];
        if (v >= 0) mymask |= 1u << k; else v = zero_row;
        s_nbr[e] = v;
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
    const uint32_t tiles_u32 = smem_u32(tiles);

    if (warp == 4) {
        // ===================== MMA issue (one thread) =====================
        const uint32_t idesc = g4_idesc_f16(kG4BM, COUT);
        const uint32_t idesc2 = g4_idesc_f16(kG4BM, 2 * COUT);
        const uint64_t desc_hi = ((uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29))) << 32;      // SBO | version | SWIZZLE_128B
        const uint32_t tiles_lo = ((tiles_u32 >> 4) & 0x3FFFu) | (1u << 16);
        const uint32_t acc_main0 = tmem_base, acc_cross = tmem_base + COUT, acc_main1 = tmem_base + 2 * COUT;
        for (int j = 0; j < nact; ++j) {
            const int s = j % C::kStages;
            const uint32_t ph = (j / C::kStages) & 1;
            if (lane == 0) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t st_lo = tiles_lo + (uint32_t)s * (C::kStage >> 4);
                const uint64_t dA = desc_hi | st_lo;
                const uint64_t dB = desc_hi | (st_lo + (C::kATile >> 4));
                if constexpr (C::kWide) {
                    const uint64_t dAl = dA + ((kG4BM * 128) >> 4);
                    const uint64_t dBh = dB + (((j & 1) ? COUT * 128 : 0) >> 4);       // the b_hi rows inside the [X ; Y] tile
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint32_t o = (uint32_t)(kk * 32) >> 4;
                        if ((j & 1) == 0) {
                            g4_mma_f16(acc_main0, dA + o, dB + o, idesc2, (j != 0 || kk != 0) ? 1u : 0u);    // [main0|cross] (+)= a_hi x [b_hi;b_lo]
                        } else if (j == 1 && kk == 0) {
                            g4_mma_f16(acc_cross, dA + o, dB + o, idesc, 1u);                                 // cross += a_hi x b_lo
                            g4_mma_f16(acc_main1, dA + o, dBh + o, idesc, 0u);                                // main1  = a_hi x b_hi
                        } else {
                            g4_mma_f16(acc_cross, dA + o, dB + o, idesc2, 1u);                                // [cross|main1] += a_hi x [b_lo;b_hi]
                        }
                        g4_mma_f16(acc_cross, dAl + o, dBh + o, idesc, 1u);                                   // cross += a_lo x b_hi
                    }
                } else {
                    const uint32_t acc_main = (j & 1) ? acc_main1 : acc_main0;
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const uint32_t o = (uint32_t)(kk * 32) >> 4;
                        g4_mma_f16(acc_main, dA + o, dB + o, idesc, (j >= 2 || kk != 0) ? 1u : 0u);           // main  (+)= a_hi x b_hi
                        g4_mma_f16(acc_cross, dA + o, dB + 4 + o, idesc, (j != 0 || kk != 0) ? 1u : 0u);      // cross (+)= a_hi x b_lo
                        g4_mma_f16(acc_cross, dA + 4 + o, dB + o, idesc, 1u);                                 // cross  += a_lo x b_hi
                    }
                }
                tc_commit(&empty[s]);
                if (j == nact - 1) tc_commit(acc_full);
            }
            __syncwarp();
        }
        if (nact == 0 && lane == 0) mbar_arrive(acc_full);     // isolated tile: nothing to accumulate
    } else {
        // ===================== TMA producers: warp w gathers rows 32w .. 32w+31 of every stage =====================
        // lanes 0-7 issue the (hi | narrow) lines of their four rows, lanes 8-15 the lo lines (wide layout)
        constexpr int kIssuers = C::kWide ? 16 : 8;
        const int g = lane & 7;                                    // gather4 group within the warp's 32 rows
        const int first_row = warp * 32 + g * 4;
        const uint32_t dst_off = (uint32_t)first_row * 128u + ((C::kWide && lane >= 8) ? (uint32_t)(kG4BM * 128) : 0u);
        const int col = (C::kWide && lane >= 8) ? 64 : 0;
        // A gather4 whose four rows are all missing only has to leave zeros behind: it is SKIPPED when the four shared-memory lines of
        // that stage are known to be zero already (each lane tracks its own lines: bit s of `dirty` = stage s holds real data or has
        // never been written), otherwise issued once with out-of-bounds indices (zero fill) to clean them.  Narrow layers only (fill
        // 5-25 %): removes a third of the TMA row requests, the resource that bounds this kernel (stress shape: 12.3 -> 9.8 ms).
        uint32_t dirty = (1u << C::kStages) - 1u;
        for (int j = 0; j < nact; ++j) {
            const int s = j % C::kStages;
            const uint32_t ph = (j / C::kStages) & 1;
            const int k = s_klist[j];
            int i0 = 0, i1 = 0, i2 = 0, i3 = 0;
            bool issue = false;
            if (lane < kIssuers) {
                const int *nb = s_nbr + first_row * kvol + k;
                i0 = nb[0]; i1 = nb[kvol]; i2 = nb[2 * kvol]; i3 = nb[3 * kvol];
                issue = true;
                if constexpr (!C::kWide) {
                    const bool all_missing = (i0 == zero_row) & (i1 == zero_row) & (i2 == zero_row) & (i3 == zero_row);
                    if (!all_missing) dirty |= 1u << s;
                    else if ((dirty >> s) & 1u) dirty &= ~(1u << s);
                    else issue = false;
                }
            }
            if constexpr (C::kWide) {
                // 64-channel layers (fill 35-90 %): skipping rarely triggers and its bookkeeping costs ~4 % -- every gather is issued
                if (lane == 0) {
                    mbar_wait(&empty[s], ph ^ 1);
                    if (warp == 0) mbar_expect_tx(&full[s], C::kATile + C::kBTile);
                }
            } else {
                const uint32_t nissue = __popc(__ballot_sync(0xffffffffu, issue));
                if (lane == 0) {
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], nissue * 512u + (warp == 0 ? (uint32_t)C::kBTile : 0u));      // arrive + this warp's bytes
                }
            }
            __syncwarp();
            if (issue) tma_gather4(tiles_u32 + (uint32_t)(s * C::kStage) + dst_off, &map_a, &full[s], col, i0, i1, i2, i3);
            if (warp == 0 && lane == 0) {
                unsigned char *b_tile = tiles + s * C::kStage + C::kATile;
                if constexpr (C::kWide) {
                    // [b_hi ; b_lo] on even steps, [b_lo ; b_hi] on odd steps (see the MMA issuer)
                    tma_load_4d(b_tile + ((j & 1) ? COUT * 128 : 0), &map_w, &full[s], 0, 0, 0, k);
                    tma_load_4d(b_tile + ((j & 1) ? 0 : COUT * 128), &map_w, &full[s], 0, 0, 1, k);
                } else {
                    tma_load_4d(b_tile, &map_w, &full[s], 0, 0, 0, k);
                }
            }
        }
        // ===================== epilogue (warps 0-3 <-> TMEM lane quadrants 0-3) =====================
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const float inv_act = 1.f / g4_act_scale(amax_in ? *amax_in : 0.f);
        if (lane == 0) mbar_wait(acc_full, 0);
        __syncwarp();
        tc_fence_after();
        const int nmain = nact < 2 ? nact : 2;
        float rmax = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < COUT; c0 += 16) {
            uint32_t acc[16], u[16];
            const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            if (nact > 0) {
                g4_tmem_ld16(lane_base, acc);
                g4_tmem_ld16(lane_base + COUT, u);                      // cross terms
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = __float_as_uint(__uint_as_float(acc[i]) + __uint_as_float(u[i]));
                if (nmain > 1) {
                    g4_tmem_ld16(lane_base + 2 * COUT, u);              // main1
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[i] = __float_as_uint(__uint_as_float(acc[i]) + __uint_as_float(u[i]));
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = 0u;
            }
            if (r < rows) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const int n = c0 + i;
                    const float4 sc = *reinterpret_cast<const float4 *>(scale + n);
                    float4 sh = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (shift) sh = *reinterpret_cast<const float4 *>(shift + n);
                    float4 o;
                    o.x = fmaf(__uint_as_float(acc[i + 0]) * inv_act, sc.x, sh.x); o.y = fmaf(__uint_as_float(acc[i + 1]) * inv_act, sc.y, sh.y);
                    o.z = fmaf(__uint_as_float(acc[i + 2]) * inv_act, sc.z, sh.z); o.w = fmaf(__uint_as_float(acc[i + 3]) * inv_act, sc.w, sh.w);
                    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    rmax = fmaxf(rmax, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
                    *reinterpret_cast<float4 *>(out_feat + (size_t)(row0 + r) * COUT + n) = o;
                }
            }
        }
        if (amax_out) {
            const unsigned m = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(rmax));
            if (lane == 0 && m != 0u) atomicMax(reinterpret_cast<unsigned *>(amax_out), m);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(C::kTmemCols) : "memory");
}

// fp32 rows -> fp16 (hi, lo) planes [row][2][cp] with the tensor-wide power-of-two scale from *amax; one thread per 4 channels
__global__ void __launch_bounds__(256) split_h2_kernel(const float *__restrict__ feat, const int *__restrict__ d_n, int max_rows, int channels,
                                                       const float *__restrict__ amax, __half *__restrict__ planes, int cp) {
    const int n = min(*d_n, max_rows);
    const int c4 = channels >> 2;
    const long long total = (long long)n * c4;
    const float s = g4_act_scale(*amax);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / c4;
        const int c = (int)(i - row * c4) * 4;
        const float4 v = __ldg(reinterpret_cast<const float4 *>(feat + row * channels + c));
        const float x[4] = {v.x * s, v.y * s, v.z * s, v.w * s};
        __align__(8) __half hi[4], lo[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            hi[t] = __float2half_rn(x[t]);
            lo[t] = __float2half_rn(x[t] - __half2float(hi[t]));
        }
        __half *dst = planes + row * (2 * cp) + c;
        *reinterpret_cast<uint2 *>(dst) = *reinterpret_cast<const uint2 *>(hi);
        *reinterpret_cast<uint2 *>(dst + cp) = *reinterpret_cast<const uint2 *>(lo);
    }
}

static int encode_map_2d_f16(CUtensorMap *m, const void *base, cuuint64_t cols, cuuint64_t rows, cuuint32_t box_cols) {
    EncodeTiledFn enc = get_tensor_map_encoder();
    if (!enc) return SESSD_EINVAL;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * 2};
    const cuuint32_t box[2] = {box_cols, 1};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 700 + (int)r;
}

static int g_sp_h2_depth = 0;      // 0 / 1: two CTAs per SM (default), 2: one CTA per SM with twice the stages

template <int CP, int COUT, int DEEP>
static int launch_spconv_h2(const void *planes, int plane_rows, int zero_row, const float *amax_in, const int *nbr, int kvol, const int *d_n,
                            int max_out, const void *w_h2, const float *sc, const float *sh, int relu, float *out, float *amax_out,
                            cudaStream_t st) {
    using C = G4Cfg<CP, COUT, DEEP>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(spconv_h2_kernel<CP, COUT, DEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem);
        if (e != cudaSuccess) return (int)e;
        attr_done = true;
    }
    CUtensorMap map_a, map_w;
    int rc = encode_map_2d_f16(&map_a, planes, 2 * CP, (cuuint64_t)plane_rows, 64);
    if (rc) return rc;
    if (C::kWide) {       // [kvol][2 (hi|lo)][Cout][64]
        const cuuint64_t dims[4] = {64, (cuuint64_t)COUT, 2, (cuuint64_t)kvol};
        const cuuint32_t box[4] = {64, (cuuint32_t)COUT, 1, 1};
        rc = encode_map_4d(&map_w, w_h2, dims, box, nullptr, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2);
    } else {              // [kvol][Cout][b_hi 32 | b_lo 32]
        const cuuint64_t dims[4] = {64, (cuuint64_t)COUT, 1, (cuuint64_t)kvol};
        const cuuint32_t box[4] = {64, (cuuint32_t)COUT, 1, 1};
        rc = encode_map_4d(&map_w, w_h2, dims, box, nullptr, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2);
    }
    if (rc) return rc;
    const int tiles = div_up(max_out, kG4BM);
    SESSD_LAUNCH((spconv_h2_kernel<CP, COUT, DEEP>), tiles, kG4Threads, C::kSmem, st, map_a, map_w, nbr, kvol, d_n, max_out, zero_row, amax_in, sc, sh,
                 relu, out, amax_out);
    return last_error();
}

}  // namespace sessd

using namespace sessd;

extern "C" void sessd_set_sp_h2_depth(int mode) { sessd::g_sp_h2_depth = mode; }

extern "C" int sessd_split_h2(const float *d_feat, const int *d_n, int max_rows, int channels, const float *d_amax, void *d_planes, int cp,
                              void *stream) {
    if (!d_feat || !d_n || !d_amax || !d_planes || max_rows < 1 || channels < 4 || (channels & 3) || cp < channels || (cp != 32 && cp != 64))
        return SESSD_EINVAL;
    SESSD_LAUNCH(split_h2_kernel, persistent_grid((long long)max_rows * (channels >> 2), 256), 256, 0, stream, d_feat, d_n, max_rows, channels,
                 d_amax, (__half *)d_planes, cp);
    return last_error();
}


// d_in_planes: [plane_rows][2][cp] fp16 from sessd_split_h2 (row `zero_row` all zero); d_weight_h2 from the host packer
// (ops.pack_weight_sp_h2): cp = 64: [kvol][2][Cout][64], cp = 32: [kvol][Cout][hi 32 | lo 32].  d_scale must already contain
// bn_scale[c] * 2^-e[c] (the per-channel weight exponent).  Supported (cp, Cout): (32,16), (32,32), (32,64), (64,64).
extern "C" int sessd_spconv_forward_h2(const void *d_in_planes, int cp, int plane_rows, int zero_row, const float *d_amax_in, const int *d_nbr,
                                       int kvol, const int *d_n_out, int max_out, const void *d_weight_h2, int cout, const float *d_scale,
                                       const float *d_shift, int relu, float *d_out_feat, float *d_amax_out, void *stream) {
    if (!d_in_planes || !d_amax_in || !d_nbr || !d_n_out || !d_weight_h2 || !d_scale || !d_out_feat || max_out < 1 || kvol < 1 || kvol > kG4MaxK ||
        plane_rows < 1)
        return SESSD_EINVAL;     // zero_row outside [0, plane_rows) = rely on the TMA's out-of-bounds zero fill (no memory traffic)
    cudaStream_t st = (cudaStream_t)stream;
    // two CTAs per SM by default.  The deep variant (one CTA per SM, twice the stages, 207 KB of shared memory) does not shorten a
    // single layer (the K loop is bound by the per-SM TMA row rate, not by latency) and its footprint keeps other streams' kernels off
    // the SM: 1400 vs 1528 frames/s in the 8-stream batch-1 bench (profiles/r1g_tune_streams_depth.log) -- kept as an explicit option.
    const bool deep = g_sp_h2_depth == 2;
#define SESSD_G4_CASE(CPV, CO)                                                                                                              \
    if (cp == CPV && cout == CO)                                                                                                            \
        return deep ? launch_spconv_h2<CPV, CO, 1>(d_in_planes, plane_rows, zero_row, d_amax_in, d_nbr, kvol, d_n_out, max_out, d_weight_h2, \
                                                   d_scale, d_shift, relu, d_out_feat, d_amax_out, st)                                       \
                    : launch_spconv_h2<CPV, CO, 0>(d_in_planes, plane_rows, zero_row, d_amax_in, d_nbr, kvol, d_n_out, max_out, d_weight_h2, \
                                                   d_scale, d_shift, relu, d_out_feat, d_amax_out, st);
    SESSD_G4_CASE(32, 16)
    SESSD_G4_CASE(32, 32)
    SESSD_G4_CASE(32, 64)
    SESSD_G4_CASE(64, 64)
#undef SESSD_G4_CASE
    return SESSD_EINVAL;
}
