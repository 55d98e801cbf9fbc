// mma_probe.cu -- micro-benchmark (profiling aid, not on the product path): sustained tcgen05.mma kind::tf32 issue rate of one CTA per
// SM for M=128, N in {64,128,256}, operands from shared memory (SS) or A from tensor memory (TS), with dependent or rotating accumulators.
#include "tc_common.cuh"

namespace sessd {

__global__ void __launch_bounds__(128, 1) mma_probe_kernel(int n, int iters, int mode, long long *out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *tiles = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint64_t dummy[4];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float *>(tiles)[i] = 0.001f * (i & 255);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int q = 0; q < 4; ++q) mbar_init(&dummy[q], 1 << 20); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 0 && lane == 0) {
        const uint32_t idesc = make_idesc_tf32(128, n);
        const uint32_t a = smem_u32(tiles), b = a + 16384;
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        long long c0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t k = (i & 3) * 32;
            const uint32_t acc = (mode & 2) ? tmem + (uint32_t)((i & 1) * n) : tmem;      // rotate between two accumulators
            if (mode & 1) tc_mma_tf32_ts(acc, tmem + 448 + (i & 3) * 8, make_sw128_desc(b + k), idesc, 1);
            else tc_mma_tf32(acc, make_sw128_desc(a + k), make_sw128_desc(b + k), idesc, 1);
            if ((mode & 4) && (i % 12) == 11) tc_commit(&dummy[(i / 12) & 3]);               // a commit per "k-step" (never completes a phase)
            if ((mode & 8) && (i % 12) == 11) { tc_commit(&dummy[(i / 12) & 3]); tc_commit(&dummy[((i / 12) + 1) & 3]); }
            if ((mode & 16) && (i % 12) == 11) { mbar_try_wait(&dummy[0], 1); tc_fence_after(); __syncwarp(0x1); }
        }
        tc_commit(&bar);
        long long c1 = clock64();
        mbar_wait(&bar, 0);
        long long c2 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (blockIdx.x == 0) { out[0] = c1 - c0; out[1] = c2 - c0; out[2] = (long long)(t1 - t0); }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512) : "memory");
}

}  // namespace sessd

// out[0] = issue cycles, out[1] = cycles until all MMAs retired, out[2] = ns (CTA 0); grid = one CTA per SM
extern "C" int sessd_mma_probe(int n, int iters, int mode, long long *d_out, void *stream) {
    using namespace sessd;
    static bool done = false;
    if (!done) { SESSD_CUDA_TRY(cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); done = true; }
    SESSD_LAUNCH(mma_probe_kernel, kNumSMs, 128, 64 * 1024, stream, n, iters, mode, d_out);
    return last_error();
}


// ---------------------------------------------------------------------------------------------------------------------------------
// kind::f16 SS probe: cycles per tcgen05.mma (M = 128, N = n, K = 16) when `iters` of them are issued back to back by one elected lane in
// warp-uniform control flow (the way the product kernels issue).  mode & 3: 0 = one accumulator (dependent chain), 1 = two, 2 = four
// accumulators in rotation; mode & 4: SWIZZLE_64B descriptors (64-byte rows) instead of SWIZZLE_128B.
namespace sessd {
__global__ void __launch_bounds__(128, 1) mma_probe_f16_kernel(int n, int iters, int mode, long long *out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *tiles = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(tiles)[i] = 0x3c003c00u;   // fp16 1.0 pairs
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const bool sw64 = (mode & 4) != 0;
        const uint64_t desc_hi = sw64 ? ((uint64_t)((512u >> 4) | (1u << 14) | (4u << 29))) << 32 : ((uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29))) << 32;
        const uint32_t a_lo = ((smem_u32(tiles) >> 4) & 0x3FFFu) | (1u << 16), b_lo = (((smem_u32(tiles) + 16384) >> 4) & 0x3FFFu) | (1u << 16);
        const int nacc = 1 << (mode & 3);
        const int kmask = sw64 ? 1 : 3;
        long long c0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t k = (uint32_t)(i & kmask) * 2u;                       // 32-byte K steps inside the swizzled row
            const uint32_t acc = tmem + (uint32_t)((i & (nacc - 1)) * (nacc == 4 ? 128 : 256));
            if (elect_one()) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(acc),
                    "l"(desc_hi | (uint64_t)(a_lo + k)), "l"(desc_hi | (uint64_t)(b_lo + k)), "r"(idesc), "r"(1)
                    : "memory");
            }
            __syncwarp();
        }
        if (elect_one()) tc_commit(&bar);
        __syncwarp();
        long long c1 = clock64();
        mbar_wait(&bar, 0);
        long long c2 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = c1 - c0; out[1] = c2 - c0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512) : "memory");
}
}  // namespace sessd

extern "C" int sessd_mma_probe_f16(int n, int iters, int mode, long long *d_out, void *stream) {
    using namespace sessd;
    if (n < 16 || n > 256 || (n & 15) || ((mode & 3) == 2 && n > 128) || ((mode & 3) == 1 && n > 256) || (mode & 3) == 3) return SESSD_EINVAL;
    static bool done = false;
    if (!done) { SESSD_CUDA_TRY(cudaFuncSetAttribute(mma_probe_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); done = true; }
    SESSD_LAUNCH(mma_probe_f16_kernel, kNumSMs, 128, 64 * 1024, stream, n, iters, mode, d_out);
    return last_error();
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Latency probe for the producer/consumer handshakes of the tensor-core kernels (one CTA, clock64 cycles averaged over `iters`):
//   out[0] tcgen05.commit (nothing outstanding) -> mbarrier phase observed by the committing thread
//   out[1] mbarrier ping-pong between two warps (arrive -> other warp's try_wait returns -> arrive back): cycles per round trip
//   out[2] tcgen05.st 32x32b.x32 + tcgen05.wait::st
//   out[3] one kind::f16 TS MMA (M128 N256 K16) + commit -> phase observed
//   out[4] four such MMAs + commit -> phase observed
//   out[5] tcgen05.ld 32x32b.x32 + wait::ld
//   out[6] commit -> phase observed by ANOTHER warp that then arrives back (commit-based ping-pong round trip)
namespace sessd {
__global__ void __launch_bounds__(128, 1) latency_probe_kernel(int iters, long long *out) {
    __shared__ __align__(1024) unsigned char btile[32 * 1024];
    __shared__ uint64_t bars[4];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int i = threadIdx.x; i < 32 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(btile)[i] = 0u;
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const uint64_t bdesc = make_sw128_desc(smem_u32(btile));
    const uint32_t idesc = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // kind::f16, fp16 x fp16 -> fp32, M128 N256
    uint32_t regs[32];
    for (int i = 0; i < 32; ++i) regs[i] = 0u;
    if (warp == 0) {
        tmem_st_32x32b_x32(tmem_base + 384, regs);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t ph = 0;
    // [0] commit -> own wait
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) { tc_commit(&bars[0]); mbar_wait(&bars[0], ph); ph ^= 1; }
        out[0] = (clock64() - t0) / iters;
    }
    __syncthreads();
    // [1] mbarrier ping-pong warp0 <-> warp1
    if (lane == 0 && warp < 2) {
        long long t0 = clock64();
        uint32_t p1 = 0;
        for (int i = 0; i < iters; ++i) {
            if (warp == 0) { mbar_arrive(&bars[1]); mbar_wait(&bars[2], p1); }
            else { mbar_wait(&bars[1], p1); mbar_arrive(&bars[2]); }
            p1 ^= 1;
        }
        if (warp == 0) out[1] = (clock64() - t0) / iters;
    }
    __syncthreads();
    // [2] tcgen05.st + wait, [5] tcgen05.ld + wait
    if (warp == 0) {
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) { tmem_st_32x32b_x32(tmem_base + 384, regs); tmem_st_wait(); }
        long long t1 = clock64();
        for (int i = 0; i < iters; ++i) tmem_ld_32x32b_x32(tmem_base, regs);
        long long t2 = clock64();
        if (lane == 0) { out[2] = (t1 - t0) / iters; out[5] = (t2 - t1) / iters + (regs[0] & 0); }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // [3] / [4] MMA(s) + commit -> wait
    if (threadIdx.x == 0) {
        for (int nm = 1; nm <= 4; nm += 3) {
            long long t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                for (int k = 0; k < nm; ++k)
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_base),
                        "r"(tmem_base + 384), "l"(bdesc), "r"(idesc), "r"(1)
                        : "memory");
                tc_commit(&bars[0]);
                mbar_wait(&bars[0], ph);
                ph ^= 1;
            }
            out[nm == 1 ? 3 : 4] = (clock64() - t0) / iters;
        }
    }
    __syncthreads();
    // [6] commit by warp 0 -> observed by warp 1 -> plain arrive back -> observed by warp 0
    if (lane == 0 && warp < 2) {
        long long t0 = clock64();
        uint32_t p1 = 0;     // bars[3] (commit target) and bars[2] start fresh parities: bars[2] has completed `iters` phases
        uint32_t p2 = iters & 1;
        for (int i = 0; i < iters; ++i) {
            if (warp == 0) { tc_commit(&bars[3]); mbar_wait(&bars[2], p2); }
            else { mbar_wait(&bars[3], p1); mbar_arrive(&bars[2]); }
            p1 ^= 1; p2 ^= 1;
        }
        if (warp == 0) out[6] = (clock64() - t0) / iters;
    }
    __syncthreads();
    // [8] tcgen05.st x32 + wait::st by warp 1 WHILE warp 0 streams MMAs (does the store wait for the tensor pipe?)  [9] same for tcgen05.ld
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < iters * 4; ++i)
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_base),
                "r"(tmem_base + 384), "l"(bdesc), "r"(idesc), "r"(1)
                : "memory");
        tc_commit(&bars[0]);
        mbar_wait(&bars[0], ph);
        ph ^= 1;
    } else if (warp == 1) {
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) { tmem_st_32x32b_x32(tmem_base + ((uint32_t)32 << 16) + 448, regs); tmem_st_wait(); }
        long long t1 = clock64();
        for (int i = 0; i < iters; ++i) tmem_ld_32x32b_x32(tmem_base + ((uint32_t)32 << 16) + 448, regs);
        long long t2 = clock64();
        if (lane == 0) { out[8] = (t1 - t0) / iters; out[9] = (t2 - t1) / iters + (regs[0] & 0); }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // [10]/[11]/[12]: 64 (N256 -> cols [0,256), N128 -> X) MMA pairs + commit -> wait, per pair: X = [128,256) overlapping the N256
    // destination (the conv kernels' [main|cross] + cross pattern), X = [256,384) disjoint, and [12] = 128 N256-only MMAs per 2
    if (threadIdx.x == 0) {
        const uint32_t idesc128 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int variant = 0; variant < 3; ++variant) {
            long long t0 = clock64();
            for (int i = 0; i < 16; ++i) {
                for (int k = 0; k < 64; ++k) {
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_base),
                        "r"(tmem_base + 384), "l"(bdesc), "r"(idesc), "r"(1)
                        : "memory");
                    if (variant < 2)
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_base + (variant == 0 ? 128u : 256u)),
                            "r"(tmem_base + 400), "l"(bdesc), "r"(idesc128), "r"(1)
                            : "memory");
                }
                tc_commit(&bars[0]);
                mbar_wait(&bars[0], ph);
                ph ^= 1;
            }
            out[10 + variant] = (clock64() - t0) / (16 * 64);
        }
    }
    __syncthreads();
    // [7] eight back-to-back commits on one barrier (count 8) -> phase observed: are commits serialised in the tensor pipe?
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 8);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        uint32_t p8 = 0;
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            for (int k = 0; k < 8; ++k) tc_commit(&bars[0]);
            mbar_wait(&bars[0], p8);
            p8 ^= 1;
        }
        out[7] = (clock64() - t0) / iters;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512) : "memory");
}
}  // namespace sessd

extern "C" int sessd_latency_probe(int iters, long long *d_out, void *stream) {
    using namespace sessd;
    SESSD_LAUNCH(latency_probe_kernel, 1, 128, 0, stream, iters, d_out);
    return last_error();
}
