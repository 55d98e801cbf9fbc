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
