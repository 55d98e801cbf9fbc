// iou3d.cu -- the iou3d_cuda operator family on B200: rotated BEV overlap / IoU / 3-D IoU matrices and the three
// NMS variants with the greedy reduction done ON DEVICE.
//
// Replaces det3d/core/iou3d/src/iou3d.cpp:34-281 and iou3d_kernel.cu:270-365,425-530.  Differences by design:
//   * explicit stream, no cudaMalloc / blocking cudaMemcpy inside (the reference allocates the N x N/64 mask and
//     copies it to the host on the legacy stream, iou3d.cpp:131-142);
//   * only the upper block-triangle of the suppression mask is computed (the reference's host loop never reads
//     the rest, iou3d.cpp:152-156);
//   * greedy scan: 64 rows at a time -- one thread resolves the diagonal word chain, the CTA ORs the kept rows.
// Compiled with -fmad=false (see rotbox.cuh).
#include "common.cuh"
#include "rotbox.cuh"

namespace sessd {

// mode: 0 overlap, 1 iou bev, 2 iou 3d
template <int MODE>
__global__ void __launch_bounds__(256) box_matrix_kernel(const float *__restrict__ a, int n, const float *__restrict__ b, int m,
                                                         float *__restrict__ out) {
    constexpr int W = (MODE == 2) ? 7 : 5;
    const long long total = (long long)n * m;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / m), j = (int)(t - (long long)i * m);
        float ba[W], bb[W];
#pragma unroll
        for (int k = 0; k < W; ++k) { ba[k] = a[(size_t)i * W + k]; bb[k] = b[(size_t)j * W + k]; }
        float v;
        if (MODE == 0) v = rot_overlap5(ba, bb);
        else if (MODE == 1) v = rot_iou_bev(ba, bb);
        else v = rot_iou_3d(ba, bb);
        out[t] = v;
    }
}

__global__ void __launch_bounds__(256) box_aligned_overlap_kernel(const float *__restrict__ a, const float *__restrict__ b, int n,
                                                                  float *__restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float ba[5], bb[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) { ba[k] = a[(size_t)i * 5 + k]; bb[k] = b[(size_t)i * 5 + k]; }
        out[i] = rot_overlap5(ba, bb);
    }
}

// suppression mask, upper block triangle.  mode 0 rot-bev, 1 3d, 2 axis-aligned.  bit (i, j) iff iou > thresh.
template <int MODE>
__global__ void __launch_bounds__(64) nms_mask_kernel(const float *__restrict__ boxes, int n, float thresh,
                                                      unsigned long long *__restrict__ mask) {
    constexpr int W = (MODE == 1) ? 7 : 5;
    const int rb = blockIdx.y, cb = blockIdx.x;
    if (cb < rb) return;
    const int col_blocks = (n + 63) / 64;
    __shared__ float sb[64 * W];
    const int ncol = min(n - cb * 64, 64);
    if ((int)threadIdx.x < ncol)
        for (int k = 0; k < W; ++k) sb[threadIdx.x * W + k] = boxes[(size_t)(cb * 64 + threadIdx.x) * W + k];
    __syncthreads();
    const int i = rb * 64 + threadIdx.x;
    if (i >= n) return;
    float me[W];
#pragma unroll
    for (int k = 0; k < W; ++k) me[k] = boxes[(size_t)i * W + k];
    unsigned long long bits = 0;
    const int start = (rb == cb) ? threadIdx.x + 1 : 0;
    for (int j = start; j < ncol; ++j) {
        float v;
        if (MODE == 0) v = rot_iou_bev(me, sb + j * W);
        else if (MODE == 1) v = rot_iou_3d(me, sb + j * W);
        else v = axis_iou(me, sb + j * W);
        if (v > thresh) bits |= 1ull << j;
    }
    mask[(size_t)i * col_blocks + cb] = bits;
}

// greedy reduction over a (row-major, upper-triangular) mask; keeps at most max_keep rows.
// One CTA.  remv lives in dynamic shared memory (col_blocks words).
__global__ void __launch_bounds__(256) nms_reduce_kernel(const unsigned long long *__restrict__ mask, int n, int max_keep,
                                                         long long *__restrict__ keep, int *__restrict__ num_keep) {
    extern __shared__ unsigned long long remv[];
    __shared__ unsigned long long diag[64];
    __shared__ unsigned long long s_keepbits;
    __shared__ int s_nkeep;
    const int col_blocks = (n + 63) / 64;
    for (int j = threadIdx.x; j < col_blocks; j += blockDim.x) remv[j] = 0;
    if (threadIdx.x == 0) s_nkeep = 0;
    __syncthreads();
    for (int b = 0; b < col_blocks; ++b) {
        const int rows = min(64, n - b * 64);
        if ((int)threadIdx.x < rows) diag[threadIdx.x] = mask[(size_t)(b * 64 + threadIdx.x) * col_blocks + b];
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long cur = remv[b], kb = 0;
            int nk = s_nkeep;
            for (int t = 0; t < rows && nk < max_keep; ++t) {
                if (!((cur >> t) & 1ull)) {
                    keep[nk++] = b * 64 + t;
                    kb |= 1ull << t;
                    cur |= diag[t];
                }
            }
            s_keepbits = kb;
            s_nkeep = nk;
        }
        __syncthreads();
        const unsigned long long kb = s_keepbits;
        if (s_nkeep >= max_keep) break;
        for (int j = b + 1 + threadIdx.x; j < col_blocks; j += blockDim.x) {
            unsigned long long acc = remv[j];
            unsigned long long bits = kb;
            while (bits) {
                const int t = __ffsll((long long)bits) - 1;
                bits &= bits - 1;
                acc |= mask[(size_t)(b * 64 + t) * col_blocks + j];
            }
            remv[j] = acc;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *num_keep = s_nkeep;
}

}  // namespace sessd

using namespace sessd;

extern "C" int sessd_boxes_overlap_bev(const float *d_a, int n, const float *d_b, int m, float *d_out, void *stream) {
    if (n < 0 || m < 0) return SESSD_EINVAL;
    if (n == 0 || m == 0) return SESSD_OK;
    SESSD_LAUNCH((box_matrix_kernel<0>), persistent_grid((long long)n * m, 256), 256, 0, stream, d_a, n, d_b, m, d_out);
    return last_error();
}

extern "C" int sessd_boxes_iou_bev(const float *d_a, int n, const float *d_b, int m, float *d_out, void *stream) {
    if (n < 0 || m < 0) return SESSD_EINVAL;
    if (n == 0 || m == 0) return SESSD_OK;
    SESSD_LAUNCH((box_matrix_kernel<1>), persistent_grid((long long)n * m, 256), 256, 0, stream, d_a, n, d_b, m, d_out);
    return last_error();
}

extern "C" int sessd_boxes_iou3d(const float *d_a, int n, const float *d_b, int m, float *d_out, void *stream) {
    if (n < 0 || m < 0) return SESSD_EINVAL;
    if (n == 0 || m == 0) return SESSD_OK;
    SESSD_LAUNCH((box_matrix_kernel<2>), persistent_grid((long long)n * m, 256), 256, 0, stream, d_a, n, d_b, m, d_out);
    return last_error();
}

extern "C" int sessd_boxes_aligned_overlap_bev(const float *d_a, const float *d_b, int n, float *d_out, void *stream) {
    if (n < 0) return SESSD_EINVAL;
    if (n == 0) return SESSD_OK;
    SESSD_LAUNCH(box_aligned_overlap_kernel, persistent_grid(n, 256), 256, 0, stream, d_a, d_b, n, d_out);
    return last_error();
}

extern "C" size_t sessd_nms_workspace_bytes(int n) {
    if (n < 0) return 0;
    const size_t cb = (size_t)(n + 63) / 64;
    return sizeof(unsigned long long) * ((size_t)n * cb + 64);
}

extern "C" int sessd_nms_sorted(const float *d_boxes, int n, float thresh, int mode, long long *d_keep, int *d_num_keep,
                                void *workspace, size_t workspace_bytes, void *stream) {
    if (n < 0 || mode < 0 || mode > 2 || !d_num_keep) return SESSD_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) { SESSD_CUDA_TRY(cudaMemsetAsync(d_num_keep, 0, sizeof(int), st)); return SESSD_OK; }
    if (workspace_bytes < sessd_nms_workspace_bytes(n) || !workspace) return SESSD_EWORKSPACE;
    const int cb = (n + 63) / 64;
    if ((size_t)cb * 8 > 200 * 1024) return SESSD_ECAPACITY;   // remv[] must fit in shared memory
    unsigned long long *mask = (unsigned long long *)workspace;
    dim3 grid(cb, cb);
    if (mode == 0) SESSD_LAUNCH((nms_mask_kernel<0>), grid, 64, 0, st, d_boxes, n, thresh, mask);
    else if (mode == 1) SESSD_LAUNCH((nms_mask_kernel<1>), grid, 64, 0, st, d_boxes, n, thresh, mask);
    else SESSD_LAUNCH((nms_mask_kernel<2>), grid, 64, 0, st, d_boxes, n, thresh, mask);
    const size_t sm = sizeof(unsigned long long) * cb;
    if (sm > 48 * 1024)
        SESSD_CUDA_TRY(cudaFuncSetAttribute(nms_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    SESSD_LAUNCH(nms_reduce_kernel, 1, 256, sm, st, mask, n, n, d_keep, d_num_keep);
    return last_error();
}
