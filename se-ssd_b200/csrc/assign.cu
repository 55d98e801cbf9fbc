// assign.cu -- IoU target assignment for the SSD head on the GPU (reference row T1).
//
// Replaces the CPU DataLoader-worker path det3d/core/anchor/target_assigner.py:68-136 (TargetAssigner.assign_v2, all GT classes
// collapsed to 1 = `enable_similar_type`, examples/second/configs/config.py:107) -> det3d/core/anchor/target_ops_v2.py:11-126
// (create_target_np) with NearestIouSimilarity (det3d/core/bbox/region_similarity.py:85-98: rbbox2d_to_near_bbox
// box_np_ops.py:354-366 + iou_jit(eps=0) :1007-1046) and second_box_encode (box_np_ops.py:52-113), batched over frames.
//
// Parity object: labels / positive sets are integer-exact.  That needs the reference's rounding of the IoU: numba forms the
// differences in fp32, then `+ eps` (a float64) promotes the rest of the expression to fp64 and the quotient is rounded once to fp32
// -- reproduced here with double arithmetic on fp32 differences (all products of two fp32 values are exact in fp64, so FMA
// contraction cannot change a bit; the file is nevertheless compiled with -fmad=false for the fp32 encode).
//
// Three launches per batch, no host synchronisation, counts stay on the device:
//   assign_colmax_kernel  : per (anchor, gt) IoU; per-GT maximum over all anchors (shared-memory atomicMax on the fp32 bits of the
//                           sparse non-zero overlaps, one global atomicMax per (CTA, gt) that saw an overlap)
//   assign_label_kernel   : recomputes the anchor's row (bit-identical code path), applies forced positives (ties with the per-GT
//                           maximum, target_ops_v2.py:61-69), matched / unmatched thresholds, re-applies the forced positives after
//                           the background rule (:102), encodes the regression targets, counts positives per CTA
//   assign_compact_kernel : ascending-anchor-order compaction of (anchor index, gt index) of the positives (`positive_gt_id`)
// Algorithmic bytes: 28 A (anchors, L2-resident across frames) + 28 M read, 40 A written per frame (labels 4, targets 28, weights 4,
// argmax scratch 4) => 4.8 MB / frame at A = 70 400: HBM-bound streaming, compute = A*M short IoUs.
#include "common.cuh"

namespace sessd {

constexpr int kAsThreads = 256;
constexpr int kAsMaxGt = 1024;

struct NearBox { float x0, y0, x1, y1; };

// rbbox2d_to_near_bbox (box_np_ops.py:354-366) of (x, y, w, l, r), every operation individually rounded in fp32 like numpy
__device__ __forceinline__ NearBox near_bbox(float x, float y, float w, float l, float r) {
    const float kPi = 3.14159274101257324f;                  // float32(np.pi)
    const float kPi4 = 0.785398185253143311f;                // float32(np.pi / 4)
    // limit_period(r, 0.5, pi) = r - floor(r / pi + 0.5) * pi   (box_np_ops.py:619-620)
    const float t = floorf(__fadd_rn(__fdiv_rn(r, kPi), 0.5f));
    const float lim = fabsf(__fsub_rn(r, __fmul_rn(t, kPi)));
    const float dx = (lim > kPi4) ? l : w;                   // swap w / l when the box is closer to "lying"
    const float dy = (lim > kPi4) ? w : l;
    const float hx = __fdiv_rn(dx, 2.f), hy = __fdiv_rn(dy, 2.f);
    NearBox b;
    b.x0 = __fsub_rn(x, hx); b.y0 = __fsub_rn(y, hy);
    b.x1 = __fadd_rn(x, hx); b.y1 = __fadd_rn(y, hy);
    return b;
}

// iou_jit(eps = 0) with the reference's mixed fp32 / fp64 rounding (see file header)
__device__ __forceinline__ float near_iou(const NearBox &a, double a_area, const NearBox &q, double q_area) {
    const float iwf = __fsub_rn(fminf(a.x1, q.x1), fmaxf(a.x0, q.x0));
    if (!(iwf > 0.f)) return 0.f;
    const float ihf = __fsub_rn(fminf(a.y1, q.y1), fmaxf(a.y0, q.y0));
    if (!(ihf > 0.f)) return 0.f;
    const double inter = (double)iwf * (double)ihf;
    const double ua = a_area + q_area - inter;
    return (float)(inter / ua);
}

__device__ __forceinline__ double near_area(const NearBox &b) {
    return (double)__fsub_rn(b.x1, b.x0) * (double)__fsub_rn(b.y1, b.y0);
}

struct GtSmem {
    NearBox box[kAsMaxGt];
    double area[kAsMaxGt];
};

__device__ __forceinline__ void stage_gt(const float *__restrict__ gt, int m, GtSmem *s) {
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        const float *g = gt + (size_t)j * 7;
        const NearBox b = near_bbox(g[0], g[1], g[3], g[4], g[6]);
        s->box[j] = b;
        s->area[j] = near_area(b);
    }
}

__global__ void __launch_bounds__(kAsThreads) assign_colmax_kernel(const float *__restrict__ anchors, int num_anchors,
                                                                   const float *__restrict__ gt, const int *__restrict__ num_gt,
                                                                   int max_gt, int *__restrict__ colmax /*[B, max_gt] fp32 bits*/) {
    extern __shared__ unsigned char smem_raw[];
    GtSmem *s = (GtSmem *)smem_raw;
    int *s_col = (int *)(s + 1);
    const int b = blockIdx.y;
    const int m = min(num_gt[b], max_gt);
    if (m <= 0) return;
    stage_gt(gt + (size_t)b * max_gt * 7, m, s);
    for (int j = threadIdx.x; j < m; j += blockDim.x) s_col[j] = 0;
    __syncthreads();
    const int i = blockIdx.x * kAsThreads + threadIdx.x;
    if (i < num_anchors) {
        const float *a = anchors + (size_t)i * 7;
        const NearBox ab = near_bbox(a[0], a[1], a[3], a[4], a[6]);
        const double aa = near_area(ab);
        for (int j = 0; j < m; ++j) {
            const float ov = near_iou(ab, aa, s->box[j], s->area[j]);
            if (ov > 0.f) atomicMax(&s_col[j], __float_as_int(ov));
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < m; j += blockDim.x)
        if (s_col[j] > 0) atomicMax(&colmax[(size_t)b * max_gt + j], s_col[j]);
}

__global__ void __launch_bounds__(kAsThreads) assign_label_kernel(const float *__restrict__ anchors, int num_anchors,
                                                                  const float *__restrict__ gt, const int *__restrict__ num_gt,
                                                                  int max_gt, const int *__restrict__ colmax, float matched_thr,
                                                                  float unmatched_thr, int *__restrict__ labels,
                                                                  float *__restrict__ targets, float *__restrict__ weights,
                                                                  int *__restrict__ argmax_gt, int *__restrict__ block_pos) {
    extern __shared__ unsigned char smem_raw[];
    GtSmem *s = (GtSmem *)smem_raw;
    float *s_col = (float *)(s + 1);
    __shared__ int s_scan[40];
    const int b = blockIdx.y;
    const int m = min(num_gt[b], max_gt);
    const float *gtb = gt + (size_t)b * max_gt * 7;
    if (m > 0) {
        stage_gt(gtb, m, s);
        for (int j = threadIdx.x; j < m; j += blockDim.x) s_col[j] = __int_as_float(colmax[(size_t)b * max_gt + j]);
    }
    __syncthreads();
    const int i = blockIdx.x * kAsThreads + threadIdx.x;
    int fg = 0;
    if (i < num_anchors) {
        const size_t o = (size_t)b * num_anchors + i;
        const float *a = anchors + (size_t)i * 7;
        int label = 0, arg = 0;                         // no GT: everything is background (target_ops_v2.py:86-87)
        float t[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (m > 0) {
            const NearBox ab = near_bbox(a[0], a[1], a[3], a[4], a[6]);
            const double aa = near_area(ab);
            float best = -1.f;
            bool force = false;
            for (int j = 0; j < m; ++j) {
                const float ov = near_iou(ab, aa, s->box[j], s->area[j]);
                if (ov > best) { best = ov; arg = j; }                 // numpy argmax: first maximum
                force |= (ov > 0.f) && (ov == s_col[j]);               // ties with the GT's best anchor (max 0 => -1: never)
            }
            label = -1;
            if (force || best >= matched_thr) { label = 1; fg = 1; }
            if (best < unmatched_thr) label = 0;
            if (force) label = 1;
            if (fg) {       // second_box_encode(gt[arg], anchor), box_np_ops.py:52-113 (smooth_dim=False, encode_angle_to_vector=False)
                const float *g = gtb + (size_t)arg * 7;
                const float xa = a[0], ya = a[1], za = a[2], wa = a[3], la = a[4], ha = a[5], ra = a[6];
                const float diag = __fsqrt_rn(__fadd_rn(__fmul_rn(la, la), __fmul_rn(wa, wa)));
                t[0] = __fdiv_rn(__fsub_rn(g[0], xa), diag);
                t[1] = __fdiv_rn(__fsub_rn(g[1], ya), diag);
                t[2] = __fdiv_rn(__fsub_rn(g[2], za), ha);
                t[3] = logf(__fdiv_rn(g[3], wa));
                t[4] = logf(__fdiv_rn(g[4], la));
                t[5] = logf(__fdiv_rn(g[5], ha));
                t[6] = __fsub_rn(g[6], ra);
            }
        }
        labels[o] = label;
        weights[o] = label > 0 ? 1.f : 0.f;
        argmax_gt[o] = arg;
#pragma unroll
        for (int k = 0; k < 7; ++k) targets[o * 7 + k] = t[k];
    }
    int tot;
    block_excl_scan(fg, s_scan, &tot);
    if (threadIdx.x == 0) block_pos[(size_t)b * gridDim.x + blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kAsThreads) assign_compact_kernel(const int *__restrict__ labels, const int *__restrict__ argmax_gt,
                                                                    int num_anchors, const int *__restrict__ block_pos,
                                                                    int *__restrict__ pos_anchor, int *__restrict__ pos_gt,
                                                                    int *__restrict__ num_pos) {
    __shared__ int s_scan[40];
    const int b = blockIdx.y;
    // offset of this CTA = positives in the CTAs before it (<= a few hundred counts: one strided pass + block reduction)
    int before = 0;
    for (int k = threadIdx.x; k < (int)blockIdx.x; k += kAsThreads) before += block_pos[(size_t)b * gridDim.x + k];
    int base;
    block_excl_scan(before, s_scan, &base);
    const int i = blockIdx.x * kAsThreads + threadIdx.x;
    const size_t o = (size_t)b * num_anchors + i;
    const int fg = (i < num_anchors) && labels[o] > 0;
    int tot;
    const int ex = block_excl_scan(fg, s_scan, &tot);
    if (fg) {
        pos_anchor[(size_t)b * num_anchors + base + ex] = i;
        pos_gt[(size_t)b * num_anchors + base + ex] = argmax_gt[o];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) num_pos[b] = base + tot;
}

}  // namespace sessd

using namespace sessd;

extern "C" size_t sessd_assign_workspace_bytes(int num_anchors, int batch, int max_gt) {
    if (num_anchors < 0 || batch < 0 || max_gt < 0) return 0;
    const size_t blocks = (size_t)div_up(num_anchors, kAsThreads);
    return sizeof(int) * ((size_t)batch * max_gt + (size_t)batch * num_anchors + (size_t)batch * blocks + 64);
}

extern "C" int sessd_assign_targets(const float *d_anchors, int num_anchors, const float *d_gt_boxes, const int *d_num_gt, int batch,
                                    int max_gt, float matched_thr, float unmatched_thr, int *d_labels, float *d_bbox_targets,
                                    float *d_bbox_outside_weights, int *d_pos_anchor, int *d_pos_gt_id, int *d_num_pos,
                                    void *workspace, size_t workspace_bytes, void *stream) {
    if (num_anchors <= 0 || batch <= 0 || max_gt < 0 || max_gt > kAsMaxGt) return SESSD_EINVAL;
    if (!d_anchors || !d_num_gt || !d_labels || !d_bbox_targets || !d_bbox_outside_weights || !d_pos_anchor || !d_pos_gt_id || !d_num_pos)
        return SESSD_EINVAL;
    if (max_gt > 0 && !d_gt_boxes) return SESSD_EINVAL;
    if (!workspace || workspace_bytes < sessd_assign_workspace_bytes(num_anchors, batch, max_gt)) return SESSD_EWORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = div_up(num_anchors, kAsThreads);
    int *colmax = (int *)workspace;
    int *argmax_gt = colmax + (size_t)batch * max_gt;
    int *block_pos = argmax_gt + (size_t)batch * num_anchors;
    const size_t smem_full = sizeof(GtSmem) + sizeof(int) * kAsMaxGt;      // 28 KB: fixed layout, below the 48 KB default limit
    dim3 grid(blocks, batch);
    if (max_gt > 0) {
        SESSD_CUDA_TRY(cudaMemsetAsync(colmax, 0, sizeof(int) * (size_t)batch * max_gt, st));
        SESSD_LAUNCH(assign_colmax_kernel, grid, kAsThreads, smem_full, st, d_anchors, num_anchors, d_gt_boxes, d_num_gt, max_gt, colmax);
    }
    SESSD_LAUNCH(assign_label_kernel, grid, kAsThreads, smem_full, st, d_anchors, num_anchors, d_gt_boxes, d_num_gt, max_gt, colmax,
                 matched_thr, unmatched_thr, d_labels, d_bbox_targets, d_bbox_outside_weights, argmax_gt, block_pos);
    SESSD_LAUNCH(assign_compact_kernel, grid, kAsThreads, 0, st, d_labels, argmax_gt, num_anchors, block_pos, d_pos_anchor, d_pos_gt_id,
                 d_num_pos);
    return last_error();
}
