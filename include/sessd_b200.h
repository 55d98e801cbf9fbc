/*
 * sessd_b200.h -- C ABI of the B200-native SE-SSD per-frame LiDAR hot path
 *                 (voxelise -> sparse 3-D conv encoder -> BEV neck/head -> rotated IoU / NMS).
 *
 * Conventions (all entry points):
 *   - plain C types only; device pointers are owned by the caller (torch's allocator in the Python host);
 *   - every call takes an explicit `stream` (a cudaStream_t passed as void*), launches asynchronously on it
 *     and performs NO hidden synchronisation or allocation, so a whole frame can be captured in a CUDA graph;
 *   - data-dependent sizes (voxel / active-site / candidate counts) live in DEVICE memory: kernels take
 *     `const int* d_count` plus a host-side capacity and are launched as persistent grids sized from the SM count;
 *   - return value 0 on success, negative SESSD_E* on argument / capacity errors, positive cudaError_t otherwise
 *     (no exit(), unlike the reference's CHECK_ERROR macro, det3d/core/iou3d/src/iou3d.cpp:13-21).
 *
 * Each function cites the reference interface (file:line under Vegeta2020/SE-SSD) it replaces.
 */
#ifndef SESSD_B200_H
#define SESSD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SESSD_OK 0
#define SESSD_EINVAL (-1)
#define SESSD_ECAPACITY (-2)
#define SESSD_EWORKSPACE (-3)

/* library / build information: returns e.g. "sessd_b200 0.1 sm_100a" */
const char *sessd_version(void);
/* number of kernel launches issued through this library since load (bench.py's gpu_launches claim) */
long long sessd_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * V1/V2/V4 + R1: voxeliser.
 * Replaces det3d/core/input/voxel_generator.py:10-32 (VoxelGenerator.generate) ->
 * det3d/ops/point_cloud/point_cloud_ops_v2.py:120-194 (points_to_voxel) and :9-62 (numba kernel),
 * batched over frames in the wire format of det3d/torchie/parallel/collate.py:154-218 (concatenated
 * voxels, coordinates with a leading batch column), with the per-voxel mean of
 * det3d/models/readers/voxel_encoder.py:205-210 (VoxelFeatureExtractorV3) fused into the same pass.
 * Results are bit-identical to the sequential reference loop (voxel id = rank of the voxel's first point in
 * input order; first max_points points kept in input order; everything after the first point that would open
 * voxel #max_voxels is dropped).
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    float voxel_size[3];      /* x, y, z */
    float range_min[3];       /* x, y, z */
    float range_max[3];
    int grid[3];              /* x, y, z cells = round((max-min)/voxel_size) in fp32 (voxel_generator.py:15-16) */
    int max_points;           /* per voxel (config: 5) */
    int max_voxels;           /* per frame (config: 20000) */
    int num_feat;             /* floats per point (4) */
} sessd_voxel_cfg;

size_t sessd_voxelize_workspace_bytes(int max_total_points, int batch, const sessd_voxel_cfg *cfg);

/* d_points [max_total_points, num_feat] f32; d_frame_off [batch+1] i32 (device; frame f owns points
 * [off[f], off[f+1])); outputs: d_voxels [batch*max_voxels, max_points, num_feat] (zero padded),
 * d_coors [batch*max_voxels, 4] i32 (b, z, y, x), d_num_points [batch*max_voxels] i32,
 * d_mean [batch*max_voxels, num_feat] f32 (nullable), d_num_voxels [batch+1] i32 (per frame; last = total). */
int sessd_voxelize(const float *d_points, const int *d_frame_off, int batch, int max_total_points,
                   const sessd_voxel_cfg *cfg, float *d_voxels, int *d_coors, int *d_num_points, float *d_mean,
                   int *d_num_voxels, void *workspace, size_t workspace_bytes, void *stream);

/* Host-buffer convenience for the numpy-level API (VoxelGenerator.generate): one frame, host in / host out,
 * allocation + H2D + D2H inside, synchronous.  Returns the voxel count (>=0) or a negative error. */
int sessd_voxelize_host(const float *h_points, int num_points, const sessd_voxel_cfg *cfg, float *h_voxels,
                        int *h_coors_zyx /*[max_voxels,3]*/, int *h_num_points);

/* ------------------------------------------------------------------------------------------------
 * S3: rulebook ("indice pairs") construction.  Replaces spconv 1.x's indice-pair builders that
 * det3d/models/backbones/scn.py:106-149,182-183 trigger (4 SubM keys + 4 strided convs per forward).
 * The rulebook is kept output-major: nbr[o, k] = input row feeding output o through kernel offset k, or -1.
 * The canonical spconv form (per offset: pairs sorted by output index) is sessd_rulebook_pairs().
 *
 * An "index" is either a hash table over given coordinates (any row order: the voxeliser's first-appearance
 * order at level 0) or a rank bitmap (rows in ascending linear index: what strided convs emit).
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    int batch;
    int shape[3];     /* D, H, W  (z, y, x) */
} sessd_grid;

/* hash index: table of `capacity` (power of two >= 2*max_rows) 64-bit slots */
size_t sessd_hash_bytes(int max_rows, int *capacity_out);
int sessd_hash_build(const int *d_coors /*[n,4] b,z,y,x*/, const int *d_n, int max_rows, sessd_grid grid,
                     uint64_t *d_table, int capacity, void *stream);

/* rank bitmap index: uint2 {bits, exclusive prefix popcount} per 32 cells; + scan scratch */
size_t sessd_bitmap_words(sessd_grid grid);
size_t sessd_scan_scratch_bytes(size_t n_items);

/* SubM rulebook over an index: nbr [max_rows, kvol] */
int sessd_subm_rulebook(const int *d_coors, const int *d_n, int max_rows, sessd_grid grid, const int ksize[3],
                        int index_kind /*0 hash, 1 bitmap*/, const void *d_index, int hash_capacity,
                        int *d_nbr, void *stream);

/* Strided sparse conv rulebook: marks reachable outputs in d_out_bitmap (zeroed by the call), ranks them
 * (ascending linear index == canonical order), emits out coords + count and the [max_out, kvol] nbr table. */
int sessd_strided_rulebook(const int *d_in_coors, const int *d_n_in, int max_in, sessd_grid in_grid,
                           int in_index_kind, const void *d_in_index, int in_hash_capacity,
                           const int ksize[3], const int stride[3], const int padding[3],
                           sessd_grid out_grid, void *d_out_bitmap /*uint2[words]*/, void *d_scan_scratch,
                           int *d_out_coors, int *d_n_out, int max_out, int *d_nbr, int *d_status, void *stream);

/* nbr table -> per-tile pair lists for sessd_spconv_forward_cg: one record of sessd_tile_list_stride(kvol) uint32 per 128 output rows:
 * [0,32) pair count per kernel offset, [32, 32 + 4 kvol) 128-bit row mask per offset, [160, ...) the pairs grouped by offset, each
 * (input row << 7) | tile row, ascending tile row (deterministic).  d_tiles: uint32 [ceil(max_out / 128)][stride].  kvol <= 27. */
int sessd_tile_list_stride(int kvol);
int sessd_rulebook_tile_lists(const int *d_nbr, int kvol, const int *d_n_out, int max_out, void *d_tiles, void *stream);
/* canonical spconv-style pairs from a nbr table: pairs_in/out [kvol, max_rows], pair_num [kvol] */
size_t sessd_rulebook_pairs_workspace_bytes(int max_rows, int kvol);
int sessd_rulebook_pairs(const int *d_nbr, const int *d_n_out, int max_rows, int kvol, int *d_pairs_in,
                         int *d_pairs_out, int *d_pair_num, void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * S4/S5: sparse convolution (gather -> GEMM -> fused BN(eval)+ReLU epilogue, output-stationary) and dense().
 * Replaces spconv.SubMConv3d / SparseConv3d forward + BatchNorm1d + ReLU (scn.py:106-149) and
 * SparseConvTensor.dense() + view (scn.py:184-187).
 * weight layout [kvol, Cin, Cout] (== spconv 1.x [kz,ky,kx,Cin,Cout] flattened); scale/shift = folded BN
 * (scale = gamma/sqrt(var+eps), shift = beta - mean*scale), nullable => identity; relu flag.
 * ------------------------------------------------------------------------------------------------ */
int sessd_spconv_forward(const float *d_in_feat, int cin, const int *d_nbr, int kvol, const int *d_n_out,
                         int max_out, const float *d_weight, int cout, const float *d_scale, const float *d_shift,
                         int relu, float *d_out_feat, void *stream);

/* dense(): out NHWC [batch, H, W, C*D] with channel index c*D + d (zero-filled by the call) */
int sessd_sparse_to_dense(const float *d_feat, const int *d_coors, const int *d_n, int max_rows, int channels,
                          sessd_grid grid, float *d_out, void *stream);
/* same result in one gather pass (no memset + scatter): the rows of every BEV cell are looked up in the level's bitmap index
 * (the `d_out_bitmap` that sessd_strided_rulebook filled for this level); writes each output byte exactly once. */
int sessd_sparse_to_dense_indexed(const float *d_feat, int max_rows, const void *d_bitmap_index, int channels, sessd_grid grid,
                                  float *d_out, void *stream);

/* S4, narrow layers (Cin <= 32; csrc/spconv_rows.cu): same contract and arguments as sessd_spconv_forward, fp32 SIMT, but the work is
 * proportional to the number of rulebook PAIRS instead of N_out x kvol row slots (a warp owns 8 output rows and visits only their valid
 * neighbours).  d_amax_out (nullable) receives the running abs-max of the output.  (Cin, Cout): (4,16) (16,16) (16,32) (32,32) (32,64). */
int sessd_spconv_forward_rows(const float *d_in_feat, int cin, const int *d_nbr, int kvol, const int *d_n_out, int max_out,
                              const float *d_weight, int cout, const float *d_scale, const float *d_shift, int relu,
                              float *d_out_feat, float *d_amax_out, void *stream);
/* ... and the output also (d_out_feat nullable: only) as fp16 (hi, lo) planes [row][2][cpo] for the tensor-core layers: d_out_info =
 * {abs-max of the output (atomicMax; zero it once per frame), plane scale S_out}; S_out = the power of two that maps the bound
 * *d_amax_in * gain + shift_max into [2^14, 2^15) (d_amax_in = abs-max of the INPUT features, gain = max_n sum_{k,c} |w[k][c][n] bn_scale[n]|,
 * shift_max = max_n |shift[n]|). */
int sessd_spconv_forward_rows_planes(const float *d_in_feat, int cin, const int *d_nbr, int kvol, const int *d_n_out, int max_out,
                                     const float *d_weight, int cout, const float *d_scale, const float *d_shift, int relu,
                                     const float *d_amax_in, float gain, float shift_max, float *d_out_feat, void *d_out_planes, int cpo,
                                     float *d_out_info, void *stream);
/* S4, the default tensor-core path of the Cin >= 32 layers (csrc/spconv_cg.cu): pair-proportional operand traffic -- the rows of the
 * neighbours that exist are copied by cp.async into the UMMA operand tiles (no row slot is spent on a missing neighbour), persistent CTAs,
 * tcgen05 kind::f16 with the two-term fp16 split of spconv_h2.cu (same weight tiles: ops.pack_weight_sp_h2).  Same contract as
 * sessd_spconv_forward (spconv 1.x gather -> GEMM -> scatter-add at det3d/models/backbones/scn.py:106-149, + folded BN + ReLU).
 * The rulebook is passed as d_tiles = the per-tile pair lists sessd_rulebook_tile_lists() makes from the nbr table (once per rulebook).
 * d_in_planes [plane_rows][2][cp] fp16, x = (hi + lo) / d_in_info[1], d_in_info[0] = abs-max of the input tensor; outputs (each nullable, at
 * least one): fp32 rows [max_out][cout]; planes [>= max_out][2][cout <= 32 ? 32 : 64] with d_out_info = {abs-max of the output (atomicMax; zero
 * it once per frame), S_out}, S_out from the bound d_in_info[0] * gain + shift_max as above.  Supported (cp, cout): (32,32) (32,64) (64,64). */
int sessd_spconv_forward_cg(const void *d_in_planes, int cp, int plane_rows, const float *d_in_info, const void *d_tiles, int kvol,
                            const int *d_n_out, int max_out, const void *d_weight_h2, int cout, const float *d_scale, const float *d_shift,
                            int relu, float gain, float shift_max, float *d_out_f32, void *d_out_planes, float *d_out_info, void *stream);
/* 1: deep pipeline (twice the stages, one CTA per SM) for launches with fewer tiles than SMs (single frames); 0 (default): two CTAs per SM */
void sessd_set_sp_cg_deep(int on);
/* *d_amax = max(*d_amax, max |d_feat[i]|) over the first *d_n rows of a [max_rows, channels] fp32 tensor */
int sessd_absmax_rows(const float *d_feat, const int *d_n, int max_rows, int channels, float *d_amax, void *stream);

/* ------------------------------------------------------------------------------------------------
 * N1/H1: BEV neck (SSFA) + head.  Replaces the cuDNN conv/deconv + BatchNorm2d + ReLU blocks of
 * det3d/models/necks/rpn_v1.py:135-235 and the four 1x1 convs of
 * det3d/models/bbox_heads/mg_head_sessd.py:202-230.  Activations are NHWC fp32.
 * One call = one tap-list convolution:  out[b, oy*os+py, ox*os+px, :] =
 *    epilogue( sum_t in[b, oy*is + dy[t], ox*is + dx[t], :] @ W[t] )   (zero outside the input)
 * with epilogue y = acc*scale + shift (nullable), optional ReLU, optional residual add AFTER the ReLU
 * (rpn_v1.py:225: deconv_block_0(x_trans_1) + x_trans_0).
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    int batch, in_h, in_w, cin;       /* input tensor  [batch, in_h, in_w, cin]  */
    int out_h, out_w, cout;           /* output tensor [batch, out_h, out_w, cout] */
    int grid_h, grid_w;               /* output positions computed by this call (per batch) */
    int in_stride;                    /* is */
    int out_stride, out_off_y, out_off_x;   /* os, py, px */
    int ntaps;
    int tap_dy[16], tap_dx[16];
    int relu;
} sessd_conv_desc;

int sessd_bev_conv(const float *d_in, const float *d_weight /*[ntaps, cin, cout]*/, const float *d_scale,
                   const float *d_shift, const float *d_residual /*nullable, same shape as out*/, float *d_out,
                   const sessd_conv_desc *desc, void *stream);

/* ------------------------------------------------------------------------------------------------
 * BEV convs from PRE-SPLIT fp16 planes (csrc/bevconv_p2.cu): the neck's default path.  Activations travel between layers as
 * __half [2 (hi|lo)][batch][H][W][C] planes with x = (hi + lo) / S, S an exact power of two; every plane tensor has a device-side
 * info pair float[2] = {abs-max of the tensor (atomically raised by its producer; zero it once per frame), S}.  Same conv semantics
 * as sessd_bev_conv / sessd_bev_deconv_tc (rpn_v1.py:135-210, mg_head_sessd.py:202-230): in_stride 1 or 2, <= 9 taps, cin % 64 == 0,
 * cout % 8 == 0.  d_weight_h2 / d_scale as for sessd_bev_conv_h2.  gain / shift_max: |out| <= amax_in * gain + shift_max
 * (+ amax of the residual) with gain = max_n sum_{tap,c} |w[tap][c][n] * bn_scale[n]|: the producer derives the OUTPUT scale from
 * this bound before it writes the first element.  Outputs: fp32 NHWC (d_out_f32) and / or planes (d_out_planes + d_out_info). */
int sessd_bev_conv_p2(const void *d_in_planes, const float *d_in_info, const void *d_weight_h2, int cout_pad, const float *d_scale,
                      const float *d_shift, const float *d_residual, const float *d_resid_info, float gain, float shift_max,
                      float *d_out_f32, void *d_out_planes, float *d_out_info, const sessd_conv_desc *desc, void *stream);
int sessd_bev_deconv_p2(const void *d_in_planes, const float *d_in_info, const void *d_weight_h2, int cout_pad, const float *d_scale,
                        const float *d_shift, const float *d_residual, const float *d_resid_info, float gain, float shift_max,
                        float *d_out_f32, void *d_out_planes, float *d_out_info, int batch, int in_h, int in_w, int cin, int cout,
                        int relu, void *stream);
/* fp32 [n] -> planes [2][n] scaled from d_info[0] (the tensor's abs-max, e.g. from sessd_absmax); writes the scale to d_info[1] */
int sessd_bev_split_planes(const float *d_x, long long n, float *d_info, void *d_planes, void *stream);
/* sessd_bev_conv_p2 / sessd_bev_deconv_p2: 0 (default) = CTA pairs (tcgen05 cta_group::2: one MMA spans two SMs, each CTA stages half of
 * every weight tile) for the layers with long K loops, single CTAs for the 1x1 convs; 1 / 2 = force single CTAs / pairs */
void sessd_set_p2_cluster(int ctas_per_cluster);
/* dense() (scn.py:184-187) straight into the planes the neck reads: d_amax = abs-max of the feature rows, d_info[2] <- {abs-max, S} */
int sessd_sparse_to_dense_planes(const float *d_feat, int max_rows, const void *d_bitmap_index, int channels, sessd_grid grid,
                                 const float *d_amax, float *d_info, void *d_planes, void *stream);
/* sessd_ssfa_fuse that also (d_out nullable: only) writes the fused map as planes; d_info0 / d_info1 [2]: abs-max of x0 / x1 */
int sessd_ssfa_fuse_planes(const float *d_x0, const float *d_x1, const float *d_w0, const float *d_w1, float s0, float t0, float s1,
                           float t1, int num_pixels, int channels, float *d_out, const float *d_info0, const float *d_info1,
                           float *d_out_info, void *d_planes, void *stream);

/* *d_amax = max(*d_amax, max_i |d_x[i]|)  (for tensors produced by kernels without an abs-max epilogue) */
int sessd_absmax(const float *d_x, long long n, float *d_amax, void *stream);


/* SSFA tail (rpn_v1.py:229-233): w_k = BN(conv1x1_{128->1}(x_k)); softmax over the pair; weighted sum */
int sessd_ssfa_fuse(const float *d_x0, const float *d_x1, const float *d_w0 /*[C]*/, const float *d_w1,
                    float s0, float t0, float s1, float t1, int num_pixels, int channels, float *d_out,
                    void *stream);

/* ------------------------------------------------------------------------------------------------
 * P1/P2/P3: decode -> sigmoid -> threshold -> IoU-rectified score -> top-k -> rotated NMS -> direction fix ->
 * range mask, all on device, no host round trip.  Replaces MultiGroupHead.predict / get_task_detections
 * (mg_head_sessd.py:893-1057), box_torch_ops.second_box_decode (:81-147), box_torch_ops.rotate_nms (:527-548),
 * nms_cpu.py:37-48 and nms_cpu.h:72-168.
 * head layout per pixel: [box 2x7 | cls 2 | dir 2x2 | iou 2] = 22 floats, row stride head_stride (22, or 24 when the
 * fused 128->22 head GEMM pads its output to a multiple of 4 channels).
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    int batch;
    int num_anchors;           /* per frame (70400) */
    int anchors_per_loc;       /* 2 */
    int head_stride;           /* floats per pixel row of d_head (>= 22) */
    float score_thresh;        /* 0.3 */
    int nms_pre_max;           /* 1000 */
    int nms_post_max;          /* 100 */
    float nms_iou_thresh;      /* 0.01 */
    int nms_ge;                /* 1: suppress when iou >= thr (nms_cpu.h:155); 0: iou > thr (iou3d nms_kernel) */
    float post_range[6];       /* 0,-40,-5,70.4,40,5 */
    float direction_offset;    /* 0 */
    int use_frustum;           /* apply the calib frustum filter (mg_head_sessd.py:1024-1030) */
} sessd_post_cfg;

size_t sessd_postprocess_workspace_bytes(const sessd_post_cfg *cfg);

/* d_head [batch, num_anchors/apl, head_stride]; d_anchors [num_anchors, 7] (shared by all frames);
 * d_frustum [batch, 6, 4] plane (a,b,c,d) per surface (nullable unless use_frustum);
 * outputs: d_boxes [batch, post_max, 7], d_scores [batch, post_max], d_labels [batch, post_max] i32,
 * d_count [batch] i32, d_aux [batch, 4] i32 (candidates, pre-NMS count, NMS-selected count, reserved),
 * d_sel_anchor [batch, post_max] i32 (anchor index of each NMS-selected box, before frustum/range masks). */
int sessd_postprocess(const float *d_head, const float *d_anchors, const float *d_frustum,
                      const sessd_post_cfg *cfg, float *d_boxes, float *d_scores, int *d_labels, int *d_count,
                      int *d_aux, int *d_sel_anchor, void *workspace, size_t workspace_bytes, void *stream);

/* Same as sessd_postprocess, plus a packed copy of the results for ONE device->host transfer per batch (the reference moves boxes,
 * scores and labels to the host separately and syncs twice per frame: box_torch_ops.py:536, mg_head_sessd.py:1026):
 * d_packed [batch, post_max, 8] = box 7 | score; d_meta [batch, 8 + post_max] i32 = count, candidates, pre-NMS count,
 * NMS-selected count, d_num_voxels[b] (0 if null), *d_status (0 if null), 0, 0, then the anchor index of every returned
 * detection (-1 beyond count). */
int sessd_postprocess_packed(const float *d_head, const float *d_anchors, const float *d_frustum,
                             const sessd_post_cfg *cfg, float *d_boxes, float *d_scores, int *d_labels, int *d_count,
                             int *d_aux, int *d_sel_anchor, float *d_packed, int *d_meta, const int *d_num_voxels,
                             const int *d_status, void *workspace, size_t workspace_bytes, void *stream);

/* stand-alone rotated NMS on [n,5] (x,y,w,l,r) + scores: box_torch_ops.rotate_nms semantics
 * (top-k pre_max by score, greedy, keep <= post_max); d_keep [post_max] i32 indices into the input */
size_t sessd_rotate_nms_workspace_bytes(int max_boxes, int pre_max);
int sessd_rotate_nms(const float *d_boxes5, const float *d_scores, const int *d_n, int max_boxes, int pre_max,
                     int post_max, float iou_thresh, int ge, int *d_keep, int *d_num_keep, void *workspace,
                     size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * I1/I2: the iou3d_cuda extension.  Replaces det3d/core/iou3d/src/iou3d.cpp:34-281 (+ iou3d_kernel.cu
 * :270-365): same box layouts ([x1,y1,x2,y2,ry] / [x1,y1,z1,x2,y2,z2,ry]), caller-allocated outputs.
 * NMS variants take boxes already sorted by descending score (iou3d_utils.py:254-306) and do the greedy
 * reduction ON DEVICE (the reference copies the N x N/64 mask to the host, iou3d.cpp:131-158).
 * ------------------------------------------------------------------------------------------------ */
int sessd_boxes_overlap_bev(const float *d_a, int n, const float *d_b, int m, float *d_out, void *stream);
int sessd_boxes_aligned_overlap_bev(const float *d_a, const float *d_b, int n, float *d_out, void *stream);
int sessd_boxes_iou_bev(const float *d_a, int n, const float *d_b, int m, float *d_out, void *stream);
int sessd_boxes_iou3d(const float *d_a, int n, const float *d_b, int m, float *d_out, void *stream);
size_t sessd_nms_workspace_bytes(int n);
/* mode 0: rotated BEV (nms_gpu), 1: 3-D (nms_3d_gpu), 2: axis-aligned (nms_normal_gpu);
 * d_keep [n] int64 (device), d_num_keep [1] i32 (device) */
int sessd_nms_sorted(const float *d_boxes, int n, float thresh, int mode, long long *d_keep, int *d_num_keep,
                     void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * T1: IoU target assignment of the SSD head, batched over frames on the device.
 * Replaces the DataLoader-worker path det3d/core/anchor/target_assigner.py:68-136 (TargetAssigner.assign_v2 with
 * enable_similar_type=True: every GT is class 1) -> det3d/core/anchor/target_ops_v2.py:11-126 (create_target_np) with
 * NearestIouSimilarity (det3d/core/bbox/region_similarity.py:85-98; box_np_ops.py:354-366, :1007-1046 iou_jit eps=0)
 * and second_box_encode (det3d/core/bbox/box_np_ops.py:52-113), called from AssignTarget
 * (det3d/datasets/pipelines/preprocess.py:286-358).
 *   d_anchors [A,7] (x,y,z,w,l,h,r), shared by all frames; d_gt_boxes [batch,max_gt,7] padded, d_num_gt [batch].
 *   d_labels [batch,A] (1 positive / 0 negative / -1 ignore), d_bbox_targets [batch,A,7] (zeros off the positives),
 *   d_bbox_outside_weights [batch,A] (1 on positives), d_pos_anchor / d_pos_gt_id [batch,A]: the first d_num_pos[b]
 *   entries of row b are the positive anchors in ascending order and their GT index (`positive_gt_id`).
 * labels / positive sets are bit-exact against the reference (its mixed fp32/fp64 IoU rounding is reproduced).
 * max_gt <= 1024.  Workspace: sessd_assign_workspace_bytes().
 * ------------------------------------------------------------------------------------------------ */
size_t sessd_assign_workspace_bytes(int num_anchors, int batch, int max_gt);
int sessd_assign_targets(const float *d_anchors, int num_anchors, const float *d_gt_boxes, const int *d_num_gt, int batch,
                         int max_gt, float matched_thr, float unmatched_thr, int *d_labels, float *d_bbox_targets,
                         float *d_bbox_outside_weights, int *d_pos_anchor, int *d_pos_gt_id, int *d_num_pos,
                         void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * SURVEY.md 8(f) row 1, first slice of the training step: the supervised SSD-head loss terms, value AND gradient in one pass.
 * Replaces (for the terms without the teacher model) det3d/models/bbox_heads/mg_head_sessd.py:706-760:
 * prepare_loss_weights/NormByNumPositives (:525-572), SigmoidFocalLoss (det3d/models/losses/losses.py:345-420, gamma = 2),
 * add_sin_difference + WeightedSmoothL1Loss (mg_head_sessd.py:39-44, losses.py:147-204), get_direction_target +
 * WeightedSoftmaxClassificationLoss (mg_head_sessd.py:62-76, losses.py:489-531) and their autograd backward.
 *   d_head [batch, A/2, head_stride]: the fused head tensor (box 2x7 | cls 2 | dir 2x2 | iou 2 | pad), as sessd_postprocess reads it;
 *   d_labels [batch, A] (1 / 0 / -1) and d_reg_targets [batch, A, 7]: outputs of sessd_assign_targets; d_anchors [A, 7].
 *   d_losses [batch, 8] = per-frame SUMS {cls, loc, dir, cls on positives, cls on negatives, 0, num_pos, num_neg} (the reference
 *   reports loss_weight * batch total / batch); d_grad_head (nullable) [batch, A/2, head_stride] = d/d_head of
 *   (w_cls * sum cls + w_loc * sum loc + w_dir * sum dir) / batch.  Sums are reduced in a fixed order (deterministic).
 * ------------------------------------------------------------------------------------------------ */
size_t sessd_head_loss_workspace_bytes(int batch);
/* IoU-prediction term (mg_head_sessd.py:755-768): smooth-L1 of the head's iou output against 2 * aligned-3D-IoU(decoded prediction,
 * decoded target) - 1 on the positives (det3d/core/iou3d/iou3d_utils.py:197-252 as the constant target).  Run AFTER sessd_head_loss on the
 * same stream: reads num_pos from d_losses[b][6], writes the per-frame sum to d_losses[b][5] and d(w_iou * sum / batch) into the iou
 * channels of d_grad_head. */
/* ODIoU box-regression loss (det3d/models/losses/odious.py:845-900 called from mg_head_sessd.py:770-778; the reference evaluates it with
 * per-box numpy loops on the CPU inside the training step): odiou = 1 - IoU3D + centre distance^2 / (min bounding rectangle diagonal^2 +
 * inter_h^2) + 1.25 (1 - |cos dr|) between the decoded prediction and the decoded target of every positive anchor, weight 1 / num_pos.
 * Run AFTER sessd_head_loss on the same stream (reads num_pos from d_losses[b][6]); d_odiou_sum [batch] receives the per-frame sums,
 * and w_odiou * d(sum over frames) / batch is ADDED to the box channels of d_grad_head (nullable).  The reference's ious_loss is
 * 2.0 * batch total / batch_size: pass w_odiou = 2.0.  Gradients are exact (forward-mode differentiation of the same arithmetic). */
size_t sessd_odiou_loss_workspace_bytes(int batch);
int sessd_odiou_loss(const float *d_head, const float *d_anchors, const int *d_labels, const float *d_reg_targets, int batch,
                     int num_anchors, int anchors_per_loc, int head_stride, float w_odiou, const float *d_losses, float *d_odiou_sum,
                     float *d_grad_head, void *workspace, size_t workspace_bytes, void *stream);
/* HOST evaluation of the identical arithmetic for n (target, prediction) box pairs [n,7]: odiou values [n] and d(odiou)/d(prediction)
 * [n,7] (nullable).  Used to pin the kernel's math to the reference's odiou_3D on the CPU; not a fallback of the device path. */
int sessd_odiou_pairs_host(const float *h_gboxes, const float *h_qboxes, int n, float *h_odiou, float *h_grad_q);
size_t sessd_iou_pred_loss_workspace_bytes(int batch);
int sessd_iou_pred_loss(const float *d_head, const float *d_anchors, const int *d_labels, const float *d_reg_targets, int batch,
                        int num_anchors, int anchors_per_loc, int head_stride, float sigma, float w_iou, float *d_losses,
                        float *d_grad_head, void *workspace, size_t workspace_bytes, void *stream);
int sessd_head_loss(const float *d_head, const float *d_anchors, const int *d_labels, const float *d_reg_targets, int batch,
                    int num_anchors, int anchors_per_loc, int head_stride, float alpha, float sigma, float dir_offset,
                    float pos_cls_weight, float neg_cls_weight, float w_cls, float w_loc, float w_dir, float *d_losses,
                    float *d_grad_head, void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Training step, optimiser side, over flat fp32 arenas (csrc/train.cu):
 * sessd_axpby  -- d_y = a d_y + b d_x: the teacher's exponential moving average of det3d/torchie/trainer/trainer_sessd.py:315-318 in one
 *                 launch (a = alpha, b = 1 - alpha), and the 1 / world_size scaling after the gradient all-reduce
 *                 (det3d/core/utils/dist_utils.py:8-29) with d_x = NULL;
 * sessd_adamw_step -- Adam with decoupled weight decay (the fastai true_wd step of det3d/solver/fastai_optim.py == torch.optim.AdamW).
 * ------------------------------------------------------------------------------------------------ */
int sessd_axpby(float *d_y, const float *d_x, float a, float b, long long n, void *stream);
int sessd_adamw_step(float *d_param, const float *d_grad, float *d_exp_avg, float *d_exp_avg_sq, long long n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, int step, void *stream);

#ifdef __cplusplus
}
#endif
#endif
