/*
 * oracle.c -- CPU restatement of the SE-SSD per-frame hot path's integer / geometry stages.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this file.
 *
 * Every function cites the reference file:line (relative to the Vegeta2020/SE-SSD tree) whose
 * algorithm it restates.  The restatement keeps the reference's fp32 operation order so that
 * results can be compared bit-for-bit with golden vectors generated from the reference itself
 * (tests/golden/make_golden.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared oracle.c -o liboracle.so -lm   (see oracle/build.py)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Voxeliser: det3d/ops/point_cloud/point_cloud_ops_v2.py:9-62 (_points_to_voxel_reverse_kernel)
 * and :120-194 (points_to_voxel wrapper: zero-initialised outputs, slice to voxel_num).
 * The reference keeps a module-global uint16 cell->voxel map (:6) with sentinel 65535; we keep a
 * lazily allocated int32 map with sentinel -1 (as the v1 twin, point_cloud_ops.py:112-184, does) so
 * that max_voxels >= 65535 (stress config) is representable.  Results are identical for
 * max_voxels < 65535.
 * ------------------------------------------------------------------------------------------ */
static int32_t *g_cell_map = NULL;
static size_t g_cell_map_cells = 0;

int oracle_points_to_voxel(const float *points, int num_points, int num_feat,
                           const float *voxel_size /*3, xyz*/, const float *range /*6*/,
                           const int *grid /*3, xyz*/, int max_points, int max_voxels,
                           float *voxels /*[max_voxels,max_points,num_feat] zeroed by caller*/,
                           int *coors /*[max_voxels,3] zyx, zeroed*/,
                           int *num_per_voxel /*[max_voxels] zeroed*/)
{
    const size_t cells = (size_t)grid[0] * grid[1] * grid[2];
    if (cells > g_cell_map_cells) {
        free(g_cell_map);
        g_cell_map = (int32_t *)malloc(cells * sizeof(int32_t));
        if (!g_cell_map) { g_cell_map_cells = 0; return -1; }
        memset(g_cell_map, 0xff, cells * sizeof(int32_t));
        g_cell_map_cells = cells;
    }
    int voxel_num = 0;
    for (int i = 0; i < num_points; ++i) {
        const float *p = points + (size_t)i * num_feat;
        int c[3]; /* x, y, z cell */
        int failed = 0;
        for (int j = 0; j < 3; ++j) {
            /* :38  fp32 subtract, fp32 divide, floor */
            float cf = floorf((p[j] - range[j]) / voxel_size[j]);
            if (cf < 0 || cf >= (float)grid[j]) { failed = 1; break; }   /* :39-41 */
            c[j] = (int)cf;
        }
        if (failed) continue;
        /* reversed (zyx) addressing, :42,:45 */
        size_t cell = ((size_t)c[2] * grid[1] + c[1]) * grid[0] + c[0];
        int vid = g_cell_map[cell];
        if (vid == -1) {                     /* :46 */
            vid = voxel_num;
            if (voxel_num >= max_voxels) break;   /* :48-49: drops ALL later points */
            voxel_num += 1;
            g_cell_map[cell] = vid;
            coors[vid * 3 + 0] = c[2];
            coors[vid * 3 + 1] = c[1];
            coors[vid * 3 + 2] = c[0];
        }
        int num = num_per_voxel[vid];
        if (num < max_points) {              /* :54-57 */
            memcpy(voxels + ((size_t)vid * max_points + num) * num_feat, p, sizeof(float) * num_feat);
            num_per_voxel[vid] = num + 1;
        }
    }
    for (int v = 0; v < voxel_num; ++v) {    /* :59-61 reset touched cells */
        size_t cell = ((size_t)coors[v * 3] * grid[1] + coors[v * 3 + 1]) * grid[0] + coors[v * 3 + 2];
        g_cell_map[cell] = -1;
    }
    return voxel_num;
}

/* ------------------------------------------------------------------------------------------
 * Rotated-rectangle overlap: det3d/core/iou3d/src/iou3d_cpu.cpp:126-245 (box_overlap, input_2d=1)
 * with helpers :36-49 (cross / check_rect_cross), :51-66 (check_in_box2d), :84-112 (intersection),
 * :114-118 (rotate_around_center), :36-38 (point_cmp).  Same fp32 operation order; the twin CUDA
 * kernel is det3d/core/iou3d/src/iou3d_kernel.cu:125-245.
 * box layout: [x1, y1, x2, y2, angle]  (axis-aligned extent before rotation + clockwise angle).
 * ------------------------------------------------------------------------------------------ */
#define OR_EPS 1e-8f

typedef struct { float x, y; } pt_t;

static inline float cross3(pt_t p1, pt_t p2, pt_t p0)
{
    return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}

static inline pt_t spin(pt_t c, float ca, float sa, pt_t p)
{
    pt_t r;
    r.x = (p.x - c.x) * ca + (p.y - c.y) * sa + c.x;
    r.y = -(p.x - c.x) * sa + (p.y - c.y) * ca + c.y;
    return r;
}

static int seg_cross(pt_t p1, pt_t p0, pt_t q1, pt_t q0, pt_t *out)
{
    /* bounding-box rejection (:41-47) */
    if (!(fminf(p0.x, p1.x) <= fmaxf(q0.x, q1.x) && fminf(q0.x, q1.x) <= fmaxf(p0.x, p1.x) &&
          fminf(p0.y, p1.y) <= fmaxf(q0.y, q1.y) && fminf(q0.y, q1.y) <= fmaxf(p0.y, p1.y)))
        return 0;
    float s1 = cross3(q0, p1, p0);
    float s2 = cross3(p1, q1, p0);
    float s3 = cross3(p0, q1, q0);
    float s4 = cross3(q1, p1, q0);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
    float s5 = cross3(q1, p1, p0);
    if (fabsf(s5 - s1) > OR_EPS) {
        out->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        out->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        float D = a0 * b1 - a1 * b0;
        out->x = (b0 * c1 - b1 * c0) / D;
        out->y = (a1 * c0 - a0 * c1) / D;
    }
    return 1;
}

static int inside_box(const float *box, pt_t p)
{
    const float margin = 1e-5f;
    float cx = (box[0] + box[2]) / 2;
    float cy = (box[1] + box[3]) / 2;
    float ca = cosf(-box[4]), sa = sinf(-box[4]);
    float rx = (p.x - cx) * ca + (p.y - cy) * sa + cx;
    float ry = -(p.x - cx) * sa + (p.y - cy) * ca + cy;
    return rx > box[0] - margin && rx < box[2] + margin && ry > box[1] - margin && ry < box[3] + margin;
}

float oracle_box_overlap(const float *a, const float *b)
{
    pt_t ca = { (a[0] + a[2]) / 2, (a[1] + a[3]) / 2 };
    pt_t cb = { (b[0] + b[2]) / 2, (b[1] + b[3]) / 2 };
    pt_t A[5] = { {a[0], a[1]}, {a[2], a[1]}, {a[2], a[3]}, {a[0], a[3]} };
    pt_t B[5] = { {b[0], b[1]}, {b[2], b[1]}, {b[2], b[3]}, {b[0], b[3]} };
    float cosa = cosf(a[4]), sina = sinf(a[4]);
    float cosb = cosf(b[4]), sinb = sinf(b[4]);
    for (int k = 0; k < 4; ++k) { A[k] = spin(ca, cosa, sina, A[k]); B[k] = spin(cb, cosb, sinb, B[k]); }
    A[4] = A[0]; B[4] = B[0];

    pt_t poly[24];
    pt_t ctr = {0.f, 0.f};
    int cnt = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            pt_t x;
            if (seg_cross(A[i + 1], A[i], B[j + 1], B[j], &x)) {
                ctr.x = ctr.x + x.x; ctr.y = ctr.y + x.y;
                poly[cnt++] = x;
            }
        }
    for (int k = 0; k < 4; ++k) {
        if (inside_box(a, B[k])) { ctr.x = ctr.x + B[k].x; ctr.y = ctr.y + B[k].y; poly[cnt++] = B[k]; }
        if (inside_box(b, A[k])) { ctr.x = ctr.x + A[k].x; ctr.y = ctr.y + A[k].y; poly[cnt++] = A[k]; }
    }
    ctr.x /= cnt; ctr.y /= cnt;
    /* bubble sort by polar angle about the centroid (:221-229) */
    for (int j = 0; j < cnt - 1; ++j)
        for (int i = 0; i < cnt - j - 1; ++i) {
            if (atan2f(poly[i].y - ctr.y, poly[i].x - ctr.x) > atan2f(poly[i + 1].y - ctr.y, poly[i + 1].x - ctr.x)) {
                pt_t t = poly[i]; poly[i] = poly[i + 1]; poly[i + 1] = t;
            }
        }
    float area = 0;
    for (int k = 0; k < cnt - 1; ++k) {
        float ux = poly[k].x - poly[0].x, uy = poly[k].y - poly[0].y;
        float vx = poly[k + 1].x - poly[0].x, vy = poly[k + 1].y - poly[0].y;
        area += ux * vy - uy * vx;
    }
    return (float)(fabsf(area) / 2.0);
}

/* iou3d_cpu.cpp:247-254 (iou_bev) */
float oracle_iou_bev(const float *a, const float *b)
{
    float sa = (a[2] - a[0]) * (a[3] - a[1]);
    float sb = (b[2] - b[0]) * (b[3] - b[1]);
    float so = oracle_box_overlap(a, b);
    return so / fmaxf(sa + sb - so, OR_EPS);
}

/* iou3d_cpu.cpp:256-281 / :283-303 : N x M matrices */
void oracle_boxes_overlap_bev(const float *a, int n, const float *b, int m, float *out)
{
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j) out[(size_t)i * m + j] = oracle_box_overlap(a + 5 * i, b + 5 * j);
}

void oracle_boxes_iou_bev(const float *a, int n, const float *b, int m, float *out)
{
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j) out[(size_t)i * m + j] = oracle_iou_bev(a + 5 * i, b + 5 * j);
}

/* iou3d_kernel.cu:256-268 (iou_3d) -- boxes [x1,y1,z1,x2,y2,z2,angle].  (The CPU twin
 * iou3d_cpu.cpp:305-336 has an index bug at :329, ans[i*num_a+j]; the CUDA kernel is normative.) */
float oracle_iou_3d(const float *a, const float *b)
{
    float va = (a[3] - a[0]) * (a[4] - a[1]) * (a[5] - a[2]);
    float vb = (b[3] - b[0]) * (b[4] - b[1]) * (b[5] - b[2]);
    float lo = fmaxf(a[2], b[2]);
    float hi = fminf(a[5], b[5]);
    float dh = fmaxf(hi - lo, OR_EPS);
    if (dh == OR_EPS) return 0.f;
    float a5[5] = { a[0], a[1], a[3], a[4], a[6] };
    float b5[5] = { b[0], b[1], b[3], b[4], b[6] };
    float vo = oracle_box_overlap(a5, b5) * dh;
    return vo / fmaxf(va + vb - vo, OR_EPS);
}

void oracle_boxes_iou_3d(const float *a, int n, const float *b, int m, float *out)
{
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j) out[(size_t)i * m + j] = oracle_iou_3d(a + 7 * i, b + 7 * j);
}

/* iou3d_kernel.cu:413-423 (iou_normal): axis-aligned IoU on [x1,y1,x2,y2,*] */
static float iou_axis(const float *a, const float *b)
{
    float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
    float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
    float w = fmaxf(right - left, 0.f), h = fmaxf(bottom - top, 0.f);
    float inter = w * h;
    float sa = (a[2] - a[0]) * (a[3] - a[1]);
    float sb = (b[2] - b[0]) * (b[3] - b[1]);
    return inter / fmaxf(sa + sb - inter, OR_EPS);
}

/* ------------------------------------------------------------------------------------------
 * iou3d NMS: iou3d_kernel.cu:323-365 (mask bit iff iou > thresh) + host greedy reduce
 * iou3d.cpp:117-164.  Boxes must already be sorted by descending score (iou3d_utils.py:254-271).
 * mode 0: rotated BEV [x1,y1,x2,y2,ry] (stride 5); mode 1: 3D (stride 7); mode 2: axis-aligned (stride 5).
 * ------------------------------------------------------------------------------------------ */
int oracle_nms_sorted(const float *boxes, int n, float thresh, int mode, int64_t *keep)
{
    unsigned char *dead = (unsigned char *)calloc((size_t)n + 1, 1);
    int stride = (mode == 1) ? 7 : 5;
    int nk = 0;
    for (int i = 0; i < n; ++i) {
        if (dead[i]) continue;
        keep[nk++] = i;
        for (int j = i + 1; j < n; ++j) {
            if (dead[j]) continue;
            float v = (mode == 0) ? oracle_iou_bev(boxes + i * stride, boxes + j * stride)
                    : (mode == 1) ? oracle_iou_3d(boxes + i * stride, boxes + j * stride)
                                  : iou_axis(boxes + i * stride, boxes + j * stride);
            if (v > thresh) dead[j] = 1;
        }
    }
    free(dead);
    return nk;
}

/* ------------------------------------------------------------------------------------------
 * The NMS the inference path actually runs: det3d/core/bbox/box_torch_ops.py:527-548 (rotate_nms)
 * -> det3d/ops/nms/nms_cpu.py:37-48 (rotate_nms_cc) -> det3d/ops/nms/nms_cpu.h:72-168
 * (rotate_non_max_suppression_cpu).  Control flow is restated exactly:
 *   order = argsort(score) descending; corners via center_to_corner_box2d (box_np_ops.py:512-532,
 *   :267-294, :433-446, clockwise rotation); stand-up AABBs (corner_to_standup_nd); pair skipped
 *   when standup IoU (iou_jit eps=0, box_np_ops.py:1007-1046) <= 0; suppressed when
 *   inter/union >= thresh.
 * boost::geometry is NOT available in this image, so the polygon inter/union areas use the
 * rotated-rectangle arithmetic of iou3d_cpu.cpp (oracle_iou_bev above) on
 * [x-w/2, y-l/2, x+w/2, y+l/2, r] (det3d/core/iou3d/utils.py:74-101).  PARITY UNPINNED against
 * boost for that one piece; an exact fp64 polygon clip is used as a second opinion
 * in tests/test_nms.py.
 * dets: [n,6] = x, y, w, l, r, score.  Ties in score are broken by lower index first.
 * ------------------------------------------------------------------------------------------ */
static const float *g_sort_scores;
static int cmp_desc(const void *pa, const void *pb)
{
    int a = *(const int *)pa, b = *(const int *)pb;
    float sa = g_sort_scores[a], sb = g_sort_scores[b];
    if (sa > sb) return -1;
    if (sa < sb) return 1;
    return (a > b) - (a < b);
}

void oracle_corners_standup(const float *det /*x,y,w,l,r*/, float *corners /*4x2*/, float *standup /*4*/)
{
    /* corners_nd (origin 0.5): (-.5,-.5), (-.5,.5), (.5,.5), (.5,-.5) scaled by (w,l) */
    static const float nx[4] = { -0.5f, -0.5f, 0.5f, 0.5f };
    static const float ny[4] = { -0.5f, 0.5f, 0.5f, -0.5f };
    float s = sinf(det[4]), c = cosf(det[4]);
    float xmin = 0, ymin = 0, xmax = 0, ymax = 0;
    for (int k = 0; k < 4; ++k) {
        float px = det[2] * nx[k], py = det[3] * ny[k];
        /* rotation_2d: einsum('aij,jka->aik', pts, [[cos,-sin],[sin,cos]]) */
        float rx = px * c + py * s;
        float ry = px * (-s) + py * c;
        rx += det[0]; ry += det[1];
        corners[2 * k] = rx; corners[2 * k + 1] = ry;
        if (k == 0) { xmin = xmax = rx; ymin = ymax = ry; }
        else { xmin = fminf(xmin, rx); xmax = fmaxf(xmax, rx); ymin = fminf(ymin, ry); ymax = fmaxf(ymax, ry); }
    }
    standup[0] = xmin; standup[1] = ymin; standup[2] = xmax; standup[3] = ymax;
}

/* iou_jit(eps=0) > 0 test for one pair (box_np_ops.py:1007-1046): boxes=n (rows), query=k (cols) */
static float standup_iou(const float *bn, const float *qk)
{
    float box_area = (qk[2] - qk[0]) * (qk[3] - qk[1]);
    float iw = fminf(bn[2], qk[2]) - fmaxf(bn[0], qk[0]);
    if (iw > 0) {
        float ih = fminf(bn[3], qk[3]) - fmaxf(bn[1], qk[1]);
        if (ih > 0) {
            float ua = (bn[2] - bn[0]) * (bn[3] - bn[1]) + box_area - iw * ih;
            return iw * ih / ua;
        }
    }
    return 0.f;
}

int oracle_rotate_nms_cc(const float *dets, int n, float thresh, int ge_mode /*1: >= (nms_cpu.h:155), 0: > */,
                         int64_t *keep)
{
    int *order = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    float *scores = (float *)malloc(sizeof(float) * (size_t)(n + 1));
    float *su = (float *)malloc(sizeof(float) * 4 * (size_t)(n + 1));
    float *bev = (float *)malloc(sizeof(float) * 5 * (size_t)(n + 1));
    unsigned char *dead = (unsigned char *)calloc((size_t)n + 1, 1);
    for (int i = 0; i < n; ++i) {
        const float *d = dets + 6 * i;
        float corners[8];
        order[i] = i; scores[i] = d[5];
        oracle_corners_standup(d, corners, su + 4 * i);
        float hw = d[2] / 2.f, hl = d[3] / 2.f;
        bev[5 * i + 0] = d[0] - hw; bev[5 * i + 1] = d[1] - hl;
        bev[5 * i + 2] = d[0] + hw; bev[5 * i + 3] = d[1] + hl; bev[5 * i + 4] = d[4];
    }
    g_sort_scores = scores;
    qsort(order, (size_t)n, sizeof(int), cmp_desc);
    int nk = 0;
    for (int _i = 0; _i < n; ++_i) {
        int i = order[_i];
        if (dead[i]) continue;
        keep[nk++] = i;
        for (int _j = _i + 1; _j < n; ++_j) {
            int j = order[_j];
            if (dead[j]) continue;
            if (standup_iou(su + 4 * i, su + 4 * j) <= 0.0f) continue;   /* nms_cpu.h:104-105 */
            float ov = oracle_iou_bev(bev + 5 * i, bev + 5 * j);
            if (ge_mode ? (ov >= thresh) : (ov > thresh)) dead[j] = 1;
        }
    }
    free(order); free(scores); free(su); free(bev); free(dead);
    return nk;
}

/* ------------------------------------------------------------------------------------------
 * Box decode: det3d/core/bbox/box_torch_ops.py:81-147 (second_box_decode, 7-dim, no vector
 * angle, exp dims).  enc/anchors/out: [n,7] = x y z w l h r.
 * ------------------------------------------------------------------------------------------ */
void oracle_box_decode(const float *enc, const float *anc, int n, float *out)
{
    for (int i = 0; i < n; ++i) {
        const float *t = enc + 7 * i, *a = anc + 7 * i;
        float *o = out + 7 * i;
        float diag = sqrtf(a[4] * a[4] + a[3] * a[3]);   /* la^2 + wa^2 */
        o[0] = t[0] * diag + a[0];
        o[1] = t[1] * diag + a[1];
        o[2] = t[2] * a[5] + a[2];
        o[3] = expf(t[3]) * a[3];
        o[4] = expf(t[4]) * a[4];
        o[5] = expf(t[5]) * a[5];
        o[6] = t[6] + a[6];
    }
}
