// odiou.cuh -- the ODIoU 3-D box regression loss of SE-SSD for ONE (target, prediction) pair, written once as a template over the scalar
// type so that the same code yields the value (T = float) and, by forward-mode automatic differentiation (T = Dual<7>: value + partials
// w.r.t. the 7 parameters of the predicted box), its exact gradient.  __host__ __device__: the device kernel (odiou.cu) and the host entry
// point sessd_odiou_pairs_host (the CPU-side check against the reference's own implementation) run the identical arithmetic.
//
// Restates det3d/models/losses/odious.py:845-900 (odiou_3D.forward) and its helpers:
//   odiou = 1 - IoU3D + |c_g - c_q|^2 / (mbr_diag_bev^2 + inter_h^2 + 1e-7) + 1.25 (1 - |cos(r_q - r_g)|)
//   * boxes (x, y, z, w, l, h, r) are clamped to [-200, 200] (:855-856); pairs with a non-positive dimension contribute 0 (:851-853);
//   * BEV corners: rbbox_to_corners (:455-486; corner order 0..3 is clockwise, positive yaw rotates clockwise);
//   * IoU3D = inter_h * inter_area / (vol_g + vol_q - inter_h * inter_area), inter_h clamped at 0 (:878-895); the reference finds the
//     intersection polygon by corner-inside tests + edge intersections + angular sort + triangle fan (:15-445, numpy loops with
//     hand-written Jacobians); here the prediction's rectangle is clipped against the four edges of the target's (Sutherland-Hodgman)
//     and the shoelace area is taken: the same polygon, differentiable by construction;
//   * mbr_diag_bev (:597-648): convex hull of the 8 corners, then for every hull edge the bounding rectangle in the edge's frame
//     (angle = |fmod(atan2(e), 3.1415926/2)|, rotation rows (cos a, cos(a - 3.1415926/2)), (cos(a + 3.1415926/2), cos a) exactly as the
//     reference writes them); the diagonal of the minimum-area one.  The reference walks hull[1:] - hull[:-1], i.e. all edges but the one
//     closing Qhull's (arbitrarily started) vertex list; all edges are used here -- identical unless that one edge is the unique minimiser.
#pragma once
#include <math.h>

#include "common.cuh"

namespace sessd {

#define SESSD_HD __host__ __device__ __forceinline__

template <int N>
struct Dual {
    float v;
    float d[N];
};

template <int N> SESSD_HD Dual<N> dual_const(float c) { Dual<N> r; r.v = c; for (int i = 0; i < N; ++i) r.d[i] = 0.f; return r; }
template <int N> SESSD_HD Dual<N> dual_var(float c, int idx) { Dual<N> r = dual_const<N>(c); r.d[idx] = 1.f; return r; }

// ---- scalar interface shared by float and Dual<N> --------------------------------------------------------------------------------
SESSD_HD float val(float a) { return a; }
template <int N> SESSD_HD float val(const Dual<N> &a) { return a.v; }

template <int N> SESSD_HD Dual<N> operator+(const Dual<N> &a, const Dual<N> &b) { Dual<N> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> SESSD_HD Dual<N> operator-(const Dual<N> &a, const Dual<N> &b) { Dual<N> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> SESSD_HD Dual<N> operator-(const Dual<N> &a) { Dual<N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
template <int N> SESSD_HD Dual<N> operator*(const Dual<N> &a, const Dual<N> &b) { Dual<N> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int N> SESSD_HD Dual<N> operator/(const Dual<N> &a, const Dual<N> &b) {
    Dual<N> r; const float inv = 1.f / b.v; r.v = a.v * inv;
    for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
    return r;
}
template <int N> SESSD_HD Dual<N> operator+(const Dual<N> &a, float b) { Dual<N> r = a; r.v += b; return r; }
template <int N> SESSD_HD Dual<N> operator-(const Dual<N> &a, float b) { Dual<N> r = a; r.v -= b; return r; }
template <int N> SESSD_HD Dual<N> operator*(const Dual<N> &a, float b) { Dual<N> r; r.v = a.v * b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b; return r; }
template <int N> SESSD_HD Dual<N> operator*(float b, const Dual<N> &a) { return a * b; }
template <int N> SESSD_HD Dual<N> operator-(float b, const Dual<N> &a) { Dual<N> r; r.v = b - a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }

SESSD_HD float t_sin(float a) { return sinf(a); }
SESSD_HD float t_cos(float a) { return cosf(a); }
SESSD_HD float t_sqrt(float a) { return sqrtf(a); }
SESSD_HD float t_abs(float a) { return fabsf(a); }
SESSD_HD float t_atan2(float y, float x) { return atan2f(y, x); }
SESSD_HD float t_fmod(float a, float m) { return fmodf(a, m); }
template <int N> SESSD_HD Dual<N> t_sin(const Dual<N> &a) { Dual<N> r; r.v = sinf(a.v); const float c = cosf(a.v); for (int i = 0; i < N; ++i) r.d[i] = c * a.d[i]; return r; }
template <int N> SESSD_HD Dual<N> t_cos(const Dual<N> &a) { Dual<N> r; r.v = cosf(a.v); const float s = -sinf(a.v); for (int i = 0; i < N; ++i) r.d[i] = s * a.d[i]; return r; }
template <int N> SESSD_HD Dual<N> t_sqrt(const Dual<N> &a) { Dual<N> r; r.v = sqrtf(a.v); const float k = r.v > 0.f ? 0.5f / r.v : 0.f; for (int i = 0; i < N; ++i) r.d[i] = k * a.d[i]; return r; }
template <int N> SESSD_HD Dual<N> t_abs(const Dual<N> &a) { return a.v < 0.f ? -a : a; }
template <int N> SESSD_HD Dual<N> t_atan2(const Dual<N> &y, const Dual<N> &x) {
    Dual<N> r; r.v = atan2f(y.v, x.v); const float n2 = x.v * x.v + y.v * y.v; const float k = n2 > 0.f ? 1.f / n2 : 0.f;
    for (int i = 0; i < N; ++i) r.d[i] = (x.v * y.d[i] - y.v * x.d[i]) * k;
    return r;
}
template <int N> SESSD_HD Dual<N> t_fmod(const Dual<N> &a, float m) { Dual<N> r = a; r.v = fmodf(a.v, m); return r; }     // d/da = 1 a.e.

template <class T> SESSD_HD T t_min(const T &a, const T &b) { return val(b) < val(a) ? b : a; }
template <class T> SESSD_HD T t_max(const T &a, const T &b) { return val(b) > val(a) ? b : a; }
SESSD_HD float t_lit(float, float c) { return c; }                                          // constant of the same scalar type
template <int N> SESSD_HD Dual<N> t_lit(const Dual<N> &, float c) { return dual_const<N>(c); }
// torch.clamp(x, lo, hi): derivative 1 inside (bounds included), 0 outside
SESSD_HD float t_clamp(float a, float lo, float hi) { return fminf(fmaxf(a, lo), hi); }
template <int N> SESSD_HD Dual<N> t_clamp(const Dual<N> &a, float lo, float hi) { return (a.v < lo) ? dual_const<N>(lo) : ((a.v > hi) ? dual_const<N>(hi) : a); }

template <class T> struct Pt { T x, y; };

// rbbox_to_corners (odious.py:455-486) of (x, y, w, l, r): 4 corners, clockwise
template <class T>
SESSD_HD void od_corners(const T &x, const T &y, const T &w, const T &l, const T &r, Pt<T> c[4]) {
    const T cs = t_cos(r), sn = t_sin(r);
    const T dxcos = w * cs * 0.5f, dxsin = w * sn * 0.5f, dycos = l * cs * 0.5f, dysin = l * sn * 0.5f;
    c[0].x = -dxcos - dysin + x; c[0].y = dxsin - dycos + y;
    c[1].x = -dxcos + dysin + x; c[1].y = dxsin + dycos + y;
    c[2].x = dxcos + dysin + x;  c[2].y = -dxsin + dycos + y;
    c[3].x = dxcos - dysin + x;  c[3].y = -dxsin - dycos + y;
}

// area of (subject polygon) n (clip rectangle), both given clockwise; Sutherland-Hodgman, <= 8 vertices
template <class T>
SESSD_HD T od_inter_area(const Pt<T> clip[4], const Pt<T> subj[4]) {
    Pt<T> a[10], b[10];
    int na = 4;
    for (int i = 0; i < 4; ++i) a[i] = subj[i];
    for (int e = 0; e < 4 && na > 0; ++e) {
        const Pt<T> p0 = clip[e], p1 = clip[(e + 1) & 3];
        const T ex = p1.x - p0.x, ey = p1.y - p0.y;
        int nb = 0;
        for (int i = 0; i < na; ++i) {
            const Pt<T> s = a[i], t = a[(i + 1) % na];
            // clockwise clip polygon: inside = on the right of the directed edge = cross(e, p - p0) <= 0
            const T cs = ex * (s.y - p0.y) - ey * (s.x - p0.x);
            const T ct = ex * (t.y - p0.y) - ey * (t.x - p0.x);
            const bool ins = val(cs) <= 0.f, int_ = val(ct) <= 0.f;
            if (ins && nb < 10) b[nb++] = s;
            if (ins != int_ && nb < 10) {      // (rect n rect has <= 8 vertices; the bound only matters for non-finite input)
                const T u = cs / (cs - ct);                   // s + u (t - s) lies on the clip edge
                Pt<T> m;
                m.x = s.x + u * (t.x - s.x);
                m.y = s.y + u * (t.y - s.y);
                b[nb++] = m;
            }
        }
        na = nb;
        for (int i = 0; i < nb; ++i) a[i] = b[i];
    }
    T s2 = t_lit(clip[0].x, 0.f);
    if (na < 3) return s2;
    for (int i = 0; i < na; ++i) {
        const Pt<T> &p = a[i], &q = a[(i + 1) % na];
        s2 = s2 + (p.x * q.y - q.x * p.y);
    }
    return t_abs(s2) * 0.5f;
}

// diagonal of the minimum-area bounding rectangle of 8 points whose orientation follows one of the convex-hull edges (odious.py:597-648)
template <class T>
SESSD_HD T od_mbr_diag(const Pt<T> pts[8]) {
    // convex hull, counter-clockwise (Andrew's monotone chain on the values; the duals ride along)
    int idx[8];
    for (int i = 0; i < 8; ++i) idx[i] = i;
    for (int i = 1; i < 8; ++i) {                                   // insertion sort by (x, y)
        const int k = idx[i];
        int j = i - 1;
        while (j >= 0 && (val(pts[idx[j]].x) > val(pts[k].x) || (val(pts[idx[j]].x) == val(pts[k].x) && val(pts[idx[j]].y) > val(pts[k].y)))) {
            idx[j + 1] = idx[j];
            --j;
        }
        idx[j + 1] = k;
    }
    int hull[18];
    int h = 0;
    auto cross = [&](int o, int a, int b) {
        return (val(pts[a].x) - val(pts[o].x)) * (val(pts[b].y) - val(pts[o].y)) - (val(pts[a].y) - val(pts[o].y)) * (val(pts[b].x) - val(pts[o].x));
    };
    for (int i = 0; i < 8; ++i) {
        while (h >= 2 && cross(hull[h - 2], hull[h - 1], idx[i]) <= 0.f) --h;
        hull[h++] = idx[i];
    }
    const int lower = h + 1;
    for (int i = 6; i >= 0; --i) {
        while (h >= lower && cross(hull[h - 2], hull[h - 1], idx[i]) <= 0.f) --h;
        hull[h++] = idx[i];
    }
    --h;                                                            // last point == first point
    const float kHalfPi = 3.1415926f / 2.0f;                        // the reference's literals
    T best_diag = t_lit(pts[0].x, 0.f);
    float best_area = 3.0e38f;
    if (h < 2) return best_diag;
    for (int e = 0; e < h; ++e) {
        const Pt<T> &p = pts[hull[e]], &q = pts[hull[(e + 1) % h]];
        const T ang = t_abs(t_fmod(t_atan2(q.y - p.y, q.x - p.x), kHalfPi));
        const T r00 = t_cos(ang), r01 = t_cos(ang - kHalfPi), r10 = t_cos(ang + kHalfPi), r11 = r00;
        T mnx, mxx, mny, mxy;
        for (int k = 0; k < h; ++k) {
            const Pt<T> &v = pts[hull[k]];
            const T rx = r00 * v.x + r01 * v.y, ry = r10 * v.x + r11 * v.y;
            if (k == 0) { mnx = mxx = rx; mny = mxy = ry; }
            else { mnx = t_min(mnx, rx); mxx = t_max(mxx, rx); mny = t_min(mny, ry); mxy = t_max(mxy, ry); }
        }
        const T dx = mxx - mnx, dy = mxy - mny;
        const float area = val(dx) * val(dy);
        if (area < best_area) { best_area = area; best_diag = t_sqrt(dx * dx + dy * dy); }
    }
    return best_diag;
}

// odiou of one pair; g = target (constants), q = prediction (the differentiated scalars)
template <class T>
SESSD_HD T odiou_pair(const float g_in[7], const T q_in[7]) {
    const T zero = t_lit(q_in[0], 0.f);
    if (!(g_in[3] > 0.f && g_in[4] > 0.f && g_in[5] > 0.f && val(q_in[3]) > 0.f && val(q_in[4]) > 0.f && val(q_in[5]) > 0.f)) return zero;
    T g[7], q[7];
    for (int i = 0; i < 7; ++i) { g[i] = t_lit(q_in[0], t_clamp(g_in[i], -200.f, 200.f)); q[i] = t_clamp(q_in[i], -200.f, 200.f); }
    const T angle_factor = 1.25f * (1.0f - t_abs(t_cos(q[6] - g[6])));
    Pt<T> cg[4], cq[4], all[8];
    od_corners(g[0], g[1], g[3], g[4], g[6], cg);
    od_corners(q[0], q[1], q[3], q[4], q[6], cq);
    for (int i = 0; i < 4; ++i) { all[i] = cg[i]; all[4 + i] = cq[i]; }
    const T inter_area = od_inter_area(cg, cq);
    const T dx = g[0] - q[0], dy = g[1] - q[1], dz = g[2] - q[2];
    const T center2 = dx * dx + dy * dy + dz * dz;
    const T diag_bev = od_mbr_diag(all);
    T inter_h = t_min(g[2] + 0.5f * g[5], q[2] + 0.5f * q[5]) - t_max(g[2] - 0.5f * g[5], q[2] - 0.5f * q[5]);
    if (val(inter_h) < 0.f) inter_h = zero;
    const T diag3 = diag_bev * diag_bev + inter_h * inter_h + 1e-7f;
    const T vol_g = g[3] * g[4] * g[5], vol_q = q[3] * q[4] * q[5];
    const T inc = inter_h * inter_area;
    const T iou = inc / (vol_g + vol_q - inc);
    return 1.0f - iou + center2 / diag3 + angle_factor;
}

}  // namespace sessd
