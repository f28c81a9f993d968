// Rotated-rectangle BEV IoU, device functions.
//
// Follows second/core/non_max_suppression/nms_gpu.py:180-415 (+564-576) of the reference
// operation for operation, including the float32/float64 precision map numba's typing creates
// there.  All float32 arithmetic uses _rn intrinsics so nvcc cannot contract
// a*b+c into an FMA (the reference's CPU-derived oracle does not contract).
//
// Beyond the reference: corners of each box are computed once per box (not once per pair), and
// pairs whose axis-aligned hulls are separated by more than a guard band skip the clip: no corner
// of one can lie in the other and no edges can cross, so the reference returns intersection 0.
#pragma once
#include <cuda_runtime.h>

namespace pp {

struct RBox {      // 8 corner floats + area + hull
    float c[8];
    float area;    // w*l in float32 (devRotateIoU line 411-412)
    float mnx, mny, mxx, mxy;
};

// rbbox_to_corners, nms_gpu.py:367-390
__device__ __forceinline__ void rbox_prepare(const float* r /*x,y,w,l,angle*/, RBox& o) {
    // float32 cos/sin as a correctly rounded libm returns them (the reference's numba path calls
    // libm cosf/sinf; CUDA's cosf/sinf are 1-2 ulp off, which is enough to flip the inclusive
    // corner tests).  Once per box, so the double-precision evaluation is off the pair loop.
    double dsn, dcs;
    sincos((double)r[4], &dsn, &dcs);
    const float a_cos = (float)dcs, a_sin = (float)dsn;
    const float cx = r[0], cy = r[1];
    const float hx = (float)((double)r[2] / 2.0), hy = (float)((double)r[3] / 2.0);
    const float xs[4] = {-hx, -hx, hx, hx};
    const float ys[4] = {-hy, hy, hy, -hy};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        o.c[2 * i] = __fadd_rn(__fadd_rn(__fmul_rn(a_cos, xs[i]), __fmul_rn(a_sin, ys[i])), cx);
        o.c[2 * i + 1] = __fadd_rn(__fadd_rn(__fmul_rn(-a_sin, xs[i]), __fmul_rn(a_cos, ys[i])), cy);
    }
    o.area = __fmul_rn(r[2], r[3]);
    o.mnx = fminf(fminf(o.c[0], o.c[2]), fminf(o.c[4], o.c[6]));
    o.mxx = fmaxf(fmaxf(o.c[0], o.c[2]), fmaxf(o.c[4], o.c[6]));
    o.mny = fminf(fminf(o.c[1], o.c[3]), fminf(o.c[5], o.c[7]));
    o.mxy = fmaxf(fmaxf(o.c[1], o.c[3]), fmaxf(o.c[5], o.c[7]));
}

// Hulls separated by more than a relative guard band => intersection is exactly 0 in the reference.
__device__ __forceinline__ bool rbox_disjoint(const RBox& a, const RBox& b) {
    const float scale = fmaxf(fmaxf(fabsf(a.mxx), fabsf(a.mnx)), fmaxf(fabsf(a.mxy), fabsf(a.mny)));
    const float eps = 1e-4f * fmaxf(1.f, scale);
    return a.mnx > b.mxx + eps || b.mnx > a.mxx + eps || a.mny > b.mxy + eps || b.mny > a.mxy + eps;
}

// Upper bound of the intersection area of two rectangles given by their corners (rbbox_to_corners order
// (-,-), (-,+), (+,+), (+,-)): the overlap of rectangle A with the bounding box of B taken along A's own axes.
// The true intersection lies inside both, so it cannot be larger.  Used by the large-N NMS to skip polygon
// clips whose IoU cannot exceed the threshold (with a 0.2 % + 1e-6 margin over float32 rounding).
// A zero-size rectangle gives NaN, which never satisfies the skip test.
__device__ __forceinline__ float rect_inter_bound(const float* a, const float* b) {
    const float ux = a[2] - a[0], uy = a[3] - a[1];  // c1 - c0
    const float vx = a[6] - a[0], vy = a[7] - a[1];  // c3 - c0
    const float lu2 = ux * ux + uy * uy, lv2 = vx * vx + vy * vy;
    float umin = 3.0e38f, umax = -3.0e38f, vmin = 3.0e38f, vmax = -3.0e38f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float dx = b[2 * k] - a[0], dy = b[2 * k + 1] - a[1];
        const float pu = dx * ux + dy * uy, pv = dx * vx + dy * vy;  // scaled by |u|, |v|
        umin = fminf(umin, pu); umax = fmaxf(umax, pu);
        vmin = fminf(vmin, pv); vmax = fmaxf(vmax, pv);
    }
    const float ou = fminf(umax, lu2) - fmaxf(umin, 0.f);
    const float ov = fminf(vmax, lv2) - fmaxf(vmin, 0.f);
    if (ou <= 0.f || ov <= 0.f) return 0.f;
    return ou * ov * rsqrtf(lu2 * lv2);
}

// true when IoU(a, b) > thresh is impossible: inter > thresh/(1+thresh) * (area_a + area_b) is needed
__device__ __forceinline__ bool rbox_cannot_exceed(const RBox& a, const RBox& b, float need_frac) {
    const float asum = a.area + b.area;
    const float need = need_frac * asum;
    const float slack = 1e-6f * asum;
    if (rect_inter_bound(a.c, b.c) * 1.002f + slack < need) return true;
    return rect_inter_bound(b.c, a.c) * 1.002f + slack < need;
}

// point_in_quadrilateral, nms_gpu.py:324-340
__device__ __forceinline__ bool pt_in_quad(float px, float py, const float* c) {
    const float ab0 = __fsub_rn(c[2], c[0]), ab1 = __fsub_rn(c[3], c[1]);
    const float ad0 = __fsub_rn(c[6], c[0]), ad1 = __fsub_rn(c[7], c[1]);
    const float ap0 = __fsub_rn(px, c[0]), ap1 = __fsub_rn(py, c[1]);
    const float abab = __fadd_rn(__fmul_rn(ab0, ab0), __fmul_rn(ab1, ab1));
    const float abap = __fadd_rn(__fmul_rn(ab0, ap0), __fmul_rn(ab1, ap1));
    const float adad = __fadd_rn(__fmul_rn(ad0, ad0), __fmul_rn(ad1, ad1));
    const float adap = __fadd_rn(__fmul_rn(ad0, ap0), __fmul_rn(ad1, ap1));
    return abab >= abap && abap >= 0.f && adad >= adap && adap >= 0.f;
}

// line_segment_intersection, nms_gpu.py:236-279
__device__ __forceinline__ bool seg_inter(const float* p1, const float* p2, int i, int j, float& ox, float& oy) {
    const float A0 = p1[2 * i], A1 = p1[2 * i + 1];
    const float B0 = p1[2 * ((i + 1) & 3)], B1 = p1[2 * ((i + 1) & 3) + 1];
    const float C0 = p2[2 * j], C1 = p2[2 * j + 1];
    const float D0 = p2[2 * ((j + 1) & 3)], D1 = p2[2 * ((j + 1) & 3) + 1];
    const float BA0 = __fsub_rn(B0, A0), BA1 = __fsub_rn(B1, A1);
    const float DA0 = __fsub_rn(D0, A0), CA0 = __fsub_rn(C0, A0);
    const float DA1 = __fsub_rn(D1, A1), CA1 = __fsub_rn(C1, A1);
    const bool acd = __fmul_rn(DA1, CA0) > __fmul_rn(CA1, DA0);
    const bool bcd = __fmul_rn(__fsub_rn(D1, B1), __fsub_rn(C0, B0)) > __fmul_rn(__fsub_rn(C1, B1), __fsub_rn(D0, B0));
    if (acd == bcd) return false;
    const bool abc = __fmul_rn(CA1, BA0) > __fmul_rn(BA1, CA0);
    const bool abd = __fmul_rn(DA1, BA0) > __fmul_rn(BA1, DA0);
    if (abc == abd) return false;
    const float DC0 = __fsub_rn(D0, C0), DC1 = __fsub_rn(D1, C1);
    const float ABBA = __fsub_rn(__fmul_rn(A0, B1), __fmul_rn(B0, A1));
    const float CDDC = __fsub_rn(__fmul_rn(C0, D1), __fmul_rn(D0, C1));
    const float DH = __fsub_rn(__fmul_rn(BA1, DC0), __fmul_rn(BA0, DC1));
    const float Dx = __fsub_rn(__fmul_rn(ABBA, DC0), __fmul_rn(BA0, CDDC));
    const float Dy = __fsub_rn(__fmul_rn(ABBA, DC1), __fmul_rn(BA1, CDDC));
    ox = __fdiv_rn(Dx, DH);
    oy = __fdiv_rn(Dy, DH);
    return true;
}

// inter(), nms_gpu.py:393-407: quadrilateral_intersection 343-364, sort_vertex_in_convex_polygon
// 196-233, area 186-193 / trangle_area 180-183.  pts1 = corners of the FIRST argument.
// Points beyond the 8th are ignored (the reference overflows its buffer there: undefined).
__device__ __forceinline__ double rbox_inter(const float* p1, const float* p2) {
    float ip[16];
    int n = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (pt_in_quad(p1[2 * i], p1[2 * i + 1], p2)) {
            if (n < 8) { ip[2 * n] = p1[2 * i]; ip[2 * n + 1] = p1[2 * i + 1]; }
            ++n;
        }
        if (pt_in_quad(p2[2 * i], p2[2 * i + 1], p1)) {
            if (n < 8) { ip[2 * n] = p2[2 * i]; ip[2 * n + 1] = p2[2 * i + 1]; }
            ++n;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float x, y;
            if (seg_inter(p1, p2, i, j, x, y)) {
                if (n < 8) { ip[2 * n] = x; ip[2 * n + 1] = y; }
                ++n;
            }
        }
    n = n > 8 ? 8 : n;
    if (n < 3) return 0.0;  // area() loops over range(n-2)

    // sort_vertex_in_convex_polygon
    float c0 = 0.f, c1 = 0.f;
    for (int i = 0; i < n; ++i) { c0 = __fadd_rn(c0, ip[2 * i]); c1 = __fadd_rn(c1, ip[2 * i + 1]); }
    c0 = (float)((double)c0 / (double)n);
    c1 = (float)((double)c1 / (double)n);
    float vs[8];
    for (int i = 0; i < n; ++i) {
        float v0 = __fsub_rn(ip[2 * i], c0), v1 = __fsub_rn(ip[2 * i + 1], c1);
        const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(v0, v0), __fmul_rn(v1, v1)));
        v0 = __fdiv_rn(v0, d);
        v1 = __fdiv_rn(v1, d);
        if (v1 < 0.f) v0 = (float)(-2.0 - (double)v0);
        vs[i] = v0;
    }
    for (int i = 1; i < n; ++i) {
        if (vs[i - 1] > vs[i]) {
            const float temp = vs[i], tx = ip[2 * i], ty = ip[2 * i + 1];
            int j = i;
            while (j > 0 && vs[j - 1] > temp) {
                vs[j] = vs[j - 1];
                ip[2 * j] = ip[2 * j - 2];
                ip[2 * j + 1] = ip[2 * j - 1];
                --j;
            }
            vs[j] = temp; ip[2 * j] = tx; ip[2 * j + 1] = ty;
        }
    }
    // area: fan from vertex 0, float64 accumulation of |float32 cross| / 2.0
    double s = 0.0;
    for (int i = 0; i < n - 2; ++i) {
        const float* a = ip; const float* b = ip + 2 * i + 2; const float* c = ip + 2 * i + 4;
        const float t0 = __fmul_rn(__fsub_rn(a[0], c[0]), __fsub_rn(b[1], c[1]));
        const float t1 = __fmul_rn(__fsub_rn(a[1], c[1]), __fsub_rn(b[0], c[0]));
        s += fabs((double)__fsub_rn(t0, t1) / 2.0);
    }
    return s;
}

// devRotateIoUEval(rbox1, rbox2, criterion), nms_gpu.py:564-576 (-1 == devRotateIoU 410-415)
__device__ __forceinline__ double rbox_iou(const RBox& b1, const RBox& b2, int criterion) {
    const double ai = rbox_disjoint(b1, b2) ? 0.0 : rbox_inter(b1.c, b2.c);
    if (criterion == -1) return ai / ((double)__fadd_rn(b1.area, b2.area) - ai);
    if (criterion == 0) return ai / (double)b1.area;
    if (criterion == 1) return ai / (double)b2.area;
    return ai;
}

}  // namespace pp
