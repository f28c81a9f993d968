// Per-box float32 arithmetic shared by boxes.cu and predict.cu.  Explicit _rn intrinsics keep numpy's
// two-step multiply/add (no FMA contraction).
#pragma once
#include "pp_common.cuh"

namespace pp {

// second_box_decode, libraries/eval_helper_functions.py:388-461 (default flags): t = encoding,
// a = anchor (x,y,z,w,l,h,r) -> o (x,y,z,w,l,h,r).  o may alias t.
__device__ __forceinline__ void box_decode_one(const float* t, const float* a, float* o) {
    const float xa = a[0], ya = a[1], wa = a[3], la = a[4], ha = a[5], ra = a[6];
    const float za = __fadd_rn(a[2], __fdiv_rn(ha, 2.f));
    const float diag = __fsqrt_rn(__fadd_rn(__fmul_rn(la, la), __fmul_rn(wa, wa)));
    const float xg = __fadd_rn(__fmul_rn(t[0], diag), xa);
    const float yg = __fadd_rn(__fmul_rn(t[1], diag), ya);
    float zg = __fadd_rn(__fmul_rn(t[2], ha), za);
    const float lg = __fmul_rn(expf(t[4]), la);
    const float wg = __fmul_rn(expf(t[3]), wa);
    const float hg = __fmul_rn(expf(t[5]), ha);
    const float rg = __fadd_rn(t[6], ra);
    zg = __fsub_rn(zg, __fdiv_rn(hg, 2.f));
    o[0] = xg; o[1] = yg; o[2] = zg; o[3] = wg; o[4] = lg; o[5] = hg; o[6] = rg;
}

// center_to_corner_box2d + corner_to_standup_nd_jit, load_data.py:1525-1594, 1330-1341:
// rotated BEV box -> (xmin, ymin, xmax, ymax)
__device__ __forceinline__ float4 rbox_standup_one(float cx, float cy, float w, float l, float r) {
    double ds, dc;
    sincos((double)r, &ds, &dc);
    const float s = (float)ds, c = (float)dc;
    const float hx[4] = {-0.5f, -0.5f, 0.5f, 0.5f};
    const float hy[4] = {-0.5f, 0.5f, 0.5f, -0.5f};
    float mnx = 0.f, mny = 0.f, mxx = 0.f, mxy = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float x = __fmul_rn(w, hx[k]), y = __fmul_rn(l, hy[k]);
        const float xr = __fadd_rn(__fadd_rn(__fmul_rn(x, c), __fmul_rn(y, s)), cx);
        const float yr = __fadd_rn(__fadd_rn(__fmul_rn(x, -s), __fmul_rn(y, c)), cy);
        if (k == 0) { mnx = mxx = xr; mny = mxy = yr; }
        else { mnx = fminf(mnx, xr); mxx = fmaxf(mxx, xr); mny = fminf(mny, yr); mxy = fmaxf(mxy, yr); }
    }
    return make_float4(mnx, mny, mxx, mxy);
}

}  // namespace pp
