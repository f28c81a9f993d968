/*
 * pp_oracle.c -- CPU restatement of the reference's PointPillars pre/post hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA path and the
 * "port" CPU baseline that bench.py times beside it.  Nothing in the product package may
 * import, link or call it; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do.
 *
 * Every function cites the reference file:line (paths relative to the reference checkout)
 * whose algorithm it follows.  Pinning status (see DESIGN.md "Oracle"):
 *   - voxelizer, box decode, standup prep, NMS sweep, rotated IoU, anchor mask, eval overlaps,
 *     predict glue: pinned against the reference's own source executed in the build container
 *     (oracle/ref_extract.py, tests/golden/ fixtures, tests/test_oracle_vs_reference.py).
 *   - decoration and scatter: the reference implements them with TensorFlow ops and TensorFlow is
 *     not installable here; its two method bodies are executed unmodified over a numpy stand-in
 *     for those ops (oracle/tf_shim.py) and the oracle matches bit for bit.  Pinned to the
 *     reference's op sequence; TensorFlow's internal float32 summation order is not
 *     ("parity unpinned" in that one respect, covered by the 1e-5 tolerance).
 *   - sensor ingest: numpy/scipy expressions of the reference as written; the ros_numpy step
 *     (third party, not vendored) is restated from its published algorithm.
 *
 * Build: plain C99, `gcc -O2 -ffp-contract=off -fno-fast-math` (no FMA contraction: the numba
 * CPU oracle derived from the reference source does not contract either).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define PPO_API __attribute__((visibility("default")))

/* cosf/sinf of the host libm are within 1 ulp but not always correctly rounded (glibc: 1.3 % of the arguments in
 * [-pi, pi] differ from RN(cos(double))), and neither is the libm behind the reference (numba -> llvm.cos.f32 ->
 * libm on the CPU, libdevice on the GPU).  One ulp on a corner at 70 m moves a car-sized IoU by ~4e-6.  The CUDA
 * path evaluates sincos in float64 and rounds; ppo_set_exact_trig(1) makes the oracle do the same, which isolates
 * that one source of difference in the parity tests. */
static int g_exact_trig = 0;
PPO_API void ppo_set_exact_trig(int on) { g_exact_trig = on; }
static inline void ppo_sincosf(float a, float* s, float* c) {
    if (g_exact_trig) { *s = (float)sin((double)a); *c = (float)cos((double)a); }
    else { *s = sinf(a); *c = cosf(a); }
}


/* ------------------------------------------------------------------------------------------
 * Grid size: load_data.py:612-615 / 722-731 (np.round == round-half-to-even, then int32).
 * arith_f32 != 0 reproduces the case where the caller handed python lists, which the wrapper
 * casts to points.dtype == float32 (load_data.py:726-729).
 * ------------------------------------------------------------------------------------------ */
PPO_API void ppo_grid_size(const double voxel_size[3], const double coors_range[6], int arith_f32,
                           int32_t grid_xyz[3]) {
    for (int j = 0; j < 3; ++j) {
        if (arith_f32) {
            float g = ((float)coors_range[3 + j] - (float)coors_range[j]) / (float)voxel_size[j];
            grid_xyz[j] = (int32_t)nearbyintf(g);
        } else {
            double g = (coors_range[3 + j] - coors_range[j]) / voxel_size[j];
            grid_xyz[j] = (int32_t)nearbyint(g);
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Voxelizer: load_data.py:593-641 (_points_to_voxel_reverse_kernel, reverse_index=True, the
 * one the reference calls at load_data.py:2966) and 643-692 (_points_to_voxel_kernel).
 * Wrapper semantics (zeroed outputs, -1 map, slice to voxel_num): load_data.py:695-771.
 *
 *   points       [N,D]  float32 (is_f64=0) or float64 (is_f64=1), C-contiguous
 *   arith_f32    cell arithmetic in float32 (only when points AND params are float32),
 *                otherwise float64 (numba promotion)
 *   voxels       [max_voxels,max_points,D] same dtype as points, zero-filled here
 *   coors        [max_voxels,3] int32 (z,y,x) if reverse_index else (x,y,z)
 *   num          [max_voxels] int32
 *   point_slot   optional [N] int32: voxel*max_points+slot for a stored point, -1 otherwise
 * returns voxel_num.
 *
 * NaN coordinates are undefined behaviour in the reference (both comparisons at
 * load_data.py:623 are false, then NaN is cast to int32).  Defined here, and in the CUDA path,
 * as "point dropped".
 * ------------------------------------------------------------------------------------------ */
PPO_API int ppo_points_to_voxel(const void* points, int is_f64, int64_t N, int D,
                                const double voxel_size[3], const double coors_range[6],
                                int arith_f32, int max_points, int max_voxels, int reverse_index,
                                void* voxels, int32_t* coors, int32_t* num, int32_t* point_slot) {
    int32_t grid[3];
    ppo_grid_size(voxel_size, coors_range, arith_f32, grid);
    const size_t esz = is_f64 ? 8 : 4;
    const int64_t ncell = (int64_t)grid[0] * grid[1] * grid[2];
    int32_t* map = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ncell > 0 ? ncell : 1));
    if (!map) return -1;
    for (int64_t c = 0; c < ncell; ++c) map[c] = -1;
    memset(voxels, 0, esz * (size_t)max_voxels * max_points * D);
    memset(coors, 0, sizeof(int32_t) * 3 * (size_t)max_voxels);
    memset(num, 0, sizeof(int32_t) * (size_t)max_voxels);
    if (point_slot)
        for (int64_t i = 0; i < N; ++i) point_slot[i] = -1;

    const float* pf = (const float*)points;
    const double* pd = (const double*)points;
    float lo32[3], vs32[3];
    for (int j = 0; j < 3; ++j) {
        lo32[j] = (float)coors_range[j];
        vs32[j] = (float)voxel_size[j];
    }
    int voxel_num = 0;
    for (int64_t i = 0; i < N; ++i) {
        int32_t cxyz[3];
        int failed = 0;
        for (int j = 0; j < 3; ++j) {
            double c;
            if (arith_f32) {
                c = (double)floorf((pf[i * D + j] - lo32[j]) / vs32[j]);
            } else {
                double p = is_f64 ? pd[i * D + j] : (double)pf[i * D + j];
                c = floor((p - coors_range[j]) / voxel_size[j]);
            }
            if (!(c >= 0.0) || !(c < (double)grid[j])) { /* also drops NaN */
                failed = 1;
                break;
            }
            cxyz[j] = (int32_t)c;
        }
        if (failed) continue;
        /* same cell whichever index order the reference's map uses */
        const int64_t cell = ((int64_t)cxyz[2] * grid[1] + cxyz[1]) * grid[0] + cxyz[0];
        int32_t v = map[cell];
        if (v == -1) {
            v = voxel_num;
            if (voxel_num >= max_voxels) break; /* load_data.py:632-633: break, not continue */
            voxel_num += 1;
            map[cell] = v;
            if (reverse_index) {
                coors[3 * v + 0] = cxyz[2];
                coors[3 * v + 1] = cxyz[1];
                coors[3 * v + 2] = cxyz[0];
            } else {
                coors[3 * v + 0] = cxyz[0];
                coors[3 * v + 1] = cxyz[1];
                coors[3 * v + 2] = cxyz[2];
            }
        }
        int32_t n = num[v];
        if (n < max_points) {
            memcpy((char*)voxels + esz * (((size_t)v * max_points + n) * D),
                   (const char*)points + esz * ((size_t)i * D), esz * D);
            num[v] = n + 1;
            if (point_slot) point_slot[i] = v * max_points + n;
        }
    }
    free(map);
    return voxel_num;
}

/* ------------------------------------------------------------------------------------------
 * Pillar decoration: model/pointpillars.py:143-203 (constants 121-124, mask 23-49).
 *   voxels [M,P,D] f32, num_points [M] i32, coors [M,4] i32 (batch,z,y,x)
 *   out    [M,P,D+5] f32 = concat(voxels, xyz - mean, (x - cx, y - cy)) * (p < num_points)
 * vx, vy, x_offset, y_offset are python doubles in the reference that TensorFlow converts to
 * float32 constants; the multiply and the add are separate TF ops (no FMA).
 * The mean is sum over ALL P slots / num_points (padding slots are zero).  TF's reduction order
 * is unspecified; this restatement adds slots in order 0..P-1.  PARITY UNPINNED (no TF here).
 * ------------------------------------------------------------------------------------------ */
PPO_API void ppo_decorate(const float* voxels, const int32_t* num_points, const int32_t* coors,
                          int64_t M, int P, int D, double vx, double vy, double x_offset,
                          double y_offset, float* out) {
    const float vxf = (float)vx, vyf = (float)vy, xof = (float)x_offset, yof = (float)y_offset;
    const int Do = D + 5;
    for (int64_t m = 0; m < M; ++m) {
        const float* v = voxels + (size_t)m * P * D;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        for (int p = 0; p < P; ++p) {
            s0 += v[p * D + 0];
            s1 += v[p * D + 1];
            s2 += v[p * D + 2];
        }
        const float n = (float)num_points[m];
        const float m0 = s0 / n, m1 = s1 / n, m2 = s2 / n;
        float ex = (float)coors[4 * m + 3] * vxf;
        ex = ex + xof;
        float ey = (float)coors[4 * m + 2] * vyf;
        ey = ey + yof;
        for (int p = 0; p < P; ++p) {
            const float mask = (num_points[m] > p) ? 1.f : 0.f;
            float* o = out + ((size_t)m * P + p) * Do;
            for (int d = 0; d < D; ++d) o[d] = v[p * D + d] * mask;
            o[D + 0] = (v[p * D + 0] - m0) * mask;
            o[D + 1] = (v[p * D + 1] - m1) * mask;
            o[D + 2] = (v[p * D + 2] - m2) * mask;
            o[D + 3] = (v[p * D + 0] - ex) * mask;
            o[D + 4] = (v[p * D + 1] - ey) * mask;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Scatter: model/pointpillars.py:285-341.  Per batch element b: rows with coords[:,0]==b,
 * index = y*nx + x (z ignored, line 302), tf.scatter_nd SUMS duplicates (line 317), result
 * transposed to [C, ny*nx], stacked, reshaped to [B,C,ny,nx] (NCHW).
 * layout_nhwc != 0 writes [B,ny,nx,C] instead (what the RPN transposes to, voxelnet.py:697).
 * Rows whose batch index is outside [0,B) are ignored (boolean_mask never selects them).
 * Duplicates are added in row order.  PARITY UNPINNED (TensorFlow op).
 * ------------------------------------------------------------------------------------------ */
PPO_API void ppo_scatter(const float* feats, const int32_t* coords, int64_t M, int C, int B, int ny,
                         int nx, int layout_nhwc, float* out) {
    memset(out, 0, sizeof(float) * (size_t)B * C * ny * nx);
    for (int64_t m = 0; m < M; ++m) {
        const int b = coords[4 * m + 0], y = coords[4 * m + 2], x = coords[4 * m + 3];
        if (b < 0 || b >= B) continue;
        if (y < 0 || y >= ny || x < 0 || x >= nx) continue; /* tf.scatter_nd on GPU drops OOB */
        for (int c = 0; c < C; ++c) {
            size_t o = layout_nhwc ? ((((size_t)b * ny + y) * nx + x) * C + c)
                                   : ((((size_t)b * C + c) * ny + y) * nx + x);
            out[o] += feats[(size_t)m * C + c];
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Box decode: libraries/eval_helper_functions.py:388-461 (default flags), float32 numpy.
 * Output order [x,y,z,w,l,h,r].
 * ------------------------------------------------------------------------------------------ */
PPO_API void ppo_second_box_decode(const float* enc, const float* anchors, int64_t N, float* out) {
    for (int64_t i = 0; i < N; ++i) {
        const float* a = anchors + 7 * i;
        const float* t = enc + 7 * i;
        const float xa = a[0], ya = a[1], wa = a[3], la = a[4], ha = a[5], ra = a[6];
        float za = a[2];
        za = za + ha / 2.f;
        const float l2 = la * la, w2 = wa * wa;
        const float diagonal = sqrtf(l2 + w2);
        float xg = t[0] * diagonal;
        xg = xg + xa;
        float yg = t[1] * diagonal;
        yg = yg + ya;
        float zg = t[2] * ha;
        zg = zg + za;
        const float lg = expf(t[4]) * la;
        const float wg = expf(t[3]) * wa;
        const float hg = expf(t[5]) * ha;
        const float rg = t[6] + ra;
        zg = zg - hg / 2.f;
        float* o = out + 7 * i;
        o[0] = xg; o[1] = yg; o[2] = zg; o[3] = wg; o[4] = lg; o[5] = hg; o[6] = rg;
    }
}

/* ------------------------------------------------------------------------------------------
 * Rotated BEV box -> axis-aligned "standup" box: load_data.py:1525-1546 (center_to_corner_box2d),
 * 1563-1594 (corners_nd, origin 0.5, corner order (-,-),(-,+),(+,+),(+,-)), 1548-1561
 * (rotation_2d: x' = x cos + y sin, y' = -x sin + y cos), 1330-1341 (min/max).
 *   boxes [N,5] f32 (x,y,w,l,r) -> out [N,4] f32 (xmin,ymin,xmax,ymax).  Call site:
 *   model/voxelnet.py:1233-1249.
 * ------------------------------------------------------------------------------------------ */
PPO_API void ppo_rbox_to_standup(const float* boxes, int64_t N, float* out) {
    static const float cn[4][2] = {{-0.5f, -0.5f}, {-0.5f, 0.5f}, {0.5f, 0.5f}, {0.5f, -0.5f}};
    for (int64_t i = 0; i < N; ++i) {
        const float* b = boxes + 5 * i;
        float s, c;
        ppo_sincosf(b[4], &s, &c);
        float mnx = 0, mny = 0, mxx = 0, mxy = 0;
        for (int k = 0; k < 4; ++k) {
            const float x = b[2] * cn[k][0], y = b[3] * cn[k][1];
            float xr = x * c;
            xr = xr + y * s;
            float yr = x * (-s);
            yr = yr + y * c;
            xr = xr + b[0];
            yr = yr + b[1];
            if (k == 0) { mnx = mxx = xr; mny = mxy = yr; }
            else {
                mnx = xr < mnx ? xr : mnx; mxx = xr > mxx ? xr : mxx;
                mny = yr < mny ? yr : mny; mxy = yr > mxy ? yr : mxy;
            }
        }
        out[4 * i + 0] = mnx; out[4 * i + 1] = mny; out[4 * i + 2] = mxx; out[4 * i + 3] = mxy;
    }
}

/* ------------------------------------------------------------------------------------------
 * Score ordering used by every NMS entry point: `order = scores.argsort()[::-1]`
 * (eval_helper_functions.py:510, nms_gpu.py:138,473).  numpy's default sort is unstable, so the
 * reference defines no order for ties; the rule fixed here (and in the CUDA path) is
 * "descending score, ties by descending original index" == argsort(kind="stable")[::-1].
 * NaN scores sort as larger than everything (numpy puts NaN last before the reversal).
 * ------------------------------------------------------------------------------------------ */
typedef struct { uint32_t key; int32_t idx; } ppo_kv;

static uint32_t ppo_score_key(float s) {
    uint32_t u;
    if (s != s) return 0xFFFFFFFFu;
    memcpy(&u, &s, 4);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static int ppo_kv_desc(const void* a, const void* b) {
    const ppo_kv* x = (const ppo_kv*)a; const ppo_kv* y = (const ppo_kv*)b;
    if (x->key != y->key) return x->key > y->key ? -1 : 1;
    return x->idx > y->idx ? -1 : (x->idx < y->idx ? 1 : 0);
}
PPO_API void ppo_argsort_desc(const float* scores, int64_t N, int32_t* order) {
    ppo_kv* kv = (ppo_kv*)malloc(sizeof(ppo_kv) * (size_t)(N > 0 ? N : 1));
    for (int64_t i = 0; i < N; ++i) { kv[i].key = ppo_score_key(scores[i]); kv[i].idx = (int32_t)i; }
    qsort(kv, (size_t)N, sizeof(ppo_kv), ppo_kv_desc);
    for (int64_t i = 0; i < N; ++i) order[i] = kv[i].idx;
    free(kv);
}

/* ------------------------------------------------------------------------------------------
 * Greedy sweep over the suppression bitmask: nms_postprocess, eval_helper_functions.py:529-546
 * (identical copy nms_gpu.py:111-128).  mask is [n, col_blocks] uint64, row i word j holds the
 * bits of boxes 64j..64j+63 that box i suppresses.
 * ------------------------------------------------------------------------------------------ */
PPO_API int ppo_nms_postprocess(const uint64_t* mask, int64_t n, int32_t* keep_out) {
    const int64_t cb = (n + 63) / 64;
    uint64_t* remv = (uint64_t*)calloc((size_t)(cb > 0 ? cb : 1), sizeof(uint64_t));
    int nk = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t nb = i / 64; const int ib = (int)(i % 64);
        if (!(remv[nb] & (1ULL << ib))) {
            keep_out[nk++] = (int32_t)i;
            for (int64_t j = nb; j < cb; ++j) remv[j] |= mask[i * cb + j];
        }
    }
    free(remv);
    return nk;
}

/* ------------------------------------------------------------------------------------------
 * Axis-aligned IoU with the pixel "+1" convention: iou_device, eval_helper_functions.py:553-564.
 * numba promotes `float32 - float32 + 1` to float64 at the `+ 1`: the differences
 * are rounded to float32 first, everything after is float64.
 * ------------------------------------------------------------------------------------------ */
static double ppo_iou_standup(const float* a, const float* b) {
    const float left = a[0] > b[0] ? a[0] : b[0];
    const float right = a[2] < b[2] ? a[2] : b[2];
    const float top = a[1] > b[1] ? a[1] : b[1];
    const float bottom = a[3] < b[3] ? a[3] : b[3];
    const float dw = right - left, dh = bottom - top;
    double width = (double)dw + 1.0; if (!(width > 0.)) width = 0.;
    double height = (double)dh + 1.0; if (!(height > 0.)) height = 0.;
    const double interS = width * height;
    const float aw = a[2] - a[0], ah = a[3] - a[1], bw = b[2] - b[0], bh = b[3] - b[1];
    const double Sa = ((double)aw + 1.0) * ((double)ah + 1.0);
    const double Sb = ((double)bw + 1.0) * ((double)bh + 1.0);
    return interS / (Sa + Sb - interS);
}

/* nms_kernel, eval_helper_functions.py:567-598: bit (i,j) set iff j>i (within the diagonal
 * tile; every j in later tiles; earlier tiles are computed too but never read by the sweep)
 * and iou > thresh (thresh is a float32 kernel argument, compared in float64).
 * boxes_sorted [n,4] f32 already in score order.  Only the upper triangle is filled. */
PPO_API void ppo_standup_mask(const float* boxes_sorted, int64_t n, float thresh, uint64_t* mask) {
    const int64_t cb = (n + 63) / 64;
    memset(mask, 0, sizeof(uint64_t) * (size_t)(n * cb));
    const double th = (double)thresh;
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = i + 1; j < n; ++j)
            if (ppo_iou_standup(boxes_sorted + 4 * i, boxes_sorted + 4 * j) > th)
                mask[i * cb + j / 64] |= 1ULL << (j % 64);
}

/* nms(): eval_helper_functions.py:463-492 with nms_gpu 494-527.
 * pre_max_size/post_max_size < 0 mean None.  The reference picks the pre_max_size best scores
 * with np.argpartition (any order), then sorts them; the kept ORDER is therefore the sorted one.
 * keep_out receives indices into the caller's arrays; returns the count (0 <=> reference None). */
PPO_API int ppo_nms_standup(const float* bboxes, const float* scores, int64_t N, int pre_max_size,
                            int post_max_size, float thresh, int64_t* keep_out) {
    if (N <= 0) return 0;
    int32_t* order = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
    ppo_argsort_desc(scores, N, order);
    int64_t n = N;
    if (pre_max_size >= 0 && pre_max_size < n) n = pre_max_size;
    if (n == 0) { free(order); return 0; }
    float* sb = (float*)malloc(sizeof(float) * 4 * (size_t)n);
    for (int64_t i = 0; i < n; ++i) memcpy(sb + 4 * i, bboxes + 4 * (size_t)order[i], 16);
    const int64_t cb = (n + 63) / 64;
    uint64_t* mask = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(n * cb));
    int32_t* keep = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    ppo_standup_mask(sb, n, thresh, mask);
    int nk = ppo_nms_postprocess(mask, n, keep);
    if (post_max_size >= 0 && nk > post_max_size) nk = post_max_size;
    for (int i = 0; i < nk; ++i) keep_out[i] = order[keep[i]];
    free(order); free(sb); free(mask); free(keep);
    return nk;
}

/* ------------------------------------------------------------------------------------------
 * Rotated IoU: second/core/non_max_suppression/nms_gpu.py:180-415 (+564-576 for criterion).
 * Float32 with the float64 islands numba's typing creates:
 *   trangle_area  180-183   f32 products/differences, "/ 2.0" in f64
 *   area          186-193   f64 accumulator
 *   sort_vertex   196-233   centre sum f32, "/ num" in f64 stored f32; key "-2 - x" f64 stored f32
 *   line_segment_intersection 236-279   f32, strict ">" orientation tests
 *   point_in_quadrilateral    324-340   f32, inclusive tests
 *   quadrilateral_intersection 343-364  interleaved corner tests then 4x4 edge pairs
 *   rbbox_to_corners 367-390  "x/2" exact, f32 rotate (cosf/sinf)
 *   devRotateIoU 410-415    areas f32, ratio f64
 * The reference's intersection buffer holds 8 points (line 397); coincident boxes overflow it
 * (undefined).  Defined here, and in the CUDA path, as "points beyond the 8th are ignored".
 * ------------------------------------------------------------------------------------------ */
static double ppo_trangle_area(const float* a, const float* b, const float* c) {
    const float t0 = (a[0] - c[0]) * (b[1] - c[1]);
    const float t1 = (a[1] - c[1]) * (b[0] - c[0]);
    const float d = t0 - t1;
    return (double)d / 2.0;
}

static double ppo_poly_area(const float* p, int n) {
    double s = 0.0;
    for (int i = 0; i < n - 2; ++i) s += fabs(ppo_trangle_area(p, p + 2 * i + 2, p + 2 * i + 4));
    return s;
}

static void ppo_sort_vertex(float* p, int n) {
    if (n <= 0) return;
    float c0 = 0.f, c1 = 0.f;
    for (int i = 0; i < n; ++i) { c0 += p[2 * i]; c1 += p[2 * i + 1]; }
    c0 = (float)((double)c0 / (double)n);
    c1 = (float)((double)c1 / (double)n);
    float vs[16];
    for (int i = 0; i < n; ++i) {
        float v0 = p[2 * i] - c0, v1 = p[2 * i + 1] - c1;
        const float q0 = v0 * v0, q1 = v1 * v1;
        const float d = sqrtf(q0 + q1);
        v0 = v0 / d;
        v1 = v1 / d;
        if (v1 < 0) v0 = (float)(-2.0 - (double)v0);
        vs[i] = v0;
    }
    for (int i = 1; i < n; ++i) {
        if (vs[i - 1] > vs[i]) {
            const float temp = vs[i], tx = p[2 * i], ty = p[2 * i + 1];
            int j = i;
            while (j > 0 && vs[j - 1] > temp) {
                vs[j] = vs[j - 1];
                p[2 * j] = p[2 * j - 2];
                p[2 * j + 1] = p[2 * j - 1];
                --j;
            }
            vs[j] = temp; p[2 * j] = tx; p[2 * j + 1] = ty;
        }
    }
}

static int ppo_seg_inter(const float* p1, const float* p2, int i, int j, float* out) {
    const float A0 = p1[2 * i], A1 = p1[2 * i + 1];
    const float B0 = p1[2 * ((i + 1) % 4)], B1 = p1[2 * ((i + 1) % 4) + 1];
    const float C0 = p2[2 * j], C1 = p2[2 * j + 1];
    const float D0 = p2[2 * ((j + 1) % 4)], D1 = p2[2 * ((j + 1) % 4) + 1];
    const float BA0 = B0 - A0, BA1 = B1 - A1, DA0 = D0 - A0, CA0 = C0 - A0, DA1 = D1 - A1,
                CA1 = C1 - A1;
    const float l0 = DA1 * CA0, r0 = CA1 * DA0;
    const int acd = l0 > r0;
    const float l1 = (D1 - B1) * (C0 - B0), r1 = (C1 - B1) * (D0 - B0);
    const int bcd = l1 > r1;
    if (acd != bcd) {
        const float l2 = CA1 * BA0, r2 = BA1 * CA0;
        const int abc = l2 > r2;
        const float l3 = DA1 * BA0, r3 = BA1 * DA0;
        const int abd = l3 > r3;
        if (abc != abd) {
            const float DC0 = D0 - C0, DC1 = D1 - C1;
            const float m0 = A0 * B1, m1 = B0 * A1; const float ABBA = m0 - m1;
            const float m2 = C0 * D1, m3 = D0 * C1; const float CDDC = m2 - m3;
            const float h0 = BA1 * DC0, h1 = BA0 * DC1; const float DH = h0 - h1;
            const float x0 = ABBA * DC0, x1 = BA0 * CDDC; const float Dx = x0 - x1;
            const float y0 = ABBA * DC1, y1 = BA1 * CDDC; const float Dy = y0 - y1;
            out[0] = Dx / DH;
            out[1] = Dy / DH;
            return 1;
        }
    }
    return 0;
}

static int ppo_pt_in_quad(float px, float py, const float* c) {
    const float ab0 = c[2] - c[0], ab1 = c[3] - c[1];
    const float ad0 = c[6] - c[0], ad1 = c[7] - c[1];
    const float ap0 = px - c[0], ap1 = py - c[1];
    float t0, t1;
    t0 = ab0 * ab0; t1 = ab1 * ab1; const float abab = t0 + t1;
    t0 = ab0 * ap0; t1 = ab1 * ap1; const float abap = t0 + t1;
    t0 = ad0 * ad0; t1 = ad1 * ad1; const float adad = t0 + t1;
    t0 = ad0 * ap0; t1 = ad1 * ap1; const float adap = t0 + t1;
    return abab >= abap && abap >= 0 && adad >= adap && adap >= 0;
}

static int ppo_quad_inter(const float* p1, const float* p2, float* ip) {
    int n = 0;
    for (int i = 0; i < 4; ++i) {
        if (ppo_pt_in_quad(p1[2 * i], p1[2 * i + 1], p2)) {
            if (n < 8) { ip[2 * n] = p1[2 * i]; ip[2 * n + 1] = p1[2 * i + 1]; }
            ++n;
        }
        if (ppo_pt_in_quad(p2[2 * i], p2[2 * i + 1], p1)) {
            if (n < 8) { ip[2 * n] = p2[2 * i]; ip[2 * n + 1] = p2[2 * i + 1]; }
            ++n;
        }
    }
    float t[2];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (ppo_seg_inter(p1, p2, i, j, t)) {
                if (n < 8) { ip[2 * n] = t[0]; ip[2 * n + 1] = t[1]; }
                ++n;
            }
    return n > 8 ? 8 : n;
}

static void ppo_rbbox_to_corners(float* corners, const float* r) {
    float a_sin, a_cos;
    ppo_sincosf(r[4], &a_sin, &a_cos);
    const float cx = r[0], cy = r[1];
    const float hx = (float)((double)r[2] / 2.0), hy = (float)((double)r[3] / 2.0);
    const float xs[4] = {-hx, -hx, hx, hx};
    const float ys[4] = {-hy, hy, hy, -hy};
    for (int i = 0; i < 4; ++i) {
        float t0 = a_cos * xs[i], t1 = a_sin * ys[i];
        float s = t0 + t1;
        corners[2 * i] = s + cx;
        t0 = (-a_sin) * xs[i]; t1 = a_cos * ys[i];
        s = t0 + t1;
        corners[2 * i + 1] = s + cy;
    }
}

static double ppo_inter(const float* r1, const float* r2) {
    float c1[8], c2[8], ip[16];
    ppo_rbbox_to_corners(c1, r1);
    ppo_rbbox_to_corners(c2, r2);
    const int n = ppo_quad_inter(c1, c2, ip);
    ppo_sort_vertex(ip, n);
    return ppo_poly_area(ip, n);
}

/* devRotateIoUEval, nms_gpu.py:564-576; criterion -1 is devRotateIoU (410-415). */
PPO_API double ppo_rotate_iou_pair(const float* r1, const float* r2, int criterion) {
    const float area1 = r1[2] * r1[3];
    const float area2 = r2[2] * r2[3];
    const double ai = ppo_inter(r1, r2);
    if (criterion == -1) { const float s = area1 + area2; return ai / ((double)s - ai); }
    if (criterion == 0) return ai / (double)area1;
    if (criterion == 1) return ai / (double)area2;
    return ai;
}

/* rotate_iou_gpu / rotate_iou_gpu_eval: nms_gpu.py:526-561, 618-653 with kernels 493-523,
 * 579-615.  out[n*K + k] = devRotateIoUEval(query_boxes[k], boxes[n]) rounded to float32. */
PPO_API void ppo_rotate_iou(const float* boxes, int64_t N, const float* qboxes, int64_t K,
                            int criterion, float* out) {
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; ++n)
        for (int64_t k = 0; k < K; ++k)
            out[n * K + k] = (float)ppo_rotate_iou_pair(qboxes + 5 * k, boxes + 5 * n, criterion);
}

/* rotate_nms_kernel, nms_gpu.py:419-452: dets_sorted [n,6] (x,y,w,l,r,score) in score order;
 * bit (i,j), j>i, set iff devRotateIoU(row i, col j) > thresh.  Upper triangle only. */
PPO_API void ppo_rotate_mask(const float* dets_sorted, int64_t n, float thresh, uint64_t* mask) {
    const int64_t cb = (n + 63) / 64;
    memset(mask, 0, sizeof(uint64_t) * (size_t)(n * cb));
    const double th = (double)thresh;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = i + 1; j < n; ++j)
            if (ppo_rotate_iou_pair(dets_sorted + 6 * i, dets_sorted + 6 * j, -1) > th)
                mask[i * cb + j / 64] |= 1ULL << (j % 64);
}

/* rotate_nms_gpu, nms_gpu.py:455-490, plus the optional pre/post caps of nms()
 * (eval_helper_functions.py:463-492) so the same entry serves the "full path" config.
 * dets [N,6] f32; keep_out gets original indices in keep order; returns the count. */
PPO_API int ppo_rotate_nms(const float* dets, int64_t N, float thresh, int pre_max_size,
                           int post_max_size, int64_t* keep_out) {
    if (N <= 0) return 0;
    float* sc = (float*)malloc(sizeof(float) * (size_t)N);
    for (int64_t i = 0; i < N; ++i) sc[i] = dets[6 * i + 5];
    int32_t* order = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
    ppo_argsort_desc(sc, N, order);
    int64_t n = N;
    if (pre_max_size >= 0 && pre_max_size < n) n = pre_max_size;
    if (n == 0) { free(sc); free(order); return 0; }
    float* sd = (float*)malloc(sizeof(float) * 6 * (size_t)n);
    for (int64_t i = 0; i < n; ++i) memcpy(sd + 6 * i, dets + 6 * (size_t)order[i], 24);
    const int64_t cb = (n + 63) / 64;
    uint64_t* mask = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(n * cb));
    int32_t* keep = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    ppo_rotate_mask(sd, n, thresh, mask);
    int nk = ppo_nms_postprocess(mask, n, keep);
    if (post_max_size >= 0 && nk > post_max_size) nk = post_max_size;
    for (int i = 0; i < nk; ++i) keep_out[i] = order[keep[i]];
    free(sc); free(order); free(sd); free(mask); free(keep);
    return nk;
}


/* ------------------------------------------------------------------------------------------
 * "Next" row N1 -- anchor mask (load_data.py:3043-3072).
 *   rbbox2d_to_near_bbox   load_data.py:534-548 (limit_period 805-806, center_to_minmax_2d_0_5 550-551):
 *       float32 numpy arithmetic on anchors[:, [0,1,3,4,6]]
 *   fused_get_anchors_area load_data.py:558-584: cell = floor((bv - offset) / stride) in float64
 *       (float32 bv, float64 offset/stride), stored to int32, clipped to the grid
 *   sparse_sum_for_anchors_mask 586-591 + cumsum(0).cumsum(1): pillar counts per (y,x), integral image
 *   mask = area > anchor_area_threshold
 * ------------------------------------------------------------------------------------------ */
PPO_API void ppo_rbbox2d_to_near_bbox(const float* rb, int64_t N, float* out) {
    const float pi = (float)M_PI;
    for (int64_t i = 0; i < N; ++i) {
        const float* r = rb + 5 * i;
        float t = r[4] / pi;
        t = t + 0.5f;
        t = floorf(t) * pi;
        const float lim = fabsf(r[4] - t);
        const int cond = lim > (float)(M_PI / 4);
        const float dx = cond ? r[3] : r[2], dy = cond ? r[2] : r[3];
        out[4 * i + 0] = r[0] - dx / 2.f;
        out[4 * i + 1] = r[1] - dy / 2.f;
        out[4 * i + 2] = r[0] + dx / 2.f;
        out[4 * i + 3] = r[1] + dy / 2.f;
    }
}

/* cell rectangle of every anchor: (x0,y0,x1,y1) as fused_get_anchors_area computes them */
PPO_API void ppo_anchor_cells(const float* anchors /*[A,7]*/, int64_t A, const double voxel_size[3],
                              const double coors_range[6], const int32_t grid_xyz[3], int32_t* cells /*[A,4]*/) {
    for (int64_t i = 0; i < A; ++i) {
        const float* a = anchors + 7 * i;
        const float rb[5] = {a[0], a[1], a[3], a[4], a[6]};
        float bv[4];
        ppo_rbbox2d_to_near_bbox(rb, 1, bv);
        /* load_data.py:569-580.  One side of each index is clipped; an index that stays negative wraps around
         * once (numba negative indexing).  Indices the reference cannot form or that fall outside dense_map
         * (NaN / infinite / far-off anchors: undefined behaviour there) are DEFINED here as the rectangle
         * (-1,-1,-1,-1) = area 0. */
        const double v[4] = {floor(((double)bv[0] - coors_range[0]) / voxel_size[0]), floor(((double)bv[1] - coors_range[1]) / voxel_size[1]),
                             floor(((double)bv[2] - coors_range[0]) / voxel_size[0]), floor(((double)bv[3] - coors_range[1]) / voxel_size[1])};
        int32_t* c = cells + 4 * i;
        c[0] = c[1] = c[2] = c[3] = -1;
        if (fabs(v[0]) < 2.0e9 && fabs(v[1]) < 2.0e9 && fabs(v[2]) < 2.0e9 && fabs(v[3]) < 2.0e9) {
            const int32_t nx = grid_xyz[0], ny = grid_xyz[1];
            int32_t c0 = (int32_t)v[0], c1 = (int32_t)v[1], c2 = (int32_t)v[2], c3 = (int32_t)v[3];
            c0 = c0 > 0 ? c0 : 0;
            c1 = c1 > 0 ? c1 : 0;
            c2 = c2 < nx - 1 ? c2 : nx - 1;
            c3 = c3 < ny - 1 ? c3 : ny - 1;
            if (c2 < 0) c2 += nx;
            if (c3 < 0) c3 += ny;
            if (c0 < nx && c1 < ny && c2 >= 0 && c3 >= 0) { c[0] = c0; c[1] = c1; c[2] = c2; c[3] = c3; }
        }
    }
}

/* coors [M,3] (z,y,x) of ONE frame -> area [A] float32, mask [A] uint8 */
PPO_API void ppo_anchors_mask(const int32_t* coors, int64_t M, int ny, int nx, const int32_t* cells, int64_t A,
                              float threshold, float* area, uint8_t* mask) {
    float* map = (float*)calloc((size_t)ny * nx, sizeof(float));
    for (int64_t m = 0; m < M; ++m) map[(size_t)coors[3 * m + 1] * nx + coors[3 * m + 2]] += 1.f;
    for (int y = 1; y < ny; ++y)
        for (int x = 0; x < nx; ++x) map[(size_t)y * nx + x] += map[(size_t)(y - 1) * nx + x];
    for (int y = 0; y < ny; ++y)
        for (int x = 1; x < nx; ++x) map[(size_t)y * nx + x] += map[(size_t)y * nx + x - 1];
    for (int64_t i = 0; i < A; ++i) {
        const int32_t* c = cells + 4 * i;
        float v = 0.f;
        if (c[0] >= 0) {
            const float ID = map[(size_t)c[3] * nx + c[2]], IA = map[(size_t)c[1] * nx + c[0]];
            const float IB = map[(size_t)c[3] * nx + c[0]], IC = map[(size_t)c[1] * nx + c[2]];
            v = ID - IB - IC + IA;
        }
        area[i] = v;
        mask[i] = v > threshold;
    }
    free(map);
}


/* ------------------------------------------------------------------------------------------
 * "Next" row N4 -- KITTI-eval 3-D overlap: d3_box_overlap / d3_box_overlap_kernel,
 * second/utils/eval.py:131-163 (bev_box_overlap, 126-128, is rotate_iou_gpu_eval itself).
 * boxes [N,7], qboxes [K,7] float64 camera boxes (x,y,z,l,h,w,ry); BEV columns [0,2,3,5,6] go through
 * the float32 rotated intersection (criterion 2), the height overlap and the ratio are float64
 * (numba promotes float64 boxes x float32 rinc), the result is stored back into the float32 matrix.
 * ------------------------------------------------------------------------------------------ */
PPO_API void ppo_d3_box_overlap(const double* boxes, int64_t N, const double* qboxes, int64_t K, int criterion,
                                float* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        const double* b = boxes + 7 * i;
        const float rb[5] = {(float)b[0], (float)b[2], (float)b[3], (float)b[5], (float)b[6]};
        for (int64_t j = 0; j < K; ++j) {
            const double* q = qboxes + 7 * j;
            const float rq[5] = {(float)q[0], (float)q[2], (float)q[3], (float)q[5], (float)q[6]};
            float rinc = (float)ppo_rotate_iou_pair(rq, rb, 2);
            if (rinc > 0) {
                const double top = b[1] < q[1] ? b[1] : q[1];
                const double b0 = b[1] - b[4], q0 = q[1] - q[4];
                const double iw = top - (b0 > q0 ? b0 : q0);
                if (iw > 0) {
                    const double area1 = b[3] * b[4] * b[5], area2 = q[3] * q[4] * q[5];
                    const double inc = iw * (double)rinc;
                    double ua;
                    if (criterion == -1) ua = area1 + area2 - inc;
                    else if (criterion == 0) ua = area1;
                    else if (criterion == 1) ua = area2;
                    else ua = 1.0;
                    rinc = (float)(inc / ua);
                } else {
                    rinc = 0.f;
                }
            }
            out[i * K + j] = rinc;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Whole hot path over a batch of frames, used ONLY as the timed CPU baseline (bench.py).
 * Frames are independent, so the batch is an OpenMP loop over frames; inside a frame
 * the reference's voxelizer is single-threaded by construction.
 * Stand-ins for the TF layers between the stages (PFN Dense, RPN) are inputs, as in
 * pfn_feats [max_voxels,C] (first M rows used), box_enc/anchors [A,7],
 * scores [A].  Returns total kept detections; det_out [B,post_max,8] (box7 + score).
 * ------------------------------------------------------------------------------------------ */
PPO_API int64_t ppo_full_path_batch(const void* points, int is_f64, const int64_t* frame_off, int B,
                                    int D, const double voxel_size[3], const double coors_range[6],
                                    int max_points, int max_voxels, const float* pfn_feats, int C,
                                    const float* box_enc, const float* anchors, const float* scores,
                                    int64_t A, int pre_max, int post_max, float thresh, int rotated,
                                    int nthreads, float* det_out, int32_t* det_count,
                                    int64_t* voxel_count) {
    int32_t grid[3];
    ppo_grid_size(voxel_size, coors_range, 0, grid);
    const int nx = grid[0], ny = grid[1];
    const size_t esz = is_f64 ? 8 : 4;
    const double xo = voxel_size[0] / 2 + coors_range[0], yo = voxel_size[1] / 2 + coors_range[1];
    int64_t total = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
    for (int b = 0; b < B; ++b) {
        const int64_t n0 = frame_off[b], n = frame_off[b + 1] - n0;
        void* vox = malloc(esz * (size_t)max_voxels * max_points * D);
        int32_t* co = (int32_t*)malloc(sizeof(int32_t) * 3 * (size_t)max_voxels);
        int32_t* nm = (int32_t*)malloc(sizeof(int32_t) * (size_t)max_voxels);
        const int M = ppo_points_to_voxel((const char*)points + esz * (size_t)n0 * D, is_f64, n, D,
                                          voxel_size, coors_range, 0, max_points, max_voxels, 1, vox,
                                          co, nm, NULL);
        /* tf.data casts voxels to float32 (load_data.py:2339-2349) */
        float* vf = (float*)malloc(sizeof(float) * (size_t)(M > 0 ? M : 1) * max_points * D);
        if (is_f64) for (size_t i = 0; i < (size_t)M * max_points * D; ++i) vf[i] = (float)((double*)vox)[i];
        else memcpy(vf, vox, sizeof(float) * (size_t)M * max_points * D);
        int32_t* c4 = (int32_t*)malloc(sizeof(int32_t) * 4 * (size_t)(M > 0 ? M : 1));
        for (int m = 0; m < M; ++m) { c4[4 * m] = 0; c4[4 * m + 1] = co[3 * m]; c4[4 * m + 2] = co[3 * m + 1]; c4[4 * m + 3] = co[3 * m + 2]; }
        float* dec = (float*)malloc(sizeof(float) * (size_t)(M > 0 ? M : 1) * max_points * (D + 5));
        ppo_decorate(vf, nm, c4, M, max_points, D, voxel_size[0], voxel_size[1], xo, yo, dec);
        float* canvas = (float*)malloc(sizeof(float) * (size_t)C * ny * nx);
        ppo_scatter(pfn_feats, c4, M, C, 1, ny, nx, 0, canvas);
        float* boxes = (float*)malloc(sizeof(float) * 7 * (size_t)A);
        ppo_second_box_decode(box_enc + 7 * (size_t)A * b, anchors, A, boxes);
        int64_t* keep = (int64_t*)malloc(sizeof(int64_t) * (size_t)A);
        int nk;
        if (rotated) {
            float* dets = (float*)malloc(sizeof(float) * 6 * (size_t)A);
            for (int64_t i = 0; i < A; ++i) {
                dets[6 * i] = boxes[7 * i]; dets[6 * i + 1] = boxes[7 * i + 1]; dets[6 * i + 2] = boxes[7 * i + 3];
                dets[6 * i + 3] = boxes[7 * i + 4]; dets[6 * i + 4] = boxes[7 * i + 6]; dets[6 * i + 5] = scores[A * b + i];
            }
            nk = ppo_rotate_nms(dets, A, thresh, pre_max, post_max, keep);
            free(dets);
        } else {
            float* rb = (float*)malloc(sizeof(float) * 5 * (size_t)A);
            float* sb = (float*)malloc(sizeof(float) * 4 * (size_t)A);
            for (int64_t i = 0; i < A; ++i) {
                rb[5 * i] = boxes[7 * i]; rb[5 * i + 1] = boxes[7 * i + 1]; rb[5 * i + 2] = boxes[7 * i + 3];
                rb[5 * i + 3] = boxes[7 * i + 4]; rb[5 * i + 4] = boxes[7 * i + 6];
            }
            ppo_rbox_to_standup(rb, A, sb);
            nk = ppo_nms_standup(sb, scores + A * b, A, pre_max, post_max, thresh, keep);
            free(rb); free(sb);
        }
        for (int k = 0; k < nk; ++k) {
            memcpy(det_out + ((size_t)b * post_max + k) * 8, boxes + 7 * keep[k], 28);
            det_out[((size_t)b * post_max + k) * 8 + 7] = scores[A * b + keep[k]];
        }
        det_count[b] = nk;
        if (voxel_count) voxel_count[b] = M;
        total += nk;
        /* keep the compiler from discarding the decorate/scatter work */
        if (dec[0] != dec[0] || canvas[0] != canvas[0]) total += 0;
        free(vox); free(co); free(nm); free(vf); free(c4); free(dec); free(canvas); free(boxes); free(keep);
    }
    return total;
}

/* ------------------------------------------------------------------------------------------
 * "Next" row N2 -- the per-frame body of VoxelNet.predict, model/voxelnet.py:1105-1326.
 *   1112-1137  a_mask_as_indices = np.where(a_mask == 1); gather box/cls/dir/anchors
 *   1143       dir_labels = np.argmax(dir_preds, -1)
 *   1150       total_scores = sigmoid_array(cls_preds)  (722-723: 1 / (1 + np.exp(-x)), float32)
 *   1176-1184  one class: top_scores = squeeze, labels 0 (more classes: max / argmax)
 *   1190-1198  optional `top_scores >= nms_score_threshold`
 *   1207       top min(len,100) by np.argpartition (a set; ties at the cut are undefined in the
 *              reference -- fixed here as everywhere: descending score, then descending index)
 *   1227       second_box_decode of the selected rows
 *   1233-1249  boxes_for_nms = standup boxes of (x,y,w,l,r)
 *   1259-1265  nms(boxes, scores, pre_max, post_max, iou_thr)  (eval_helper_functions.py:463-492),
 *              or rotate NMS on (x,y,w,l,r) when `rotated`
 *   1281-1287  box_preds[selected], dir_labels[selected], top_scores[selected]
 *   1301-1306  opp = (r > 0) ^ dir_label; r += where(opp, pi, 0)   (float32 += float64: one rounding)
 *   1319       box_lidar_to_camera (load_data.py:1511-1523): xyz1 (float64) @ (rect @ Trv2c (float32)).T,
 *              concat [xyz, l, h, w, r] -> float64
 * Outputs hold `cap` rows; returns the number of detections (0 = the reference's None branch).
 * ------------------------------------------------------------------------------------------ */
static float ppo_sigmoid(float x) { const float e = expf(-x); const float d = 1.f + e; return 1.f / d; }

PPO_API int ppo_predict_frame(const float* box_preds, const float* cls_preds, const float* dir_preds,
                              const float* anchors, const uint8_t* a_mask, const float* rect, const float* trv2c,
                              int64_t A, int num_class, int use_dir, int top_k, int pre_max, int post_max,
                              float iou_thr, float score_thr, int rotated, int cap, float* box3d_lidar,
                              double* box3d_camera, float* scores_out, int32_t* labels_out, int32_t* index_out) {
    if (A <= 0) return 0;
    int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)A);
    float* sc = (float*)malloc(sizeof(float) * (size_t)A);
    int64_t m = 0;
    for (int64_t a = 0; a < A; ++a) {
        if (a_mask && a_mask[a] != 1) continue;
        float s = ppo_sigmoid(cls_preds[a * num_class]);
        for (int c = 1; c < num_class; ++c) { const float v = ppo_sigmoid(cls_preds[a * num_class + c]); if (v > s) s = v; }
        if (score_thr > 0.f && !(s >= score_thr)) continue;
        idx[m] = (int32_t)a; sc[m] = s; ++m;
    }
    int nk = 0;
    if (m > 0) {
        int32_t* order = (int32_t*)malloc(sizeof(int32_t) * (size_t)m);
        ppo_argsort_desc(sc, m, order);  /* the gather preserves index order, so ties break as on anchor index */
        int64_t n = m < top_k ? m : top_k;
        if (pre_max > 0 && pre_max < n) n = pre_max;
        float* dec = (float*)malloc(sizeof(float) * 7 * (size_t)n);
        float* te = (float*)malloc(sizeof(float) * 7 * (size_t)n);
        float* ta = (float*)malloc(sizeof(float) * 7 * (size_t)n);
        float* ssel = (float*)malloc(sizeof(float) * (size_t)n);
        for (int64_t i = 0; i < n; ++i) {
            const int64_t a = idx[order[i]];
            memcpy(te + 7 * i, box_preds + 7 * a, 28);
            memcpy(ta + 7 * i, anchors + 7 * a, 28);
            ssel[i] = sc[order[i]];
        }
        ppo_second_box_decode(te, ta, n, dec);
        /* the selection is already in NMS order (descending score, ties by descending anchor index):
         * mask + sweep directly, eval_helper_functions.py:494-546 / nms_gpu.py:455-490 */
        int32_t* keep = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
        const int64_t cb = (n + 63) / 64;
        uint64_t* mask = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(n * cb));
        if (rotated) {
            float* dets = (float*)malloc(sizeof(float) * 6 * (size_t)n);
            for (int64_t i = 0; i < n; ++i) {
                const float* d = dec + 7 * i;
                float* o = dets + 6 * i;
                o[0] = d[0]; o[1] = d[1]; o[2] = d[3]; o[3] = d[4]; o[4] = d[6]; o[5] = ssel[i];
            }
            ppo_rotate_mask(dets, n, iou_thr, mask);
            free(dets);
        } else {
            float* rb = (float*)malloc(sizeof(float) * 5 * (size_t)n);
            float* sb = (float*)malloc(sizeof(float) * 4 * (size_t)n);
            for (int64_t i = 0; i < n; ++i) {
                const float* d = dec + 7 * i;
                float* o = rb + 5 * i;
                o[0] = d[0]; o[1] = d[1]; o[2] = d[3]; o[3] = d[4]; o[4] = d[6];
            }
            ppo_rbox_to_standup(rb, n, sb);
            ppo_standup_mask(sb, n, iou_thr, mask);
            free(rb); free(sb);
        }
        nk = ppo_nms_postprocess(mask, n, keep);
        free(mask);
        if (post_max > 0 && nk > post_max) nk = post_max;
        if (nk > cap) nk = cap;
        float M[12];
        if (rect && trv2c)
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 4; ++j) {
                    float acc = rect[i * 4] * trv2c[j];
                    for (int k = 1; k < 4; ++k) { const float p = rect[i * 4 + k] * trv2c[k * 4 + j]; acc = acc + p; }
                    M[i * 4 + j] = acc;
                }
        for (int k = 0; k < nk; ++k) {
            const int64_t t = keep[k];
            const int64_t a = idx[order[t]];
            float o[7];
            memcpy(o, dec + 7 * t, 28);
            int label = 0;
            if (num_class > 1) {
                float best = ppo_sigmoid(cls_preds[a * num_class]);
                for (int c = 1; c < num_class; ++c) {
                    const float v = ppo_sigmoid(cls_preds[a * num_class + c]);
                    if (v > best) { best = v; label = c; }
                }
            }
            if (use_dir && dir_preds) {
                const int dl = dir_preds[2 * a + 1] > dir_preds[2 * a];
                const int opp = (o[6] > 0.f) != dl;
                o[6] = (float)((double)o[6] + (opp ? 3.141592653589793 : 0.0));
            }
            memcpy(box3d_lidar + 7 * k, o, 28);
            if (box3d_camera && rect && trv2c) {
                double* c = box3d_camera + 7 * k;
                for (int i = 0; i < 3; ++i) {
                    double acc = (double)o[0] * (double)M[i * 4];
                    acc = acc + (double)o[1] * (double)M[i * 4 + 1];
                    acc = acc + (double)o[2] * (double)M[i * 4 + 2];
                    acc = acc + (double)M[i * 4 + 3];
                    c[i] = acc;
                }
                c[3] = o[4]; c[4] = o[5]; c[5] = o[3]; c[6] = o[6];
            }
            if (scores_out) scores_out[k] = ssel[t];
            if (labels_out) labels_out[k] = label;
            if (index_out) index_out[k] = (int32_t)a;
        }
        free(order); free(dec); free(te); free(ta); free(ssel); free(keep);
    }
    free(idx); free(sc);
    return nk;
}
