// "Next" row N1: anchor mask on the device, from the voxelizer's own coors.
//
// Reference (load_data.py:3043-3072): pillar counts per BEV cell (sparse_sum_for_anchors_mask,
// 586-591), 2-D inclusive prefix sums (numpy cumsum(0).cumsum(1)), then per anchor the count inside
// its axis-aligned footprint from four corner lookups (fused_get_anchors_area, 558-584) compared
// with anchor_area_threshold.  The footprint cells depend only on the anchors and the grid, so they
// are computed once (anchor_cells); per batch of frames the work is histogram -> two scans -> lookup.
// Counts are small integers: int32 here, float32 in the reference, identical values.
#include <math.h>

#include "pp_common.cuh"

namespace pp {

// rbbox2d_to_near_bbox (load_data.py:534-548, limit_period 805-806) in float32, then
// floor((bv - offset) / stride) in float64 and the clip of fused_get_anchors_area.
__global__ void __launch_bounds__(256)
anchor_cells_kernel(const float* __restrict__ anchors, int64_t A, double vsx, double vsy, double lox, double loy,
                    int nx, int ny, int4* __restrict__ cells) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= A) return;
    const float* a = anchors + 7 * i;
    const float x = a[0], y = a[1], w = a[3], l = a[4], r = a[6];
    const float pi = 3.14159274101257324f;  // float32(np.pi)
    const float t = __fmul_rn(floorf(__fadd_rn(__fdiv_rn(r, pi), 0.5f)), pi);
    const bool cond = fabsf(__fsub_rn(r, t)) > 0.785398185253143311f;  // float32(np.pi / 4)
    const float dx = cond ? l : w, dy = cond ? w : l;
    const float hx = __fdiv_rn(dx, 2.f), hy = __fdiv_rn(dy, 2.f);
    const float b0 = __fsub_rn(x, hx), b1 = __fsub_rn(y, hy), b2 = __fadd_rn(x, hx), b3 = __fadd_rn(y, hy);
    // The reference clips one side of each index only (load_data.py:577-580) and then indexes dense_map:
    // an index that stays negative wraps around once (numba's negative indexing); anything else outside the
    // map, and NaN / infinite coordinates, is undefined behaviour there.  Here those anchors get the cell
    // rectangle (-1,-1,-1,-1), which the lookup turns into area 0.
    const double v0 = floor(__ddiv_rn(__dsub_rn((double)b0, lox), vsx)), v1 = floor(__ddiv_rn(__dsub_rn((double)b1, loy), vsy));
    const double v2 = floor(__ddiv_rn(__dsub_rn((double)b2, lox), vsx)), v3 = floor(__ddiv_rn(__dsub_rn((double)b3, loy), vsy));
    int4 c = make_int4(-1, -1, -1, -1);
    if (fabs(v0) < 2.0e9 && fabs(v1) < 2.0e9 && fabs(v2) < 2.0e9 && fabs(v3) < 2.0e9) {
        int c0 = max((int)v0, 0), c1 = max((int)v1, 0), c2 = min((int)v2, nx - 1), c3 = min((int)v3, ny - 1);
        c2 += c2 < 0 ? nx : 0;
        c3 += c3 < 0 ? ny : 0;
        if (c0 < nx && c1 < ny && c2 >= 0 && c3 >= 0) c = make_int4(c0, c1, c2, c3);
    }
    cells[i] = c;
}

__global__ void __launch_bounds__(256)
amask_hist_kernel(const int* __restrict__ coors, int cols, int64_t M, const int* __restrict__ M_dev, int B, int ny,
                  int nx, int* __restrict__ map) {
    const int64_t Mv = M_dev ? min((int64_t)*M_dev, M) : M;
    for (int64_t m = (int64_t)blockIdx.x * 256 + threadIdx.x; m < Mv; m += (int64_t)gridDim.x * 256) {
        const int* c = coors + m * cols;
        const int b = cols == 4 ? c[0] : 0;
        const int y = c[cols - 2], x = c[cols - 1];
        if (b < 0 || b >= B || y < 0 || y >= ny || x < 0 || x >= nx) continue;
        atomicAdd(&map[((int64_t)b * ny + y) * nx + x], 1);
    }
}

// inclusive scan along x: one warp per (frame,row)
__global__ void __launch_bounds__(256)
amask_scan_x_kernel(int* __restrict__ map, int64_t rows, int nx) {
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = lane_id();
    int* r = map + row * nx;
    int carry = 0;
    for (int x0 = 0; x0 < nx; x0 += 32) {
        const int x = x0 + lane;
        int v = x < nx ? r[x] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        v += carry;
        if (x < nx) r[x] = v;
        carry = __shfl_sync(0xffffffffu, v, 31);
    }
}

// inclusive scan along y: one thread per (frame,column), consecutive threads on consecutive x
__global__ void __launch_bounds__(256)
amask_scan_y_kernel(int* __restrict__ map, int B, int ny, int nx) {
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= (int64_t)B * nx) return;
    const int b = (int)(t / nx), x = (int)(t - (int64_t)b * nx);
    int* col = map + (int64_t)b * ny * nx + x;
    int acc = 0;
    int y = 0;
    for (; y + 8 <= ny; y += 8) {  // eight independent loads in flight, the running sum is the only chain
        int v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = col[(int64_t)(y + k) * nx];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            acc += v[k];
            col[(int64_t)(y + k) * nx] = acc;
        }
    }
    for (; y < ny; ++y) {
        acc += col[(int64_t)y * nx];
        col[(int64_t)y * nx] = acc;
    }
}

__global__ void __launch_bounds__(256)
amask_lookup_kernel(const int* __restrict__ map, int ny, int nx, const int4* __restrict__ cells, int64_t A,
                    float threshold, const float* __restrict__ scores, float* __restrict__ area,
                    unsigned char* __restrict__ mask, float* __restrict__ masked_scores) {
    const int b = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= A) return;
    const int4 c = cells[i];
    const int* m = map + (int64_t)b * ny * nx;
    float v = 0.f;  // anchors whose footprint the reference cannot index (anchor_cells_kernel)
    if (c.x >= 0) {
        const int ID = m[(int64_t)c.w * nx + c.z], IA = m[(int64_t)c.y * nx + c.x];
        const int IB = m[(int64_t)c.w * nx + c.x], IC = m[(int64_t)c.y * nx + c.z];
        v = (float)(ID - IB - IC + IA);
    }
    const bool on = v > threshold;
    const int64_t o = (int64_t)b * A + i;
    if (area) area[o] = v;
    if (mask) mask[o] = on;
    if (masked_scores) masked_scores[o] = on ? scores[o] : -INFINITY;
}

}  // namespace pp

using namespace pp;

extern "C" int pp_anchor_cells_dev(const float* anchors, int64_t A, const double voxel_size[3],
                                   const double coors_range[6], int32_t* cells, void* stream) {
    PP_CHECK_ARG(A >= 0 && voxel_size && coors_range, "pp_anchor_cells_dev: bad arguments");
    if (A == 0) return PP_OK;
    PP_CHECK_ARG(anchors && cells && (reinterpret_cast<uintptr_t>(cells) & 15) == 0, "pp_anchor_cells_dev: null/unaligned");
    int32_t grid[3];
    pp_grid_size(voxel_size, coors_range, 0, grid);
    PP_TIMED("anchor_cells", static_cast<cudaStream_t>(stream));
    anchor_cells_kernel<<<(unsigned)ceil_div(A, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        anchors, A, voxel_size[0], voxel_size[1], coors_range[0], coors_range[1], grid[0], grid[1],
        reinterpret_cast<int4*>(cells));
    PP_LAUNCHED();
    return PP_OK;
}

extern "C" size_t pp_anchor_mask_workspace_bytes(int B, int ny, int nx) {
    if (B <= 0 || ny <= 0 || nx <= 0) return 0;
    return align_up((size_t)B * ny * nx * sizeof(int), 256) + 256;
}

extern "C" int pp_anchor_mask_dev(const int32_t* coors, int coors_cols, int64_t M, const int32_t* M_dev, int B,
                                  int ny, int nx, const int32_t* cells, int64_t A, float threshold,
                                  const float* scores, float* area, uint8_t* mask, float* masked_scores,
                                  void* workspace, size_t workspace_bytes, void* stream) {
    PP_CHECK_ARG(B > 0 && B <= 65535 && ny > 0 && nx > 0 && A >= 0 && M >= 0, "pp_anchor_mask_dev: bad shape");
    PP_CHECK_ARG(coors_cols == 3 || coors_cols == 4, "pp_anchor_mask_dev: coors_cols must be 3 or 4");
    PP_CHECK_ARG(coors_cols == 4 || B == 1, "pp_anchor_mask_dev: 3-column coors describe a single frame");
    PP_CHECK_ARG(workspace && (M == 0 || coors) && (A == 0 || cells), "pp_anchor_mask_dev: null argument");
    PP_CHECK_ARG(!masked_scores || scores, "pp_anchor_mask_dev: masked_scores needs scores");
    if (pp_anchor_mask_workspace_bytes(B, ny, nx) > workspace_bytes) {
        set_error("pp_anchor_mask_dev: workspace %zu < required %zu", workspace_bytes, pp_anchor_mask_workspace_bytes(B, ny, nx));
        return PP_E_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int* map = static_cast<int*>(workspace);
    PP_CUDA(cudaMemsetAsync(map, 0, (size_t)B * ny * nx * sizeof(int), st));
    if (M > 0) {
        int64_t blocks = ceil_div(M, 256);
        if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
        PP_TIMED("amask_hist", st);
        amask_hist_kernel<<<(unsigned)blocks, 256, 0, st>>>(coors, coors_cols, M, M_dev, B, ny, nx, map);
        PP_LAUNCHED();
    }
    {
        PP_TIMED("amask_scan_x", st);
        amask_scan_x_kernel<<<(unsigned)ceil_div((int64_t)B * ny, 8), 256, 0, st>>>(map, (int64_t)B * ny, nx);
        PP_LAUNCHED();
    }
    {
        PP_TIMED("amask_scan_y", st);
        amask_scan_y_kernel<<<(unsigned)ceil_div((int64_t)B * nx, 256), 256, 0, st>>>(map, B, ny, nx);
        PP_LAUNCHED();
    }
    if (A > 0) {
        const dim3 g((unsigned)ceil_div(A, 256), B);
        PP_TIMED("amask_lookup", st);
        amask_lookup_kernel<<<g, 256, 0, st>>>(map, ny, nx, reinterpret_cast<const int4*>(cells), A, threshold, scores,
                                              area, mask, masked_scores);
        PP_LAUNCHED();
    }
    return PP_OK;
}
