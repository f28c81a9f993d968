// Pieces shared by the two voxelizer paths (voxelize.cu: any grid; voxelize_small.cu: grids whose
// per-cell tables fit in shared memory).
#pragma once
#include <math.h>

#include "pp_common.cuh"

namespace pp {

// n / d for 0 <= n < 2^31 with one 32x32->64 multiply (Granlund-Montgomery round-up magic)
struct FastDiv {
    unsigned d, m, s;
    __host__ __device__ FastDiv() : d(1), m(0x80000000u), s(31) {}
    __host__ explicit FastDiv(unsigned dd) : d(dd) {
        unsigned l = 0;
        while ((1ull << l) < dd) ++l;
        s = 31 + l;
        m = (unsigned)(((1ull << s) + dd - 1) / dd);
    }
    __device__ __forceinline__ int div(int n) const { return (int)(((unsigned long long)(unsigned)n * m) >> s); }
};

struct VoxParams {
    double lo[3], vs[3], inv[3];
    float lo32[3], vs32[3], inv32[3];
    int grid[3];  // nx, ny, nz
    int ncell;
    FastDiv div_nx, div_nxny;
    int max_points, max_voxels, reverse_index, arith_f32;
    int D;
    // decoration constants (model/pointpillars.py:121-124), float32 like TF constants
    float vx, vy, x_off, y_off;
};

// ---------------------------------------------------------------------------------------------
// cell id in the reference's arithmetic (load_data.py:620-626).  -1: outside the grid or NaN.
// floor((p - lo) / vs) must equal the reference's correctly rounded IEEE division.
// The quotient is first formed with a reciprocal multiply (relative error < 2^-51 in float64,
// < 2^-22 in float32); only when it lands within a 2^-48 (2^-20) relative band of an integer --
// where the two roundings could fall on different sides -- is the exact division evaluated.
template <typename T, bool A32>
__device__ __forceinline__ int cell_of(const T* q, const VoxParams& p) {
    int c[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        if (A32) {
            const float d = __fsub_rn((float)q[j], p.lo32[j]);
            const float qq = __fmul_rn(d, p.inv32[j]);
            float v = floorf(qq);
            const float frac = __fsub_rn(qq, v), tol = fabsf(qq) * 0x1p-20f + 1e-30f;
            if (frac < tol || frac > 1.f - tol) v = floorf(__fdiv_rn(d, p.vs32[j]));
            if (!(v >= 0.f) || !((double)v < (double)p.grid[j])) return -1;
            c[j] = (int)v;
        } else {
            const double d = __dsub_rn((double)q[j], p.lo[j]);
            const double qq = __dmul_rn(d, p.inv[j]);
            if (!(fabs(qq) < 2.0e9)) return -1;  // far outside any grid, inf or NaN
            // floor without the conversion (XU) pipe: qq + 1.5*2^52 rounded down leaves floor(qq) in the
            // low mantissa bits (two's complement), and subtracting the constant gives it back as a double
            const double kMagic = 6755399441055744.0;
            const double t = __dadd_rd(qq, kMagic);
            int ci = __double2loint(t);
            const double fl = __dsub_rn(t, kMagic);
            const double frac = __dsub_rn(qq, fl), tol = fabs(qq) * 0x1p-48 + 1e-300;
            if (frac < tol || frac > 1.0 - tol) {
                const double v = floor(__ddiv_rn(d, p.vs[j]));  // the reference's exact quotient (rare path)
                if (!(v >= 0.0) || !(v < (double)p.grid[j])) return -1;
                ci = (int)v;
            }
            if (ci < 0 || ci >= p.grid[j]) return -1;
            c[j] = ci;
        }
    }
    return (c[2] * p.grid[1] + c[1]) * p.grid[0] + c[0];
}

// Same result as cell_of<T, false> for grids of at most 2047 cells per axis, in a third of the instructions.
// qq + 1.5*2^32 (round to nearest) leaves qq in fixed point with 20 fractional bits in the low mantissa word:
// the cell is the word >> 20, and only when the fraction field is all zeros or all ones -- |frac| < 2^-20, a
// band 2^20 times wider than the 2^-40 by which the reciprocal multiply can differ from the reference's
// division -- is the exact quotient evaluated.
template <typename T>
__device__ __forceinline__ int cell_of_fast20(const T* q, const VoxParams& p) {
    int c[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double d = __dsub_rn((double)q[j], p.lo[j]);
        const double qq = __dmul_rn(d, p.inv[j]);
        if (!(fabs(qq) < 2047.0)) return -1;  // outside any grid this variant serves, inf or NaN
        const int fx = __double2loint(__dadd_rn(qq, 6442450944.0));
        int ci = fx >> 20;
        const int fr = fx & 0xfffff;
        if (fr == 0 || fr == 0xfffff) {
            const double v = floor(__ddiv_rn(d, p.vs[j]));  // the reference's exact quotient (rare path)
            if (!(v >= 0.0) || !(v < (double)p.grid[j])) return -1;
            ci = (int)v;
        }
        if ((unsigned)ci >= (unsigned)p.grid[j]) return -1;
        c[j] = ci;
    }
    return (c[2] * p.grid[1] + c[1]) * p.grid[0] + c[0];
}

__device__ __forceinline__ void warp_store_row(float* __restrict__ dst, const float* __restrict__ src, int n, int lane) {
    // dst: global, src: shared (16-byte aligned); n floats
    const uintptr_t a = reinterpret_cast<uintptr_t>(dst);
    if ((a & 15) == 0 && (n & 3) == 0) {
        for (int k = lane; k < (n >> 2); k += 32)
            reinterpret_cast<float4*>(dst)[k] = reinterpret_cast<const float4*>(src)[k];
    } else if ((a & 7) == 0 && (n & 1) == 0) {
        for (int k = lane; k < (n >> 1); k += 32)
            reinterpret_cast<float2*>(dst)[k] = reinterpret_cast<const float2*>(src)[k];
    } else {
        for (int k = lane; k < n; k += 32) dst[k] = src[k];
    }
}

// same, but elements at or past `lim` are stored as zero (src need not be initialised there)
__device__ __forceinline__ void warp_store_row_padded(float* __restrict__ dst, const float* __restrict__ src, int n, int lim, int lane) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(dst);
    if ((a & 15) == 0 && (n & 3) == 0) {
        for (int k = lane; k < (n >> 2); k += 32) {
            float4 v = reinterpret_cast<const float4*>(src)[k];
            const int e = k << 2;
            v.x = e < lim ? v.x : 0.f; v.y = e + 1 < lim ? v.y : 0.f; v.z = e + 2 < lim ? v.z : 0.f; v.w = e + 3 < lim ? v.w : 0.f;
            reinterpret_cast<float4*>(dst)[k] = v;
        }
    } else if ((a & 7) == 0 && (n & 1) == 0) {
        for (int k = lane; k < (n >> 1); k += 32) {
            float2 v = reinterpret_cast<const float2*>(src)[k];
            const int e = k << 1;
            v.x = e < lim ? v.x : 0.f; v.y = e + 1 < lim ? v.y : 0.f;
            reinterpret_cast<float2*>(dst)[k] = v;
        }
    } else {
        for (int k = lane; k < n; k += 32) dst[k] = k < lim ? src[k] : 0.f;
    }
}

}  // namespace pp
