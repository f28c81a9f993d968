// Sensor ingest on the device ("next" row N3): the production branch of dataLoader.__getitem__,
// load_data.py:2434-2443 (same sequence in scripts/realsense_make_dataset.py:382-414):
//
//     points = ros_numpy.point_cloud2.pointcloud2_to_xyz_array(pc)[1::4]   # finite x,y,z rows only, float64
//     points = np.dot(points, r); points = np.dot(points, r2)              # two float64 3x3 rotations
//     points = points + [0.0, 0.0, 1.0]
//
// ros_numpy (third party, not vendored by the reference) drops every row with a non-finite x, y or z
// BEFORE the [1::4] slice, so the subsample is a slice of an ordered stream compaction.  Three launches
// for a batch of frames: per-tile finite counts, a scan over tiles per frame, and a pass that ranks the
// finite rows (four consecutive records per thread, one block scan per 1024-record tile), keeps
// rank = start + j*step, applies the rotations and the translation in float64 (k = 0,1,2 in order,
// separate multiply and add) and stores row j.  Rows past the frame's
// count are filled with NaN, which the voxelizer drops, so the [B, cap, 3] output feeds pp_voxelize_dev
// with fixed frame offsets and no host round trip for the counts.
//
// With the reference's matrices (entries 0, +-1 and +-2^-52, scipy's from_euler(+-90 deg)) every output
// coordinate is the sum of at most two non-zero exact products, so the result is bit-identical to numpy's
// BLAS dgemm whatever its summation order or FMA use.  For general matrices it agrees to 1 ulp per dot.
#include "pp_common.cuh"

namespace pp {

constexpr int kIngestThreads = 256;
constexpr int kIngestTile = 1024;  // rows per CTA
constexpr int kIngestMaxRot = 4;

struct IngestXform {
    double r[kIngestMaxRot][9];
    double t[3];
    int n_rot;
};

__device__ __forceinline__ bool finite3(float x, float y, float z) {
    return isfinite(x) && isfinite(y) && isfinite(z);
}

// Thread t of a tile owns the four consecutive records 4t..4t+3.  PACKED (point_step 12, 16-byte aligned
// frames): they are 48 contiguous bytes = three 16-byte loads; otherwise twelve 4-byte loads.
template <bool PACKED>
__device__ __forceinline__ void load_rows4(const unsigned char* __restrict__ fc, int64_t row0, int64_t n_in, int point_step,
                                           int ox, int oy, int oz, float (&x)[4], float (&y)[4], float (&z)[4],
                                           unsigned& okmask) {
    okmask = 0;
    if (PACKED && row0 + 4 <= n_in) {
        const float4* p = reinterpret_cast<const float4*>(fc + row0 * 12);
        const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
        x[0] = a.x; y[0] = a.y; z[0] = a.z; x[1] = a.w; y[1] = b.x; z[1] = b.y;
        x[2] = b.z; y[2] = b.w; z[2] = c.x; x[3] = c.y; y[3] = c.z; z[3] = c.w;
#pragma unroll
        for (int k = 0; k < 4; ++k) okmask |= finite3(x[k], y[k], z[k]) ? 1u << k : 0u;
        return;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        x[k] = y[k] = z[k] = 0.f;
        if (row0 + k < n_in) {
            const unsigned char* p = fc + (row0 + k) * point_step;
            x[k] = *reinterpret_cast<const float*>(p + ox);
            y[k] = *reinterpret_cast<const float*>(p + oy);
            z[k] = *reinterpret_cast<const float*>(p + oz);
            okmask |= finite3(x[k], y[k], z[k]) ? 1u << k : 0u;
        }
    }
}

template <bool PACKED>
__global__ void __launch_bounds__(kIngestThreads)
ingest_count_kernel(const unsigned char* __restrict__ cloud, int64_t n_in, int point_step, int ox, int oy, int oz,
                    int tiles_per_frame, int* __restrict__ tile_count) {
    const int b = blockIdx.y, tile = blockIdx.x;
    const unsigned char* fc = cloud + (int64_t)b * n_in * point_step;
    float x[4], y[4], z[4];
    unsigned ok;
    load_rows4<PACKED>(fc, (int64_t)tile * kIngestTile + 4 * threadIdx.x, n_in, point_step, ox, oy, oz, x, y, z, ok);
    int c = __popc(ok);
    __shared__ int s_w[kIngestThreads / 32];
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane_id() == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < kIngestThreads / 32; ++w) s += s_w[w];
        tile_count[(int64_t)b * tiles_per_frame + tile] = s;
    }
}

// one CTA per frame: exclusive scan of the tile counts in place; n_out[b] = rows kept by [start::step]
__global__ void __launch_bounds__(1024)
ingest_scan_kernel(int* __restrict__ tile_count, int tiles_per_frame, int start, int step, int64_t cap,
                   int* __restrict__ n_out) {
    __shared__ int sm[33];
    __shared__ int s_run;
    const int b = blockIdx.x;
    int* tc = tile_count + (int64_t)b * tiles_per_frame;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int t0 = 0; t0 < tiles_per_frame; t0 += 1024) {
        const int t = t0 + threadIdx.x;
        const int v = t < tiles_per_frame ? tc[t] : 0;
        int tot;
        const int ex = block_excl_scan(v, &tot, sm);
        const int run = s_run;
        if (t < tiles_per_frame) tc[t] = run + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_run = run + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int total = s_run;
        int64_t n = total > start ? ((int64_t)total - start + step - 1) / step : 0;
        if (n > cap) n = cap;
        n_out[b] = (int)n;
    }
}

template <bool PACKED>
__global__ void __launch_bounds__(kIngestThreads)
ingest_write_kernel(const unsigned char* __restrict__ cloud, int64_t n_in, int point_step, int ox, int oy, int oz,
                    int tiles_per_frame, const int* __restrict__ tile_base, int start, int step, IngestXform xf,
                    double* __restrict__ out, int64_t cap, const int* __restrict__ n_out) {
    __shared__ int sm[33];
    const int b = blockIdx.y, tile = blockIdx.x;
    const unsigned char* fc = cloud + (int64_t)b * n_in * point_step;
    double* fo = out + (int64_t)b * cap * 3;
    float x[4], y[4], z[4];
    unsigned ok;
    load_rows4<PACKED>(fc, (int64_t)tile * kIngestTile + 4 * threadIdx.x, n_in, point_step, ox, oy, oz, x, y, z, ok);
    int tot;
    const int rank0 = tile_base[(int64_t)b * tiles_per_frame + tile] + block_excl_scan(__popc(ok), &tot, sm);
    // Ranks of this thread's finite rows are rank0, rank0+1, ...: one division finds the first kept rank
    // (start + j*step) at or after rank0, the rest is counting.  The kept rows (at most one per thread when
    // step >= 4) are then transformed in a loop all lanes walk together -- running the float64 transform
    // inside the four-row loop would execute it four times per warp with a quarter of the lanes each.
    const int rel0 = rank0 - start;
    int64_t j = rel0 <= 0 ? 0 : (rel0 + step - 1) / step;
    int next_sel = start + (int)j * step;
    unsigned sel = 0;
    {
        int rank = rank0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if ((ok >> k) & 1u) {
                if (rank == next_sel) { sel |= 1u << k; next_sel += step; }
                ++rank;
            }
        }
    }
    while (sel) {
        const int k = __ffs(sel) - 1;
        sel &= sel - 1;
        if (j < cap) {
            const float xs = k == 0 ? x[0] : k == 1 ? x[1] : k == 2 ? x[2] : x[3];
            const float ys = k == 0 ? y[0] : k == 1 ? y[1] : k == 2 ? y[2] : y[3];
            const float zs = k == 0 ? z[0] : k == 1 ? z[1] : k == 2 ? z[2] : z[3];
            double p[3] = {(double)xs, (double)ys, (double)zs};
            for (int m = 0; m < xf.n_rot; ++m) {
                const double* r = xf.r[m];
                double o[3];
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    o[c] = __dadd_rn(__dadd_rn(__dmul_rn(p[0], r[c]), __dmul_rn(p[1], r[3 + c])), __dmul_rn(p[2], r[6 + c]));
                p[0] = o[0]; p[1] = o[1]; p[2] = o[2];
            }
            fo[j * 3 + 0] = __dadd_rn(p[0], xf.t[0]);
            fo[j * 3 + 1] = __dadd_rn(p[1], xf.t[1]);
            fo[j * 3 + 2] = __dadd_rn(p[2], xf.t[2]);
        }
        ++j;
    }
    // rows [n_out[b], cap) = NaN (dropped by the voxelizer): every CTA of the frame pads an equal slice
    const int64_t first = (int64_t)n_out[b] * 3, len = cap * 3 - first;
    if (len > 0) {
        const int64_t per = (len + tiles_per_frame - 1) / tiles_per_frame;
        const int64_t lo = first + (int64_t)tile * per, hi = min(lo + per, cap * 3);
        for (int64_t k = lo + threadIdx.x; k < hi; k += kIngestThreads) fo[k] = __longlong_as_double(0x7ff8000000000000ll);
    }
}

// n_in == 0: every row of the output is padding
__global__ void __launch_bounds__(256)
ingest_pad_kernel(double* __restrict__ out, int64_t cap, const int* __restrict__ n_out) {
    const int b = blockIdx.y;
    const int64_t first = (int64_t)n_out[b] * 3;
    const int64_t k = first + (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (k < cap * 3) out[(int64_t)b * cap * 3 + k] = __longlong_as_double(0x7ff8000000000000ll);
}

}  // namespace pp

using namespace pp;

static int ingest_tiles(int64_t n_in) { return (int)((n_in + kIngestTile - 1) / kIngestTile); }

extern "C" size_t pp_ingest_workspace_bytes(int B, int64_t n_in) {
    if (B <= 0 || n_in < 0) return 0;
    Carver c(nullptr);
    c.take<int>((size_t)B * (ingest_tiles(n_in) + 1));
    return c.used() + 256;
}

extern "C" int pp_ingest_dev(const void* cloud, int B, int64_t n_in, int point_step, int off_x, int off_y, int off_z,
                             int start, int step, const double* rotations, int n_rot, const double* translation,
                             double* points_out, int64_t cap, int32_t* n_out, void* workspace, size_t workspace_bytes,
                             void* stream) {
    PP_CHECK_ARG(B > 0 && B <= 65535 && n_in >= 0 && n_in < ((int64_t)1 << 31), "pp_ingest_dev: bad B/n_in");
    PP_CHECK_ARG(point_step >= 12 && (point_step & 3) == 0 && off_x >= 0 && off_y >= 0 && off_z >= 0 &&
                     ((off_x | off_y | off_z) & 3) == 0 && off_x + 4 <= point_step && off_y + 4 <= point_step &&
                     off_z + 4 <= point_step,
                 "pp_ingest_dev: bad point_step / field offsets (float32 fields, 4-byte aligned)");
    PP_CHECK_ARG(start >= 0 && step >= 1, "pp_ingest_dev: bad slice");
    PP_CHECK_ARG(n_rot >= 0 && n_rot <= kIngestMaxRot && (n_rot == 0 || rotations), "pp_ingest_dev: 0..%d rotations", kIngestMaxRot);
    PP_CHECK_ARG(points_out && n_out && workspace && cap >= 0, "pp_ingest_dev: null argument");
    PP_CHECK_ARG((reinterpret_cast<uintptr_t>(cloud) & 3) == 0, "pp_ingest_dev: cloud must be 4-byte aligned");
    if (workspace_bytes < pp_ingest_workspace_bytes(B, n_in)) {
        set_error("pp_ingest_dev: workspace %zu < %zu", workspace_bytes, pp_ingest_workspace_bytes(B, n_in));
        return PP_E_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int tiles = ingest_tiles(n_in);
    Carver c(workspace);
    int* tile_count = c.take<int>((size_t)B * (tiles + 1));
    IngestXform xf;
    xf.n_rot = n_rot;
    for (int m = 0; m < n_rot; ++m)
        for (int k = 0; k < 9; ++k) xf.r[m][k] = rotations[m * 9 + k];
    for (int k = 0; k < 3; ++k) xf.t[k] = translation ? translation[k] : 0.0;
    const unsigned char* cl = static_cast<const unsigned char*>(cloud);
    const bool packed = point_step == 12 && off_x == 0 && off_y == 4 && off_z == 8 &&
                        (reinterpret_cast<uintptr_t>(cloud) & 15) == 0 && ((n_in * 12) & 15) == 0;
    if (tiles > 0) {
        PP_CHECK_ARG(cloud, "pp_ingest_dev: null cloud");
        PP_TIMED("ingest_count", st);
        if (packed) ingest_count_kernel<true><<<dim3(tiles, B), kIngestThreads, 0, st>>>(cl, n_in, point_step, off_x, off_y, off_z, tiles, tile_count);
        else ingest_count_kernel<false><<<dim3(tiles, B), kIngestThreads, 0, st>>>(cl, n_in, point_step, off_x, off_y, off_z, tiles, tile_count);
        PP_LAUNCHED();
    }
    {
        PP_TIMED("ingest_scan", st);
        ingest_scan_kernel<<<B, 1024, 0, st>>>(tile_count, tiles, start, step, cap, n_out);
        PP_LAUNCHED();
    }
    if (tiles > 0 && cap > 0) {
        PP_TIMED("ingest_write", st);
        if (packed) ingest_write_kernel<true><<<dim3(tiles, B), kIngestThreads, 0, st>>>(cl, n_in, point_step, off_x, off_y, off_z, tiles, tile_count, start, step, xf, points_out, cap, n_out);
        else ingest_write_kernel<false><<<dim3(tiles, B), kIngestThreads, 0, st>>>(cl, n_in, point_step, off_x, off_y, off_z, tiles, tile_count, start, step, xf, points_out, cap, n_out);
        PP_LAUNCHED();
    } else if (cap > 0) {
        PP_TIMED("ingest_pad", st);
        ingest_pad_kernel<<<dim3((unsigned)ceil_div(cap * 3, 256), B), 256, 0, st>>>(points_out, cap, n_out);
        PP_LAUNCHED();
    }
    return PP_OK;
}
