// PointPillarsScatter (model/pointpillars.py:285-341 of the reference) as a GATHER: the canvas is
// written exactly once, with full-line stores, instead of memset + scattered atomics (+ the three
// transposes and the python loop over the batch the reference does).
//
//   link   (pillars)  head[b,y,x] <- pillar row (atomicExch), next[row] <- previous head.
//                     z is ignored; pillars that share (b,y,x) -- the two z slabs of the reference
//                     grid, or any duplicate coords -- form a chain (lines 302, 317: scatter_nd adds).
//   canvas (tiles)    one block per (b, y, 32 x): chains are sorted by row so the float32 sum has
//                     a fixed order (ascending row, the order a sequential scatter-add uses), feature
//                     rows are read with consecutive lanes on consecutive channels into a shared
//                     [C][33] tile, and the tile is stored NCHW (drop-in) or NHWC (what the RPN
//                     transposes to).  Empty tiles are pure zero stores.
#include "pp_common.cuh"

namespace pp {

constexpr int kTileX = 32;
// The canvas pass is bound by its dependent loads (head -> next -> feature row) before any store: more,
// smaller CTAs keep more of those chains in flight.  Measured on B200, 64 frames: KITTI (C=64, 8 KB of
// output per tile) 1170 us at 256 threads, 865 at 128, 822 at 64 (765 with the division-free store loop, 838 at 32);
// D435 (C=128) 90 / 97 / 121 us.
constexpr int kScThreadsWide = 256;    // C > 64
constexpr int kScThreadsNarrow = 64;   // C <= 64
constexpr int kChainMax = 4;

__global__ void __launch_bounds__(256)
scatter_link_kernel(const int* __restrict__ coords, int64_t M, const int* __restrict__ M_dev, int B,
                    int ny, int nx, int* __restrict__ head, int* __restrict__ next, unsigned char* __restrict__ multi) {
    const int64_t Mv = M_dev ? min((int64_t)*M_dev, M) : M;
    for (int64_t m = (int64_t)blockIdx.x * 256 + threadIdx.x; m < Mv; m += (int64_t)gridDim.x * 256) {
        const int4 c = *reinterpret_cast<const int4*>(coords + 4 * m);  // (b, z, y, x)
        if (c.x < 0 || c.x >= B || c.z < 0 || c.z >= ny || c.w < 0 || c.w >= nx) continue;
        const int64_t cell = ((int64_t)c.x * ny + c.z) * nx + c.w;
        const int old = atomicExch(&head[cell], (int)m);
        next[m] = old;
        if (old >= 0) multi[cell] = 1;  // more than one pillar on this (b,y,x): the canvas pass must walk the chain
    }
}

// CELLMAP: `head` is the voxelizer's cell -> row map [B][nz][ny][nx] (pp_voxelize_dev's cell_voxel) and the chain of a
// canvas cell is read straight from its nz <= kChainMax slabs -- no link pass, no list walk (pp_scatter_cells_dev).
template <bool NHWC, int kScThreads, bool CELLMAP>
__global__ void __launch_bounds__(kScThreads)
scatter_canvas_kernel(const float* __restrict__ feats, const int* __restrict__ head,
                      const int* __restrict__ next, const unsigned char* __restrict__ multi, int nz, int C, int ny, int nx,
                      float* __restrict__ out) {
    extern __shared__ float tile[];  // [C][33]
    __shared__ int s_chain[kTileX][kChainMax];
    __shared__ int s_len[kTileX];
    const int xt = blockIdx.x, y = blockIdx.y, b = blockIdx.z;
    const int x0 = xt * kTileX;
    const int wx = min(kTileX, nx - x0);
    const int lane = lane_id(), w = threadIdx.x >> 5;
    const int64_t cellbase = ((int64_t)b * ny + y) * nx + x0;

    int any = 0;
    if (threadIdx.x < kTileX) {
        int len = 0;
        if (CELLMAP) {
            if (threadIdx.x < wx) {
                int rows[kChainMax];
#pragma unroll
                for (int z = 0; z < kChainMax; ++z)
                    rows[z] = z < nz ? __ldg(&head[(((int64_t)b * nz + z) * ny + y) * nx + x0 + threadIdx.x]) : -1;
#pragma unroll
                for (int z = 0; z < kChainMax; ++z) {
                    const int m = rows[z];
                    if (m < 0) continue;
                    int j = len;
                    while (j > 0 && s_chain[threadIdx.x][j - 1] > m) {
                        s_chain[threadIdx.x][j] = s_chain[threadIdx.x][j - 1];
                        --j;
                    }
                    s_chain[threadIdx.x][j] = m;
                    ++len;
                }
            }
        } else if (threadIdx.x < wx) {
            int m = head[cellbase + threadIdx.x];
            if (m >= 0 && !multi[cellbase + threadIdx.x]) {
                // the usual case, one pillar on the cell: no dependent next[] round trip
                s_chain[threadIdx.x][0] = m;
                len = 1;
                m = -1;
            }
            // insertion sort of the chain by row index (ascending)
            while (m >= 0) {
                if (len < kChainMax) {
                    int j = len;
                    while (j > 0 && s_chain[threadIdx.x][j - 1] > m) {
                        s_chain[threadIdx.x][j] = s_chain[threadIdx.x][j - 1];
                        --j;
                    }
                    s_chain[threadIdx.x][j] = m;
                }
                ++len;
                m = next[m];
            }
        }
        s_len[threadIdx.x] = len;
        any = len > 0;
    }
    any = __syncthreads_or(any);

    if (any) {
        // Each warp owns whole x columns of the tile and writes every channel of them (zeros for an
        // empty cell), so the tile needs no clearing.  Sums run in registers in ascending row order.
        const bool vec = ((C & 3) == 0) && ((reinterpret_cast<uintptr_t>(feats) & 15) == 0);
        for (int x = w; x < wx; x += kScThreads / 32) {
            const int len = s_len[x];
            if (vec) {
                for (int c4 = lane; c4 < (C >> 2); c4 += 32) {
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (len <= kChainMax) {
                        for (int j = 0; j < len; ++j) {
                            const float4 v = __ldg(reinterpret_cast<const float4*>(feats + (int64_t)s_chain[x][j] * C) + c4);
                            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                        }
                    } else {
                        int last = -1;
                        for (int j = 0; j < len; ++j) {
                            int best = 0x7fffffff;
                            for (int m = head[cellbase + x]; m >= 0; m = next[m])
                                if (m > last && m < best) best = m;
                            const float4 v = __ldg(reinterpret_cast<const float4*>(feats + (int64_t)best * C) + c4);
                            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                            last = best;
                        }
                    }
                    float* t = tile + (c4 << 2) * 33 + x;
                    t[0] = acc.x; t[33] = acc.y; t[66] = acc.z; t[99] = acc.w;
                }
            } else {
                for (int c = lane; c < C; c += 32) {
                    float acc = 0.f;
                    if (len <= kChainMax) {
                        for (int j = 0; j < len; ++j) acc += feats[(int64_t)s_chain[x][j] * C + c];
                    } else {
                        int last = -1;
                        for (int j = 0; j < len; ++j) {
                            int best = 0x7fffffff;
                            for (int m = head[cellbase + x]; m >= 0; m = next[m])
                                if (m > last && m < best) best = m;
                            acc += feats[(int64_t)best * C + c];
                            last = best;
                        }
                    }
                    tile[c * 33 + x] = acc;
                }
            }
        }
        __syncthreads();
    }

    if (NHWC) {
        // [b, y, x0..x0+wx, C] is one contiguous run of wx*C floats
        float* dst = out + cellbase * C;
        const int n = wx * C;
        if (!any) {
            if ((C & 3) == 0) {
                float4* d4 = reinterpret_cast<float4*>(dst);
                for (int k = threadIdx.x; k < n / 4; k += kScThreads) d4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                for (int k = threadIdx.x; k < n; k += kScThreads) dst[k] = 0.f;
            }
        } else if ((C & 3) == 0 && (C & (C - 1)) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
            // power-of-two channel count: float4 over channels, no division
            const int sh = 31 - __clz(C >> 2);  // log2(C / 4)
            float4* d4 = reinterpret_cast<float4*>(dst);
            for (int k = threadIdx.x; k < (n >> 2); k += kScThreads) {
                const int x = k >> sh, c = (k - (x << sh)) << 2;
                const float* t = tile + c * 33 + x;
                d4[k] = make_float4(t[0], t[33], t[66], t[99]);
            }
        } else {
            for (int k = threadIdx.x; k < n; k += kScThreads) {
                const int x = k / C, c = k - x * C;
                dst[k] = tile[c * 33 + x];
            }
        }
    } else {
        // NCHW: channel c row segment at ((b*C + c)*ny + y)*nx + x0, wx floats
        const bool vec = ((nx & 3) == 0) && ((wx & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
        if (vec && wx == kTileX) {
            // full tile: 8 float4 per channel row; thread -> (channel c0 + j*step, float4 i) with no division and one
            // pointer bump per store (ncu: the generic loop below made this pass 75 % issue-active on KITTI)
            const int i = threadIdx.x & 7;
            const int64_t plane = (int64_t)ny * nx;
            float* dst = out + (((int64_t)b * C + (threadIdx.x >> 3)) * ny + y) * nx + x0 + (i << 2);
            const float* t = tile + (threadIdx.x >> 3) * 33 + (i << 2);
            constexpr int kStep = kScThreads >> 3;
            if (any) {
                for (int c = threadIdx.x >> 3; c < C; c += kStep, dst += kStep * plane, t += kStep * 33)
                    *reinterpret_cast<float4*>(dst) = make_float4(t[0], t[1], t[2], t[3]);
            } else {
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int c = threadIdx.x >> 3; c < C; c += kStep, dst += kStep * plane) *reinterpret_cast<float4*>(dst) = z;
            }
        } else if (vec) {
            const int q = wx >> 2;  // float4 per channel row
            for (int k = threadIdx.x; k < C * q; k += kScThreads) {
                const int c = k / q, i = k - c * q;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (any) {
                    const float* t = tile + c * 33 + (i << 2);
                    v = make_float4(t[0], t[1], t[2], t[3]);
                }
                reinterpret_cast<float4*>(out + (((int64_t)b * C + c) * ny + y) * nx + x0)[i] = v;
            }
        } else {
            for (int c = w; c < C; c += kScThreads / 32) {
                float* dst = out + (((int64_t)b * C + c) * ny + y) * nx + x0;
                if (lane < wx) dst[lane] = any ? tile[c * 33 + lane] : 0.f;
            }
        }
    }
}

}  // namespace pp

using namespace pp;

extern "C" size_t pp_scatter_workspace_bytes(int B, int ny, int nx, int64_t M) {
    if (B <= 0 || ny <= 0 || nx <= 0 || M < 0) return 0;
    return align_up((size_t)B * ny * nx * sizeof(int), 256) + align_up((size_t)(M + 1) * sizeof(int), 256) +
           align_up((size_t)B * ny * nx, 256) + 256;
}

extern "C" int pp_scatter_dev(const float* feats, const int32_t* coords, int64_t M, const int32_t* M_dev,
                              int C, int B, int ny, int nx, int layout, float* out, void* workspace,
                              size_t workspace_bytes, void* stream) {
    PP_CHECK_ARG(B > 0 && ny > 0 && nx > 0 && C > 0 && M >= 0, "pp_scatter_dev: bad shape");
    PP_CHECK_ARG(B <= 65535 && ny <= 65535, "pp_scatter_dev: B and ny must be <= 65535");
    PP_CHECK_ARG(layout == PP_LAYOUT_NCHW || layout == PP_LAYOUT_NHWC, "pp_scatter_dev: bad layout");
    PP_CHECK_ARG(out && workspace && (M == 0 || (feats && coords)), "pp_scatter_dev: null argument");
    PP_CHECK_ARG(M < ((int64_t)1 << 31), "pp_scatter_dev: M must be < 2^31");
    // coords rows are read as int4 and the canvas is stored with 16-byte vectors
    PP_CHECK_ARG((reinterpret_cast<uintptr_t>(coords) & 15) == 0, "pp_scatter_dev: coords must be 16-byte aligned");
    PP_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0, "pp_scatter_dev: out must be 16-byte aligned");
    const size_t smem = (size_t)C * 33 * sizeof(float);
    PP_CHECK_ARG(smem <= 200 * 1024, "pp_scatter_dev: C=%d too large", C);
    if (pp_scatter_workspace_bytes(B, ny, nx, M) > workspace_bytes) {
        set_error("pp_scatter_dev: workspace %zu < required %zu", workspace_bytes,
                  pp_scatter_workspace_bytes(B, ny, nx, M));
        return PP_E_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Carver cv(workspace);
    int* head = cv.take<int>((size_t)B * ny * nx);
    int* next = cv.take<int>((size_t)M + 1);
    unsigned char* multi = cv.take<unsigned char>((size_t)B * ny * nx);
    {
        PP_TIMED("scatter_memset", st);
        PP_CUDA(cudaMemsetAsync(head, 0xff, (size_t)B * ny * nx * sizeof(int), st));
        PP_CUDA(cudaMemsetAsync(multi, 0, (size_t)B * ny * nx, st));
    }
    if (M > 0) {
        int64_t blocks = ceil_div(M, 256);
        if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
        PP_TIMED("scatter_link", st);
        scatter_link_kernel<<<(unsigned)blocks, 256, 0, st>>>(coords, M, M_dev, B, ny, nx, head, next, multi);
        PP_LAUNCHED();
    }
    const dim3 g((unsigned)ceil_div(nx, kTileX), ny, B);
    PP_TIMED("scatter_canvas", st);
#define PP_CANVAS(NHWC_, T_)                                                                                          \
    do {                                                                                                              \
        auto k = scatter_canvas_kernel<NHWC_, T_, false>;                                                             \
        int per_sm__ = 0;                                                                                             \
        PP_TRY_RC(kernel_config(reinterpret_cast<const void*>(k), T_, smem, &per_sm__));                              \
        k<<<g, T_, smem, st>>>(feats, head, next, multi, 1, C, ny, nx, out);                                          \
    } while (0)
    const bool nhwc = layout == PP_LAYOUT_NHWC;
    if (C <= 64) { if (nhwc) PP_CANVAS(true, kScThreadsNarrow); else PP_CANVAS(false, kScThreadsNarrow); }
    else { if (nhwc) PP_CANVAS(true, kScThreadsWide); else PP_CANVAS(false, kScThreadsWide); }
#undef PP_CANVAS
    PP_LAUNCHED();
    return PP_OK;
}

extern "C" int pp_scatter_cells_dev(const float* feats, const int32_t* cell_voxel, int nz, int C, int B, int ny, int nx,
                                    int layout, float* out, void* stream) {
    PP_CHECK_ARG(B > 0 && ny > 0 && nx > 0 && C > 0, "pp_scatter_cells_dev: bad shape");
    PP_CHECK_ARG(B <= 65535 && ny <= 65535, "pp_scatter_cells_dev: B and ny must be <= 65535");
    PP_CHECK_ARG(nz >= 1 && nz <= kChainMax, "pp_scatter_cells_dev: nz=%d, at most %d slabs (use pp_scatter_dev)", nz, kChainMax);
    PP_CHECK_ARG(layout == PP_LAYOUT_NCHW || layout == PP_LAYOUT_NHWC, "pp_scatter_cells_dev: bad layout");
    PP_CHECK_ARG(out && feats && cell_voxel, "pp_scatter_cells_dev: null argument");
    PP_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0, "pp_scatter_cells_dev: out must be 16-byte aligned");
    const size_t smem = (size_t)C * 33 * sizeof(float);
    PP_CHECK_ARG(smem <= 200 * 1024, "pp_scatter_cells_dev: C=%d too large", C);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dim3 g((unsigned)ceil_div(nx, kTileX), ny, B);
    PP_TIMED("scatter_canvas", st);
#define PP_CANVAS(NHWC_, T_)                                                                                          \
    do {                                                                                                              \
        auto k = scatter_canvas_kernel<NHWC_, T_, true>;                                                              \
        int per_sm__ = 0;                                                                                             \
        PP_TRY_RC(kernel_config(reinterpret_cast<const void*>(k), T_, smem, &per_sm__));                              \
        k<<<g, T_, smem, st>>>(feats, cell_voxel, nullptr, nullptr, nz, C, ny, nx, out);                              \
    } while (0)
    const bool nhwc = layout == PP_LAYOUT_NHWC;
    if (C <= 64) { if (nhwc) PP_CANVAS(true, kScThreadsNarrow); else PP_CANVAS(false, kScThreadsNarrow); }
    else { if (nhwc) PP_CANVAS(true, kScThreadsWide); else PP_CANVAS(false, kScThreadsWide); }
#undef PP_CANVAS
    PP_LAUNCHED();
    return PP_OK;
}
