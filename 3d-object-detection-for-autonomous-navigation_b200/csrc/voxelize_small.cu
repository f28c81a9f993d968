// Deterministic first-come pillarization for grids whose per-cell tables fit in shared memory
// (the d435i grid of configs/train.yaml: 80 x 64 x 2 = 10 240 cells).  Same results, bit for bit, as the
// reference's sequential loop (load_data.py:593-692) and as the any-grid path in voxelize.cu, but built
// as a two-level counting sort: no pass issues a global atomic per point, none sorts, none ranks a bitmap.
//
//   scan    (frame, chunk of 16 384 points)  rows staged by TMA bulk copies, four points per thread; cell id in
//           the reference's arithmetic; in-range points are compacted IN INDEX ORDER into 16/32-byte records
//           {coordinates in the output type, cell | index in chunk}; the chunk's per-cell counts live in SHARED
//           memory (warp-aggregated ATOMS) and are written out once per chunk.
//   prefix  (frame, 256 cells)  per cell: exclusive prefix of the chunk counts = the slot base of every chunk
//           (uint8, saturated: a base >= max_points means "full"); the chunk in which a cell first appears is
//           counted, which gives every chunk the number of voxels opened before it.
//   place   (frame, chunk)  ONE WARP walks the chunk's records in order, 32 per step, with the chunk's running
//           per-cell counts in shared memory: slot = count[cell] + rank among the step's earlier records of the
//           cell (match_any).  A record that finds count 0 opens its cell: because the walk is in index order,
//           the running number of such records IS the voxel id (order of first touch), and the record that
//           would open voxel number max_voxels is the reference's `break` position (load_data.py:630-634).
//           Records go to a per-cell staging row with one 16-byte store; no barrier, no atomic.
//   finish  (pillar)  one warp per kept voxel: staged records before the break position -> zero-padded voxel
//           row, num_points, coors, point->slot map and the fused PillarFeatureNet decoration
//           (model/pointpillars.py:143-203).
#include "pp_common.cuh"
#include "vox_common.cuh"
#include "vox_internal.h"

namespace pp {

constexpr int kChunk = 16384;  // points per chunk; positions inside a chunk are 14-bit
constexpr int kChunkShift = 14;
constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kScanPPT = 4;
constexpr int kScanTile = kScanThreads * kScanPPT;
constexpr int kPrefixThreads = 256;
constexpr int kMaxChunks = 64;                     // per frame: max_frame_points <= 2^20
constexpr int kMaxCellsSmall = 16384;              // cell ids are 14-bit in the record word
constexpr int kMaxFramePointsSmall = kMaxChunks * kChunk;
constexpr int kMaxPointsSmall = 254;               // uint8 counts saturate at 255
constexpr int kPlaceUnroll = 4;
constexpr int kFinishWarps = 8;
constexpr int kNoCut = 0x7fffffff;

static_assert(kChunk == 1 << kChunkShift && kChunk % kScanTile == 0, "chunk size");

__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_bulk_g2s_hint(void* dst_smem, const void* src_gmem, unsigned bytes,
                                                  unsigned long long* bar, unsigned long long policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// record layout: kRec16 16-byte units; when D == 3 the spare element carries the 32-bit tag
template <typename TO, int DS> struct RecFmt {
    static constexpr int kBytes = (int)(((DS * sizeof(TO) + 15) / 16) * 16);
    static constexpr int kRec16 = kBytes / 16;
    static constexpr bool kTagInside = DS == 3;
};
template <int REC16> __device__ __forceinline__ unsigned rec_tag(const uint4 (&rv)[REC16]) {
    return REC16 == 1 ? rv[0].w : rv[REC16 - 1].z;
}
template <int REC16> __device__ __forceinline__ void rec_set_tag(uint4 (&rv)[REC16], unsigned tag) {
    if (REC16 == 1) rv[0].w = tag; else rv[REC16 - 1].z = tag;
}

// ---------------------------------------------------------------------------------------------
// Pass 1.  Warp w of a tile owns its points [128 w, 128 w + 128); round r of lane l is point 128 w + 32 r + l,
// so (warp, round, lane) enumerates the tile in index order and shared-memory row reads are conflict free.
template <typename T, bool A32, bool FAST, typename TO, int DS>
__global__ void __launch_bounds__(kScanThreads)
vox_scan_kernel(const T* __restrict__ points, const int64_t* __restrict__ frame_off, VoxParams p, int64_t total_points,
                int aligned16, int S, int ncellp, uint4* __restrict__ crec, unsigned* __restrict__ cmeta,
                unsigned* __restrict__ hist_out, int* __restrict__ nvalid, int* __restrict__ newcount,
                int* __restrict__ cutoff, int* __restrict__ point_slot, int* __restrict__ done_counter) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ int s_wtot[kScanWarps];
    constexpr int row_bytes = DS * (int)sizeof(T);
    constexpr int stage_bytes = kScanTile * row_bytes + 32;  // multiple of 16
    constexpr int rec16 = RecFmt<TO, DS>::kRec16;
    unsigned* hist32 = reinterpret_cast<unsigned*>(smem + stage_bytes);  // packed uint16 counts
    const int tid = threadIdx.x, lane = lane_id(), w = tid >> 5;
    const int b = blockIdx.y, s = blockIdx.x;
    if (tid == 0) {
        newcount[b * S + s] = 0;
        if (s == 0) cutoff[b] = kNoCut;
        if (b == 0 && s == 0) *done_counter = 0;
    }
    const int64_t f0 = frame_off[b];
    const int n = (int)(frame_off[b + 1] - f0);
    const int c0 = s * kChunk;
    if (c0 >= n) return;
    const int cn = min(kChunk, n - c0);
    for (int k = tid; k < (ncellp >> 1); k += kScanThreads) hist32[k] = 0u;
    if (tid == 0) mbar_init(&s_bar, 1);
    __syncthreads();

    const int64_t total_bytes = total_points * (int64_t)row_bytes;
    const int64_t tail0 = total_bytes & ~(int64_t)15;  // end of the buffer's last full 16-byte unit
    const unsigned char* src = reinterpret_cast<const unsigned char*>(points);
    const unsigned long long policy = l2_evict_first_policy();  // the cloud is read exactly once
    const int ntiles = (cn + kScanTile - 1) / kScanTile;
    unsigned phase = 0;
    int crun = 0;  // in-range points of the chunk so far (uniform over the CTA)
    const int64_t gbase = f0 + c0;

    for (int j = 0; j < ntiles; ++j) {
        const int base = j * kScanTile;  // first point of the tile, relative to the chunk
        const int m = min(kScanTile, cn - base);
        const int64_t start = (gbase + base) * (int64_t)row_bytes;
        const int64_t end = start + (int64_t)m * row_bytes;
        const int64_t a0 = start & ~(int64_t)15;
        int64_t a1 = (end + 15) & ~(int64_t)15;
        if (a1 > tail0) a1 = tail0 > a0 ? tail0 : a0;
        int shift = 0;
        if (aligned16) {
            shift = (int)(start - a0);
            const unsigned bulk = (unsigned)(a1 - a0);
            if (bulk) {
                if (tid == 0) {
                    mbar_arrive_expect_tx(&s_bar, bulk);
                    tma_bulk_g2s_hint(smem, src + a0, bulk, &s_bar, policy);
                }
                mbar_wait(&s_bar, phase);
                phase ^= 1;
            }
            if (a1 < end) {
                // bytes past the buffer's last full unit (only the very last tile of the buffer)
                for (int64_t a = a1 + (int64_t)tid * (int)sizeof(T); a < end && a < total_bytes;
                     a += (int64_t)kScanThreads * (int)sizeof(T))
                    *reinterpret_cast<T*>(smem + (a - a0)) = *reinterpret_cast<const T*>(src + a);
                __syncthreads();
            }
        } else {
            const int nel = m * DS;
            for (int k = tid; k < nel; k += kScanThreads)
                reinterpret_cast<T*>(smem)[k] = points[(gbase + base) * DS + k];
            __syncthreads();
        }
        const unsigned char* rows = smem + shift + (size_t)(w * 128 + lane) * row_bytes;

        int cell[kScanPPT];
        unsigned bal[kScanPPT];
        int wtot = 0;
#pragma unroll
        for (int r = 0; r < kScanPPT; ++r) {
            const int t = w * 128 + r * 32 + lane;
            const T* q = reinterpret_cast<const T*>(rows + (size_t)r * 32 * row_bytes);
            cell[r] = -1;
            if (t < m) cell[r] = FAST ? cell_of_fast20<T>(q, p) : cell_of<T, A32>(q, p);
            bal[r] = __ballot_sync(0xffffffffu, cell[r] >= 0);
            wtot += __popc(bal[r]);
        }
        if (lane == 0) s_wtot[w] = wtot;
        __syncthreads();
        // where this warp's in-range points go: the counts of the lower warps
        int woff = 0, tile_total = 0;
#pragma unroll
        for (int k = 0; k < kScanWarps; ++k) {
            const int v = s_wtot[k];
            woff += k < w ? v : 0;
            tile_total += v;
        }
        int pos = crun + woff;
#pragma unroll
        for (int r = 0; r < kScanPPT; ++r) {
            const int t = w * 128 + r * 32 + lane;
            const unsigned peers = __match_any_sync(0xffffffffu, cell[r]);
            if (cell[r] >= 0) {
                if ((int)lane == __ffs(peers) - 1)
                    atomicAdd(&hist32[cell[r] >> 1], (unsigned)__popc(peers) << ((cell[r] & 1) * 16));
                const T* q = reinterpret_cast<const T*>(rows + (size_t)r * 32 * row_bytes);
                const int64_t gi = gbase + pos + __popc(bal[r] & lanemask_lt());
                const unsigned tag = (unsigned)cell[r] | ((unsigned)(base + t) << 16);
                uint4* dst = crec + gi * rec16;
                if (sizeof(TO) == 4) {
                    const float e3 = DS == 4 ? (float)q[DS - 1] : __uint_as_float(tag);
                    *reinterpret_cast<float4*>(dst) = make_float4((float)q[0], (float)q[1], (float)q[2], e3);
                } else {
                    const double e3 = DS == 4 ? (double)q[DS - 1] : __longlong_as_double((long long)tag);
                    reinterpret_cast<double2*>(dst)[0] = make_double2((double)q[0], (double)q[1]);
                    reinterpret_cast<double2*>(dst)[1] = make_double2((double)q[2], e3);
                }
                if (DS == 4) cmeta[gi] = tag;
            }
            if (point_slot && t < m) point_slot[gbase + base + t] = -1;
            pos += __popc(bal[r]);
        }
        crun += tile_total;
        __syncthreads();  // every thread is done with the stage and with s_wtot
    }
    // the chunk's counts (uint16 pairs, as they lie in shared memory)
    const size_t tb = ((size_t)b * S + s) * (size_t)(ncellp >> 1);
    for (int k = tid; k < (ncellp >> 1); k += kScanThreads) hist_out[tb + k] = hist32[k];
    if (tid == 0) nvalid[b * S + s] = crun;
}

// ---------------------------------------------------------------------------------------------
// Pass 2: one thread per cell.  base8[s][cell] = min(255, points of the cell in chunks < s); a cell is counted
// as a new voxel of the first chunk that holds it.  The last CTA turns the per-chunk counts into voxel_num
// and voxel_base (rows of a batch are packed back to back, merge_second_batch layout).
__global__ void __launch_bounds__(kPrefixThreads)
vox_prefix_kernel(const int64_t* __restrict__ frame_off, int S, int ncell, int ncellp,
                  const unsigned short* __restrict__ hist, unsigned char* __restrict__ base8,
                  unsigned char* __restrict__ tot8, int* __restrict__ newcount, int max_voxels, int B,
                  int* __restrict__ voxel_num, int* __restrict__ voxel_base, int* __restrict__ done_counter) {
    __shared__ int s_new[kMaxChunks];
    __shared__ int sm[33];
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int cell = blockIdx.x * kPrefixThreads + tid;
    const int n = (int)(frame_off[b + 1] - frame_off[b]);
    const int Sb = (n + kChunk - 1) >> kChunkShift;
    if (tid < kMaxChunks) s_new[tid] = 0;
    __syncthreads();
    if (cell < ncell) {
        int run = 0;
        const size_t t0 = (size_t)b * S * ncellp + cell;
        for (int s0 = 0; s0 < Sb; s0 += 8) {
            int h[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) h[k] = s0 + k < Sb ? (int)__ldcg(&hist[t0 + (size_t)(s0 + k) * ncellp]) : 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (s0 + k < Sb) {
                    base8[t0 + (size_t)(s0 + k) * ncellp] = (unsigned char)min(run, 255);
                    if (run == 0 && h[k] > 0) atomicAdd(&s_new[s0 + k], 1);
                    run += h[k];
                }
            }
        }
        tot8[(size_t)b * ncellp + cell] = (unsigned char)min(run, 255);
    }
    __syncthreads();
    if (tid < Sb && s_new[tid]) atomicAdd(&newcount[b * S + tid], s_new[tid]);
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(done_counter, 1) == (int)(gridDim.x * gridDim.y) - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    int running = 0;
    for (int b0 = 0; b0 < B; b0 += kPrefixThreads) {
        const int i = b0 + tid;
        int v = 0;
        if (i < B) {
            for (int s = 0; s < S; ++s) v += __ldcg(&newcount[i * S + s]);  // chunks past the frame's end hold 0
            v = min(v, max_voxels);
            voxel_num[i] = v;
        }
        int tot;
        const int e = running + block_excl_scan(v, &tot, sm);
        if (i < B) voxel_base[i] = e;
        running += tot;
    }
    if (tid == 0) {
        voxel_base[B] = running;
        *done_counter = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// Pass 3: one warp per (frame, chunk).  tbl: running count per cell, uint8, starts at the chunk's base.
template <int REC16, bool TAG_INSIDE>
__global__ void __launch_bounds__(32)
vox_place_kernel(const int64_t* __restrict__ frame_off, int S, int ncell, int ncellp, int P, int max_voxels,
                 const uint4* __restrict__ crec, const unsigned* __restrict__ cmeta,
                 const unsigned char* __restrict__ base8, const int* __restrict__ nvalid,
                 const int* __restrict__ newcount, uint4* __restrict__ staging, unsigned* __restrict__ stag,
                 int* __restrict__ rowcell, int* __restrict__ cutoff) {
    extern __shared__ __align__(16) unsigned char tbl[];  // [ncellp]
    const int lane = lane_id();
    const int b = blockIdx.y, s = blockIdx.x;
    const int64_t f0 = frame_off[b];
    const int n = (int)(frame_off[b + 1] - f0);
    const int Sb = (n + kChunk - 1) >> kChunkShift;
    if (s >= Sb) return;
    const int nval = nvalid[b * S + s];
    // voxels opened by the earlier chunks of the frame
    int newrun = 0;
    for (int k = lane; k < s; k += 32) newrun += newcount[b * S + k];
#pragma unroll
    for (int o = 16; o; o >>= 1) newrun += __shfl_xor_sync(0xffffffffu, newrun, o);
    {
        const uint4* srcb = reinterpret_cast<const uint4*>(base8 + ((size_t)b * S + s) * ncellp);
        uint4* dst = reinterpret_cast<uint4*>(tbl);
        for (int k = lane; k < (ncellp >> 4); k += 32) dst[k] = __ldcg(&srcb[k]);
    }
    __syncwarp();
    const int64_t g0 = f0 + (int64_t)s * kChunk;
    const size_t cellrow0 = (size_t)b * ncell;

    // records of the next kPlaceUnroll steps are requested while the current ones are processed
    uint4 nx[kPlaceUnroll][REC16];
    unsigned nt[kPlaceUnroll];
    auto fetch = [&](int g) {
#pragma unroll
        for (int u = 0; u < kPlaceUnroll; ++u) {
            const int i = g + u * 32 + lane;
            nt[u] = 0u;
            if (i < nval) {
#pragma unroll
                for (int k = 0; k < REC16; ++k) nx[u][k] = __ldcs(&crec[(g0 + i) * REC16 + k]);
                if (!TAG_INSIDE) nt[u] = __ldcs(&cmeta[g0 + i]);
            }
        }
    };
    fetch(0);
    for (int g = 0; g < nval; g += kPlaceUnroll * 32) {
        uint4 rv[kPlaceUnroll][REC16];
        unsigned tg[kPlaceUnroll];
#pragma unroll
        for (int u = 0; u < kPlaceUnroll; ++u) {
#pragma unroll
            for (int k = 0; k < REC16; ++k) rv[u][k] = nx[u][k];
            tg[u] = TAG_INSIDE ? rec_tag<REC16>(nx[u]) : nt[u];
        }
        if (g + kPlaceUnroll * 32 < nval) fetch(g + kPlaceUnroll * 32);
#pragma unroll
        for (int u = 0; u < kPlaceUnroll; ++u) {
            if (g + u * 32 < nval) {  // uniform
                const int i = g + u * 32 + lane;
                const bool act = i < nval;
                const int c = act ? (int)(tg[u] & 0xffffu) : (0x10000 | lane);
                const unsigned peers = __match_any_sync(0xffffffffu, c);
                const int rk = __popc(peers & lanemask_lt());
                const bool isldr = act && rk == 0;
                const int cnt = act ? (int)tbl[c] : 0;
                // a count of 0 means no earlier point of the frame fell in this cell: this record opens a voxel
                const unsigned opens = __ballot_sync(0xffffffffu, isldr && cnt == 0);
                const int orig = (s << kChunkShift) | (int)(tg[u] >> 16);  // index of the point in its frame
                if (isldr && cnt == 0) {
                    const int rank = newrun + __popc(opens & lanemask_lt());
                    if (rank < max_voxels) rowcell[cellrow0 + rank] = c;
                    else if (rank == max_voxels) cutoff[b] = orig;  // the reference's break position
                }
                newrun += __popc(opens);
                if (isldr) tbl[c] = (unsigned char)min(cnt + __popc(peers), 255);
                __syncwarp();
                const int slot = cnt + rk;
                if (act && slot < P) {
                    const size_t si = (cellrow0 + c) * (size_t)P + slot;
                    if (TAG_INSIDE) rec_set_tag<REC16>(rv[u], (unsigned)orig);
                    else stag[si] = (unsigned)orig;
#pragma unroll
                    for (int k = 0; k < REC16; ++k) staging[si * REC16 + k] = rv[u][k];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Pass 4: one warp per pillar.
template <typename TO, int DS>
__global__ void __launch_bounds__(kFinishWarps * 32)
vox_finish_kernel(const int64_t* __restrict__ frame_off, VoxParams p, int ncellp, const int* __restrict__ rowcell,
                  const unsigned char* __restrict__ tot8, const int* __restrict__ voxel_num,
                  const int* __restrict__ voxel_base, const int* __restrict__ cutoff, int64_t cap_rows,
                  const uint4* __restrict__ staging, const unsigned* __restrict__ stag, TO* __restrict__ voxels,
                  float* __restrict__ decorated, int* __restrict__ coors, int coors_cols, int* __restrict__ num_points,
                  int* __restrict__ point_slot, int* __restrict__ cell_voxel) {
    extern __shared__ __align__(16) unsigned char fsm[];
    constexpr int D = DS, Do = DS + 5;
    constexpr int rec16 = RecFmt<TO, DS>::kRec16;
    constexpr bool tag_inside = RecFmt<TO, DS>::kTagInside;
    const int P = p.max_points;
    const int lane = lane_id(), w = threadIdx.x >> 5;
    const int nvox = (P * D + 3) & ~3, ndec = DS == 3 ? 0 : ((P * Do + 3) & ~3);
    float* vrow = reinterpret_cast<float*>(fsm) + (size_t)w * (nvox + ndec);
    float* dsm = vrow + nvox;
    const int b = blockIdx.y;
    const int M = voxel_num[b], vb = voxel_base[b];
    const int cut = cutoff[b];
    const int64_t f0 = frame_off[b];
    for (int rank = blockIdx.x * kFinishWarps + w; rank < M; rank += gridDim.x * kFinishWarps) {
        const int64_t row = (int64_t)vb + rank;
        if (row >= cap_rows) break;
        const int cell = rowcell[(size_t)b * p.ncell + rank];
        const int tot = min((int)tot8[(size_t)b * ncellp + cell], P);
        const int cz = p.div_nxny.div(cell);
        const int rem = cell - cz * p.grid[0] * p.grid[1];
        const int cy = p.div_nx.div(rem), cx = rem - cy * p.grid[0];
        const size_t si0 = ((size_t)b * p.ncell + cell) * (size_t)P;
        float sx = 0.f, sy = 0.f, sz = 0.f;
        int nsel = 0;
        for (int s0 = 0; s0 < tot; s0 += 32) {
            const int s = s0 + lane;
            uint4 rv[rec16];
            unsigned tag = 0x7fffffffu;
            if (s < tot) {
#pragma unroll
                for (int k = 0; k < rec16; ++k) rv[k] = __ldcs(&staging[(si0 + s) * rec16 + k]);
                tag = tag_inside ? rec_tag<rec16>(rv) : __ldcs(&stag[si0 + s]);
            }
            // slots are in index order, so the records before the break position are a prefix of the row
            const bool ok = s < tot && (int)tag < cut;
            nsel += __popc(__ballot_sync(0xffffffffu, ok));
            if (ok) {
                float c[DS];
                if (sizeof(TO) == 4) {
                    c[0] = __uint_as_float(rv[0].x); c[1] = __uint_as_float(rv[0].y); c[2] = __uint_as_float(rv[0].z);
                    if (DS == 4) c[DS - 1] = __uint_as_float(rv[0].w);
                } else {
                    const double v0x = __hiloint2double((int)rv[0].y, (int)rv[0].x), v0y = __hiloint2double((int)rv[0].w, (int)rv[0].z);
                    const double v1x = __hiloint2double((int)rv[rec16 - 1].y, (int)rv[rec16 - 1].x);
                    const double v1y = __hiloint2double((int)rv[rec16 - 1].w, (int)rv[rec16 - 1].z);
                    TO* vo = voxels + (row * (int64_t)P + s) * D;
                    vo[0] = (TO)v0x; vo[1] = (TO)v0y; vo[2] = (TO)v1x;
                    if (DS == 4) vo[DS - 1] = (TO)v1y;
                    c[0] = (float)v0x; c[1] = (float)v0y; c[2] = (float)v1x;
                    if (DS == 4) c[DS - 1] = (float)v1y;
                }
#pragma unroll
                for (int d = 0; d < DS; ++d) vrow[s * D + d] = c[d];
                sx += c[0]; sy += c[1]; sz += c[2];
                if (point_slot) point_slot[f0 + tag] = rank * P + s;
            }
        }
        if (lane == 0) {
            num_points[row] = nsel;
            int* co = coors + row * coors_cols;
            if (coors_cols == 4) *co++ = b;
            if (p.reverse_index) { co[0] = cz; co[1] = cy; co[2] = cx; }
            else { co[0] = cx; co[1] = cy; co[2] = cz; }
            if (cell_voxel) cell_voxel[(size_t)b * p.ncell + cell] = (int)row;
        }
        if (sizeof(TO) == 8) {
            TO* vo = voxels + row * (int64_t)P * D;
            for (int k = nsel * D + lane; k < P * D; k += 32) vo[k] = (TO)0;
        }
        __syncwarp();
        if (voxels && sizeof(TO) == 4)  // padding (k >= nsel*D) is written as zeros without touching smem
            warp_store_row_padded(reinterpret_cast<float*>(voxels) + row * (int64_t)P * D, vrow, P * D, nsel * D, lane);
        if (decorated) {
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                sx += __shfl_xor_sync(0xffffffffu, sx, o);
                sy += __shfl_xor_sync(0xffffffffu, sy, o);
                sz += __shfl_xor_sync(0xffffffffu, sz, o);
            }
            const float nf = (float)nsel;
            const float mx = __fdiv_rn(sx, nf), my = __fdiv_rn(sy, nf), mz = __fdiv_rn(sz, nf);
            const float ex = __fadd_rn(__fmul_rn((float)cx, p.vx), p.x_off);
            const float ey = __fadd_rn(__fmul_rn((float)cy, p.vy), p.y_off);
            float* drow = decorated + row * (int64_t)P * Do;
            if (DS == 3 && (reinterpret_cast<uintptr_t>(decorated) & 15) == 0) {
                // 8 floats per point = two float4: (x,y,z,x-mx) and (y-my,z-mz,x-ex,y-ey)
                float4* d4 = reinterpret_cast<float4*>(drow);
                for (int s = lane; s < P; s += 32) {
                    float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
                    if (s < nsel) {
                        const float q0 = vrow[s * 3], q1 = vrow[s * 3 + 1], q2 = vrow[s * 3 + 2];
                        o0 = make_float4(q0, q1, q2, q0 - mx);
                        o1 = make_float4(q1 - my, q2 - mz, q0 - ex, q1 - ey);
                    }
                    d4[2 * s] = o0;
                    d4[2 * s + 1] = o1;
                }
            } else if (DS == 3) {
                for (int k = lane; k < P * 8; k += 32) {
                    const int s = k >> 3, d = k & 7;
                    float o = 0.f;
                    if (s < nsel) {
                        const float* q = vrow + s * 3;
                        o = d < 3 ? q[d] : d == 3 ? q[0] - mx : d == 4 ? q[1] - my : d == 5 ? q[2] - mz : d == 6 ? q[0] - ex : q[1] - ey;
                    }
                    drow[k] = o;
                }
            } else {
                // one point per lane into shared memory (stride Do words), then 16-byte row stores
                for (int s = lane; s < P; s += 32) {
                    float* o = dsm + s * Do;
                    if (s < nsel) {
                        const float* q = vrow + s * D;
#pragma unroll
                        for (int d = 0; d < DS; ++d) o[d] = q[d];
                        o[D] = q[0] - mx; o[D + 1] = q[1] - my; o[D + 2] = q[2] - mz;
                        o[D + 3] = q[0] - ex; o[D + 4] = q[1] - ey;
                    } else {
                        for (int d = 0; d < Do; ++d) o[d] = 0.f;
                    }
                }
                __syncwarp();
                warp_store_row(drow, dsm, P * Do, lane);
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
struct SmallWs {
    uint4* crec;              // [total_points + 1] records, compacted per chunk at the chunk's own offset
    unsigned* cmeta;          // [total_points + 1] cell | index in chunk << 16 (only when the record has no spare element)
    unsigned short* hist;     // [B*S*ncellp]  chunk counts
    unsigned char* base8;     // [B*S*ncellp]  chunk bases
    unsigned char* tot8;      // [B*ncellp]    points per cell, saturated
    int* nvalid;              // [B*S] records per chunk
    int* newcount;            // [B*S] voxels opened per chunk
    int* rowcell;             // [B*ncell] voxel id in frame -> cell
    int* cutoff;              // [B] break position (point index in frame) or kNoCut
    int* done_counter;        // [1]
    uint4* staging;           // [B*ncell*P] records per (cell, slot)
    unsigned* stag;           // [B*ncell*P] point index per (cell, slot) (only when the record has no spare element)
    size_t total;
};

static int rec_bytes_of(int D, int out_dtype) { return (int)(((size_t)D * (out_dtype == PP_F64 ? 8 : 4) + 15) / 16 * 16); }
static int chunks_of(int64_t max_frame_points) { return (int)(max_frame_points > 0 ? ceil_div(max_frame_points, kChunk) : 1); }

static SmallWs carve_small(void* ws, const pp_voxel_cfg* cfg, int64_t ncell, int n_frames, int64_t total_points,
                           int64_t max_frame_points, int D, int out_dtype) {
    SmallWs w;
    Carver c(ws);
    const int S = chunks_of(max_frame_points);
    const size_t ncellp = (size_t)align_up((size_t)ncell, 16);
    const size_t rb = (size_t)rec_bytes_of(D, out_dtype);
    const bool tag_inside = D == 3;
    w.crec = reinterpret_cast<uint4*>(c.take<unsigned char>(((size_t)total_points + 1) * rb));
    w.cmeta = tag_inside ? nullptr : c.take<unsigned>((size_t)total_points + 1);
    w.hist = c.take<unsigned short>((size_t)n_frames * S * ncellp);
    w.base8 = c.take<unsigned char>((size_t)n_frames * S * ncellp);
    w.tot8 = c.take<unsigned char>((size_t)n_frames * ncellp);
    w.nvalid = c.take<int>((size_t)n_frames * S);
    w.newcount = c.take<int>((size_t)n_frames * S);
    w.rowcell = c.take<int>((size_t)n_frames * ncell);
    w.cutoff = c.take<int>(n_frames);
    w.done_counter = c.take<int>(1);
    const size_t slots = (size_t)n_frames * ncell * cfg->max_points;
    w.staging = reinterpret_cast<uint4*>(c.take<unsigned char>(slots * rb));
    w.stag = tag_inside ? nullptr : c.take<unsigned>(slots);
    w.total = c.used();
    return w;
}

bool vox_small_eligible(const pp_voxel_cfg* cfg, int64_t ncell, int n_frames, int64_t total_points,
                        int64_t max_frame_points, int D) {
    if (ncell > kMaxCellsSmall || D > 4 || max_frame_points > kMaxFramePointsSmall || cfg->max_points > kMaxPointsSmall)
        return false;
    // the per-chunk tables and the per-cell staging rows must stay small next to the points themselves (callers
    // that do not know the largest frame pass the batch total, which sizes one table set per 16 384 points for
    // every frame)
    const double ncellp = (double)align_up((size_t)ncell, 16);
    const double tables = (double)n_frames * chunks_of(max_frame_points) * ncellp * 3.0;
    const double staging = (double)n_frames * (double)ncell * cfg->max_points * 36.0;
    const double budget = 64.0 * (double)total_points + (double)(256 << 20);
    return tables + staging <= budget;
}

size_t vox_small_workspace_bytes(const pp_voxel_cfg* cfg, int64_t ncell, int n_frames, int64_t total_points,
                                 int64_t max_frame_points, int D, int out_dtype) {
    return carve_small(nullptr, cfg, ncell, n_frames, total_points, max_frame_points, D, out_dtype).total;
}

template <typename T, bool A32, bool FAST, typename TO, int DS>
static int launch_scan(const VoxParams& p, const SmallWs& w, const void* points, const int64_t* frame_off, int64_t total_points,
                       int S, int ncellp, int n_frames, int32_t* point_slot, cudaStream_t st) {
    const size_t smem = ((size_t)kScanTile * DS * sizeof(T) + 32) + (size_t)ncellp * 2;
    auto kern = vox_scan_kernel<T, A32, FAST, TO, DS>;
    int per_sm = 0;
    PP_TRY_RC(kernel_config(reinterpret_cast<const void*>(kern), kScanThreads, smem, &per_sm));
    const int aligned16 = (reinterpret_cast<uintptr_t>(points) & 15) == 0;
    PP_TIMED("vox_scan", st);
    kern<<<dim3((unsigned)S, (unsigned)n_frames), kScanThreads, smem, st>>>(
        static_cast<const T*>(points), frame_off, p, total_points, aligned16, S, ncellp, w.crec, w.cmeta,
        reinterpret_cast<unsigned*>(w.hist), w.nvalid, w.newcount, w.cutoff, point_slot, w.done_counter);
    PP_LAUNCHED();
    return PP_OK;
}

template <typename TO, int DS>
static int launch_finish(const VoxParams& p, const SmallWs& w, const int64_t* frame_off, int ncellp, int n_frames,
                         int64_t rows_per_frame, const int32_t* voxel_num, const int32_t* voxel_base, int64_t cap_rows,
                         void* voxels, float* decorated, int32_t* coors, int coors_cols, int32_t* num_points,
                         int32_t* point_slot, int32_t* cell_voxel, cudaStream_t st) {
    const int P = p.max_points;
    const size_t per_warp = (size_t)(((P * DS + 3) & ~3) + (DS == 3 ? 0 : ((P * (DS + 5) + 3) & ~3))) * 4;
    const size_t smem = kFinishWarps * per_warp;
    auto kern = vox_finish_kernel<TO, DS>;
    int per_sm = 0;
    PP_TRY_RC(kernel_config(reinterpret_cast<const void*>(kern), kFinishWarps * 32, smem, &per_sm));
    // about four pillars per warp; the grid's y dimension is the frame
    int64_t gx = ceil_div(rows_per_frame, kFinishWarps * 4);
    const int64_t resident = (int64_t)num_sms() * (per_sm > 0 ? per_sm : 1);
    const int64_t min_gx = ceil_div(resident, n_frames);
    if (gx < min_gx) gx = min_gx;  // few frames: spread each frame over the whole GPU
    if (gx > ceil_div(rows_per_frame, kFinishWarps)) gx = ceil_div(rows_per_frame, kFinishWarps);
    if (gx < 1) gx = 1;
    PP_TIMED("vox_finish", st);
    kern<<<dim3((unsigned)gx, (unsigned)n_frames), kFinishWarps * 32, smem, st>>>(
        frame_off, p, ncellp, w.rowcell, w.tot8, voxel_num, voxel_base, w.cutoff, cap_rows, w.staging, w.stag,
        static_cast<TO*>(voxels), decorated, coors, coors_cols, num_points, point_slot, cell_voxel);
    PP_LAUNCHED();
    return PP_OK;
}

int vox_small_run(const pp_voxel_cfg* cfg, const VoxParams& p, const void* points, int point_dtype,
                  const int64_t* frame_offsets, int n_frames, int64_t total_points, int64_t max_frame_points,
                  int out_dtype, void* voxels, float* decorated, int32_t* coors, int coors_cols, int32_t* num_points,
                  int64_t cap_rows, int32_t* voxel_num, int32_t* voxel_base, int32_t* point_slot, int32_t* cell_voxel,
                  void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const int D = p.D, P = p.max_points;
    const int64_t ncell = p.ncell;
    const int ncellp = (int)align_up((size_t)ncell, 16);
    const int S = chunks_of(max_frame_points);
    const SmallWs w = carve_small(workspace, cfg, ncell, n_frames, total_points, max_frame_points, D, out_dtype);
    if (w.total > workspace_bytes) {
        set_error("pp_voxelize_dev: workspace %zu < required %zu", workspace_bytes, w.total);
        return PP_E_WORKSPACE;
    }
    const bool fast = !cfg->arith_f32 && p.grid[0] <= 2047 && p.grid[1] <= 2047 && p.grid[2] <= 2047;
    if (cell_voxel) PP_CUDA(cudaMemsetAsync(cell_voxel, 0xff, (size_t)n_frames * ncell * sizeof(int32_t), st));

    int rc;
#define PP_SCAN(T, A32, FAST, TO, DS) \
    launch_scan<T, A32, FAST, TO, DS>(p, w, points, frame_offsets, total_points, S, ncellp, n_frames, point_slot, st)
#define PP_SCAN_D(T, A32, FAST, TO) (D == 3 ? PP_SCAN(T, A32, FAST, TO, 3) : PP_SCAN(T, A32, FAST, TO, 4))
    if (point_dtype == PP_F64 && out_dtype == PP_F64) rc = fast ? PP_SCAN_D(double, false, true, double) : PP_SCAN_D(double, false, false, double);
    else if (point_dtype == PP_F64) rc = fast ? PP_SCAN_D(double, false, true, float) : PP_SCAN_D(double, false, false, float);
    else if (cfg->arith_f32) rc = PP_SCAN_D(float, true, false, float);
    else rc = fast ? PP_SCAN_D(float, false, true, float) : PP_SCAN_D(float, false, false, float);
#undef PP_SCAN_D
#undef PP_SCAN
    if (rc) return rc;
    {
        const dim3 g((unsigned)ceil_div(ncell, kPrefixThreads), (unsigned)n_frames);
        PP_TIMED("vox_prefix", st);
        vox_prefix_kernel<<<g, kPrefixThreads, 0, st>>>(frame_offsets, S, (int)ncell, ncellp, w.hist, w.base8, w.tot8,
                                                        w.newcount, cfg->max_voxels, n_frames, voxel_num, voxel_base,
                                                        w.done_counter);
        PP_LAUNCHED();
    }
    if (cap_rows <= 0) return PP_OK;
    {
        const size_t smem = (size_t)ncellp;
        const dim3 g((unsigned)S, (unsigned)n_frames);
        const int rec16 = rec_bytes_of(D, out_dtype) / 16;
        PP_TIMED("vox_place", st);
#define PP_PLACE(R, TI)                                                                                                  \
    vox_place_kernel<R, TI><<<g, 32, smem, st>>>(frame_offsets, S, (int)ncell, ncellp, P, cfg->max_voxels, w.crec, w.cmeta, \
                                                 w.base8, w.nvalid, w.newcount, w.staging, w.stag, w.rowcell, w.cutoff)
        if (rec16 == 1) { if (D == 3) PP_PLACE(1, true); else PP_PLACE(1, false); }
        else { if (D == 3) PP_PLACE(2, true); else PP_PLACE(2, false); }
#undef PP_PLACE
        PP_LAUNCHED();
    }
    const int64_t rows_per_frame = cfg->max_voxels < ncell ? cfg->max_voxels : ncell;
    if (rows_per_frame <= 0) return PP_OK;
#define PP_FINISH(TO, DS)                                                                                            \
    launch_finish<TO, DS>(p, w, frame_offsets, ncellp, n_frames, rows_per_frame, voxel_num, voxel_base, cap_rows,  \
                          voxels, decorated, coors, coors_cols, num_points, point_slot, cell_voxel, st)
    if (out_dtype == PP_F64) rc = D == 3 ? PP_FINISH(double, 3) : PP_FINISH(double, 4);
    else rc = D == 3 ? PP_FINISH(float, 3) : PP_FINISH(float, 4);
#undef PP_FINISH
    return rc;
}

}  // namespace pp
