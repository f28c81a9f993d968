// Deterministic first-come pillarization on sm_100a.
//
// Replaces the sequential numba loop of the reference (load_data.py:593-692, wrapper 695-771)
// with six data-parallel passes, batched over frames, that reproduce its results bit for bit:
//
//   mark    (points)  persistent CTAs, 1024-point tiles staged by TMA bulk copies (two stages);
//                     cell id per point in the reference's float64/float32 arithmetic;
//                     atomicMin(first_idx[cell], i); pos = atomicAdd(count[cell]) (warp-aggregated)
//   cell    (cells)   occupied cells: set bit first_idx in a per-frame bitmap over point indices,
//                     reserve count[cell] bucket entries, emit a 16-byte descriptor per occupied cell
//   rank    (frames)  popcount prefix over the bitmap: voxel id of a cell = number of set bits
//                     below its first_idx == order of first touch.  The bit of rank max_voxels is
//                     the reference's `break` position i*: every point >= i* is dropped
//                     (load_data.py:630-634).  Last block scans voxel counts into packed row bases.
//   rowmap  (cells)   rank -> packed output row (or -1 past the cap); coors; optional cell->row map
//   bucket  (points)  bucket[offset[cell] + pos] = i      (unordered inside a cell)
//   gather  (voxels)  one warp per kept voxel, two-deep software pipeline: bucket into registers, cut
//                     at i*, register bitonic sort => slot order (the max_points smallest indices),
//                     gather the points once, write the zero-padded voxel row + fused decoration.
//
// The only nondeterminism is the order inside a bucket, which the sort in the gather pass removes.
#include <math.h>
#include <stdlib.h>

#include "pp_common.cuh"
#include "vox_common.cuh"
#include "vox_internal.h"

namespace pp {

constexpr int kMarkThreads = 256;
constexpr int kCellThreads = 256;
constexpr int kRankThreads = 1024;
#ifndef PP_GATHER_WARPS
#define PP_GATHER_WARPS 8
#endif
#ifndef PP_GATHER_MINBLOCKS
#define PP_GATHER_MINBLOCKS 5
#endif
constexpr int kGatherWarps = PP_GATHER_WARPS;

__device__ __forceinline__ int64_t word_base(const int64_t* frame_off, int b) {
    return (frame_off[b] >> 5) + b;
}

// ---------------------------------------------------------------------------------------------
// Pass 1: kMarkPPT points per thread.  The block's rows are staged through shared memory with
// 16-byte loads (rows are 12/16/24/32 bytes, so per-thread row loads would be strided); round r
// of thread t handles point base + r*256 + t, so lanes of a warp always hold consecutive,
// increasing point indices.  The rounds are independent: their atomics are in flight together.
#ifndef PP_MARK_PPT
#define PP_MARK_PPT 4
#endif
constexpr int kMarkPPT = PP_MARK_PPT;
constexpr int kMarkTile = kMarkThreads * kMarkPPT;

// Persistent CTAs, two shared-memory stages: while a tile is being processed the TMA bulk copy
// of the CTA's next tile is already in flight (the un-pipelined version spent a third of its
// stall samples waiting on the tile load).  Tiles are numbered frame-major:
// tile = frame_in_chunk * tiles_per_frame + tile_in_frame.
template <typename T, bool A32>
__global__ void __launch_bounds__(kMarkThreads)
vox_mark_kernel(const T* __restrict__ points, const int64_t* __restrict__ frame_off, VoxParams p,
                int64_t total_points, int aligned16, int b0, int n_frames, int tiles_per_frame,
                unsigned* __restrict__ first_idx, int* __restrict__ cnt, int2* __restrict__ cellpos,
                int* __restrict__ point_slot) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) unsigned long long s_bar[2];
    const int row_bytes = p.D * (int)sizeof(T);
    const int stage_bytes = kMarkTile * row_bytes + 32;  // multiple of 16
    const int64_t total_bytes = total_points * (int64_t)row_bytes;
    const int64_t tail0 = total_bytes & ~(int64_t)15;  // end of the buffer's last full 16-byte chunk
    const unsigned char* src = reinterpret_cast<const unsigned char*>(points);
    const int total_tiles = n_frames * tiles_per_frame;

    if (threadIdx.x == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); }
    __syncthreads();

    // tile -> (frame, first point, count, byte range); all threads compute the same values
    struct Tile { int b, base, m; int64_t f0, a0, a1; int shift; };
    auto locate = [&](int tile, Tile& t) -> bool {
        t.b = b0 + tile / tiles_per_frame;
        t.f0 = frame_off[t.b];
        const int n = (int)(frame_off[t.b + 1] - t.f0);
        t.base = (tile % tiles_per_frame) * kMarkTile;
        if (t.base >= n) return false;
        t.m = min(kMarkTile, n - t.base);
        const int64_t start = (t.f0 + t.base) * (int64_t)row_bytes;
        const int64_t end = start + (int64_t)t.m * row_bytes;
        t.a0 = start & ~(int64_t)15;
        t.shift = (int)(start - t.a0);
        t.a1 = (end + 15) & ~(int64_t)15;
        if (t.a1 > tail0) t.a1 = tail0 > t.a0 ? tail0 : t.a0;
        return true;
    };
    auto issue = [&](const Tile& t, int stage) {
        const unsigned bulk = (unsigned)(t.a1 - t.a0);
        if (aligned16 && bulk && threadIdx.x == 0) {
            mbar_arrive_expect_tx(&s_bar[stage], bulk);
            tma_bulk_g2s(smem + (size_t)stage * stage_bytes, src + t.a0, bulk, &s_bar[stage]);
        }
    };

    unsigned phase0 = 0, phase1 = 0;
    int tile = blockIdx.x;
    Tile cur, nxt;
    bool have_cur = false;
    while (tile < total_tiles && !(have_cur = locate(tile, cur))) tile += gridDim.x;
    if (have_cur) issue(cur, 0);
    for (int it = 0; have_cur; ++it) {
        const int stage = it & 1;
        // look up and start loading the next non-empty tile of this CTA
        int ntile = tile + gridDim.x;
        bool have_nxt = false;
        while (ntile < total_tiles && !(have_nxt = locate(ntile, nxt))) ntile += gridDim.x;
        if (have_nxt) issue(nxt, stage ^ 1);

        unsigned char* buf = smem + (size_t)stage * stage_bytes;
        const int64_t end = (cur.f0 + cur.base + cur.m) * (int64_t)row_bytes;
        if (aligned16) {
            if (cur.a1 > cur.a0) {
                mbar_wait(&s_bar[stage], stage ? phase1 : phase0);
                if (stage) phase1 ^= 1; else phase0 ^= 1;
            }
            if (cur.a1 < end) {
                // bytes past the buffer's last full chunk (only the very last tile of the buffer)
                for (int64_t a = cur.a1 + (int64_t)threadIdx.x * (int)sizeof(T); a < end && a < total_bytes;
                     a += (int64_t)kMarkThreads * (int)sizeof(T))
                    *reinterpret_cast<T*>(buf + (a - cur.a0)) = *reinterpret_cast<const T*>(src + a);
                __syncthreads();
            }
        } else {
            const int nel = cur.m * p.D;
            for (int k = threadIdx.x; k < nel; k += kMarkThreads)
                reinterpret_cast<T*>(buf)[k] = points[(cur.f0 + cur.base) * p.D + k];
            __syncthreads();
        }
        const int shift = aligned16 ? cur.shift : 0;
        const int b = cur.b, base = cur.base, m = cur.m;
        const int64_t f0 = cur.f0;

        int cell[kMarkPPT];
#pragma unroll
        for (int r = 0; r < kMarkPPT; ++r) {
            const int t = r * kMarkThreads + threadIdx.x;
            cell[r] = t < m ? cell_of<T, A32>(reinterpret_cast<const T*>(buf + shift + (size_t)t * row_bytes), p) : -1;
        }
        unsigned peers[kMarkPPT];
        int basepos[kMarkPPT];
#pragma unroll
        for (int r = 0; r < kMarkPPT; ++r) {
            peers[r] = __match_any_sync(0xffffffffu, cell[r]);
            basepos[r] = 0;
            if (cell[r] >= 0 && (int)lane_id() == __ffs(peers[r]) - 1) {
                // lanes are in index order, so the leader carries the group's smallest index
                const size_t gc = (size_t)b * p.ncell + cell[r];
                // (guarding the atomicMin with a load of the current minimum was measured slower: 433 vs 316 us)
                atomicMin(&first_idx[gc], (unsigned)(base + r * kMarkThreads + threadIdx.x));
                basepos[r] = atomicAdd(&cnt[gc], __popc(peers[r]));
            }
        }
#pragma unroll
        for (int r = 0; r < kMarkPPT; ++r) {
            const int t = r * kMarkThreads + threadIdx.x;
            const int bp = __shfl_sync(0xffffffffu, basepos[r], __ffs(peers[r]) - 1);
            if (t < m) {
                cellpos[f0 + base + t] = make_int2(cell[r], bp + __popc(peers[r] & lanemask_lt()));
                if (point_slot) point_slot[f0 + base + t] = -1;
            }
        }
        __syncthreads();  // every thread is done with this stage before it is refilled
        cur = nxt;
        tile = ntile;
        have_cur = have_nxt;
    }
}

// ---------------------------------------------------------------------------------------------
// Pass 2: four consecutive cells per thread.  Blocks that see no occupied cell (most of a KITTI
// grid) leave after one barrier.
constexpr int kCellPerThread = 4;
__global__ void __launch_bounds__(kCellThreads)
vox_cell_kernel(const unsigned* __restrict__ first_idx, const int* __restrict__ cnt,
                const int64_t* __restrict__ frame_off, int ncell, int b0, unsigned* __restrict__ bitmap,
                int* __restrict__ cell_off, int* __restrict__ frame_cursor,
                int4* __restrict__ occ_desc, int* __restrict__ frame_occ,
                int* __restrict__ cell_voxel) {
    __shared__ int sm[33];
    __shared__ int s_base, s_obase;
    const int b = b0 + blockIdx.y;
    const int cell0 = (blockIdx.x * kCellThreads + threadIdx.x) * kCellPerThread;
    const size_t gc0 = (size_t)b * ncell + cell0;
    int c[kCellPerThread];
    if (cell0 + kCellPerThread <= ncell && ((gc0 & 3) == 0)) {
        const int4 v = *reinterpret_cast<const int4*>(cnt + gc0);
        c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
    } else {
#pragma unroll
        for (int j = 0; j < kCellPerThread; ++j) c[j] = cell0 + j < ncell ? cnt[gc0 + j] : 0;
    }
    if (cell_voxel) {
#pragma unroll
        for (int j = 0; j < kCellPerThread; ++j) if (cell0 + j < ncell) cell_voxel[gc0 + j] = -1;
    }
    int tcnt = 0, tocc = 0;
#pragma unroll
    for (int j = 0; j < kCellPerThread; ++j) { tcnt += c[j]; tocc += c[j] > 0; }
    if (!__syncthreads_or(tocc)) return;
    int tot, otot;
    int ex = block_excl_scan(tcnt, &tot, sm);
    int oex = block_excl_scan(tocc, &otot, sm);
    if (threadIdx.x == 0) {
        s_base = atomicAdd(&frame_cursor[b], tot);
        s_obase = atomicAdd(&frame_occ[b], otot);
    }
    __syncthreads();
    ex += s_base;
    oex += s_obase;
    const int64_t f0 = frame_off[b];
    const int64_t wb = (f0 >> 5) + b;
#pragma unroll
    for (int j = 0; j < kCellPerThread; ++j) {
        if (c[j] > 0) {
            const unsigned f = first_idx[gc0 + j];
            atomicOr(&bitmap[wb + (f >> 5)], 1u << (f & 31));
            cell_off[gc0 + j] = ex;  // frame-local offset into the frame's bucket range
            // one 16-byte descriptor per occupied cell, listed frame by frame: everything the later
            // passes need comes from a single coalesced load {cell, first index, count, bucket offset}
            occ_desc[(size_t)b * ncell + oex] = make_int4(cell0 + j, (int)f, c[j], (int)(f0 + ex));
            ex += c[j];
            ++oex;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Pass 3: one block per frame: exclusive popcount prefix over the frame's bitmap words, voxel
// count, break position; the last block to finish scans the voxel counts of all frames.
__global__ void __launch_bounds__(kRankThreads)
vox_rank_kernel(const unsigned* __restrict__ bitmap, unsigned* __restrict__ word_prefix,
                const int64_t* __restrict__ frame_off, int b0, int B, int max_voxels,
                int* __restrict__ voxel_num, int* __restrict__ cutoff, int* __restrict__ voxel_base,
                const int* __restrict__ frame_occ, int* __restrict__ occ_base, int* __restrict__ done_counter) {
    __shared__ int sm[33];
    __shared__ int s_cut, s_last;
    const int b = b0 + blockIdx.x;
    const int n = (int)(frame_off[b + 1] - frame_off[b]);
    const int words = (n + 31) >> 5;
    const int64_t wb = word_base(frame_off, b);
    if (threadIdx.x == 0) s_cut = 0x7fffffff;
    __syncthreads();
    // each thread owns a contiguous run of words: one block scan per frame instead of one per 1024 words
    const int wpt = (words + kRankThreads - 1) / kRankThreads;
    const int w0 = threadIdx.x * wpt, w1 = min(words, w0 + wpt);
    int mine = 0;
    for (int w = w0; w < w1; ++w) mine += __popc(bitmap[wb + w]);
    int running;
    int ex = block_excl_scan(mine, &running, sm);
    for (int w = w0; w < w1; ++w) {
        const unsigned bits = bitmap[wb + w];
        const int pc = __popc(bits);
        word_prefix[wb + w] = (unsigned)ex;
        if (ex <= max_voxels && max_voxels < ex + pc) s_cut = 32 * w + (int)__fns(bits, 0, max_voxels - ex + 1);
        ex += pc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        voxel_num[b] = min(running, max_voxels);
        cutoff[b] = s_cut;
        __threadfence();
        s_last = (atomicAdd(done_counter, 1) == B - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // rows continue where the previous chunk of frames stopped (chunks run in stream order)
    running = b0 ? __ldcg(&voxel_base[b0]) : 0;
    for (int s = 0; s < B; s += kRankThreads) {
        const int i = s + threadIdx.x;
        const int v = i < B ? __ldcg(&voxel_num[b0 + i]) : 0;
        int tot;
        const int ex = running + block_excl_scan(v, &tot, sm);
        if (i < B) voxel_base[b0 + i] = ex;
        running += tot;
    }
    if (threadIdx.x == 0) voxel_base[b0 + B] = running;
    // occupied-cell list offsets of the chunk's frames (relative to the chunk)
    running = 0;
    for (int s = 0; s < B; s += kRankThreads) {
        const int i = s + threadIdx.x;
        const int v = i < B ? __ldcg(&frame_occ[b0 + i]) : 0;
        int tot;
        const int ex = running + block_excl_scan(v, &tot, sm);
        if (i < B) occ_base[i] = ex;  // occ_base points at this chunk's [B+1] slice
        running += tot;
    }
    if (threadIdx.x == 0) {
        occ_base[B] = running;
        *done_counter = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// Pass 3b: one thread per occupied cell: voxel id (rank of its first index) -> packed output row,
// coordinates of the row, optional cell->row map.  The descriptor's `first` field becomes the row
// (-1 for cells past the max_voxels cap).
__global__ void __launch_bounds__(256)
vox_rowmap_kernel(int4* __restrict__ occ_desc, const int* __restrict__ occ_base, const int64_t* __restrict__ frame_off,
                  VoxParams p, int b0, const unsigned* __restrict__ bitmap, const unsigned* __restrict__ word_prefix,
                  const int* __restrict__ voxel_base, int64_t cap_rows, int* __restrict__ coors, int coors_cols,
                  int* __restrict__ cell_voxel) {
    const int bl = blockIdx.y, b = b0 + bl;
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= occ_base[bl + 1] - occ_base[bl]) return;
    int4 d = occ_desc[(size_t)b * p.ncell + k];
    const unsigned f = (unsigned)d.y;
    const int64_t wb = word_base(frame_off, b);
    const int rank = (int)word_prefix[wb + (f >> 5)] + __popc(bitmap[wb + (f >> 5)] & ((1u << (f & 31)) - 1u));
    int64_t row = rank < p.max_voxels ? (int64_t)voxel_base[b] + rank : -1;
    if (row >= cap_rows) row = -1;
    d.y = (int)row;
    occ_desc[(size_t)b * p.ncell + k] = d;
    if (row < 0) return;
    const int cell = d.x;
    const int cz = p.div_nxny.div(cell);
    const int rem = cell - cz * p.grid[0] * p.grid[1];
    const int cy = p.div_nx.div(rem), cx = rem - cy * p.grid[0];
    int* co = coors + row * coors_cols;
    if (coors_cols == 4) *co++ = b;
    if (p.reverse_index) { co[0] = cz; co[1] = cy; co[2] = cx; }
    else { co[0] = cx; co[1] = cy; co[2] = cz; }
    if (cell_voxel) cell_voxel[(size_t)b * p.ncell + cell] = (int)row;
}

// ---------------------------------------------------------------------------------------------
// Pass 4: bucket fill, 4 independent points per thread.
constexpr int kBucketPPT = 4;
__global__ void __launch_bounds__(256)
vox_bucket_kernel(const int2* __restrict__ cellpos, const int64_t* __restrict__ frame_off, int ncell, int b0,
                  const int* __restrict__ cell_off, int* __restrict__ bucket) {
    const int b = b0 + blockIdx.y;
    const int64_t f0 = frame_off[b];
    const int n = (int)(frame_off[b + 1] - f0);
    const int base = blockIdx.x * 256 * kBucketPPT + threadIdx.x;
    if (base - (int)threadIdx.x >= n) return;
    int2 cp[kBucketPPT];
    int off[kBucketPPT];
#pragma unroll
    for (int r = 0; r < kBucketPPT; ++r) {
        const int i = base + r * 256;
        cp[r] = i < n ? cellpos[f0 + i] : make_int2(-1, 0);
    }
#pragma unroll
    for (int r = 0; r < kBucketPPT; ++r) off[r] = cp[r].x >= 0 ? cell_off[(size_t)b * ncell + cp[r].x] : 0;
#pragma unroll
    for (int r = 0; r < kBucketPPT; ++r)
        if (cp[r].x >= 0) bucket[f0 + off[r] + cp[r].y] = base + r * 256;
}

// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// Pass 5: one warp per occupied cell.
//
// A cell's bucket (unordered point indices) is loaded into registers, striped over the warp
// (element r*32+lane), cut at the break position, and sorted ascending with a register bitonic
// network: after the sort, slot s of the voxel is element s, so arrival order needs no shared
// memory and no search.  Buckets longer than 256 indices (heavy skew) take a streaming path:
// binary search on the index threshold, compaction, rank-by-counting.
// The selected points are gathered from global memory exactly once (all loads of a round in
// flight together) into a float32 row in shared memory; the voxel row and the fused
// PillarFeatureNet decoration (model/pointpillars.py:143-203) are streamed from it with 16-byte
// stores.  DS = compile-time point width (3, 4) or 0 for a runtime width.
constexpr int kSegRegs = 8;  // register path handles buckets up to 32*kSegRegs indices
constexpr int kIdxInf = 0x7fffffff;

// Zero the floats [begin, end) of a global row with TMA bulk stores (BULK instantiation of the gather pass).
// Pillar rows are mostly padding (D435 46 % of the slots, KITTI 95 %).  For long rows the padding is not pushed
// through the LSU lane by lane: the 16-byte aligned middle goes out as bulk stores from a block of zeros in shared
// memory, issued by one lane (SASS UBLKCP); the <= 3 floats of an unaligned head / tail are ordinary stores.
// Measured on B200, 64 frames: KITTI rows (3.6 KB decorated + 1.6 KB voxel row) 1043 -> 922 us; D435 rows
// (1.6 KB + 0.6 KB) 481 -> 708 us -- a bulk store only amortises over multi-KB runs, so the launcher picks the
// BULK instantiation by row length and the other instantiation is compiled without any of this.
constexpr int kZeroFloats = 1024;      // 4 KB of zeros per CTA
constexpr int kBulkMinRowBytes = 2048;  // decorated-row length from which the BULK instantiation is used
__device__ __forceinline__ void warp_zero_fill(float* __restrict__ begin, float* __restrict__ end, int lane,
                                               const float* __restrict__ s_zero) {
    if (end <= begin) return;
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(begin), a1 = reinterpret_cast<uintptr_t>(end);
    const uintptr_t m0 = (a0 + 15) & ~(uintptr_t)15, m1 = a1 & ~(uintptr_t)15;
    if (m1 <= m0) {  // shorter than one aligned chunk
        for (float* q = begin + lane; q < end; q += 32) *q = 0.f;
        return;
    }
    const int head = (int)((m0 - a0) >> 2), tail = (int)((a1 - m1) >> 2);
    if (lane < head) begin[lane] = 0.f;
    if (lane < tail) reinterpret_cast<float*>(m1)[lane] = 0.f;
    if (lane == 0) {
        for (uintptr_t q = m0; q < m1; q += kZeroFloats * 4) {
            const unsigned bytes = (unsigned)((m1 - q) < (uintptr_t)(kZeroFloats * 4) ? (m1 - q) : (uintptr_t)(kZeroFloats * 4));
            tma_bulk_s2g(reinterpret_cast<void*>(q), s_zero, bytes);
        }
        tma_bulk_commit();
    }
}

// ascending bitonic sort of 32*R ints, element index i = r*32 + lane
template <int R>
__device__ __forceinline__ void warp_bitonic_sort(int (&v)[R], int lane) {
#pragma unroll
    for (int size = 2; size <= 32 * R; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32) {
                const int rs = stride >> 5;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if ((r & rs) == 0) {
                        const bool up = (((r * 32) & size) == 0);
                        const int a = v[r], b = v[r | rs];
                        const int lo = min(a, b), hi = max(a, b);
                        v[r] = up ? lo : hi;
                        v[r | rs] = up ? hi : lo;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int o = __shfl_xor_sync(0xffffffffu, v[r], stride);
                    const bool up = (((r * 32 + lane) & size) == 0);
                    const bool lower = ((lane & stride) == 0);
                    v[r] = (up == lower) ? min(v[r], o) : max(v[r], o);
                }
            }
        }
    }
}

// load + cut + sort a bucket of up to 32*R indices; returns the number of valid entries
template <int R>
__device__ __forceinline__ int load_sort_bucket(const int* __restrict__ seg, int L, int cut, int lane, int (&v)[R]) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int k = lane + 32 * r;
        const int x = k < L ? seg[k] : kIdxInf;
        v[r] = x < cut ? x : kIdxInf;
    }
    warp_bitonic_sort<R>(v, lane);
    int n = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) n += __popc(__ballot_sync(0xffffffffu, v[r] != kIdxInf));
    return n;
}
// same for values already in registers
template <int R>
__device__ __forceinline__ int sort_bucket_regs(int cut, int lane, int (&v)[R]) {
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = v[r] < cut ? v[r] : kIdxInf;
    warp_bitonic_sort<R>(v, lane);
    int n = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) n += __popc(__ballot_sync(0xffffffffu, v[r] != kIdxInf));
    return n;
}

template <typename T, typename TO, int DS, bool BULK>
__global__ void __launch_bounds__(kGatherWarps * 32, PP_GATHER_MINBLOCKS)
vox_gather_kernel(const T* __restrict__ points, const int64_t* __restrict__ frame_off, VoxParams p, int b0, int nb,
                  const int4* __restrict__ occ_desc, const int* __restrict__ occ_base,
                  const int* __restrict__ bucket, const int* __restrict__ cutoff,
                  const int* __restrict__ voxel_base, TO* __restrict__ voxels, float* __restrict__ decorated,
                  int* __restrict__ num_points, int* __restrict__ point_slot) {
    extern __shared__ __align__(16) unsigned char gsm_raw[];
    const float* s_zero = nullptr;
    if constexpr (BULK) {
        __shared__ __align__(16) float s_zero_buf[kZeroFloats];
        for (int k = threadIdx.x; k < kZeroFloats; k += kGatherWarps * 32) s_zero_buf[k] = 0.f;
        fence_proxy_async_smem();
        __syncthreads();
        s_zero = s_zero_buf;
    }
    const int P = p.max_points;
    const int D = DS ? DS : p.D;
    const int Do = D + 5;
    const int lane = lane_id(), w = threadIdx.x >> 5;
    // per-warp carve: ord[P] (int) | vrow[P*D] | drow[P*Do] (float, only when DS != 3); every region
    // a multiple of 16 bytes
    const int nord = (P + 3) & ~3, nvox = (P * D + 3) & ~3, ndec = DS == 3 ? 0 : ((P * Do + 3) & ~3);
    int* ord = reinterpret_cast<int*>(gsm_raw + (size_t)w * (nord + nvox + ndec) * 4);
    float* vrow = reinterpret_cast<float*>(ord + nord);
    float* dsm = vrow + nvox;

    // Occupied cells are listed frame by frame as 16-byte descriptors {cell, row, count, bucket offset}.
    // Two-deep software pipeline per warp: while cell e is processed, the descriptor of cell e+2*stride
    // and the first 64 bucket entries (+ frame constants) of cell e+stride are already in flight, so a
    // cell costs one exposed memory round trip (the point gather) instead of three.
    const int nocc = occ_base[nb];
    const int nwarps = gridDim.x * kGatherWarps;
    auto load_desc = [&](int ee, int& bl) -> int4 {
        while (ee >= occ_base[bl + 1]) ++bl;
        return occ_desc[(size_t)(b0 + bl) * p.ncell + (ee - occ_base[bl])];
    };
    auto load_head = [&](const int4& d, int bl, int& v0, int& v1, int& cut, int64_t& f0) {
        const int* seg = bucket + d.w;
        v0 = (d.y >= 0 && lane < d.z) ? seg[lane] : kIdxInf;
        v1 = (d.y >= 0 && lane + 32 < d.z) ? seg[lane + 32] : kIdxInf;
        cut = cutoff[b0 + bl];
        f0 = frame_off[b0 + bl];
    };
    int e = blockIdx.x * kGatherWarps + w;
    int4 d0 = make_int4(0, -1, 0, 0), d1 = d0;
    int bl0 = 0, bl1 = 0, bl2 = 0;
    int hv0 = kIdxInf, hv1 = kIdxInf, hcut = 0;
    int64_t hf0 = 0;
    if (e < nocc) d0 = load_desc(e, bl0);
    bl1 = bl0;
    if (e + nwarps < nocc) d1 = load_desc(e + nwarps, bl1);
    bl2 = bl1;
    if (e < nocc) load_head(d0, bl0, hv0, hv1, hcut, hf0);

    for (; e < nocc; e += nwarps) {
        int4 d2 = d1;
        if (e + 2 * nwarps < nocc) d2 = load_desc(e + 2 * nwarps, bl2);
        int nv0 = kIdxInf, nv1 = kIdxInf, ncut = 0;
        int64_t nf0 = 0;
        if (e + nwarps < nocc) load_head(d1, bl1, nv0, nv1, ncut, nf0);

      if (d0.y >= 0) {
        const int cell = d0.x, L = d0.z, b = b0 + bl0;
        const int64_t row = d0.y;
        const int64_t f0 = hf0;
        const int* seg = bucket + d0.w;
        const int cut = hcut;
        int nsel;

        // ---- ord[s] = point index of slot s
        if (L <= 32) {
            int v[1] = {hv0};
            nsel = min(sort_bucket_regs<1>(cut, lane, v), P);
            if (lane < nsel) ord[lane] = v[0];
        } else if (L <= 64) {
            int v[2] = {hv0, hv1};
            nsel = min(sort_bucket_regs<2>(cut, lane, v), P);
#pragma unroll
            for (int r = 0; r < 2; ++r) if (r * 32 + lane < nsel) ord[r * 32 + lane] = v[r];
        } else if (L <= 128) {
            int v[4];
            nsel = min(load_sort_bucket<4>(seg, L, cut, lane, v), P);
#pragma unroll
            for (int r = 0; r < 4; ++r) if (r * 32 + lane < nsel) ord[r * 32 + lane] = v[r];
        } else if (L <= 32 * kSegRegs) {
            int v[kSegRegs];
            nsel = min(load_sort_bucket<kSegRegs>(seg, L, cut, lane, v), P);
#pragma unroll
            for (int r = 0; r < kSegRegs; ++r) if (r * 32 + lane < nsel) ord[r * 32 + lane] = v[r];
        } else {
            // ---- long bucket: threshold search streaming the bucket, then rank by counting.
            // vrow doubles as the unordered selection buffer (P ints <= P*D floats).
            int* sel = reinterpret_cast<int*>(vrow);
            int Lc = 0;
            for (int k = lane; k < L; k += 32) Lc += seg[k] < cut;
#pragma unroll
            for (int o = 16; o; o >>= 1) Lc += __shfl_xor_sync(0xffffffffu, Lc, o);
            int thr = cut;
            if (Lc > P) {
                // smallest thr with #{idx < thr} >= P (distinct indices => exactly P)
                int lo = 0, hi = cut;
                while (lo < hi) {
                    const int mid = lo + ((hi - lo) >> 1);
                    int g = 0;
                    for (int k = lane; k < L; k += 32) g += seg[k] < mid;
#pragma unroll
                    for (int o = 16; o; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
                    if (g >= P) hi = mid; else lo = mid + 1;
                }
                thr = lo;
            }
            nsel = min(Lc, P);
            int nb_ = 0;
            for (int k0 = 0; k0 < L; k0 += 32) {
                const int k = k0 + lane;
                const int x = k < L ? seg[k] : kIdxInf;
                const bool pr = x < thr;
                const unsigned bal = __ballot_sync(0xffffffffu, pr);
                if (pr) sel[nb_ + __popc(bal & lanemask_lt())] = x;
                nb_ += __popc(bal);
            }
            __syncwarp();
            for (int j = lane; j < nsel; j += 32) {
                const int x = sel[j];
                int r = 0;
                for (int q = 0; q < nsel; ++q) r += sel[q] < x;
                ord[r] = x;
            }
        }
        __syncwarp();

        const int rem = cell - p.div_nxny.div(cell) * p.grid[0] * p.grid[1];
        const int cy = p.div_nx.div(rem), cx = rem - cy * p.grid[0];
        if (lane == 0) num_points[row] = nsel;
        const int rank = point_slot ? (int)(row - voxel_base[b]) : 0;
        const T* fp = points + f0 * D;
        // ---- gather each selected point once; zero the padding of the row
        float sx = 0.f, sy = 0.f, sz = 0.f;
        if (P <= 64 && DS != 0) {
            // both rounds of point loads are issued before either is consumed (one exposed round trip)
            T raw[2][DS ? DS : 1];
            int pi[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int s = r * 32 + lane;
                pi[r] = s < nsel ? ord[s] : 0;
                const T* q = fp + (int64_t)pi[r] * D;
#pragma unroll
                for (int d = 0; d < (DS ? DS : 1); ++d) raw[r][d] = s < nsel ? q[d] : (T)0;
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int s = r * 32 + lane;
                if (s < nsel) {
                    if (point_slot) point_slot[f0 + pi[r]] = rank * P + s;
                    if (sizeof(TO) == 8) {
                        TO* vo = voxels + (row * (int64_t)P + s) * D;
#pragma unroll
                        for (int d = 0; d < (DS ? DS : 1); ++d) vo[d] = (TO)raw[r][d];
                    }
                    float c[DS ? DS : 1];
#pragma unroll
                    for (int d = 0; d < (DS ? DS : 1); ++d) c[d] = (float)raw[r][d];
#pragma unroll
                    for (int d = 0; d < (DS ? DS : 1); ++d) vrow[s * D + d] = c[d];
                    sx += c[0]; sy += c[1 % (DS ? DS : 1)]; sz += c[2 % (DS ? DS : 1)];
                }
            }
        } else
        for (int s = lane; s < nsel; s += 32) {
            const int pi = ord[s];
            const T* q = fp + (int64_t)pi * D;
            if (point_slot) point_slot[f0 + pi] = rank * P + s;
            if (sizeof(TO) == 8) {
                TO* vo = voxels + (row * (int64_t)P + s) * D;
#pragma unroll
                for (int d = 0; d < (DS ? DS : 16); ++d) if (d < D) vo[d] = (TO)q[d];
            }
            float c[DS ? DS : 16];
#pragma unroll
            for (int d = 0; d < (DS ? DS : 16); ++d) if (d < D) c[d] = (float)q[d];
#pragma unroll
            for (int d = 0; d < (DS ? DS : 16); ++d) if (d < D) vrow[s * D + d] = c[d];
            sx += c[0]; sy += c[1]; sz += c[2];
        }
        if (sizeof(TO) == 8) {
            TO* vo = voxels + row * (int64_t)P * D;
            for (int k = nsel * D + lane; k < P * D; k += 32) vo[k] = (TO)0;
        }
        __syncwarp();
        if constexpr (BULK) {
            if (voxels && sizeof(TO) == 4) {  // data floats from the staged row, padding as bulk zero stores
                float* vr = reinterpret_cast<float*>(voxels) + row * (int64_t)P * D;
                const int nd = nsel * D, nd4 = min((nd + 3) & ~3, P * D);  // data rounded up to whole float4
                warp_store_row_padded(vr, vrow, nd4, nd, lane);
                warp_zero_fill(vr + nd4, vr + P * D, lane, s_zero);
            }
        } else
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
                if constexpr (BULK) {
                    for (int s = lane; s < nsel; s += 32) {
                        const float q0 = vrow[s * 3], q1 = vrow[s * 3 + 1], q2 = vrow[s * 3 + 2];
                        d4[2 * s] = make_float4(q0, q1, q2, q0 - mx);
                        d4[2 * s + 1] = make_float4(q1 - my, q2 - mz, q0 - ex, q1 - ey);
                    }
                    warp_zero_fill(drow + nsel * 8, drow + P * 8, lane, s_zero);
                } else
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
                if constexpr (BULK) {
                    const int nd = nsel * Do, nd4 = min((nd + 3) & ~3, P * Do);
                    for (int s = lane; s < nsel; s += 32) {
                        float* o = dsm + s * Do;
                        const float* q = vrow + s * D;
#pragma unroll
                        for (int d = 0; d < (DS ? DS : 16); ++d) if (d < D) o[d] = q[d];
                        o[D] = q[0] - mx; o[D + 1] = q[1] - my; o[D + 2] = q[2] - mz;
                        o[D + 3] = q[0] - ex; o[D + 4] = q[1] - ey;
                    }
                    if (lane < nd4 - nd) dsm[nd + lane] = 0.f;  // round the data up to whole float4
                    __syncwarp();
                    warp_store_row(drow, dsm, nd4, lane);
                    warp_zero_fill(drow + nd4, drow + P * Do, lane, s_zero);
                } else {
                for (int s = lane; s < P; s += 32) {
                    float* o = dsm + s * Do;
                    if (s < nsel) {
                        const float* q = vrow + s * D;
#pragma unroll
                        for (int d = 0; d < (DS ? DS : 16); ++d) if (d < D) o[d] = q[d];
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
        }
        __syncwarp();
      }
        // rotate the pipeline
        d0 = d1; d1 = d2;
        bl0 = bl1; bl1 = bl2;
        hv0 = nv0; hv1 = nv1; hcut = ncut; hf0 = nf0;
    }
    if constexpr (BULK) {
        if (lane == 0) tma_bulk_wait_all();  // this lane's bulk stores have completed before the CTA retires
    }
}

// ---------------------------------------------------------------------------------------------
struct VoxWorkspace {
    unsigned* first_idx;  // [B*ncell]  0xff init
    int* cnt;             // [B*ncell]  zero init   -- zero region starts here
    unsigned* bitmap;     // [nwords]
    int* frame_cursor;    // [B]
    int* frame_occ;       // [B] occupied cells per frame
    int* done_counter;    // [1]        -- zero region ends here
    unsigned* word_prefix;  // [nwords]
    int* cell_off;        // [B*ncell]
    int4* occ_desc;       // [B*ncell] descriptors of occupied cells, frame-major
    int* cutoff;          // [B]
    int* occ_base;        // [B+1] exclusive scan of frame_occ
    int2* cellpos;        // [total_points]
    int* bucket;          // [total_points]
    size_t zero_begin, zero_end, total;
};

static VoxWorkspace carve(void* ws, int64_t ncell, int64_t total_points, int B) {
    VoxWorkspace w;
    Carver c(ws);
    const size_t nc = (size_t)B * ncell;
    const size_t nwords = (size_t)(total_points >> 5) + B + 2;
    w.first_idx = c.take<unsigned>(nc);
    w.zero_begin = c.used();
    w.cnt = c.take<int>(nc);
    w.bitmap = c.take<unsigned>(nwords);
    w.frame_cursor = c.take<int>(B);
    w.frame_occ = c.take<int>(B);
    w.done_counter = c.take<int>(1);
    w.zero_end = c.used();
    w.word_prefix = c.take<unsigned>(nwords);
    w.cell_off = c.take<int>(nc);
    w.occ_desc = c.take<int4>(nc + 1);
    w.cutoff = c.take<int>(B);
    w.occ_base = c.take<int>((size_t)B + 2);
    w.cellpos = c.take<int2>((size_t)total_points + 1);
    w.bucket = c.take<int>((size_t)total_points + 1);
    w.total = c.used();
    return w;
}

static int64_t ncell_of(const pp_voxel_cfg* cfg, int32_t grid[3]) {
    pp_grid_size(cfg->voxel_size, cfg->coors_range, cfg->arith_f32, grid);
    return (int64_t)grid[0] * grid[1] * grid[2];
}

}  // namespace pp

using namespace pp;

extern "C" int pp_grid_size(const double voxel_size[3], const double coors_range[6], int arith_f32,
                            int32_t grid_xyz[3]) {
    for (int j = 0; j < 3; ++j) {
        if (arith_f32) {
            const float g = ((float)coors_range[3 + j] - (float)coors_range[j]) / (float)voxel_size[j];
            grid_xyz[j] = (int32_t)nearbyintf(g);
        } else {
            const double g = (coors_range[3 + j] - coors_range[j]) / voxel_size[j];
            grid_xyz[j] = (int32_t)nearbyint(g);
        }
    }
    return PP_OK;
}

extern "C" size_t pp_voxelize_workspace_bytes(const pp_voxel_cfg* cfg, int64_t total_points, int n_frames,
                                              int64_t max_frame_points, int D, int out_dtype) {
    if (!cfg || total_points < 0 || n_frames <= 0 || max_frame_points < 0 || D < 3) return 0;
    int32_t grid[3];
    const int64_t ncell = ncell_of(cfg, grid);
    if (ncell <= 0) return 0;
    if (max_frame_points > total_points) max_frame_points = total_points;
    // both paths are bit-identical; the call takes the shared-memory path when the workspace allows it, so size for both
    size_t need = carve(nullptr, ncell, total_points, n_frames).total;
    if (vox_small_eligible(cfg, ncell, n_frames, total_points, max_frame_points, D)) {
        const size_t small = vox_small_workspace_bytes(cfg, ncell, n_frames, total_points, max_frame_points, D, out_dtype);
        if (small > need) need = small;
    }
    return need + 256;
}

template <typename T, typename TO, int DS>
static int launch_gather(const VoxParams& p, const VoxWorkspace& w, const void* points,
                         const int64_t* frame_off, void* voxels, float* decorated, int32_t* num_points,
                         const int32_t* voxel_base, int32_t* point_slot, int nb, int64_t max_occ, cudaStream_t st) {
    const int P_ = p.max_points, D_ = p.D;
    const size_t per_warp = (size_t)(((P_ + 3) & ~3) + ((P_ * D_ + 3) & ~3) + (DS == 3 ? 0 : ((P_ * (D_ + 5) + 3) & ~3))) * 4;
    const size_t smem = (size_t)kGatherWarps * per_warp;
    PP_CHECK_ARG(smem <= 200 * 1024, "pp_voxelize_dev: max_points * D too large for the gather pass");
    // long float32 rows: the instantiation that writes the padding with TMA bulk stores
    const bool bulk = sizeof(TO) == 4 && (size_t)P_ * (D_ + 5) * 4 >= (size_t)kBulkMinRowBytes;
    auto kern = bulk ? vox_gather_kernel<T, TO, DS, true> : vox_gather_kernel<T, TO, DS, false>;
    int per_sm = 0;
    PP_TRY_RC(kernel_config(reinterpret_cast<const void*>(kern), kGatherWarps * 32, smem, &per_sm));
    int64_t blocks = ceil_div(max_occ, kGatherWarps);
    const int64_t cap = (int64_t)num_sms() * (per_sm > 0 ? per_sm : 1);  // persistent: one resident wave
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    PP_TIMED("vox_gather", st);
    kern<<<(unsigned)blocks, kGatherWarps * 32, smem, st>>>(
        static_cast<const T*>(points), frame_off, p, 0, nb, w.occ_desc, w.occ_base, w.bucket, w.cutoff, voxel_base,
        static_cast<TO*>(voxels), decorated, num_points, point_slot);
    PP_LAUNCHED();
    return PP_OK;
}

extern "C" int pp_voxelize_dev(const pp_voxel_cfg* cfg, const void* points, int point_dtype, int D,
                               const int64_t* frame_offsets, int n_frames, int64_t total_points,
                               int64_t max_frame_points, int out_dtype, void* voxels,
                               float* decorated, int32_t* coors, int coors_cols,
                               int32_t* num_points, int64_t cap_rows, int32_t* voxel_num,
                               int32_t* voxel_base, int32_t* point_slot, int32_t* cell_voxel,
                               void* workspace, size_t workspace_bytes, void* stream) {
    PP_CHECK_ARG(cfg && frame_offsets && coors && num_points && voxel_num && voxel_base && workspace,
                 "pp_voxelize_dev: null argument");
    PP_CHECK_ARG(point_dtype == PP_F32 || point_dtype == PP_F64, "point_dtype must be PP_F32/PP_F64");
    PP_CHECK_ARG(out_dtype == PP_F32 || (out_dtype == PP_F64 && point_dtype == PP_F64),
                 "out_dtype must be PP_F32, or PP_F64 for float64 points");
    PP_CHECK_ARG(D >= 3 && D <= 16, "D=%d outside [3,16]", D);
    PP_CHECK_ARG(coors_cols == 3 || coors_cols == 4, "coors_cols must be 3 or 4");
    PP_CHECK_ARG(n_frames > 0 && n_frames <= 65535, "n_frames=%d outside [1,65535]", n_frames);
    PP_CHECK_ARG(total_points >= 0 && max_frame_points >= 0 && max_frame_points <= total_points,
                 "bad point counts");
    PP_CHECK_ARG(max_frame_points < (int64_t)1 << 31, "a frame may hold < 2^31 points");
    PP_CHECK_ARG(cfg->max_points >= 1 && cfg->max_points <= 2048, "max_points=%d outside [1,2048]", cfg->max_points);
    PP_CHECK_ARG(cfg->max_voxels >= 0, "max_voxels < 0");
    PP_CHECK_ARG(!(cfg->arith_f32 && point_dtype == PP_F64), "arith_f32 needs float32 points");
    PP_CHECK_ARG(total_points == 0 || points, "points is null");
    PP_CHECK_ARG(voxels || decorated, "pp_voxelize_dev: voxels and decorated are both null");
    PP_CHECK_ARG(!(out_dtype == PP_F64 && !voxels), "pp_voxelize_dev: float64 output needs voxels");
    int32_t grid[3];
    const int64_t ncell = ncell_of(cfg, grid);
    PP_CHECK_ARG(grid[0] > 0 && grid[1] > 0 && grid[2] > 0, "empty grid %d x %d x %d", grid[0], grid[1], grid[2]);
    PP_CHECK_ARG(ncell * n_frames < ((int64_t)1 << 31), "n_frames * cells must be < 2^31");
    PP_CHECK_ARG(total_points < ((int64_t)1 << 31), "a batch may hold < 2^31 points (bucket offsets are int32)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    VoxParams p;
    for (int j = 0; j < 3; ++j) {
        p.lo[j] = cfg->coors_range[j]; p.vs[j] = cfg->voxel_size[j];
        p.lo32[j] = (float)cfg->coors_range[j]; p.vs32[j] = (float)cfg->voxel_size[j];
        p.inv[j] = 1.0 / p.vs[j]; p.inv32[j] = 1.0f / p.vs32[j];
        p.grid[j] = grid[j];
    }
    p.div_nx = FastDiv((unsigned)grid[0]); p.div_nxny = FastDiv((unsigned)(grid[0] * grid[1]));
    p.ncell = (int)ncell; p.max_points = cfg->max_points; p.max_voxels = cfg->max_voxels;
    p.reverse_index = cfg->reverse_index; p.arith_f32 = cfg->arith_f32; p.D = D;
    p.vx = (float)cfg->voxel_size[0]; p.vy = (float)cfg->voxel_size[1];
    p.x_off = (float)(cfg->voxel_size[0] / 2 + cfg->coors_range[0]);
    p.y_off = (float)(cfg->voxel_size[1] / 2 + cfg->coors_range[1]);

    const int esz = point_dtype == PP_F64 ? 8 : 4;
    // Grids whose per-cell tables fit in shared memory (the d435i grid: 10 240 cells) take the
    // chunk-privatised path of voxelize_small.cu: no global atomics, no bucket pass, no sort.
    if (vox_small_eligible(cfg, ncell, n_frames, total_points, max_frame_points, D) &&
        vox_small_workspace_bytes(cfg, ncell, n_frames, total_points, max_frame_points, D, out_dtype) <= workspace_bytes)
        return vox_small_run(cfg, p, points, point_dtype, frame_offsets, n_frames, total_points, max_frame_points,
                             out_dtype, voxels, decorated, coors, coors_cols, num_points, cap_rows, voxel_num,
                             voxel_base, point_slot, cell_voxel, workspace, workspace_bytes, st);

    const VoxWorkspace w = carve(workspace, ncell, total_points, n_frames);
    if (w.total > workspace_bytes) {
        set_error("pp_voxelize_dev: workspace %zu < required %zu", workspace_bytes, w.total);
        return PP_E_WORKSPACE;
    }
    const size_t nc = (size_t)n_frames * ncell;
    {
        PP_TIMED("vox_memset", st);
        PP_CUDA(cudaMemsetAsync(w.first_idx, 0xff, nc * sizeof(unsigned), st));
        PP_CUDA(cudaMemsetAsync(static_cast<char*>(workspace) + w.zero_begin, 0, w.zero_end - w.zero_begin, st));
    }
    // (Processing the frames in L2-sized chunks was measured and rejected: per-chunk launch tails cost more
    //  than the L2 hits save, profiles/r01_notes.md.)
    const int nb = n_frames;
    const size_t mark_smem = 2 * ((size_t)kMarkTile * D * esz + 32);  // two pipeline stages
    const int aligned16 = (reinterpret_cast<uintptr_t>(points) & 15) == 0;
    PP_CHECK_ARG(mark_smem <= 200 * 1024, "pp_voxelize_dev: D too large for the mark pass");
    if (max_frame_points > 0) {
        const void* kern = point_dtype == PP_F64 ? reinterpret_cast<const void*>(vox_mark_kernel<double, false>)
                           : cfg->arith_f32     ? reinterpret_cast<const void*>(vox_mark_kernel<float, true>)
                                                : reinterpret_cast<const void*>(vox_mark_kernel<float, false>);
        int mark_ctas_per_sm = 0;
        PP_TRY_RC(kernel_config(kern, kMarkThreads, mark_smem, &mark_ctas_per_sm));
        if (mark_ctas_per_sm > 8) mark_ctas_per_sm = 8;
        if (mark_ctas_per_sm < 1) mark_ctas_per_sm = 1;
        const int tpf = (int)ceil_div(max_frame_points, kMarkTile);
        const int64_t tiles = (int64_t)tpf * nb;
        int64_t g = (int64_t)num_sms() * mark_ctas_per_sm;
        if (g > tiles) g = tiles;
        PP_TIMED("vox_mark", st);
        if (point_dtype == PP_F64)
            vox_mark_kernel<double, false><<<(unsigned)g, kMarkThreads, mark_smem, st>>>(
                static_cast<const double*>(points), frame_offsets, p, total_points, aligned16, 0, nb, tpf,
                w.first_idx, w.cnt, w.cellpos, point_slot);
        else if (cfg->arith_f32)
            vox_mark_kernel<float, true><<<(unsigned)g, kMarkThreads, mark_smem, st>>>(
                static_cast<const float*>(points), frame_offsets, p, total_points, aligned16, 0, nb, tpf,
                w.first_idx, w.cnt, w.cellpos, point_slot);
        else
            vox_mark_kernel<float, false><<<(unsigned)g, kMarkThreads, mark_smem, st>>>(
                static_cast<const float*>(points), frame_offsets, p, total_points, aligned16, 0, nb, tpf,
                w.first_idx, w.cnt, w.cellpos, point_slot);
        PP_LAUNCHED();
    }
    {
        const dim3 g((unsigned)ceil_div(ncell, kCellThreads * kCellPerThread), nb);
        PP_TIMED("vox_cell", st);
        vox_cell_kernel<<<g, kCellThreads, 0, st>>>(w.first_idx, w.cnt, frame_offsets, (int)ncell, 0,
                                                    w.bitmap, w.cell_off, w.frame_cursor, w.occ_desc,
                                                    w.frame_occ, cell_voxel);
        PP_LAUNCHED();
    }
    {
        PP_TIMED("vox_rank", st);
        vox_rank_kernel<<<nb, kRankThreads, 0, st>>>(w.bitmap, w.word_prefix, frame_offsets, 0, nb,
                                                     cfg->max_voxels, voxel_num, w.cutoff, voxel_base,
                                                     w.frame_occ, w.occ_base, w.done_counter);
        PP_LAUNCHED();
    }
    const int64_t chunk_pts = (int64_t)nb * max_frame_points;
    const int64_t frame_occ_max = ncell < max_frame_points ? ncell : max_frame_points;
    if (frame_occ_max > 0) {
        const dim3 g((unsigned)ceil_div(frame_occ_max, 256), nb);
        PP_TIMED("vox_rowmap", st);
        vox_rowmap_kernel<<<g, 256, 0, st>>>(w.occ_desc, w.occ_base, frame_offsets, p, 0, w.bitmap, w.word_prefix,
                                             voxel_base, cap_rows, coors, coors_cols, cell_voxel);
        PP_LAUNCHED();
    }
    if (max_frame_points > 0) {
        // (a variant that staged the frame's offset table in shared memory was measured slower: 185 vs 113 us)
        const dim3 g((unsigned)ceil_div(max_frame_points, 256 * kBucketPPT), nb);
        PP_TIMED("vox_bucket", st);
        vox_bucket_kernel<<<g, 256, 0, st>>>(w.cellpos, frame_offsets, (int)ncell, 0, w.cell_off, w.bucket);
        PP_LAUNCHED();
    }
    const int64_t max_occ = (int64_t)nb * ncell < chunk_pts ? (int64_t)nb * ncell : chunk_pts;
    if (max_occ > 0 && cap_rows > 0) {
        int rc;
#define PP_GATHER(T, TO, DS)                                                                                   \
    launch_gather<T, TO, DS>(p, w, points, frame_offsets, voxels, decorated, num_points, voxel_base,          \
                             point_slot, nb, max_occ, st)
        if (point_dtype == PP_F64 && out_dtype == PP_F64) rc = PP_GATHER(double, double, 0);
        else if (point_dtype == PP_F64) rc = D == 3 ? PP_GATHER(double, float, 3) : D == 4 ? PP_GATHER(double, float, 4) : PP_GATHER(double, float, 0);
        else rc = D == 3 ? PP_GATHER(float, float, 3) : D == 4 ? PP_GATHER(float, float, 4) : PP_GATHER(float, float, 0);
#undef PP_GATHER
        if (rc) return rc;
    }
    return PP_OK;
}

extern "C" int pp_voxelize_set_small_path_min_points(int64_t n) {
    pp::vox_small_set_min_points(n);  // n < 0: back to the default
    return PP_OK;
}
