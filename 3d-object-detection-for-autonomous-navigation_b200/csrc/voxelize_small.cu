// Deterministic first-come pillarization for grids whose per-cell tables fit in shared memory
// (the d435i grid of configs/train.yaml: 80 x 64 x 2 = 10 240 cells).  Same results, bit for bit, as the
// reference's sequential loop (load_data.py:593-692) and as the any-grid path in voxelize.cu, but built
// as a two-level counting sort: no pass issues a global atomic per point, none sorts, none ranks a bitmap.
//
//   scan    (frame, chunk of 16 384 points)  rows staged by TMA bulk copies, four points per thread; cell id in
//           the reference's arithmetic; in-range points are compacted IN INDEX ORDER into 16/32-byte records
//           (coordinates in the output type) and 4-byte tags {cell | index in chunk}; the chunk's per-cell counts
//           live in SHARED memory (ATOMS) and are written out once per chunk.
//   prefix  (frame, 256 cells)  per cell: exclusive prefix of the chunk counts = the slot base of every chunk
//           (uint8, saturated: a base >= max_points means "full"); the chunk in which a cell first appears is
//           counted, which gives every chunk the number of voxels opened before it.
//   place   (frame, chunk)  ONE WARP walks the chunk's record tags in order, 32 per step, with the chunk's running
//           per-cell counts in shared memory: slot = count[cell] + rank among the step's earlier records of the
//           cell (match_any).  A record that finds count 0 opens its cell: because the walk is in index order,
//           the running number of such records IS the voxel id (order of first touch), and the record that
//           would open voxel number max_voxels is the reference's `break` position (load_data.py:630-634).
//           The record's position goes to slot (cell, slot) of a 4-byte index table (2 MB per d435i frame: it
//           stays in L2, where scattering the 16-byte records themselves over 16 MB per frame was measured
//           DRAM-random-write bound: 384 us against 158 us without the stores); no barrier, no atomic.
//   finish  (pillar)  32 pillars per CTA: the records a pillar's slots point at (before the break position) ->
//           zero-padded voxel row, num_points, coors, point->slot map and the fused PillarFeatureNet decoration
//           (model/pointpillars.py:143-203), streamed out with 16/32-byte stores.
#include "pp_common.cuh"
#include "vox_common.cuh"
#include "vox_internal.h"

namespace pp {

#ifndef PP_CHUNK_SHIFT
#define PP_CHUNK_SHIFT 14
#endif
constexpr int kChunkShift = PP_CHUNK_SHIFT;  // points per chunk = 2^shift; positions inside a chunk fit 16 bits
constexpr int kChunk = 1 << kChunkShift;
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
constexpr int kNoCut = 0x7fffffff;

static_assert(kChunk % kScanTile == 0 && kChunk <= 65536, "chunk size");

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

// record: the point's coordinates in the output type, padded to 16-byte units
template <typename TO, int DS> struct RecFmt {
    static constexpr int kBytes = (int)(((DS * sizeof(TO) + 15) / 16) * 16);
    static constexpr int kRec16 = kBytes / 16;
};

// ---------------------------------------------------------------------------------------------
// Pass 1.  Warp w of a tile owns its points [128 w, 128 w + 128); round r of lane l is point 128 w + 32 r + l,
// so (warp, round, lane) enumerates the tile in index order and shared-memory row reads are conflict free.
// kScanStages shared-memory stages: with two, the TMA bulk copy of tile j+1 is in flight while tile j is processed.
#ifndef PP_SCAN_STAGES
#define PP_SCAN_STAGES 2
#endif
constexpr int kScanStages = PP_SCAN_STAGES;

template <typename T, bool A32, bool FAST, typename TO, int DS>
__global__ void __launch_bounds__(kScanThreads)
vox_scan_kernel(const T* __restrict__ points, const int64_t* __restrict__ frame_off, VoxParams p, int64_t total_points,
                int aligned16, int S, int ncellp, uint4* __restrict__ crec, unsigned* __restrict__ ctag,
                unsigned* __restrict__ hist_out, int* __restrict__ nvalid, int* __restrict__ newcount,
                int* __restrict__ cutoff, int* __restrict__ point_slot, int* __restrict__ done_counter,
                int* __restrict__ frame_done) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) unsigned long long s_bar[kScanStages];
    __shared__ int s_wtot[kScanWarps];
    constexpr int row_bytes = DS * (int)sizeof(T);
    constexpr int stage_bytes = kScanTile * row_bytes + 32;  // multiple of 16
    constexpr int rec16 = RecFmt<TO, DS>::kRec16;
    unsigned* hist32 = reinterpret_cast<unsigned*>(smem + kScanStages * stage_bytes);  // packed uint16 counts
    const int tid = threadIdx.x, lane = lane_id(), w = tid >> 5;
    const int b = blockIdx.y, s = blockIdx.x;
    if (tid == 0) {
        newcount[b * S + s] = 0;
        if (s == 0) { cutoff[b] = kNoCut; frame_done[b] = 0; }
        if (b == 0 && s == 0) *done_counter = 0;
    }
    const int64_t f0 = frame_off[b];
    const int n = (int)(frame_off[b + 1] - f0);
    const int c0 = s * kChunk;
    if (c0 >= n) return;
    const int cn = min(kChunk, n - c0);
    for (int k = tid; k < (ncellp >> 1); k += kScanThreads) hist32[k] = 0u;
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < kScanStages; ++k) mbar_init(&s_bar[k], 1);
    }
    __syncthreads();

    const int64_t total_bytes = total_points * (int64_t)row_bytes;
    const int64_t tail0 = total_bytes & ~(int64_t)15;  // end of the buffer's last full 16-byte unit
    const unsigned char* src = reinterpret_cast<const unsigned char*>(points);
    const unsigned long long policy = l2_evict_first_policy();  // the cloud is read exactly once
    const int ntiles = (cn + kScanTile - 1) / kScanTile;
    const int64_t gbase = f0 + c0;
    // byte range [a0, a1) of tile j rounded out to 16-byte units (clipped to the buffer's last full unit)
    auto span = [&](int j, int64_t& a0, int64_t& a1, int64_t& end) {
        const int base = j * kScanTile;
        const int m = min(kScanTile, cn - base);
        const int64_t start = (gbase + base) * (int64_t)row_bytes;
        end = start + (int64_t)m * row_bytes;
        a0 = start & ~(int64_t)15;
        a1 = (end + 15) & ~(int64_t)15;
        if (a1 > tail0) a1 = tail0 > a0 ? tail0 : a0;
    };
    auto issue = [&](int j) {
        if (!aligned16 || tid != 0 || j >= ntiles) return;
        int64_t a0, a1, end;
        span(j, a0, a1, end);
        const unsigned bulk = (unsigned)(a1 - a0);
        if (bulk) {
            const int st = j % kScanStages;
            mbar_arrive_expect_tx(&s_bar[st], bulk);
            tma_bulk_g2s_hint(smem + (size_t)st * stage_bytes, src + a0, bulk, &s_bar[st], policy);
        }
    };
    unsigned phase[kScanStages];
#pragma unroll
    for (int k = 0; k < kScanStages; ++k) phase[k] = 0;
    int crun = 0;  // in-range points of the chunk so far (uniform over the CTA)
#pragma unroll
    for (int k = 0; k < kScanStages - 1; ++k) issue(k);

    for (int j = 0; j < ntiles; ++j) {
        const int stg = j % kScanStages;
        issue(j + kScanStages - 1);  // its stage was released by the barrier that ended tile j-1
        unsigned char* buf = smem + (size_t)stg * stage_bytes;
        const int base = j * kScanTile;  // first point of the tile, relative to the chunk
        const int m = min(kScanTile, cn - base);
        int64_t a0, a1, end;
        span(j, a0, a1, end);
        int shift = 0;
        if (aligned16) {
            shift = (int)((gbase + base) * (int64_t)row_bytes - a0);
            if (a1 > a0) {
#pragma unroll
                for (int k = 0; k < kScanStages; ++k)
                    if (k == stg) { mbar_wait(&s_bar[k], phase[k]); phase[k] ^= 1; }
            }
            if (a1 < end) {
                // bytes past the buffer's last full unit (only the very last tile of the buffer)
                for (int64_t a = a1 + (int64_t)tid * (int)sizeof(T); a < end && a < total_bytes;
                     a += (int64_t)kScanThreads * (int)sizeof(T))
                    *reinterpret_cast<T*>(buf + (a - a0)) = *reinterpret_cast<const T*>(src + a);
                __syncthreads();
            }
        } else {
            const int nel = m * DS;
            for (int k = tid; k < nel; k += kScanThreads)
                reinterpret_cast<T*>(buf)[k] = points[(gbase + base) * DS + k];
            __syncthreads();
        }
        const unsigned char* rows = buf + shift + (size_t)(w * 128 + lane) * row_bytes;

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
            // one shared-memory atomic per in-range point (lanes that share a cell are serialised by the bank,
            // 32 cycles at worst; match_any aggregation costs ~11 cycles per distinct cell, 350 on scattered clouds)
            if (cell[r] >= 0) atomicAdd(&hist32[cell[r] >> 1], 1u << ((cell[r] & 1) * 16));
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
            if (cell[r] >= 0) {
                const T* q = reinterpret_cast<const T*>(rows + (size_t)r * 32 * row_bytes);
                const int64_t gi = gbase + pos + __popc(bal[r] & lanemask_lt());
                uint4* dst = crec + gi * rec16;
                if (sizeof(TO) == 4) {
                    *reinterpret_cast<float4*>(dst) = make_float4((float)q[0], (float)q[1], (float)q[2], DS == 4 ? (float)q[DS - 1] : 0.f);
                } else {
                    reinterpret_cast<double2*>(dst)[0] = make_double2((double)q[0], (double)q[1]);
                    reinterpret_cast<double2*>(dst)[1] = make_double2((double)q[2], DS == 4 ? (double)q[DS - 1] : 0.0);
                }
                ctag[gi] = (unsigned)cell[r] | ((unsigned)(base + t) << 16);
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
vox_prefix_kernel(const int64_t* __restrict__ frame_off, int S, int ncell, int ncellp, int P,
                  const unsigned short* __restrict__ hist, unsigned char* __restrict__ base8,
                  unsigned* __restrict__ cellinfo, int* __restrict__ newcount, int max_voxels, int B,
                  int* __restrict__ voxel_num, int* __restrict__ voxel_base, int* __restrict__ done_counter,
                  int* __restrict__ frame_done) {
    __shared__ int s_new[kMaxChunks];
    __shared__ int sm[33];
    __shared__ int s_last, s_flast;
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
        cellinfo[(size_t)b * ncellp + cell] = (unsigned)min(run, P) << 24;  // slots of the cell; offset added below
    }
    __syncthreads();
    if (tid < Sb && s_new[tid]) atomicAdd(&newcount[b * S + tid], s_new[tid]);
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        s_flast = (atomicAdd(&frame_done[b], 1) == (int)gridDim.x - 1);
        s_last = (atomicAdd(done_counter, 1) == (int)(gridDim.x * gridDim.y) - 1);
    }
    __syncthreads();
    if (s_flast) {
        // last CTA of the frame: exclusive prefix of the cells' slot counts = where each cell's run of the dense
        // slot -> record table starts (cellinfo = offset | slots << 24)
        __threadfence();
        unsigned* ci = cellinfo + (size_t)b * ncellp;
        constexpr int kPer = kMaxCellsSmall / kPrefixThreads;  // cells per thread, all loaded before the scan
        const int per = (ncell + kPrefixThreads - 1) / kPrefixThreads;
        const int c0 = tid * per;
        unsigned v[kPer];
        int mine = 0;
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            v[k] = (k < per && c0 + k < ncell) ? __ldcg(&ci[c0 + k]) : 0u;
        }
#pragma unroll
        for (int k = 0; k < kPer; ++k) mine += (int)(v[k] >> 24);
        int tot;
        int ex = block_excl_scan(mine, &tot, sm);
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            if (k < per && c0 + k < ncell) ci[c0 + k] = v[k] | (unsigned)ex;
            ex += (int)(v[k] >> 24);
        }
    }
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
// ---------------------------------------------------------------------------------------------
// 32-byte accesses (one full L2 sector per request; SASS STG.E.ENL2.256 / LDG.E.ENL2.256 on sm_100)
__device__ __forceinline__ void st_global_256(void* ptr, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
                 "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}

// Pass 3: one warp per (frame, chunk).  tbl: running count per cell, uint8, starts at the chunk's base.
// The walk is sequential over steps of 32 records, but only the table update is a true dependence between steps:
// a group of kPlaceUnroll steps is processed in three stages -- (A) cell of every record and the lanes that share
// it (match_any: ~11 cycles per distinct value, 350 on scattered clouds, but independent across steps, so the
// group's matches overlap), (B) the chain count = tbl[cell]; tbl[cell] += n, (C) voxel ids of the cells that were
// opened, slots and the index stores -- and the tags of the next two groups are already in flight.
__global__ void __launch_bounds__(32)
vox_place_kernel(const int64_t* __restrict__ frame_off, int S, int ncell, int ncellp, int P, int max_voxels,
                 const unsigned* __restrict__ ctag, const unsigned char* __restrict__ base8,
                 const unsigned* __restrict__ cellinfo, const int* __restrict__ nvalid,
                 const int* __restrict__ newcount, unsigned* __restrict__ sidx, uint2* __restrict__ rowinfo,
                 int* __restrict__ cutoff) {
    extern __shared__ __align__(16) unsigned char tbl[];  // [ncellp]
    constexpr int U = kPlaceUnroll;
    const int lane = lane_id();
    // frames in reverse order: the scan pass wrote the last frames' tags last, so they are still in L2
    const int b = (int)gridDim.y - 1 - (int)blockIdx.y, s = blockIdx.x;
    const int64_t f0 = frame_off[b];
    const int n = (int)(frame_off[b + 1] - f0);
    const int Sb = (n + kChunk - 1) >> kChunkShift;
    if (s >= Sb) return;
    const int nval = nvalid[b * S + s];
    if (nval <= 0) return;
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
    const unsigned* tags = ctag + f0 + (int64_t)s * kChunk;
    const size_t cellrow0 = (size_t)b * ncell;
    const unsigned* ci_b = cellinfo + (size_t)b * ncellp;
    unsigned* sidx_b = sidx + cellrow0 * (size_t)P;  // the frame's dense slot -> record table

    struct Group { unsigned tg[U]; };
    // unconditional loads (index clamped to the last record): a predicated load makes the compiler merge the
    // loaded registers with their old contents right behind the load, which waits for it and defeats the prefetch
    auto fetch = [&](Group& gr, int g) {
#pragma unroll
        for (int u = 0; u < U; ++u) gr.tg[u] = __ldcs(&tags[min(g + u * 32 + lane, nval - 1)]);
    };
    auto process = [&](const Group& cur, int g) {
        if (g >= nval) return;  // uniform
        int c[U], rk[U], npeer[U], cnt[U];
        unsigned civ[U];
        // (A) independent of the table
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool act = g + u * 32 + lane < nval;
            c[u] = act ? (int)(cur.tg[u] & 0xffffu) : (0x10000 | lane);
            civ[u] = act ? __ldg(&ci_b[c[u]]) : 0u;  // start of the cell's slot run | slots << 24; needed in (C)
            const unsigned peers = __match_any_sync(0xffffffffu, c[u]);
            rk[u] = act ? __popc(peers & lanemask_lt()) : -1;  // rk 0: first record of its cell in the step
            npeer[u] = __popc(peers);
        }
        // (B) the sequential part
#pragma unroll
        for (int u = 0; u < U; ++u) {
            cnt[u] = rk[u] >= 0 ? (int)tbl[c[u]] : 1;
            if (rk[u] == 0) tbl[c[u]] = (unsigned char)min(cnt[u] + npeer[u], 255);
            __syncwarp();
        }
        // (C) a count of 0 means no earlier point of the frame fell in the cell: the record opens a voxel, and
        // since the walk is in index order the running number of opened cells is the voxel id
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool opener = rk[u] == 0 && cnt[u] == 0;
            const unsigned opens = __ballot_sync(0xffffffffu, opener);
            // position of the record in its frame's record array: same order as the point indices
            const unsigned cp = (unsigned)((s << kChunkShift) + g + u * 32 + lane);
            if (opener) {
                const int rank = newrun + __popc(opens & lanemask_lt());
                // the finish pass finds cell and point count of a voxel in one word
                if (rank < max_voxels) rowinfo[cellrow0 + rank] = make_uint2((unsigned)c[u], civ[u]);
                else if (rank == max_voxels) cutoff[b] = (int)cp;  // the reference's break position
            }
            newrun += __popc(opens);
            const int slot = cnt[u] + rk[u];
#ifndef PP_EXP_NOSTORE
            if (rk[u] >= 0 && slot < P) sidx_b[(civ[u] & 0xffffffu) + slot] = cp;
#endif
        }
    };
    // three register buffers used in rotation (no copies: a register move of a load that is still in flight
    // would wait for it), each fetched two groups before it is processed
    constexpr int GS = U * 32;
    Group ga, gb, gc;
    fetch(ga, 0);
    fetch(gb, GS);
    for (int g = 0; g < nval; g += 3 * GS) {
        fetch(gc, g + 2 * GS);
        process(ga, g);
        fetch(ga, g + 3 * GS);
        process(gb, g + GS);
        fetch(gb, g + 4 * GS);
        process(gc, g + 2 * GS);
    }
}

// ---------------------------------------------------------------------------------------------
// 32-byte store that does not displace resident lines: final outputs are written once and not read again here
__device__ __forceinline__ void st_global_256_cs(void* ptr, const uint4& a, const uint4& b) {
    asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
                 "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}

// Pass 4: one CTA per kFinishRows consecutive pillars of a frame.  Phase 1, kFinishRPW pillars per warp with the
// loads of all of them issued before the first is consumed: staged records before the break position -> float32
// coordinates in shared memory laid out exactly like the output rows (zero padded), mean, num_points, coors,
// point->slot map.  Phase 2, the whole CTA: the voxel rows and the decorated rows of the pillars are two contiguous
// runs of global memory and are streamed out with 16/32-byte stores.
constexpr int kFinishWarps = 8;
constexpr int kFinishRPW = 4;
constexpr int kFinishRows = kFinishWarps * kFinishRPW;

template <typename TO, int DS>
__global__ void __launch_bounds__(kFinishWarps * 32)
vox_finish_kernel(const int64_t* __restrict__ frame_off, VoxParams p, FastDiv div_P, int ncellp,
                  const uint2* __restrict__ rowinfo, const int* __restrict__ voxel_num,
                  const int* __restrict__ voxel_base, const int* __restrict__ cutoff, int64_t cap_rows,
                  const unsigned* __restrict__ sidx, const uint4* __restrict__ crec, const unsigned* __restrict__ ctag,
                  const int* __restrict__ nvalid, int S, TO* __restrict__ voxels, float* __restrict__ decorated, int* __restrict__ coors, int coors_cols,
                  int* __restrict__ num_points, int* __restrict__ point_slot, int* __restrict__ cell_voxel) {
    extern __shared__ __align__(16) float xyz[];  // [kFinishRows][P*D] float32 coordinates, output layout
    __shared__ float s_mean[kFinishRows][5];      // mx, my, mz, pillar centre x, y
    __shared__ int s_n[kFinishRows];
    constexpr int D = DS, Do = DS + 5;
    constexpr int rec16 = RecFmt<TO, DS>::kRec16;
    constexpr int NT = kFinishWarps * 32;
    const int P = p.max_points, PD = P * D;
    const int tid = threadIdx.x, lane = lane_id(), w = tid >> 5;
    const int b = blockIdx.y;
    const int M = voxel_num[b], vb = voxel_base[b];
    const int r0 = blockIdx.x * kFinishRows;
    if (r0 >= M || (int64_t)vb + r0 >= cap_rows) return;
    int nrows = min(kFinishRows, M - r0);
    if ((int64_t)vb + r0 + nrows > cap_rows) nrows = (int)(cap_rows - vb - r0);
    const int64_t row0 = (int64_t)vb + r0;
    const int cut = cutoff[b];
    const int64_t f0 = frame_off[b];
    const size_t cellrow0 = (size_t)b * p.ncell;
#ifndef PP_EXP_NOPREFETCH
    {
        // The pillars' records are gathered at random from the frame's record array.  The CTAs of a frame run at
        // about the same time, so each first asks L2 for one contiguous slice of that array (128-byte lines, only
        // the compacted part of every chunk): the array then comes from DRAM as a sequential stream instead of
        // sector by sector in gather order, and the gathers below find it in L2 or in flight.
        const int n = (int)(frame_off[b + 1] - f0);
        const int64_t lines = ((int64_t)n * rec16 * 16 + 127) >> 7;
        const int64_t per_cta = (lines + gridDim.x - 1) / gridDim.x;
        const char* base = reinterpret_cast<const char*>(crec + f0 * rec16);
        for (int64_t l = blockIdx.x * per_cta + tid; l < min(lines, (int64_t)(blockIdx.x + 1) * per_cta); l += NT) {
            const int rec = (int)((l << 7) / (rec16 * 16));  // first record of the line
            if ((rec & (kChunk - 1)) < __ldg(&nvalid[b * S + (rec >> kChunkShift)]))
                asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (l << 7)));
        }
    }
#endif

    // ---- phase 1: rows w, w + 8, w + 16, w + 24 of the block; the fast path holds a pillar in two registers sets
    uint2 info[kFinishRPW];  // cell, start of the cell's slot run | slots << 24
    const unsigned* sidx_b = sidx + cellrow0 * (size_t)P;
#pragma unroll
    for (int k = 0; k < kFinishRPW; ++k) {
        const int lr = w + k * kFinishWarps;
        info[k] = lr < nrows ? __ldg(&rowinfo[cellrow0 + r0 + lr]) : make_uint2(0u, 0u);
    }
    if (P <= 64) {
        // slot -> position of the record in the frame's record array (4-byte index table, L2 resident), then the
        // records themselves: both rounds of loads of all four pillars are in flight together
        unsigned tg[kFinishRPW][2];
#pragma unroll
        for (int k = 0; k < kFinishRPW; ++k) {
            const int tot = (int)(info[k].y >> 24);
            const unsigned si0 = info[k].y & 0xffffffu;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int s = h * 32 + lane;
                tg[k][h] = (w + k * kFinishWarps < nrows && s < tot) ? __ldcg(&sidx_b[si0 + s]) : 0x7fffffffu;
            }
        }
        uint4 rv[kFinishRPW][2][rec16];
#pragma unroll
        for (int k = 0; k < kFinishRPW; ++k) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                // slots are in index order, so the records before the break position are a prefix of the row
                if ((int)tg[k][h] < cut) {
#pragma unroll
                    for (int q = 0; q < rec16; ++q) rv[k][h][q] = __ldg(&crec[(f0 + tg[k][h]) * rec16 + q]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kFinishRPW; ++k) {
            const int lr = w + k * kFinishWarps;
            if (lr >= nrows) continue;  // uniform per warp
            const int rank = r0 + lr;
            const int64_t row = row0 + lr;
            const int cell = (int)info[k].x;
            float* vrow = xyz + (size_t)lr * PD;
            float sx = 0.f, sy = 0.f, sz = 0.f;
            int nsel = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int s = h * 32 + lane;
                // slots are in index order, so the records before the break position are a prefix of the row
                const bool ok = (int)tg[k][h] < cut;
                nsel += __popc(__ballot_sync(0xffffffffu, ok));
                if (ok) {
                    float c[DS];
                    if (sizeof(TO) == 4) {
                        c[0] = __uint_as_float(rv[k][h][0].x); c[1] = __uint_as_float(rv[k][h][0].y); c[2] = __uint_as_float(rv[k][h][0].z);
                        if (DS == 4) c[DS - 1] = __uint_as_float(rv[k][h][0].w);
                    } else {
                        const uint4 &u0 = rv[k][h][0], &u1 = rv[k][h][rec16 - 1];
                        const double v0x = __hiloint2double((int)u0.y, (int)u0.x), v0y = __hiloint2double((int)u0.w, (int)u0.z);
                        const double v1x = __hiloint2double((int)u1.y, (int)u1.x), v1y = __hiloint2double((int)u1.w, (int)u1.z);
                        TO* vo = voxels + (row * (int64_t)P + s) * D;
                        vo[0] = (TO)v0x; vo[1] = (TO)v0y; vo[2] = (TO)v1x;
                        if (DS == 4) vo[DS - 1] = (TO)v1y;
                        c[0] = (float)v0x; c[1] = (float)v0y; c[2] = (float)v1x;
                        if (DS == 4) c[DS - 1] = (float)v1y;
                    }
#pragma unroll
                    for (int d = 0; d < DS; ++d) vrow[s * D + d] = c[d];
                    sx += c[0]; sy += c[1]; sz += c[2];
                    if (point_slot) {
                        const unsigned cp = tg[k][h];
                        const int64_t orig = (int64_t)(cp & ~(unsigned)(kChunk - 1)) + (__ldg(&ctag[f0 + cp]) >> 16);
                        point_slot[f0 + orig] = rank * P + s;
                    }
                }
            }
            for (int q = nsel * D + lane; q < PD; q += 32) vrow[q] = 0.f;  // zero padding of the row
            if (sizeof(TO) == 8) {
                TO* vo = voxels + row * (int64_t)PD;
                for (int q = nsel * D + lane; q < PD; q += 32) vo[q] = (TO)0;
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                sx += __shfl_xor_sync(0xffffffffu, sx, o);
                sy += __shfl_xor_sync(0xffffffffu, sy, o);
                sz += __shfl_xor_sync(0xffffffffu, sz, o);
            }
            if (lane == 0) {
                const int cz = p.div_nxny.div(cell);
                const int rem = cell - cz * p.grid[0] * p.grid[1];
                const int cy = p.div_nx.div(rem), cx = rem - cy * p.grid[0];
                const float nf = (float)nsel;
                s_mean[lr][0] = __fdiv_rn(sx, nf); s_mean[lr][1] = __fdiv_rn(sy, nf); s_mean[lr][2] = __fdiv_rn(sz, nf);
                s_mean[lr][3] = __fadd_rn(__fmul_rn((float)cx, p.vx), p.x_off);
                s_mean[lr][4] = __fadd_rn(__fmul_rn((float)cy, p.vy), p.y_off);
                s_n[lr] = nsel;
                num_points[row] = nsel;
                int* co = coors + row * coors_cols;
                if (coors_cols == 4) *co++ = b;
                if (p.reverse_index) { co[0] = cz; co[1] = cy; co[2] = cx; }
                else { co[0] = cx; co[1] = cy; co[2] = cz; }
                if (cell_voxel) cell_voxel[cellrow0 + cell] = (int)row;
            }
        }
    } else {
        // general max_points: one pillar at a time, slots in rounds of 32
        for (int k = 0; k < kFinishRPW; ++k) {
            const int lr = w + k * kFinishWarps;
            if (lr >= nrows) continue;
            const int rank = r0 + lr;
            const int64_t row = row0 + lr;
            const int cell = (int)info[k].x, tot = (int)(info[k].y >> 24);
            const unsigned si0 = info[k].y & 0xffffffu;
            float* vrow = xyz + (size_t)lr * PD;
            float sx = 0.f, sy = 0.f, sz = 0.f;
            int nsel = 0;
            for (int s0 = 0; s0 < tot; s0 += 32) {
                const int s = s0 + lane;
                uint4 rv[rec16];
                const unsigned tag = s < tot ? __ldcg(&sidx_b[si0 + s]) : 0x7fffffffu;
                if ((int)tag < cut) {
#pragma unroll
                    for (int q = 0; q < rec16; ++q) rv[q] = __ldg(&crec[(f0 + tag) * rec16 + q]);
                }
                const bool ok = (int)tag < cut;
                nsel += __popc(__ballot_sync(0xffffffffu, ok));
                if (ok) {
                    float c[DS];
                    if (sizeof(TO) == 4) {
                        c[0] = __uint_as_float(rv[0].x); c[1] = __uint_as_float(rv[0].y); c[2] = __uint_as_float(rv[0].z);
                        if (DS == 4) c[DS - 1] = __uint_as_float(rv[0].w);
                    } else {
                        const uint4 &u0 = rv[0], &u1 = rv[rec16 - 1];
                        const double v0x = __hiloint2double((int)u0.y, (int)u0.x), v0y = __hiloint2double((int)u0.w, (int)u0.z);
                        const double v1x = __hiloint2double((int)u1.y, (int)u1.x), v1y = __hiloint2double((int)u1.w, (int)u1.z);
                        TO* vo = voxels + (row * (int64_t)P + s) * D;
                        vo[0] = (TO)v0x; vo[1] = (TO)v0y; vo[2] = (TO)v1x;
                        if (DS == 4) vo[DS - 1] = (TO)v1y;
                        c[0] = (float)v0x; c[1] = (float)v0y; c[2] = (float)v1x;
                        if (DS == 4) c[DS - 1] = (float)v1y;
                    }
#pragma unroll
                    for (int d = 0; d < DS; ++d) vrow[s * D + d] = c[d];
                    sx += c[0]; sy += c[1]; sz += c[2];
                    if (point_slot) {
                        const int64_t orig = (int64_t)(tag & ~(unsigned)(kChunk - 1)) + (__ldg(&ctag[f0 + tag]) >> 16);
                        point_slot[f0 + orig] = rank * P + s;
                    }
                }
            }
            for (int q = nsel * D + lane; q < PD; q += 32) vrow[q] = 0.f;
            if (sizeof(TO) == 8) {
                TO* vo = voxels + row * (int64_t)PD;
                for (int q = nsel * D + lane; q < PD; q += 32) vo[q] = (TO)0;
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                sx += __shfl_xor_sync(0xffffffffu, sx, o);
                sy += __shfl_xor_sync(0xffffffffu, sy, o);
                sz += __shfl_xor_sync(0xffffffffu, sz, o);
            }
            if (lane == 0) {
                const int cz = p.div_nxny.div(cell);
                const int rem = cell - cz * p.grid[0] * p.grid[1];
                const int cy = p.div_nx.div(rem), cx = rem - cy * p.grid[0];
                const float nf = (float)nsel;
                s_mean[lr][0] = __fdiv_rn(sx, nf); s_mean[lr][1] = __fdiv_rn(sy, nf); s_mean[lr][2] = __fdiv_rn(sz, nf);
                s_mean[lr][3] = __fadd_rn(__fmul_rn((float)cx, p.vx), p.x_off);
                s_mean[lr][4] = __fadd_rn(__fmul_rn((float)cy, p.vy), p.y_off);
                s_n[lr] = nsel;
                num_points[row] = nsel;
                int* co = coors + row * coors_cols;
                if (coors_cols == 4) *co++ = b;
                if (p.reverse_index) { co[0] = cz; co[1] = cy; co[2] = cx; }
                else { co[0] = cx; co[1] = cy; co[2] = cz; }
                if (cell_voxel) cell_voxel[cellrow0 + cell] = (int)row;
            }
        }
    }
    __syncthreads();

    // ---- phase 2a: voxel rows, nrows*P*D floats with the same layout in shared and in global memory
    if (voxels && sizeof(TO) == 4) {
        float* dst = reinterpret_cast<float*>(voxels) + row0 * (int64_t)PD;
        const int nf = nrows * PD;
        const uintptr_t a = reinterpret_cast<uintptr_t>(dst);
        if ((a & 15) == 0) {
            for (int k = tid; k < (nf >> 2); k += NT) __stcs(reinterpret_cast<float4*>(dst) + k, reinterpret_cast<const float4*>(xyz)[k]);
            for (int k = (nf & ~3) + tid; k < nf; k += NT) dst[k] = xyz[k];
        } else if ((a & 7) == 0) {
            // rows of an odd row0 start 8 bytes into a 16-byte unit: one float2, then float4 from shared float2 pairs
            if (tid == 0) *reinterpret_cast<float2*>(dst) = *reinterpret_cast<const float2*>(xyz);
            const int n4 = (nf - 2) >> 2;
            for (int k = tid; k < n4; k += NT) {
                const float2 lo = *reinterpret_cast<const float2*>(xyz + 2 + 4 * k), hi = *reinterpret_cast<const float2*>(xyz + 4 + 4 * k);
                __stcs(reinterpret_cast<float4*>(dst + 2) + k, make_float4(lo.x, lo.y, hi.x, hi.y));
            }
            for (int k = 2 + 4 * n4 + tid; k < nf; k += NT) dst[k] = xyz[k];
        } else {
            for (int k = tid; k < nf; k += NT) dst[k] = xyz[k];
        }
    }
    // ---- phase 2b: decorated rows (model/pointpillars.py:143-203), nrows*P points of D+5 floats
    if (decorated) {
        float* drow = decorated + row0 * (int64_t)P * Do;
        const int npts = nrows * P;
        if (DS == 3 && (reinterpret_cast<uintptr_t>(drow) & 31) == 0) {
            // 8 floats per point = one 32-byte store: (x,y,z,x-mx, y-my,z-mz,x-ex,y-ey)
            for (int q = tid; q < npts; q += NT) {
                const int r = div_P.div(q), sl = q - r * P;
                uint4 o0 = make_uint4(0u, 0u, 0u, 0u), o1 = o0;
                if (sl < s_n[r]) {
                    const float* v = xyz + (size_t)r * PD + sl * 3;
                    const float q0 = v[0], q1 = v[1], q2 = v[2];
                    o0 = make_uint4(__float_as_uint(q0), __float_as_uint(q1), __float_as_uint(q2), __float_as_uint(q0 - s_mean[r][0]));
                    o1 = make_uint4(__float_as_uint(q1 - s_mean[r][1]), __float_as_uint(q2 - s_mean[r][2]),
                                    __float_as_uint(q0 - s_mean[r][3]), __float_as_uint(q1 - s_mean[r][4]));
                }
                st_global_256_cs(drow + (size_t)q * 8, o0, o1);
            }
        } else {
            const int nfl = npts * Do;
            for (int k = tid; k < nfl; k += NT) {
                const int q = k / Do, d = k - q * Do;  // Do is a compile-time constant
                const int r = div_P.div(q), sl = q - r * P;
                float o = 0.f;
                if (sl < s_n[r]) {
                    const float* v = xyz + (size_t)r * PD + sl * D;
                    o = d < D ? v[d] : d < D + 3 ? v[d - D] - s_mean[r][d - D] : v[d - D - 3] - s_mean[r][d - D];
                }
                drow[k] = o;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
struct SmallWs {
    uint4* crec;              // [total_points + 1] records, compacted per chunk at the chunk's own offset
    unsigned* ctag;           // [total_points + 1] cell | index in chunk << 16, same positions
    unsigned short* hist;     // [B*S*ncellp]  chunk counts
    unsigned char* base8;     // [B*S*ncellp]  chunk bases
    unsigned* cellinfo;       // [B*ncellp]    start of the cell's run in the frame's slot table | min(P, points) << 24
    int* frame_done;          // [B]
    int* nvalid;              // [B*S] records per chunk
    int* newcount;            // [B*S] voxels opened per chunk
    uint2* rowinfo;           // [B*ncell] voxel id in frame -> {cell, cellinfo of the cell}
    int* cutoff;              // [B] break position (record position in frame) or kNoCut
    int* done_counter;        // [1]
    unsigned* sidx;           // [B][ncell*P] slot table: the cells' runs back to back -> record position in frame
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
    w.crec = reinterpret_cast<uint4*>(c.take<unsigned char>(((size_t)total_points + 1) * rb));
    w.ctag = c.take<unsigned>((size_t)total_points + 1);
    w.hist = c.take<unsigned short>((size_t)n_frames * S * ncellp);
    w.base8 = c.take<unsigned char>((size_t)n_frames * S * ncellp);
    w.cellinfo = c.take<unsigned>((size_t)n_frames * ncellp);
    w.frame_done = c.take<int>(n_frames);
    w.nvalid = c.take<int>((size_t)n_frames * S);
    w.newcount = c.take<int>((size_t)n_frames * S);
    w.rowinfo = c.take<uint2>((size_t)n_frames * ncell);
    w.cutoff = c.take<int>(n_frames);
    w.done_counter = c.take<int>(1);
    w.sidx = c.take<unsigned>((size_t)n_frames * ncell * cfg->max_points);
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
    const double staging = (double)n_frames * (double)ncell * cfg->max_points * 4.0;
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
    const size_t smem = kScanStages * ((size_t)kScanTile * DS * sizeof(T) + 32) + (size_t)ncellp * 2;
    auto kern = vox_scan_kernel<T, A32, FAST, TO, DS>;
    int per_sm = 0;
    PP_TRY_RC(kernel_config(reinterpret_cast<const void*>(kern), kScanThreads, smem, &per_sm));
    const int aligned16 = (reinterpret_cast<uintptr_t>(points) & 15) == 0;
    PP_TIMED("vox_scan", st);
    kern<<<dim3((unsigned)S, (unsigned)n_frames), kScanThreads, smem, st>>>(
        static_cast<const T*>(points), frame_off, p, total_points, aligned16, S, ncellp, w.crec, w.ctag,
        reinterpret_cast<unsigned*>(w.hist), w.nvalid, w.newcount, w.cutoff, point_slot, w.done_counter, w.frame_done);
    PP_LAUNCHED();
    return PP_OK;
}

template <typename TO, int DS>
static int launch_finish(const VoxParams& p, const SmallWs& w, const int64_t* frame_off, int ncellp, int S, int n_frames,
                         int64_t rows_per_frame, const int32_t* voxel_num, const int32_t* voxel_base, int64_t cap_rows,
                         void* voxels, float* decorated, int32_t* coors, int coors_cols, int32_t* num_points,
                         int32_t* point_slot, int32_t* cell_voxel, cudaStream_t st) {
    const int P = p.max_points;
    const size_t smem = (size_t)kFinishRows * P * DS * sizeof(float);
    auto kern = vox_finish_kernel<TO, DS>;
    int per_sm = 0;
    PP_CHECK_ARG(smem <= 200 * 1024, "pp_voxelize_dev: max_points too large for the finish pass");
    PP_TRY_RC(kernel_config(reinterpret_cast<const void*>(kern), kFinishWarps * 32, smem, &per_sm));
    const dim3 g((unsigned)ceil_div(rows_per_frame, kFinishRows), (unsigned)n_frames);
    PP_TIMED("vox_finish", st);
    kern<<<g, kFinishWarps * 32, smem, st>>>(
        frame_off, p, FastDiv((unsigned)P), ncellp, w.rowinfo, voxel_num, voxel_base, w.cutoff, cap_rows, w.sidx, w.crec,
        w.ctag, w.nvalid, S, static_cast<TO*>(voxels), decorated, coors, coors_cols, num_points, point_slot, cell_voxel);
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
        vox_prefix_kernel<<<g, kPrefixThreads, 0, st>>>(frame_offsets, S, (int)ncell, ncellp, P, w.hist, w.base8, w.cellinfo,
                                                        w.newcount, cfg->max_voxels, n_frames, voxel_num, voxel_base,
                                                        w.done_counter, w.frame_done);
        PP_LAUNCHED();
    }
    if (cap_rows <= 0) return PP_OK;
    {
        const size_t smem = (size_t)ncellp;
        const dim3 g((unsigned)S, (unsigned)n_frames);
        PP_TIMED("vox_place", st);
        vox_place_kernel<<<g, 32, smem, st>>>(frame_offsets, S, (int)ncell, ncellp, P, cfg->max_voxels, w.ctag, w.base8, w.cellinfo,
                                              w.nvalid, w.newcount, w.sidx, w.rowinfo, w.cutoff);
        PP_LAUNCHED();
    }
    const int64_t rows_per_frame = cfg->max_voxels < ncell ? cfg->max_voxels : ncell;
    if (rows_per_frame <= 0) return PP_OK;
#define PP_FINISH(TO, DS)                                                                                            \
    launch_finish<TO, DS>(p, w, frame_offsets, ncellp, S, n_frames, rows_per_frame, voxel_num, voxel_base, cap_rows,  \
                          voxels, decorated, coors, coors_cols, num_points, point_slot, cell_voxel, st)
    if (out_dtype == PP_F64) rc = D == 3 ? PP_FINISH(double, 3) : PP_FINISH(double, 4);
    else rc = D == 3 ? PP_FINISH(float, 3) : PP_FINISH(float, 4);
#undef PP_FINISH
    return rc;
}

}  // namespace pp
