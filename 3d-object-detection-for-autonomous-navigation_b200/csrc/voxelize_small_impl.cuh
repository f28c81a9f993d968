// Body of the table path (see voxelize_small.cu), compiled once per chunk size:
//   PP_VS_NS     namespace of this instance
//   PP_VS_SHIFT  log2 of the points per chunk
//   PP_VS_PPT    points per thread and scan tile (tile = 256 threads x this)
//   PP_VS_STAGES shared-memory stages of the scan pass (tiles in flight)
//   PP_VS_PLACE_THREADS  threads of a place CTA (one warp walks, the others only help with the prologue)
//   PP_VS_KB     16-byte table units per lane that the place pass loads per round trip
//   PP_VS_SUB    parts of a chunk that the place pass walks independently (a part is a whole number of scan tiles)
namespace pp {
namespace PP_VS_NS {


constexpr int kChunkShift = PP_VS_SHIFT;  // points per chunk = 2^shift; positions inside a chunk fit 16 bits
constexpr int kChunk = 1 << kChunkShift;
constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kScanPPT = PP_VS_PPT;  // points per thread and tile
constexpr int kWarpPts = 32 * kScanPPT;  // consecutive points of a tile that one warp owns
constexpr int kScanTile = kScanThreads * kScanPPT;
constexpr int kPrefixThreads = 256;
// prefix CTA = kPrefixCells cells x kPrefixParts ranges of the frame's chunks.  Batches: one thread per cell walks all
// chunks (64 frames x 25 chunks: 43 us; split eight ways the same batch takes 103 us, all CTA turnover).  Short batches
// (small chunks, 100 per d435i frame, few CTAs): eight ranges per cell, 21 -> 15 us for one frame.
constexpr int kPrefixParts = kChunkShift <= 12 ? 8 : 1;
constexpr int kPrefixCells = kPrefixThreads / kPrefixParts;
constexpr int kMaxChunks = (1 << 20) >> kChunkShift;  // per frame: max_frame_points <= 2^20
static_assert(kMaxChunks <= 1024, "s_new of the prefix pass");
constexpr int kMaxCellsSmall = 16384;              // cell ids are 14-bit in the record word
constexpr int kMaxFramePointsSmall = kMaxChunks * kChunk;
constexpr int kMaxPointsSmall = 254;               // uint8 counts saturate at 255
constexpr int kPlaceUnroll = 4;
constexpr int kSub = PP_VS_SUB;                    // parts (quarters) of a chunk that the place pass walks independently
constexpr int kSubTiles = kChunk / kScanTile / kSub;  // scan tiles per quarter
constexpr int kNoCut = 0x7fffffff;

static_assert(kChunk % (kScanTile * kSub) == 0 && kChunk <= 65536, "chunk size");

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
// Pass 1.  Warp w of a tile owns kWarpPts consecutive points; round r of lane l is point kWarpPts w + 32 r + l,
// so (warp, round, lane) enumerates the tile in index order and shared-memory row reads are conflict free.
// kScanStages shared-memory stages: with two, the TMA bulk copy of tile j+1 is in flight while tile j is processed.
constexpr int kScanStages = PP_VS_STAGES;

template <typename T, bool A32, bool FAST, typename TO, int DS>
__global__ void __launch_bounds__(kScanThreads)
vox_scan_kernel(const T* __restrict__ points, const int64_t* __restrict__ frame_off, VoxParams p, int64_t total_points,
                int aligned16, int S, int ncellp, uint4* __restrict__ crec, unsigned* __restrict__ ctag,
                unsigned* __restrict__ hist_out, int* __restrict__ nvalid, int* __restrict__ newcount,
                int* __restrict__ cutoff, int* __restrict__ point_slot, int* __restrict__ done_counter,
                unsigned* __restrict__ snap, int* __restrict__ subv) {
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
        if (s == 0) cutoff[b] = kNoCut;
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
        if (j % kSubTiles == 0 && j > 0) {
            // quarter boundary: the counts so far (saturated to uint8) and the number of records so far let the
            // place pass start a walk here, independently of the quarters before
            const int q = j / kSubTiles;
            unsigned* dst = snap + (((size_t)b * S + s) * (kSub - 1) + (q - 1)) * (size_t)(ncellp >> 2);
            for (int k = tid; k < (ncellp >> 2); k += kScanThreads) {
                const unsigned lo = hist32[2 * k], hi = hist32[2 * k + 1];
                dst[k] = min(lo & 0xffffu, 255u) | (min(lo >> 16, 255u) << 8) | (min(hi & 0xffffu, 255u) << 16) | (min(hi >> 16, 255u) << 24);
            }
            if (tid == 0) subv[(b * S + s) * kSub + q] = crun;
            __syncthreads();  // before the next tile's atomics change the counts
        }
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
        const unsigned char* rows = buf + shift + (size_t)(w * kWarpPts + lane) * row_bytes;

        int cell[kScanPPT];
        unsigned bal[kScanPPT];
        int wtot = 0;
#pragma unroll
        for (int r = 0; r < kScanPPT; ++r) {
            const int t = w * kWarpPts + r * 32 + lane;
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
            const int t = w * kWarpPts + r * 32 + lane;
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
    if (tid == 0) {
        nvalid[b * S + s] = crun;
        subv[(b * S + s) * kSub] = 0;
        for (int q = (ntiles + kSubTiles - 1) / kSubTiles; q < kSub; ++q) subv[(b * S + s) * kSub + q] = crun;  // empty quarters
    }
}


// ---------------------------------------------------------------------------------------------
// Pass 2: one thread per cell.  base8[s][cell] = min(255, points of the cell in chunks < s) = the slot of the
// cell's first record of chunk s; min(points, max_points) goes to the last word of the cell's row of the slot table;
// a cell is counted as a new voxel of the first chunk that holds it.  The last CTA of the grid turns the per-chunk voxel counts into voxel_num and
// voxel_base (rows of a batch are packed back to back, merge_second_batch layout).
__global__ void __launch_bounds__(kPrefixThreads)
vox_prefix_kernel(const int64_t* __restrict__ frame_off, int S, int ncell, int ncellp, int P,
                  const unsigned short* __restrict__ hist, unsigned char* __restrict__ base8,
                  unsigned* __restrict__ sidx, int* __restrict__ newcount, int max_voxels, int B,
                  int* __restrict__ voxel_num, int* __restrict__ voxel_base, int* __restrict__ done_counter,
                  int* __restrict__ cell_voxel) {
    __shared__ int s_new[kMaxChunks];
    __shared__ int s_part[kPrefixParts][kPrefixCells];
    __shared__ int sm[33];
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid % kPrefixCells, part = tid / kPrefixCells;
    const int b = blockIdx.y;
    const int cell = blockIdx.x * kPrefixCells + lane;
    const int n = (int)(frame_off[b + 1] - frame_off[b]);
    const int Sb = (n + kChunk - 1) >> kChunkShift;
    for (int i = tid; i < kMaxChunks; i += kPrefixThreads) s_new[i] = 0;
    // the chunks of a frame are split into kPrefixParts ranges: every thread sums its range, the ranges are combined
    // through shared memory, then every thread walks its range again with the right start
    const int per = (Sb + kPrefixParts - 1) / kPrefixParts;
    const int s_lo = part * per, s_hi = min(Sb, s_lo + per);
    const size_t t0 = (size_t)b * S * ncellp + cell;
    int sum = 0;
    if (kPrefixParts > 1 && cell < ncell) {
        for (int s0 = s_lo; s0 < s_hi; s0 += 16) {
            int h[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) h[k] = s0 + k < s_hi ? (int)__ldcg(&hist[t0 + (size_t)(s0 + k) * ncellp]) : 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) sum += h[k];
        }
    }
    s_part[part][lane] = sum;
    __syncthreads();
    if (cell < ncell) {
        if (cell_voxel && part == 0) cell_voxel[(size_t)b * ncell + cell] = -1;  // the finish pass fills in the cells that get a row
        int run = 0, first = -1;
        for (int q = 0; q < part; ++q) run += s_part[q][lane];
        // eight loads in flight per round trip (32 at once, a whole d435i frame of 25 chunks, costs occupancy: 46 -> 62 us)
        for (int s0 = s_lo; s0 < s_hi; s0 += 8) {
            int h[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) h[k] = s0 + k < s_hi ? (int)__ldcg(&hist[t0 + (size_t)(s0 + k) * ncellp]) : 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (s0 + k < s_hi) {
                    base8[t0 + (size_t)(s0 + k) * ncellp] = (unsigned char)min(run, 255);
                    if (run == 0 && h[k] > 0) first = s0 + k;
                    run += h[k];
                }
            }
        }
        if (part == kPrefixParts - 1) sidx[((size_t)b * ncell + cell) * (size_t)(P + 1) + P] = (unsigned)min(run, P);
        if (first >= 0) atomicAdd(&s_new[first], 1);
    }
    __syncthreads();
    for (int i = tid; i < Sb; i += kPrefixThreads)
        if (s_new[i]) atomicAdd(&newcount[b * S + i], s_new[i]);
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(done_counter, 1) == (int)(gridDim.x * gridDim.y) - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // voxels per frame: one warp per frame sums the frame's chunks (chunks past the frame's end hold 0)
    for (int i = tid >> 5; i < B; i += kPrefixThreads / 32) {
        int v = 0;
        for (int s = tid & 31; s < S; s += 32) v += __ldcg(&newcount[i * S + s]);
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) voxel_num[i] = min(v, max_voxels);
    }
    __syncthreads();
    int running = 0;
    for (int b0 = 0; b0 < B; b0 += kPrefixThreads) {
        const int i = b0 + tid;
        const int v = i < B ? voxel_num[i] : 0;
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
// Pass 3: one warp per (frame, chunk, quarter) walks the quarter's record tags in index order, 32 per step, with
// the running per-cell counts (uint8) in shared memory: slot = count[cell] + rank among the step's earlier
// records of the same cell.  The counts start at the chunk base plus the chunk's own counts at the
// quarter boundary, which the scan pass snapshotted -- so the four quarters of a chunk are walked independently,
// and a 64-frame batch has 6 400 short dependent chains in flight instead of 1 600 long ones (the walk is bound
// by per-warp instruction latency, not by bandwidth).  A record that finds count 0 opens its cell: the walk is in
// index order, so the running number of such records IS the voxel id (order of first touch), and the record
// that would open voxel number max_voxels is the reference's `break` position (load_data.py:630-634).  The
// record's position goes to entry (cell, slot) of the frame's slot table (4 bytes per kept point, L2 resident;
// max_points entries per cell, so the address needs no lookup: the pass is bound by its scattered sector
// accesses, one store per record).  Tags are fetched two groups ahead.  No atomic, no block barrier.
constexpr unsigned kNone = 0xffffffffu;
constexpr int kPlaceScratch = 1024;  // words of the per-warp peer table (a power of two)
// Threads of a place CTA.  ONE warp walks; three more warps help with the prologue, the 20 KB of chunk tables the walk
// starts from, and leave -- ncu had 45 % of the pass's samples of a single frame (20 % for a 64-frame batch) waiting for
// those loads with one warp's worth in flight.
constexpr int kPlaceThreads = PP_VS_PLACE_THREADS;

__global__ void __launch_bounds__(kPlaceThreads)
vox_place_kernel(const int64_t* __restrict__ frame_off, int S, int ncell, int ncellp, int P, int max_voxels,
                 const unsigned* __restrict__ ctag, const unsigned char* __restrict__ base8,
                 const int* __restrict__ nvalid,
                 const int* __restrict__ newcount, unsigned* __restrict__ sidx, unsigned* __restrict__ rowinfo,
                 int* __restrict__ cutoff, const unsigned char* __restrict__ snap, const int* __restrict__ subv) {
    extern __shared__ __align__(16) unsigned char tbl[];  // [ncellp] running counts
    __shared__ unsigned scr[kPlaceScratch];                // lanes of the current step, keyed by cell (all zero between steps)
    __shared__ int s_run[kPlaceThreads / 32];
    constexpr int U = kPlaceUnroll;
    const int lane = lane_id(), tid = threadIdx.x, nthr = blockDim.x;  // 32, or up to kPlaceThreads when the launch has few walks
    for (int k = tid; k < kPlaceScratch; k += nthr) scr[k] = 0u;
    // frames in reverse order: the scan pass wrote the last frames' tags last, so they are still in L2
    const int b = (int)gridDim.y - 1 - (int)blockIdx.y, s = blockIdx.x / kSub, q = blockIdx.x % kSub;
    const int64_t f0 = frame_off[b];
    const int n = (int)(frame_off[b + 1] - f0);
    const int Sb = (n + kChunk - 1) >> kChunkShift;
    if (s >= Sb) return;
    // the quarter's records
    const int rec0 = subv[(b * S + s) * kSub + q];
    const int nval = q + 1 < kSub ? subv[(b * S + s) * kSub + q + 1] : nvalid[b * S + s];
    if (nval <= rec0) return;
    // voxels opened by the earlier chunks of the frame
    int newrun = 0;
    for (int k = tid; k < s; k += nthr) newrun += newcount[b * S + k];
    {
        // counts at the start of the quarter = chunk base + the chunk's counts at the quarter boundary; a cell
        // that the chunk base does not hold but the boundary counts do was opened by an earlier quarter
        const uint4* srcb = reinterpret_cast<const uint4*>(base8 + ((size_t)b * S + s) * ncellp);
        const uint4* srcs = reinterpret_cast<const uint4*>(snap + (((size_t)b * S + s) * (kSub - 1) + (q ? q - 1 : 0)) * (size_t)ncellp);
        uint4* dst = reinterpret_cast<uint4*>(tbl);
        constexpr int kB = PP_VS_KB;  // 16-byte units per lane and batch: all loads of a batch are in flight together
        const int nu = ncellp >> 4;
        for (int k0 = tid; k0 < nu; k0 += nthr * kB) {
            uint4 a[kB], c[kB];
#pragma unroll
            for (int i = 0; i < kB; ++i) {
                const int k = min(k0 + nthr * i, nu - 1);
                a[i] = __ldcg(&srcb[k]);
                c[i] = q ? __ldcs(&srcs[k]) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int i = 0; i < kB; ++i) {
                if (k0 + nthr * i >= nu) break;
                newrun += (__popc(__vcmpeq4(a[i].x, 0u) & __vcmpne4(c[i].x, 0u)) + __popc(__vcmpeq4(a[i].y, 0u) & __vcmpne4(c[i].y, 0u)) +
                           __popc(__vcmpeq4(a[i].z, 0u) & __vcmpne4(c[i].z, 0u)) + __popc(__vcmpeq4(a[i].w, 0u) & __vcmpne4(c[i].w, 0u))) >> 3;
                dst[k0 + nthr * i] = make_uint4(__vaddus4(a[i].x, c[i].x), __vaddus4(a[i].y, c[i].y), __vaddus4(a[i].z, c[i].z),
                                              __vaddus4(a[i].w, c[i].w));
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) newrun += __shfl_xor_sync(0xffffffffu, newrun, o);
    if (kPlaceThreads > 32 && nthr > 32) {
        if (lane == 0) s_run[tid >> 5] = newrun;
        __syncthreads();
        if (tid >= 32) return;  // the helpers of the prologue are done
        newrun = 0;
        for (int k = 0; k < (nthr >> 5); ++k) newrun += s_run[k];
    }
    __syncwarp();
    const unsigned* tags = ctag + f0 + (int64_t)s * kChunk;
    const size_t cellrow0 = (size_t)b * ncell;
    unsigned* sidx_b = sidx + cellrow0 * (size_t)(P + 1);  // the frame's slot table: max_points entries (+ the count) per cell

    struct Tags { unsigned tg[U]; };
    // unconditional loads (index clamped to the last record): a predicated load makes the compiler merge the
    // loaded registers with their old contents right behind the load, which waits for it and defeats the prefetch
    auto fetch_tags = [&](Tags& t, int g) {
#pragma unroll
        for (int u = 0; u < U; ++u) t.tg[u] = __ldcs(&tags[max(0, min(g + u * 32 + lane, nval - 1))]);
    };
    auto process = [&](const Tags& cur, int g) {
        if (g >= nval) return;  // uniform
        int c[U], rk[U], npeer[U];
        // Lanes of a step that share a cell: every lane ORs its bit into a shared-memory word keyed by the low
        // bits of its cell and reads the word back -- 4.4 cycles per step and SM where match_any takes 60 (it
        // costs ~2 cycles per distinct value on a unit the SM's warps share: tools/micro/match_tput.cu).  Two
        // different cells under one key are caught by comparing with the word's first lane; such a step (rare:
        // neighbouring points differ in the low bits) is redone with match_any.
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool act = g + u * 32 + lane < nval;
            c[u] = (int)(cur.tg[u] & 0xffffu);
            unsigned* slot = scr + (c[u] & (kPlaceScratch - 1));
            if (act) atomicOr(slot, 1u << lane);
            __syncwarp();
            unsigned peers = act ? *slot : 1u << lane;
            __syncwarp();
            if (act) *slot = 0u;
            const int first = __shfl_sync(0xffffffffu, c[u], __ffs(peers) - 1);
            if (__any_sync(0xffffffffu, act && first != c[u]))
                peers = __match_any_sync(0xffffffffu, act ? c[u] : (0x10000 | lane));
            __syncwarp();
            rk[u] = act ? __popc(peers & lanemask_lt()) : -1;  // rk 0: first record of its cell in the step
            npeer[u] = __popc(peers);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (g + u * 32 >= nval) break;  // uniform
            // the sequential part
            const int cnt = rk[u] >= 0 ? (int)tbl[c[u]] : 1;
            if (rk[u] == 0) tbl[c[u]] = (unsigned char)min(cnt + npeer[u], 255);
            __syncwarp();
            // a count of 0 means no earlier point of the frame fell in the cell: the record opens a voxel, and
            // since the walk is in index order the running number of opened cells is the voxel id
            const bool opener = rk[u] == 0 && cnt == 0;
            const unsigned opens = __ballot_sync(0xffffffffu, opener);
            // position of the record in its frame's record array: same order as the point indices
            const unsigned cp = (unsigned)((s << kChunkShift) + g + u * 32 + lane);
            if (opener) {
                const int rank = newrun + __popc(opens & lanemask_lt());
                // the finish pass finds cell and point count of a voxel in one word
                if (rank < max_voxels) rowinfo[cellrow0 + rank] = (unsigned)c[u];
                else if (rank == max_voxels) cutoff[b] = (int)cp;  // the reference's break position
            }
            newrun += __popc(opens);
            const int slot = cnt + rk[u];
            if (rk[u] >= 0 && slot < P) sidx_b[c[u] * (P + 1) + slot] = cp;
        }
    };
    // register buffers used in rotation (no copies: a register move of a load that is still in flight would
    // wait for it): the tags of group i+3 are requested before group i is processed
    constexpr int GS = U * 32;
    Tags t0, t1, t2, t3;
    fetch_tags(t0, rec0);
    fetch_tags(t1, rec0 + GS);
    fetch_tags(t2, rec0 + 2 * GS);
    for (int g = rec0; g < nval; g += 4 * GS) {
        fetch_tags(t3, g + 3 * GS);
        process(t0, g);
        fetch_tags(t0, g + 4 * GS);
        process(t1, g + GS);
        fetch_tags(t1, g + 5 * GS);
        process(t2, g + 2 * GS);
        fetch_tags(t2, g + 6 * GS);
        process(t3, g + 3 * GS);
    }
}

// ---------------------------------------------------------------------------------------------
// 32-byte store that does not displace resident lines: final outputs are written once and not read again here
// (one full L2 sector per lane; SASS STG.E.ENL2.256 on sm_100)
__device__ __forceinline__ void st_global_256_cs(void* ptr, const float4& a, const float4& b) {
    asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w),
                 "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
                 : "memory");
}

// Pass 4: one warp per two consecutive pillars of a frame, lane = slot (NR slots per lane).  The loads of both
// pillars are issued before the first is consumed: their {cell, run} words, the runs' entries of the slot table
// (already in index order), the records they point at (before the break position).  Every lane owns its slot of
// the output rows whether or not a point sits in it -- lanes of empty slots write the zero padding -- so the
// voxel rows and the fused PillarFeatureNet decoration (model/pointpillars.py:143-203) go out as dense,
// coalesced stores straight from registers (a decorated d435i point is one 32-byte store).  The sums are reduced
// so that half-warp h ends up with pillar h's: everything that is per pillar (mean, cell -> coordinates -> pillar
// centre, num_points, coors) is computed once per warp, for both pillars.  No shared memory, few registers: the
// three dependent round trips of a pillar are hidden by the other warps of the SM.
constexpr int kFinishWarps = 4;
constexpr int kFinishRows = 2 * kFinishWarps;

#ifndef PP_FINISH_MINB
#define PP_FINISH_MINB 1
#endif
template <typename TO, int DS, int NR>
__global__ void __launch_bounds__(kFinishWarps * 32, PP_FINISH_MINB)
vox_finish_kernel(const int64_t* __restrict__ frame_off, VoxParams p, const unsigned* __restrict__ rowinfo,
                  const int* __restrict__ voxel_num, const int* __restrict__ voxel_base,
                  const int* __restrict__ cutoff, int64_t cap_rows, const unsigned* __restrict__ sidx,
                  const uint4* __restrict__ crec, const unsigned* __restrict__ ctag, const int* __restrict__ nvalid,
                  int S, TO* __restrict__ voxels, float* __restrict__ decorated, int* __restrict__ coors,
                  int coors_cols, int* __restrict__ num_points, int* __restrict__ point_slot,
                  int* __restrict__ cell_voxel) {
    constexpr int D = DS, Do = DS + 5;
    constexpr int rec16 = RecFmt<TO, DS>::kRec16;
    constexpr int NT = kFinishWarps * 32;
    const int P = p.max_points;
    const int tid = threadIdx.x, lane = lane_id(), w = tid >> 5;
    const int b = blockIdx.y;
    const int M = voxel_num[b], vb = voxel_base[b];
    const int r0 = blockIdx.x * kFinishRows;
    if (r0 >= M || (int64_t)vb + r0 >= cap_rows) return;
    int nrows = min(kFinishRows, M - r0);
    if ((int64_t)vb + r0 + nrows > cap_rows) nrows = (int)(cap_rows - vb - r0);
    const unsigned cut = (unsigned)cutoff[b];
    const int64_t f0 = frame_off[b];
    {
        // The pillars' records are gathered at random from the frame's record array.  The CTAs of a frame run at
        // about the same time, so each first asks L2 for one contiguous slice of that array (128-byte lines, only
        // the compacted part of every chunk): the array then comes from DRAM as a sequential stream instead of
        // sector by sector in gather order, and the gathers below find it in L2 or in flight.
        const int n = (int)(frame_off[b + 1] - f0);
        const int lines = (int)(((int64_t)n * rec16 * 16 + 127) >> 7);  // n <= 2^20 points
        const int per_cta = (lines + (int)gridDim.x - 1) / (int)gridDim.x;
        const char* base = reinterpret_cast<const char*>(crec + f0 * rec16);
        for (int l = blockIdx.x * per_cta + tid; l < min(lines, (int)(blockIdx.x + 1) * per_cta); l += NT) {
            const int rec = (l << 3) / rec16;  // first record of the line
            if ((rec & (kChunk - 1)) < __ldg(&nvalid[b * S + (rec >> kChunkShift)]))
                asm volatile("prefetch.global.L2 [%0];" ::"l"(base + ((int64_t)l << 7)));
        }
    }
    const int lr0 = 2 * w;
    if (lr0 >= nrows) return;  // uniform per warp
    const bool two = lr0 + 1 < nrows;
    const int64_t row0 = (int64_t)vb + r0 + lr0;  // output row of the warp's first pillar
    // ---- loads: cells -> rows of the slot table (entries + count, requested together) -> records
    unsigned cellk[2];
    {
        const unsigned* ri = rowinfo + (size_t)b * p.ncell + r0 + lr0;
        cellk[0] = __ldg(&ri[0]);
        cellk[1] = two ? __ldg(&ri[1]) : cellk[0];
    }
    unsigned key[2][NR];
    int len[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const unsigned* run = sidx + ((size_t)b * p.ncell + cellk[k]) * (size_t)(P + 1);
        len[k] = (int)__ldcg(&run[P]);
#pragma unroll
        for (int r = 0; r < NR; ++r) key[k][r] = __ldcg(&run[min(r * 32 + lane, P)]);  // entries past the count are not used
    }
    if (!two) len[1] = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
#pragma unroll
        for (int r = 0; r < NR; ++r)
            if (r * 32 + lane >= len[k]) key[k][r] = kNone;
    }
    uint4 rv[2][NR][rec16];
    {
        const uint4* rec = crec + f0 * rec16;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                // the run is in index order: the records before the break position are a prefix of it
                if (key[k][r] < cut) {
#pragma unroll
                    for (int q = 0; q < rec16; ++q) rv[k][r][q] = __ldg(&rec[(size_t)key[k][r] * rec16 + q]);
                }
            }
        }
    }
    // ---- 1. coordinates, voxel rows, sums
    float c[2][NR][DS];
    float sx[2], sy[2], sz[2];
    int nsel[2];
    unsigned okm = 0u;  // bit k*NR + r: this lane holds a point of pillar k in slot r*32 + lane
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        sx[k] = sy[k] = sz[k] = 0.f;
        nsel[k] = 0;
        TO* vrow = voxels ? voxels + ((row0 + k) * (int64_t)P + lane) * D : nullptr;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int s = r * 32 + lane;
            const bool ok = key[k][r] < cut;
            nsel[k] += __popc(__ballot_sync(0xffffffffu, ok));
            okm |= ok ? 1u << (k * NR + r) : 0u;
            TO v[DS];
#pragma unroll
            for (int dd = 0; dd < DS; ++dd) { v[dd] = (TO)0; c[k][r][dd] = 0.f; }
            if (ok) {
                if (sizeof(TO) == 4) {
                    v[0] = (TO)__uint_as_float(rv[k][r][0].x); v[1] = (TO)__uint_as_float(rv[k][r][0].y); v[2] = (TO)__uint_as_float(rv[k][r][0].z);
                    if (DS == 4) v[DS - 1] = (TO)__uint_as_float(rv[k][r][0].w);
                } else {
                    const uint4 &u0 = rv[k][r][0], &u1 = rv[k][r][rec16 - 1];
                    v[0] = (TO)__hiloint2double((int)u0.y, (int)u0.x); v[1] = (TO)__hiloint2double((int)u0.w, (int)u0.z);
                    v[2] = (TO)__hiloint2double((int)u1.y, (int)u1.x);
                    if (DS == 4) v[DS - 1] = (TO)__hiloint2double((int)u1.w, (int)u1.z);
                }
#pragma unroll
                for (int dd = 0; dd < DS; ++dd) c[k][r][dd] = (float)v[dd];
                if (point_slot) {
                    const unsigned cp = key[k][r];
                    const int64_t orig = (int64_t)(cp & ~(unsigned)(kChunk - 1)) + (__ldg(&ctag[f0 + cp]) >> 16);
                    point_slot[f0 + orig] = (r0 + lr0 + k) * P + s;
                }
            }
            if (vrow && s < P && (k == 0 || two)) {  // lanes of empty slots write the padding
#pragma unroll
                for (int dd = 0; dd < DS; ++dd) __stcs(&vrow[(size_t)r * 32 * D + dd], v[dd]);
            }
            sx[k] += c[k][r][0]; sy[k] += c[k][r][1]; sz[k] += c[k][r][2];
        }
    }
    // ---- 2. half-warp h takes pillar h: its sums, then the per-pillar values once for both pillars
    const int h = lane >> 4;
    float mx, my, mz;
    {
        const float ax = h ? sx[0] : sx[1], ay = h ? sy[0] : sy[1], az = h ? sz[0] : sz[1];
        mx = (h ? sx[1] : sx[0]) + __shfl_xor_sync(0xffffffffu, ax, 16);
        my = (h ? sy[1] : sy[0]) + __shfl_xor_sync(0xffffffffu, ay, 16);
        mz = (h ? sz[1] : sz[0]) + __shfl_xor_sync(0xffffffffu, az, 16);
#pragma unroll
        for (int o = 8; o; o >>= 1) {
            mx += __shfl_xor_sync(0xffffffffu, mx, o);
            my += __shfl_xor_sync(0xffffffffu, my, o);
            mz += __shfl_xor_sync(0xffffffffu, mz, o);
        }
    }
    const int n_mine = h ? nsel[1] : nsel[0];
    const int cell = (int)(h ? cellk[1] : cellk[0]);
    const float nf = (float)n_mine;
    mx = __fdiv_rn(mx, nf); my = __fdiv_rn(my, nf); mz = __fdiv_rn(mz, nf);
    const int cz = p.div_nxny.div(cell);
    const int rem = cell - cz * p.grid[0] * p.grid[1];
    const int cy = p.div_nx.div(rem), cx = rem - cy * p.grid[0];
    const float ex = __fadd_rn(__fmul_rn((float)cx, p.vx), p.x_off);
    const float ey = __fadd_rn(__fmul_rn((float)cy, p.vy), p.y_off);
    if ((lane & 15) == 0 && (h == 0 || two)) {
        const int64_t row = row0 + h;
        num_points[row] = n_mine;
        int* co = coors + row * coors_cols;
        if (coors_cols == 4) {
            *reinterpret_cast<int4*>(co) = p.reverse_index ? make_int4(b, cz, cy, cx) : make_int4(b, cx, cy, cz);
        } else if (p.reverse_index) { co[0] = cz; co[1] = cy; co[2] = cx; }
        else { co[0] = cx; co[1] = cy; co[2] = cz; }
        if (cell_voxel) cell_voxel[(size_t)b * p.ncell + cell] = (int)row;
    }
    // ---- 3. decorated rows
    if (decorated) {
        const bool vec = DS == 3 && (reinterpret_cast<uintptr_t>(decorated) & 31u) == 0 && ((P * Do) & 7) == 0;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (k == 1 && !two) break;
            const float kmx = __shfl_sync(0xffffffffu, mx, k * 16), kmy = __shfl_sync(0xffffffffu, my, k * 16);
            const float kmz = __shfl_sync(0xffffffffu, mz, k * 16);
            const float kex = __shfl_sync(0xffffffffu, ex, k * 16), key_ = __shfl_sync(0xffffffffu, ey, k * 16);
            float* drow = decorated + ((row0 + k) * (int64_t)P + lane) * Do;
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                if (r * 32 + lane >= P) continue;
                const bool ok = (okm >> (k * NR + r)) & 1u;
                const float x = c[k][r][0], y = c[k][r][1], z = c[k][r][2];
                float o[Do];
#pragma unroll
                for (int dd = 0; dd < DS; ++dd) o[dd] = c[k][r][dd];
                o[D] = ok ? x - kmx : 0.f; o[D + 1] = ok ? y - kmy : 0.f; o[D + 2] = ok ? z - kmz : 0.f;
                o[D + 3] = ok ? x - kex : 0.f; o[D + 4] = ok ? y - key_ : 0.f;
                float* dp = drow + (size_t)r * 32 * Do;
                if (vec) {
                    st_global_256_cs(dp, make_float4(o[0], o[1], o[2], o[3]), make_float4(o[4], o[5], o[6], o[Do - 1]));
                } else {
#pragma unroll
                    for (int dd = 0; dd < Do; ++dd) __stcs(&dp[dd], o[dd]);
                }
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
    unsigned char* snap;      // [B*S][kSub-1][ncellp] chunk counts at the quarter boundaries, saturated
    int* subv;                // [B*S][kSub] records of the chunk before each quarter
    int* nvalid;              // [B*S] records per chunk
    int* newcount;            // [B*S] voxels opened per chunk
    unsigned* rowinfo;        // [B*ncell] voxel id in frame -> cell
    int* cutoff;              // [B] break position (record position in frame) or kNoCut
    int* done_counter;        // [1]
    unsigned* sidx;           // [B][ncell][max_points + 1] slot table -> record position in frame; last word: min(points, max_points)
    size_t total;
};

static int rec_bytes_of(int D, int out_dtype) { return (int)(((size_t)D * (out_dtype == PP_F64 ? 8 : 4) + 15) / 16 * 16); }
static int chunks_of(int64_t max_frame_points) { return (int)(max_frame_points > 0 ? ceil_div(max_frame_points, kChunk) : 1); }

// n_chunks: frames x chunks of the largest frame (the per-chunk arrays are indexed [frame * S + chunk])
static SmallWs carve_small(void* ws, int64_t ncell, int n_frames, int64_t n_chunks, int64_t total_points, int D,
                           int out_dtype, int P) {
    SmallWs w;
    Carver c(ws);
    const size_t ncellp = (size_t)align_up((size_t)ncell, 16);
    const size_t rb = (size_t)rec_bytes_of(D, out_dtype);
    w.crec = reinterpret_cast<uint4*>(c.take<unsigned char>(((size_t)total_points + 1) * rb));
    w.ctag = c.take<unsigned>((size_t)total_points + 1);
    w.hist = c.take<unsigned short>((size_t)n_chunks * ncellp);
    w.base8 = c.take<unsigned char>((size_t)n_chunks * ncellp);
    w.snap = c.take<unsigned char>((size_t)n_chunks * (kSub - 1) * ncellp + 16);
    w.subv = c.take<int>((size_t)n_chunks * kSub);
    w.nvalid = c.take<int>((size_t)n_chunks);
    w.newcount = c.take<int>((size_t)n_chunks);
    w.rowinfo = c.take<unsigned>((size_t)n_frames * ncell);
    w.cutoff = c.take<int>(n_frames);
    w.done_counter = c.take<int>(1);
    w.sidx = c.take<unsigned>((size_t)n_frames * ncell * (size_t)(P + 1));
    w.total = c.used();
    return w;
}

bool eligible(const pp_voxel_cfg* cfg, int64_t ncell, int n_frames, int64_t total_points,
                        int64_t max_frame_points, int D) {
    if (ncell > kMaxCellsSmall || D > 4 || max_frame_points > kMaxFramePointsSmall || cfg->max_points > kMaxPointsSmall)
        return false;
    // the per-chunk tables must stay small next to the points themselves (callers that do not know the largest
    // frame pass the batch total, which sizes one table set per 16 384 points for every frame)
    const double ncellp = (double)align_up((size_t)ncell, 16);
    const double tables = (double)n_frames * chunks_of(max_frame_points) * ncellp * (3.0 + kSub - 1) +
                          (double)n_frames * (double)ncell * (cfg->max_points * 4.0 + 8.0);
    const double budget = 64.0 * (double)total_points + (double)(256 << 20);
    return tables <= budget;
}

size_t workspace_bytes_of(const pp_voxel_cfg* cfg, int64_t ncell, int n_frames, int64_t n_chunks, int64_t total_points, int D,
                          int out_dtype) {
    return carve_small(nullptr, ncell, n_frames, n_chunks, total_points, D, out_dtype, cfg->max_points).total;
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
        reinterpret_cast<unsigned*>(w.hist), w.nvalid, w.newcount, w.cutoff, point_slot, w.done_counter,
        reinterpret_cast<unsigned*>(w.snap), w.subv);
    PP_LAUNCHED();
    return PP_OK;
}

template <typename TO, int DS, int NR>
static int launch_finish(const VoxParams& p, const SmallWs& w, const int64_t* frame_off, int S, int n_frames,
                         int64_t rows_per_frame, const int32_t* voxel_num, const int32_t* voxel_base, int64_t cap_rows,
                         void* voxels, float* decorated, int32_t* coors, int coors_cols, int32_t* num_points,
                         int32_t* point_slot, int32_t* cell_voxel, cudaStream_t st) {
    const dim3 g((unsigned)ceil_div(rows_per_frame, kFinishRows), (unsigned)n_frames);
    PP_TIMED("vox_finish", st);
    vox_finish_kernel<TO, DS, NR><<<g, kFinishWarps * 32, 0, st>>>(
        frame_off, p, w.rowinfo, voxel_num, voxel_base, w.cutoff, cap_rows, w.sidx, w.crec, w.ctag, w.nvalid, S,
        static_cast<TO*>(voxels), decorated, coors, coors_cols, num_points, point_slot, cell_voxel);
    PP_LAUNCHED();
    return PP_OK;
}

int run(const pp_voxel_cfg* cfg, const VoxParams& p, const void* points, int point_dtype,
                  const int64_t* frame_offsets, int n_frames, int64_t total_points, int64_t max_frame_points,
                  int out_dtype, void* voxels, float* decorated, int32_t* coors, int coors_cols, int32_t* num_points,
                  int64_t cap_rows, int32_t* voxel_num, int32_t* voxel_base, int32_t* point_slot, int32_t* cell_voxel,
                  void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const int D = p.D, P = p.max_points;
    const int64_t ncell = p.ncell;
    const int ncellp = (int)align_up((size_t)ncell, 16);
    const int S = chunks_of(max_frame_points);
    const SmallWs w = carve_small(workspace, ncell, n_frames, (int64_t)n_frames * S, total_points, D, out_dtype, P);
    if (w.total > workspace_bytes) {
        set_error("pp_voxelize_dev: workspace %zu < required %zu", workspace_bytes, w.total);
        return PP_E_WORKSPACE;
    }
    const bool fast = !cfg->arith_f32 && p.grid[0] <= 2047 && p.grid[1] <= 2047 && p.grid[2] <= 2047;

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
        const dim3 g((unsigned)ceil_div(ncell, kPrefixCells), (unsigned)n_frames);
        PP_TIMED("vox_prefix", st);
        vox_prefix_kernel<<<g, kPrefixThreads, 0, st>>>(frame_offsets, S, (int)ncell, ncellp, P, w.hist, w.base8, w.sidx,
                                                        w.newcount, cfg->max_voxels, n_frames, voxel_num, voxel_base,
                                                        w.done_counter, cell_voxel);
        PP_LAUNCHED();
    }
    if (cap_rows <= 0) return PP_OK;
    {
        const size_t smem = (size_t)ncellp;
        int per_sm = 0;
        PP_TRY_RC(kernel_config(reinterpret_cast<const void*>(vox_place_kernel), kPlaceThreads, smem, &per_sm));
        const dim3 g((unsigned)(S * kSub), (unsigned)n_frames);
        PP_TIMED("vox_place", st);
        // helper warps for the prologue: always in the batch instance (64 frames: 160 -> 153 us), in the short-batch instance
        // while the launch leaves SMs idle (one frame: 20 -> 15 us; 1 600 walks of four frames: 21 -> 23 us with them)
        const int pthreads = (kChunkShift >= 14 || (int64_t)S * kSub * n_frames <= 4 * (int64_t)num_sms()) ? kPlaceThreads : 32;
        vox_place_kernel<<<g, pthreads, smem, st>>>(frame_offsets, S, (int)ncell, ncellp, P, cfg->max_voxels, w.ctag, w.base8, w.nvalid,
                                              w.newcount, w.sidx, w.rowinfo, w.cutoff, w.snap, w.subv);
        PP_LAUNCHED();
    }
    const int64_t rows_per_frame = cfg->max_voxels < ncell ? cfg->max_voxels : ncell;
    if (rows_per_frame <= 0) return PP_OK;
#define PP_FINISH(TO, DS, NR)                                                                                           \
    launch_finish<TO, DS, NR>(p, w, frame_offsets, S, n_frames, rows_per_frame, voxel_num, voxel_base, cap_rows, voxels, \
                              decorated, coors, coors_cols, num_points, point_slot, cell_voxel, st)
#define PP_FINISH_P(TO, DS) (P <= 64 ? PP_FINISH(TO, DS, 2) : P <= 128 ? PP_FINISH(TO, DS, 4) : PP_FINISH(TO, DS, 8))
    if (out_dtype == PP_F64) rc = D == 3 ? PP_FINISH_P(double, 3) : PP_FINISH_P(double, 4);
    else rc = D == 3 ? PP_FINISH_P(float, 3) : PP_FINISH_P(float, 4);
#undef PP_FINISH_P
#undef PP_FINISH
    return rc;
}

}  // namespace PP_VS_NS
}  // namespace pp
