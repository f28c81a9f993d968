// Batched NMS on sm_100a: score ordering, pairwise suppression bitmask, greedy sweep.
//
//   standup kind  nms / nms_gpu / nms_kernel / nms_postprocess,
//                 libraries/eval_helper_functions.py:463-598 of the reference
//   rotated kind  rotate_nms_gpu / rotate_nms_kernel, second/core/non_max_suppression/nms_gpu.py:419-490
//   IoU matrix    rotate_iou_gpu(_eval), nms_gpu.py:493-653
//
// The reference sorts on the host, runs a 64-thread block per 64x64 tile (every tile, both
// triangles, corners recomputed per pair), copies the N*N/64 mask to the host and sweeps there.
// Here everything stays on the device, one frame per grid.z / per block:
//   order   top-k radix select + bitonic sort (pre_max_size <= 1024), else a stable LSD radix sort;
//           total order = descending score, ties by descending index.
//   prep    per sorted box: corners, area, hull (rotated) once per box.
//   mask    upper-triangle tiles only, 64 rows x 32 col-blocks per CTA, hull pre-reject, candidate
//           bits first then polygon clips, rows of the tile stored as 256-byte runs.
//   sweep   one CTA per frame, removed-bitmap in shared memory, early exit at post_max_size.
#include <cooperative_groups.h>
#include <type_traits>

#include "box_math.cuh"
#include "nms_common.cuh"

namespace pp {


__global__ void __launch_bounds__(kSortThreads)
nms_topk_kernel(const float* __restrict__ scores, const int* __restrict__ n_valid, int64_t N, int k,
                int* __restrict__ order, int64_t order_stride, int* __restrict__ n_sorted) {
    __shared__ unsigned long long skey[kSelectMaxK];
    const int b = blockIdx.x;
    const float* sc = scores + (int64_t)b * N;
    const int nv = n_valid ? max(0, min(n_valid[b], (int)N)) : (int)N;
    const int kk = block_topk(sc, nv, k, skey);
    if (threadIdx.x == 0) n_sorted[b] = kk;
    for (int i = threadIdx.x; i < kk; i += kSortThreads)
        order[(int64_t)b * order_stride + i] = (int)(skey[i] & 0xffffffffu);
}

// ---------------------------------------------------------------------------------------------
// Top-k of a long score list (KITTI: 107 k anchors per frame) by a thread-block CLUSTER: the frame is
// split over kTopkCluster CTAs, every CTA histograms its slice in its own shared memory, and after a
// cluster barrier each CTA sums the eight histograms through distributed shared memory and picks the
// digit -- all CTAs reach the same decision, so nothing has to be broadcast back.  The selected
// (key, index) pairs are appended to the rank-0 CTA's shared-memory list with DSMEM atomics and sorted
// there.  Same total order and tie rule as block_topk (descending score, then descending index).
constexpr int kTopkCluster = 8;
constexpr int64_t kLongScores = kLongScoreList;  // score lists from this length on are selected by the cluster kernel
// 256 threads per CTA: at 64 registers a 1024-thread CTA owns a whole SM's register file, i.e. 16 clusters in flight
// on the chip and four waves for 64 frames (198 us); with 256 threads 74 clusters are resident and 64 frames are one wave
#ifndef PP_TOPK_THREADS
#define PP_TOPK_THREADS 256
#endif
constexpr int kTopkThreads = PP_TOPK_THREADS;

template <int kT>
__global__ void __cluster_dims__(kTopkCluster, 1, 1) __launch_bounds__(kT)
nms_topk_cluster_kernel(const float* __restrict__ scores, const int* __restrict__ n_valid, int64_t N, int k,
                        int* __restrict__ order, int64_t order_stride, int* __restrict__ n_sorted) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ unsigned long long skey[kSelectMaxK];   // used on rank 0
    __shared__ unsigned hist[256];                      // this CTA's slice
    __shared__ unsigned tot[256];                       // cluster-wide
    __shared__ unsigned s_prefix, s_remaining, s_count, s_scratch, s_fill, s_present;
    const unsigned rank = cluster.block_rank();
    const int b = blockIdx.x / kTopkCluster;
    const float* sc = scores + (int64_t)b * N;
    const int nv = n_valid ? max(0, min(n_valid[b], (int)N)) : (int)N;
    const int per = (nv + kTopkCluster - 1) / kTopkCluster;
    const int lo = min(nv, (int)rank * per), hi = min(nv, lo + per);

    // cluster-wide histogram step: local -> DSMEM sum -> digit choice (identical in every CTA)
    auto reduce_and_select = [&](unsigned prefix, int shift, unsigned* cnt_out) {
        cluster.sync();  // every slice histogram is complete
        if (threadIdx.x < 256) {
            unsigned t = 0;
#pragma unroll
            for (int r = 0; r < kTopkCluster; ++r) t += cluster.map_shared_rank(hist, r)[threadIdx.x];
            tot[threadIdx.x] = t;
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            const unsigned rem = s_remaining;
            __syncwarp();
            select_digit(tot, rem, prefix, shift, &s_prefix, &s_remaining, cnt_out);
        }
        cluster.sync();  // remote reads done before any CTA clears its histogram again
    };

    // present scores (not -inf) over the whole frame
    if (threadIdx.x == 0) s_present = 0;
    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    __syncthreads();
    {
        int c = 0;
#pragma unroll 4
        for (int i = lo + threadIdx.x; i < hi; i += kT) c += sc[i] != -INFINITY;
#pragma unroll
        for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane_id() == 0 && c) atomicAdd(&hist[0], (unsigned)c);
    }
    __syncthreads();
    cluster.sync();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int r = 0; r < kTopkCluster; ++r) t += cluster.map_shared_rank(hist, r)[0];
        s_present = t;
    }
    __syncthreads();
    cluster.sync();
    const int kk = min(k, (int)s_present);
    if (kk == 0) {  // uniform over the cluster
        if (rank == 0 && threadIdx.x == 0) n_sorted[b] = 0;
        return;
    }

    unsigned T = 0, Tidx = 0;
    if (kk < nv) {
        if (threadIdx.x == 0) { s_prefix = 0; s_remaining = (unsigned)kk; }
        for (int shift = 24; shift >= 0; shift -= 8) {
            if (threadIdx.x < 256) hist[threadIdx.x] = 0;
            __syncthreads();
            const unsigned prefix = s_prefix;
            // four independent loads per trip: the scan is bound by the L2 latency of its one dependent load otherwise
            for (int i0 = lo; i0 < hi; i0 += 4 * kT) {
                float v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * kT + threadIdx.x;
                    v[u] = i < hi ? sc[i] : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * kT + threadIdx.x;
                    int d = -1;
                    if (i < hi) {
                        const unsigned key = score_key(v[u]);
                        if (shift == 24 || ((key ^ prefix) >> (shift + 8)) == 0) d = (int)((key >> shift) & 255u);
                    }
                    const unsigned peers = __match_any_sync(0xffffffffu, d);
                    if (d >= 0 && (int)lane_id() == __ffs(peers) - 1) atomicAdd(&hist[d], (unsigned)__popc(peers));
                }
            }
            __syncthreads();
            reduce_and_select(prefix, shift, &s_count);
        }
        T = s_prefix;
        const unsigned need_eq = s_remaining, have_eq = s_count;
        __syncthreads();
        if (have_eq > need_eq) {  // uniform over the cluster
            // among keys == T keep the need_eq largest indices (tie rule: descending index)
            if (threadIdx.x == 0) { s_prefix = 0; s_remaining = need_eq; }
            for (int shift = 24; shift >= 0; shift -= 8) {
                if (threadIdx.x < 256) hist[threadIdx.x] = 0;
                __syncthreads();
                const unsigned prefix = s_prefix;
                for (int i = lo + threadIdx.x; i < hi; i += kT) {
                    if (score_key(sc[i]) != T) continue;
                    const unsigned key = (unsigned)i;
                    if (shift == 24 || ((key ^ prefix) >> (shift + 8)) == 0) atomicAdd(&hist[(key >> shift) & 255u], 1u);
                }
                __syncthreads();
                reduce_and_select(prefix, shift, &s_scratch);
            }
            Tidx = s_prefix;
            __syncthreads();
        }
    }
    // compaction into rank 0's list (arbitrary order), then bitonic sort there
    int np2 = 1;
    while (np2 < kk) np2 <<= 1;
    if (rank == 0) {
        if (threadIdx.x == 0) s_fill = 0;
        for (int i = threadIdx.x; i < np2; i += kT) skey[i] = 0ull;
    }
    cluster.sync();
    {
        unsigned long long* rkey = cluster.map_shared_rank(skey, 0);
        unsigned* rfill = cluster.map_shared_rank(&s_fill, 0);
        for (int i0 = lo; i0 < hi; i0 += 4 * kT) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * kT + threadIdx.x;
                v[u] = i < hi ? sc[i] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * kT + threadIdx.x;
                if (i >= hi) continue;
                const unsigned key = score_key(v[u]);
                if (kk == nv || key > T || (key == T && (unsigned)i >= Tidx)) {
                    const unsigned pos = atomicAdd(rfill, 1u);
                    if (pos < (unsigned)kSelectMaxK) rkey[pos] = ((unsigned long long)key << 32) | (unsigned)i;
                }
            }
        }
    }
    cluster.sync();
    if (rank != 0) return;
    for (int size = 2; size <= np2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < (np2 >> 1); t += kT) {
                const int l = 2 * t - (t & (stride - 1));
                const int h = l + stride;
                const bool desc = ((l & size) == 0);
                const unsigned long long a = skey[l], c = skey[h];
                if ((a < c) == desc) { skey[l] = c; skey[h] = a; }
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) n_sorted[b] = kk;
    for (int i = threadIdx.x; i < kk; i += kT)
        order[(int64_t)b * order_stride + i] = (int)(skey[i] & 0xffffffffu);
}

// 256-thread CTAs keep 64 clusters in one wave; a few frames have the SMs to themselves and take 1024-thread CTAs
static int nms_topk_cluster_launch(const float* scores, const int* n_valid, int B, int64_t N, int k, int* order, int64_t order_stride,
                                   int* n_sorted, cudaStream_t st) {
    if ((int64_t)B * kTopkCluster <= num_sms())
        nms_topk_cluster_kernel<1024><<<B * kTopkCluster, 1024, 0, st>>>(scores, n_valid, N, k, order, order_stride, n_sorted);
    else
        nms_topk_cluster_kernel<kTopkThreads><<<B * kTopkCluster, kTopkThreads, 0, st>>>(scores, n_valid, N, k, order, order_stride, n_sorted);
    PP_LAUNCHED();
    return PP_OK;
}

// ---------------------------------------------------------------------------------------------
// Full stable LSD radix sort per frame (one CTA), ascending key / ascending index, read reversed.
__global__ void __launch_bounds__(kSortThreads)
nms_sort_kernel(const float* __restrict__ scores, const int* __restrict__ n_valid, int64_t N, int limit,
                unsigned* __restrict__ kbuf /*[B][2][N]*/, int* __restrict__ ibuf /*[B][2][N]*/,
                int* __restrict__ order, int64_t order_stride, int* __restrict__ n_sorted) {
    __shared__ unsigned hist[256];
    __shared__ unsigned base[256];
    __shared__ unsigned short wcount[32][256];
    __shared__ int sm[33];
    const int b = blockIdx.x;
    const float* sc = scores + (int64_t)b * N;
    const int nv = n_valid ? max(0, min(n_valid[b], (int)N)) : (int)N;
    const int present = block_count_present(sc, nv);
    const int n = limit > 0 ? min(limit, present) : present;
    if (threadIdx.x == 0) n_sorted[b] = n;
    if (n == 0) return;
    unsigned* k0 = kbuf + (int64_t)b * 2 * N; unsigned* k1 = k0 + N;
    int* i0 = ibuf + (int64_t)b * 2 * N; int* i1 = i0 + N;
    const int lane = lane_id(), w = threadIdx.x >> 5;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 8 * pass;
        const unsigned* kin = (pass & 1) ? k1 : k0; unsigned* kout = (pass & 1) ? k0 : k1;
        const int* iin = (pass & 1) ? i1 : i0; int* iout = (pass & 1) ? i0 : i1;
        if (threadIdx.x < 256) hist[threadIdx.x] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < nv; i += kSortThreads) {
            const unsigned key = pass == 0 ? score_key(sc[i]) : kin[i];
            atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        {
            int tot;
            const int v = threadIdx.x < 256 ? (int)hist[threadIdx.x] : 0;
            const int ex = block_excl_scan(v, &tot, sm);
            if (threadIdx.x < 256) base[threadIdx.x] = (unsigned)ex;
        }
        __syncthreads();
        for (int t0 = 0; t0 < nv; t0 += kSortThreads) {
            for (int q = threadIdx.x; q < 32 * 256; q += kSortThreads) (&wcount[0][0])[q] = 0;
            __syncthreads();
            const int i = t0 + threadIdx.x;
            const bool valid = i < nv;
            unsigned key = 0; int idx = 0;
            if (valid) { key = pass == 0 ? score_key(sc[i]) : kin[i]; idx = pass == 0 ? i : iin[i]; }
            const int d = valid ? (int)((key >> shift) & 255u) : 256;
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            const int rank = __popc(peers & lanemask_lt());
            if (valid && rank == 0) wcount[w][d] = (unsigned short)__popc(peers);
            __syncthreads();
            if (threadIdx.x < 256) {
                unsigned run = 0;
                for (int ww = 0; ww < 32; ++ww) {
                    const unsigned c = wcount[ww][threadIdx.x];
                    wcount[ww][threadIdx.x] = (unsigned short)run;
                    run += c;
                }
                hist[threadIdx.x] = run;  // tile total for this digit
            }
            __syncthreads();
            if (valid) {
                const unsigned pos = base[d] + wcount[w][d] + rank;
                kout[pos] = key;
                iout[pos] = idx;
            }
            __syncthreads();
            if (threadIdx.x < 256) base[threadIdx.x] += hist[threadIdx.x];
            __syncthreads();
        }
        (void)lane;
    }
    // after 4 passes the result is back in buffer 0
    for (int i = threadIdx.x; i < n; i += kSortThreads)
        order[(int64_t)b * order_stride + i] = i0[nv - 1 - i];
}


// Where NMS reads a box from.  anchors == nullptr: `boxes` holds boxes (stride floats per row: 4 standup, 5 BEV,
// 7 decoded).  anchors != nullptr: `boxes` holds box ENCODINGS [B,N,7] and the box is decoded on the fly
// (second_box_decode, then BEV columns 0,1,3,4,6 / their standup box): only boxes that reach NMS are ever decoded,
// the order of the reference's live path (top-k at model/voxelnet.py:1207, decode at 1227, NMS at 1259).
struct BoxSrc {
    const float* boxes;
    int stride;
    const float* anchors;
    int64_t period;  // > 0: anchors hold `period` rows reused cyclically over the batch
};

__device__ __forceinline__ void src_decoded(const BoxSrc& s, int64_t row, float* d /*7*/) {
    const int64_t ar = s.period > 0 ? row % s.period : row;
    box_decode_one(s.boxes + row * 7, s.anchors + ar * 7, d);
}
__device__ __forceinline__ void src_bev(const BoxSrc& s, int64_t row, float* r /*x,y,w,l,angle*/) {
    if (s.anchors) {
        float d[7];
        src_decoded(s, row, d);
        r[0] = d[0]; r[1] = d[1]; r[2] = d[3]; r[3] = d[4]; r[4] = d[6];
        return;
    }
    const float* src = s.boxes + row * s.stride;
    const bool dec7 = s.stride == 7;  // decoded boxes (x,y,z,w,l,h,r) -> BEV columns 0,1,3,4,6 (model/voxelnet.py:1233)
    r[0] = src[0]; r[1] = src[1]; r[2] = dec7 ? src[3] : src[2]; r[3] = dec7 ? src[4] : src[3]; r[4] = dec7 ? src[6] : src[4];
}
__device__ __forceinline__ float4 src_standup(const BoxSrc& s, int64_t row) {
    if (s.anchors) {
        float d[7];
        src_decoded(s, row, d);
        return rbox_standup_one(d[0], d[1], d[3], d[4], d[6]);  // model/voxelnet.py:1233-1249
    }
    const float* src = s.boxes + row * s.stride;
    return make_float4(src[0], src[1], src[2], src[3]);
}

// gather boxes into score order; rotated: corners/area/hull once per box
template <bool ROTATED>
__global__ void __launch_bounds__(256)
nms_prep_kernel(BoxSrc bs, int64_t N, const int* __restrict__ order,
                int64_t order_stride, const int* __restrict__ n_sorted, void* __restrict__ sorted,
                int64_t sorted_stride) {
    const int b = blockIdx.y;
    const int n = n_sorted[b];
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int64_t row = (int64_t)b * N + order[(int64_t)b * order_stride + i];
    if (ROTATED) {
        float r[5];
        src_bev(bs, row, r);
        RBox rb;
        rbox_prepare(r, rb);
        float4* dst = reinterpret_cast<float4*>(static_cast<RBoxG*>(sorted) + (int64_t)b * sorted_stride + i);
        dst[0] = make_float4(rb.c[0], rb.c[1], rb.c[2], rb.c[3]);
        dst[1] = make_float4(rb.c[4], rb.c[5], rb.c[6], rb.c[7]);
        dst[2] = make_float4(rb.area, rb.mnx, rb.mny, rb.mxx);
        dst[3] = make_float4(rb.mxy, 0.f, 0.f, 0.f);
    } else {
        static_cast<float4*>(sorted)[(int64_t)b * sorted_stride + i] = src_standup(bs, row);
    }
}


constexpr int kMaskGroup = 32;  // col blocks per CTA
constexpr int kMaskQ = 4;       // col blocks processed concurrently

template <bool ROTATED>
__global__ void __launch_bounds__(64 * kMaskQ)
nms_mask_kernel(const void* __restrict__ sorted, int64_t sorted_stride, const int* __restrict__ n_sorted,
                float thresh, unsigned long long* __restrict__ mask, int64_t mask_stride /*words per frame*/) {
    using BoxG = typename std::conditional<ROTATED, RBoxG, float4>::type;
    __shared__ BoxG s_col[kMaskQ * 64];
    __shared__ unsigned long long s_tile[64][kMaskGroup + 1];
    const int b = blockIdx.z, rb = blockIdx.y, grp = blockIdx.x;
    const int n = n_sorted[b];
    const int cb = (n + 63) >> 6;
    if (rb >= cb) return;
    const int cb0 = grp * kMaskGroup;
    if (cb0 >= cb || cb0 + kMaskGroup - 1 < rb) return;  // nothing of the upper triangle here
    const int r = threadIdx.x & 63, q = threadIdx.x >> 6;
    const int row = rb * 64 + r;
    const BoxG* sb = static_cast<const BoxG*>(sorted) + (int64_t)b * sorted_stride;
    const double th = (double)thresh;
    const float need_frac = thresh / (1.f + thresh);

    RBox rrow; float4 frow = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < n) {
        if constexpr (ROTATED) load_rbox(reinterpret_cast<const RBoxG*>(sb) + row, rrow);
        else frow = reinterpret_cast<const float4*>(sb)[row];
    }
    for (int s = 0; s < kMaskGroup / kMaskQ; ++s) {
        const int cbase = cb0 + s * kMaskQ;  // first col block of this step
        __syncthreads();
        {
            const int col = cbase * 64 + threadIdx.x;
            if (col < n) s_col[threadIdx.x] = sb[col];
        }
        __syncthreads();
        const int cbk = cbase + q;
        unsigned long long word = 0ull;
        if (row < n && cbk < cb && cbk >= rb) {
            const int jn = min(64, n - cbk * 64);
            const int j0 = (cbk == rb) ? r + 1 : 0;
            if constexpr (ROTATED) {
                unsigned long long cand = 0ull;
                for (int j = j0; j < jn; ++j) {
                    const RBoxG& c = s_col[q * 64 + j];
                    const float scale = fmaxf(fmaxf(fabsf(rrow.mxx), fabsf(rrow.mnx)), fmaxf(fabsf(rrow.mxy), fabsf(rrow.mny)));
                    const float eps = 1e-4f * fmaxf(1.f, scale);
                    const bool dis = rrow.mnx > c.mxx + eps || c.mnx > rrow.mxx + eps ||
                                     rrow.mny > c.mxy + eps || c.mny > rrow.mxy + eps;
                    if (!dis) cand |= 1ull << j;
                }
                while (cand) {
                    const int j = __ffsll((long long)cand) - 1;
                    cand &= cand - 1;
                    RBox cbx;
                    load_rbox(&s_col[q * 64 + j], cbx);
                    // a pair whose intersection-area upper bound cannot reach thresh/(1+thresh)*(a1+a2) is not clipped
                    if (rbox_cannot_exceed(rrow, cbx, need_frac)) continue;
                    // devRotateIoU(row box, col box), nms_gpu.py:445-449
                    const double ai = rbox_inter(rrow.c, cbx.c);
                    const double iou = ai / ((double)__fadd_rn(rrow.area, cbx.area) - ai);
                    if (iou > th) word |= 1ull << j;
                }
            } else {
                for (int j = j0; j < jn; ++j)
                    if (standup_iou(frow, reinterpret_cast<const float4*>(s_col)[q * 64 + j]) > th) word |= 1ull << j;
            }
        }
        s_tile[r][s * kMaskQ + q] = word;
    }
    __syncthreads();
    // rows of the tile as contiguous runs of up to 32 words
    unsigned long long* mb = mask + (int64_t)b * mask_stride;
    for (int k = threadIdx.x; k < 64 * kMaskGroup; k += 64 * kMaskQ) {
        const int rr = k / kMaskGroup, cc = k - rr * kMaskGroup;
        const int grow = rb * 64 + rr, gcb = cb0 + cc;
        if (grow < n && gcb < cb && gcb >= rb) mb[(int64_t)grow * cb + gcb] = s_tile[rr][cc];
    }
}

// Greedy sweep (nms_postprocess, eval_helper_functions.py:529-546): one CTA per frame.
constexpr int kSweepThreads = 512;
__global__ void __launch_bounds__(kSweepThreads)
nms_sweep_kernel(const unsigned long long* __restrict__ mask, int64_t mask_stride,
                 const int* __restrict__ n_sorted, const int* __restrict__ order, int64_t order_stride,
                 int post_max, int* __restrict__ keep, int64_t keep_stride, int* __restrict__ keep_count) {
    extern __shared__ unsigned long long remv[];
    __shared__ unsigned long long s_diag[64];
    __shared__ unsigned long long s_kept;
    __shared__ int s_nkeep;
    const int b = blockIdx.x;
    const int n = n_sorted[b];
    const int cb = (n + 63) >> 6;
    const unsigned long long* mb = mask + (int64_t)b * mask_stride;
    const int* ord = order + (int64_t)b * order_stride;
    int* kp = keep + (int64_t)b * keep_stride;
    int limit = (int)min((int64_t)(post_max > 0 ? post_max : n), keep_stride);
    for (int j = threadIdx.x; j < cb; j += kSweepThreads) remv[j] = 0ull;
    if (threadIdx.x == 0) s_nkeep = 0;
    __syncthreads();
    for (int k = 0; k < cb; ++k) {
        const int base = k * 64;
        const int cnt = min(64, n - base);
        if (threadIdx.x < 64) s_diag[threadIdx.x] = threadIdx.x < cnt ? mb[(int64_t)(base + threadIdx.x) * cb + k] : 0ull;
        __syncthreads();
        const int nk_before = s_nkeep;
        if (threadIdx.x == 0) {
            // the serial part touches shared memory only (a global load of order[] per kept box sat on this chain)
            unsigned long long rm = remv[k], kept = 0ull;
            int nk = nk_before;
            for (int i = 0; i < cnt && nk < limit; ++i) {
                if (!((rm >> i) & 1ull)) {
                    kept |= 1ull << i;
                    ++nk;
                    rm |= s_diag[i];
                }
            }
            s_kept = kept;
            s_nkeep = nk;
        }
        __syncthreads();
        const unsigned long long kept = s_kept;
        if (threadIdx.x < 64 && ((kept >> threadIdx.x) & 1ull))  // kept box i of this block is output nk_before + rank(i)
            kp[nk_before + __popcll(kept & ((1ull << threadIdx.x) - 1ull))] = ord[base + threadIdx.x];
        if (s_nkeep >= limit) break;
        if (kept) {
            for (int j = k + 1 + threadIdx.x; j < cb; j += kSweepThreads) {
                unsigned long long acc = 0ull, kk = kept;
                while (kk) {
                    const int i = __ffsll((long long)kk) - 1;
                    kk &= kk - 1;
                    acc |= mb[(int64_t)(base + i) * cb + j];
                }
                remv[j] |= acc;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) keep_count[b] = s_nkeep;
}

// Sweep for at most 1024 boxes (16 mask words per row): the frame's whole mask (<= 128 KB) is copied into shared
// memory with one coalesced pass, then one warp walks the rows -- lane w owns removed-word w -- without touching
// global memory again.  (nms_sweep_kernel pays an L2 round trip per 64-row block and more for the propagation.)
constexpr int kSweepSmemWords = 16;
__global__ void __launch_bounds__(512)
nms_sweep_smem_kernel(const unsigned long long* __restrict__ mask, int64_t mask_stride, int cb_stride,
                      const int* __restrict__ n_sorted, const int* __restrict__ order, int64_t order_stride,
                      int post_max, int* __restrict__ keep, int64_t keep_stride, int* __restrict__ keep_count) {
    extern __shared__ unsigned long long s_m[];  // [n][cb_stride] then the kept list
    __shared__ int s_nk;
    const int b = blockIdx.x;
    const int n = n_sorted[b];
    const int cb = (n + 63) >> 6;
    int* s_list = reinterpret_cast<int*>(s_m + (size_t)kSweepSmemWords * 64 * cb_stride);
    const unsigned long long* mb = mask + (int64_t)b * mask_stride;
    // rows are packed with the frame's own word count cb (nms_mask_kernel), not with the capacity
    for (int k = threadIdx.x; k < n * cb; k += 512) s_m[k] = mb[k];
    __syncthreads();
    const int limit = (int)min((int64_t)(post_max > 0 ? post_max : n), keep_stride);
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        unsigned long long rm = 0ull;
        int nk = 0;
        for (int i = 0; i < n && nk < limit; ++i) {
            const unsigned long long wv = __shfl_sync(0xffffffffu, rm, i >> 6);
            if (!((wv >> (i & 63)) & 1ull)) {
                if (lane == 0) s_list[nk] = i;
                ++nk;
                if (lane < cb && lane >= (i >> 6)) rm |= s_m[(size_t)i * cb + lane];
            }
        }
        if (lane == 0) s_nk = nk;
    }
    __syncthreads();
    const int nk = s_nk;
    const int* ord = order + (int64_t)b * order_stride;
    int* kp = keep + (int64_t)b * keep_stride;
    for (int k = threadIdx.x; k < nk; k += 512) kp[k] = ord[s_list[k]];
    if (threadIdx.x == 0) keep_count[b] = nk;
}

// rotate_iou_kernel(_eval), nms_gpu.py:493-523 / 579-615: out[n,k] = f(query k, box n)
__global__ void __launch_bounds__(256)
rotate_iou_matrix_kernel(const float* __restrict__ boxes, int64_t N, const float* __restrict__ qboxes,
                         int64_t K, int criterion, float* __restrict__ out) {
    __shared__ RBoxG s_q[64];
    __shared__ RBoxG s_b[64];
    const int64_t n0 = (int64_t)blockIdx.y * 64, k0 = (int64_t)blockIdx.x * 64;
    if (threadIdx.x < 128) {
        const bool isq = threadIdx.x < 64;
        const int t = threadIdx.x & 63;
        const int64_t g = (isq ? k0 : n0) + t;
        if (g < (isq ? K : N)) {
            const float* src = (isq ? qboxes : boxes) + g * 5;
            const float r[5] = {src[0], src[1], src[2], src[3], src[4]};
            RBox rb;
            rbox_prepare(r, rb);
            RBoxG& d = isq ? s_q[t] : s_b[t];
#pragma unroll
            for (int i = 0; i < 8; ++i) d.c[i] = rb.c[i];
            d.area = rb.area; d.mnx = rb.mnx; d.mny = rb.mny; d.mxx = rb.mxx; d.mxy = rb.mxy;
        }
    }
    __syncthreads();
    // thread -> (box row, 16 consecutive queries): consecutive threads write consecutive k
    for (int e = threadIdx.x; e < 64 * 64; e += 256) {
        const int bn = e >> 6, qk = e & 63;
        if (n0 + bn >= N || k0 + qk >= K) continue;
        RBox q, bx;
        load_rbox(&s_q[qk], q);
        load_rbox(&s_b[bn], bx);
        out[(n0 + bn) * K + k0 + qk] = (float)rbox_iou(q, bx, criterion);
    }
}

// ---------------------------------------------------------------------------------------------
// Alive-stripe NMS for large box counts (> kStripeMin after pre_max_size).
//
// Greedy NMS keeps box j iff no KEPT box of higher score overlaps it.  The all-pairs bitmask of the
// reference (N^2/128 bytes: 1.25 GB at 100 k boxes) spends most of its work on rows of boxes that
// end up suppressed.  Here every box carries a dead flag, all boxes are binned ONCE into a uniform
// grid over their centres (bin edge >= the largest hull, so overlapping boxes sit in the same or in
// adjacent bins), and the score-ordered boxes are consumed in stripes of the next kAlive boxes that
// are still ALIVE:
//   select  per frame: the next <= kAlive alive positions after the frame's cursor
//   mask    upper-triangle bitmask inside the stripe (every row is alive)
//   sweep   greedy sweep of the stripe; kept boxes go to the output and to the new-kept list
//   push    one warp per newly kept box: the boxes of its 3x3 bin neighbourhood that lie after the
//           cursor and are still alive are tested (hull pre-reject, survivors queued per warp so
//           the polygon clips run with full lanes) and flagged dead
// Every (kept, later) pair is looked at once at most; suppressed boxes never enter a stripe, so the
// number of stripes follows the number of boxes alive at their turn (about a third of 100 k random
// boxes), not N.  The result is identical to the all-pairs algorithm (same IoU function, same argument
// order (higher score first), same strict > test).
constexpr int kStripe = 2048;           // mask rows per frame = kMaskGroup * 64
#ifndef PP_NMS_ALIVE
#define PP_NMS_ALIVE 1024
#endif
constexpr int kAlive = PP_NMS_ALIVE;    // alive boxes per stripe (<= kStripe): the in-stripe mask costs kAlive/2 tests per box
constexpr int kStripeMin = 16384;       // use the stripe path above this many boxes per frame
constexpr int kPushWarps = 8;

struct StripeGrid { float ox, oy, inv_s; int gx, gy; };
constexpr int kGridMax = 64;            // bins per axis (bin table: kGridMax^2 + 1 ints per frame)
constexpr int kGridBins = kGridMax * kGridMax;

template <bool ROTATED>
__device__ __forceinline__ void box_centre_extent(const void* sorted, int64_t i, float& cx, float& cy, float& ext) {
    if constexpr (ROTATED) {
        const RBoxG* g = static_cast<const RBoxG*>(sorted) + i;
        cx = 0.5f * (g->mnx + g->mxx); cy = 0.5f * (g->mny + g->mxy);
        ext = fmaxf(g->mxx - g->mnx, g->mxy - g->mny);
    } else {
        const float4 v = static_cast<const float4*>(sorted)[i];
        cx = 0.5f * (v.x + v.z); cy = 0.5f * (v.y + v.w);
        ext = fmaxf(v.z - v.x, v.w - v.y) + 1.f;  // the "+1" convention makes boxes within one unit overlap
    }
}

__device__ __forceinline__ void grid_bin(const StripeGrid& g, float cx, float cy, int& bx, int& by) {
    // NaN / out-of-bounds centres clamp into the grid (NaN boxes can neither suppress nor be suppressed)
    const float fx = (cx - g.ox) * g.inv_s, fy = (cy - g.oy) * g.inv_s;
    bx = fx >= 0.f ? min((int)fx, g.gx - 1) : 0;
    by = fy >= 0.f ? min((int)fy, g.gy - 1) : 0;
}

// one CTA per frame: bounds of the centres and the largest extent -> grid; counting sort of the
// frame's boxes by bin (bin_start[kGridBins+1], bin_items[n] = score-order positions)
template <bool ROTATED>
__global__ void __launch_bounds__(1024)
nms_bin_build_kernel(const void* __restrict__ sorted, int64_t sorted_stride, const int* __restrict__ n_sorted,
                     StripeGrid* __restrict__ grid, int* __restrict__ bin_start, int* __restrict__ bin_items,
                     float4* __restrict__ bin_hull, float* __restrict__ bin_area, int* __restrict__ slot_of) {
    __shared__ float s_red[5][32];
    __shared__ int s_hist[kGridBins];
    __shared__ int sm[33];
    __shared__ StripeGrid s_g;
    const int b = blockIdx.x;
    const int n = n_sorted[b];
    using BoxG = typename std::conditional<ROTATED, RBoxG, float4>::type;
    const BoxG* sb = static_cast<const BoxG*>(sorted) + (int64_t)b * sorted_stride;
    float mnx = 3.0e38f, mny = 3.0e38f, mxx = -3.0e38f, mxy = -3.0e38f, ext = 0.f;
    for (int i = threadIdx.x; i < n; i += 1024) {
        float cx, cy, e;
        box_centre_extent<ROTATED>(sb, i, cx, cy, e);
        if (cx == cx && cy == cy && e == e) {  // NaN boxes never overlap anything: leave them out of the bounds
            mnx = fminf(mnx, cx); mxx = fmaxf(mxx, cx); mny = fminf(mny, cy); mxy = fmaxf(mxy, cy); ext = fmaxf(ext, e);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o)); mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
        ext = fmaxf(ext, __shfl_xor_sync(0xffffffffu, ext, o));
    }
    const int lane = lane_id(), w = threadIdx.x >> 5;
    if (lane == 0) { s_red[0][w] = mnx; s_red[1][w] = mny; s_red[2][w] = mxx; s_red[3][w] = mxy; s_red[4][w] = ext; }
    for (int k = threadIdx.x; k < kGridBins; k += 1024) s_hist[k] = 0;
    __syncthreads();
    if (w == 0) {
        mnx = s_red[0][lane]; mny = s_red[1][lane]; mxx = s_red[2][lane]; mxy = s_red[3][lane]; ext = s_red[4][lane];
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
            mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o)); mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
            ext = fmaxf(ext, __shfl_xor_sync(0xffffffffu, ext, o));
        }
        if (lane == 0) {
            StripeGrid g;
            if (!(mxx >= mnx)) { mnx = mny = 0.f; mxx = mxy = 0.f; }
            const float scale = fmaxf(fmaxf(fabsf(mnx), fabsf(mxx)), fmaxf(fabsf(mny), fabsf(mxy))) + ext;
            // edge: largest extent + the hull test's guard band (1e-4 relative) with a wide margin
            float edge = ext * 1.01f + 1e-3f * fmaxf(1.f, scale);
            edge = fmaxf(edge, fmaxf(mxx - mnx, mxy - mny) / (float)(kGridMax - 2));
            g.ox = mnx; g.oy = mny; g.inv_s = 1.f / edge;
            g.gx = min(kGridMax, (int)((mxx - mnx) * g.inv_s) + 2);
            g.gy = min(kGridMax, (int)((mxy - mny) * g.inv_s) + 2);
            grid[b] = g;
            s_g = g;
        }
    }
    __syncthreads();
    const StripeGrid g = s_g;
    for (int i = threadIdx.x; i < n; i += 1024) {
        float cx, cy, e;
        box_centre_extent<ROTATED>(sb, i, cx, cy, e);
        int bx, by;
        grid_bin(g, cx, cy, bx, by);
        atomicAdd(&s_hist[by * kGridMax + bx], 1);
    }
    __syncthreads();
    // exclusive scan of the histogram, four bins per thread
    int* bs = bin_start + (int64_t)b * (kGridBins + 1);
    {
        int v[4], t = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) { v[q] = s_hist[threadIdx.x * 4 + q]; t += v[q]; }
        int tot;
        int ex = block_excl_scan(t, &tot, sm);
#pragma unroll
        for (int q = 0; q < 4; ++q) { s_hist[threadIdx.x * 4 + q] = ex; bs[threadIdx.x * 4 + q] = ex; ex += v[q]; }
        if (threadIdx.x == 0) bs[kGridBins] = tot;
    }
    __syncthreads();
    // scatter in chunks of 1024 score positions with a barrier in between: inside a bin the items are then
    // ordered by chunk, so the push stage can skip everything before the cursor's chunk with a binary search
    int* items = bin_items + (int64_t)b * sorted_stride;
    float4* hull = bin_hull + (int64_t)b * sorted_stride;
    float* area = bin_area + (int64_t)b * sorted_stride;
    int* slots = slot_of + (int64_t)b * sorted_stride;
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        if (i < n) {
            float cx, cy, e;
            box_centre_extent<ROTATED>(sb, i, cx, cy, e);
            int bx, by;
            grid_bin(g, cx, cy, bx, by);
            const int t = atomicAdd(&s_hist[by * kGridMax + bx], 1);
            items[t] = i;
            slots[i] = t;
            if constexpr (ROTATED) {
                const RBoxG* gb = reinterpret_cast<const RBoxG*>(sb) + i;
                hull[t] = make_float4(gb->mnx, gb->mny, gb->mxx, gb->mxy);
                area[t] = gb->area;
            } else {
                hull[t] = reinterpret_cast<const float4*>(sb)[i];
            }
        }
        __syncthreads();
    }
}

// per frame: the next <= kAlive alive positions at or after the cursor
__global__ void __launch_bounds__(1024)
nms_select_kernel(const int* __restrict__ n_sorted, const unsigned char* __restrict__ dead,
                  const int* __restrict__ slot_of, int64_t sorted_stride, const int* __restrict__ kept_cnt, int limit,
                  int* __restrict__ cursor, int* __restrict__ s_idx, int* __restrict__ s_n) {
    __shared__ int sm[33];
    __shared__ int s_stop;
    const int b = blockIdx.x;
    const int n = n_sorted[b];
    int cur = cursor[b];
    if (kept_cnt[b] >= limit || cur >= n) {
        if (threadIdx.x == 0) { s_n[b] = 0; cursor[b] = n; }
        return;
    }
    const unsigned char* df = dead + (int64_t)b * sorted_stride;   // indexed by bin slot
    const int* slots = slot_of + (int64_t)b * sorted_stride;
    int* out = s_idx + (int64_t)b * kStripe;
    int have = 0;
    while (cur < n && have < kAlive) {
        const int pos = cur + threadIdx.x;
        const int alive = pos < n && !df[slots[pos]];
        int tot;
        const int r = block_excl_scan(alive, &tot, sm);
        if (threadIdx.x == 0) s_stop = min(n, cur + 1024);
        __syncthreads();
        if (alive && have + r < kAlive) {
            out[have + r] = pos;
            if (have + r == kAlive - 1) s_stop = pos + 1;  // the stripe is full: the cursor stops right behind it
        }
        __syncthreads();
        have = min(kAlive, have + tot);
        cur = s_stop;
        __syncthreads();
    }
    if (threadIdx.x == 0) { s_n[b] = have; cursor[b] = cur; }
}

// upper-triangle mask inside one stripe of alive boxes (layout [kStripe][kMaskGroup] words per frame)
template <bool ROTATED>
__global__ void __launch_bounds__(64 * kMaskQ)
nms_stripe_mask_kernel(const void* __restrict__ sorted, int64_t sorted_stride, const int* __restrict__ s_idx,
                       const int* __restrict__ s_n, float thresh, unsigned long long* __restrict__ mask) {
    using BoxG = typename std::conditional<ROTATED, RBoxG, float4>::type;
    __shared__ BoxG s_col[kMaskQ * 64];
    const int b = blockIdx.y, rb = blockIdx.x;
    const int n = s_n[b];
    const int cb = (n + 63) >> 6;
    if (n <= 0 || rb >= cb) return;
    const int r = threadIdx.x & 63, q = threadIdx.x >> 6;
    const int row = rb * 64 + r;
    const BoxG* sb = static_cast<const BoxG*>(sorted) + (int64_t)b * sorted_stride;
    const int* idx = s_idx + (int64_t)b * kStripe;
    unsigned long long* mb = mask + (int64_t)b * kStripe * kMaskGroup;
    const double th = (double)thresh;
    const float need_frac = thresh / (1.f + thresh);
    const bool live = row < n;
    RBox rrow; float4 frow = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
        if constexpr (ROTATED) load_rbox(reinterpret_cast<const RBoxG*>(sb) + idx[row], rrow);
        else frow = reinterpret_cast<const float4*>(sb)[idx[row]];
    }
    for (int cbase = rb & ~(kMaskQ - 1); cbase < cb; cbase += kMaskQ) {
        __syncthreads();
        {
            const int col = cbase * 64 + threadIdx.x;
            if (col < n) s_col[threadIdx.x] = sb[idx[col]];
        }
        __syncthreads();
        const int cbk = cbase + q;
        if (live && cbk < cb && cbk >= rb) {
            unsigned long long word = 0ull;
            const int jn = min(64, n - cbk * 64);
            const int j0 = (cbk == rb) ? r + 1 : 0;
            for (int jj = j0; jj < jn; ++jj) {
                bool sup;
                if constexpr (ROTATED) {
                    const RBoxG& c = s_col[q * 64 + jj];
                    const float scale = fmaxf(fmaxf(fabsf(rrow.mxx), fabsf(rrow.mnx)), fmaxf(fabsf(rrow.mxy), fabsf(rrow.mny)));
                    const float eps = 1e-4f * fmaxf(1.f, scale);
                    if (rrow.mnx > c.mxx + eps || c.mnx > rrow.mxx + eps || rrow.mny > c.mxy + eps || c.mny > rrow.mxy + eps) continue;
                    RBox cbx;
                    load_rbox(&s_col[q * 64 + jj], cbx);
                    if (rbox_cannot_exceed(rrow, cbx, need_frac)) continue;
                    const double ai = rbox_inter(rrow.c, cbx.c);
                    sup = ai / ((double)__fadd_rn(rrow.area, cbx.area) - ai) > th;
                } else {
                    sup = standup_iou(frow, reinterpret_cast<const float4*>(s_col)[q * 64 + jj]) > th;
                }
                if (sup) word |= 1ull << jj;
            }
            mb[(int64_t)row * kMaskGroup + cbk] = word;
        }
    }
}

// greedy sweep of one stripe; kept boxes go to the output and to the new-kept list of the push stage
__global__ void __launch_bounds__(256)
nms_stripe_sweep_kernel(const int* __restrict__ s_idx, const int* __restrict__ s_n, int64_t sorted_stride,
                        const unsigned long long* __restrict__ mask, const int* __restrict__ order, int limit,
                        int* __restrict__ kept_cnt, int* __restrict__ new_kept, int* __restrict__ n_new,
                        int* __restrict__ keep, int64_t keep_stride) {
    __shared__ int s_list[kStripe];
    __shared__ int s_nk;
    const int b = blockIdx.x;
    const int n = s_n[b];
    const int nk0 = kept_cnt[b];
    if (n <= 0 || nk0 >= limit) {
        if (threadIdx.x == 0) n_new[b] = 0;
        return;
    }
    const int cb = (n + 63) >> 6;
    const unsigned long long* mb = mask + (int64_t)b * kStripe * kMaskGroup;
    if (threadIdx.x < 32) {
        // one warp: lane w owns removed-word w; the kept test is broadcast from the owning lane
        const int lane = threadIdx.x;
        unsigned long long rm = 0ull;
        int nk = 0;
        // rows are fetched eight at a time ahead of the (serial) keep test, so the L2 latency of a kept
        // row's mask is not on the dependency chain
        for (int i0 = 0; i0 < n && nk0 + nk < limit; i0 += 8) {
            unsigned long long rows[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                rows[u] = (i0 + u < n && lane < cb && lane >= ((i0 + u) >> 6)) ? mb[(int64_t)(i0 + u) * kMaskGroup + lane] : 0ull;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u;
                if (i < n && nk0 + nk < limit) {
                    const unsigned long long wv = __shfl_sync(0xffffffffu, rm, i >> 6);
                    if (!((wv >> (i & 63)) & 1ull)) {
                        if (lane == 0) s_list[nk] = i;
                        ++nk;
                        rm |= rows[u];
                    }
                }
            }
        }
        if (lane == 0) s_nk = nk;
    }
    __syncthreads();
    const int nk = s_nk;
    const int* idx = s_idx + (int64_t)b * kStripe;
    const int* ord = order + (int64_t)b * sorted_stride;
    int* kp = keep + (int64_t)b * keep_stride;
    int* nkp = new_kept + (int64_t)b * kStripe;
    for (int k = threadIdx.x; k < nk; k += 256) {
        const int pos = idx[s_list[k]];
        nkp[k] = pos;
        if (nk0 + k < keep_stride) kp[nk0 + k] = ord[pos];
    }
    if (threadIdx.x == 0) { kept_cnt[b] = nk0 + nk; n_new[b] = nk; }
}

// push stage: one warp per newly kept box flags the later, still alive boxes it suppresses.  Everything it
// scans (item positions, dead flags, hulls) is stored in bin order, so the scan is coalesced.
#ifndef PP_PUSH_MINBLOCKS
#define PP_PUSH_MINBLOCKS 4
#endif
template <bool ROTATED>
__global__ void __launch_bounds__(kPushWarps * 32, PP_PUSH_MINBLOCKS)
nms_push_kernel(const void* __restrict__ sorted, int64_t sorted_stride, const int* __restrict__ new_kept,
                const int* __restrict__ n_new, const int* __restrict__ cursor, const int* __restrict__ n_sorted,
                const int* __restrict__ kept_cnt, int limit, float thresh, const StripeGrid* __restrict__ grid,
                const int* __restrict__ bin_start, const int* __restrict__ bin_items,
                const float4* __restrict__ bin_hull, const float* __restrict__ bin_area,
                unsigned char* __restrict__ dead) {
    using BoxG = typename std::conditional<ROTATED, RBoxG, float4>::type;
    __shared__ int s_q[kPushWarps][64];
    const int b = blockIdx.y;
    const int lane = lane_id(), w = threadIdx.x >> 5;
    const int k = blockIdx.x * kPushWarps + w;
    const int cur = cursor[b];
    if (k >= n_new[b] || cur >= n_sorted[b] || kept_cnt[b] >= limit) return;  // warp-uniform
    const BoxG* sb = static_cast<const BoxG*>(sorted) + (int64_t)b * sorted_stride;
    const int* bs = bin_start + (int64_t)b * (kGridBins + 1);
    const int* items = bin_items + (int64_t)b * sorted_stride;
    const float4* hull = bin_hull + (int64_t)b * sorted_stride;
    const float* area = bin_area + (int64_t)b * sorted_stride;
    unsigned char* df = dead + (int64_t)b * sorted_stride;
    const StripeGrid g = grid[b];
    const int pos = new_kept[(int64_t)b * kStripe + k];
    RBox me; float4 mef = make_float4(0.f, 0.f, 0.f, 0.f);
    float cx, cy, e;
    box_centre_extent<ROTATED>(sb, pos, cx, cy, e);
    if constexpr (ROTATED) load_rbox(reinterpret_cast<const RBoxG*>(sb) + pos, me);
    else mef = reinterpret_cast<const float4*>(sb)[pos];
    int bx, by;
    grid_bin(g, cx, cy, bx, by);
    const double th = (double)thresh;
    const float need_frac = thresh / (1.f + thresh);
    const int cur_chunk = cur & ~1023;  // items of a bin are ordered by 1024-position chunk (nms_bin_build_kernel)
    int qn = 0;
    int* q = s_q[w];
    float eps = 0.f;
    if constexpr (ROTATED) eps = 1e-4f * fmaxf(1.f, fmaxf(fmaxf(fabsf(me.mxx), fabsf(me.mnx)), fmaxf(fabsf(me.mxy), fabsf(me.mny))));
    auto clip = [&](int t) {  // full test of the candidate in slot t (this kept box has the higher score: first argument)
        RBox c;
        load_rbox(reinterpret_cast<const RBoxG*>(sb) + items[t], c);
        if (rbox_cannot_exceed(me, c, need_frac)) return;
        const double ai = rbox_inter(me.c, c.c);
        if (ai / ((double)__fadd_rn(me.area, c.area) - ai) > th) df[t] = 1;
    };
    for (int dy = -1; dy <= 1; ++dy) {
        const int yy = by + dy;
        if (yy < 0 || yy >= g.gy) continue;
        for (int xx = max(bx - 1, 0); xx <= min(bx + 1, g.gx - 1); ++xx) {
            int t0 = bs[yy * kGridMax + xx];
            const int t1 = bs[yy * kGridMax + xx + 1];
            // first item whose chunk is not before the cursor's
            {
                int lo = t0, hi = t1;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (items[mid] < cur_chunk) lo = mid + 1; else hi = mid;
                }
                t0 = lo;
            }
            for (int tb = t0; tb < t1; tb += 32) {
                const int t = tb + lane;
                bool cand = false;
                if constexpr (ROTATED) {
                    if (t < t1) {
                        // the four loads are independent: one round trip per 32 items
                        const int it = items[t];
                        const unsigned char dd = df[t];
                        const float4 h = hull[t];  // mnx, mny, mxx, mxy
                        const float ar = area[t];
                        cand = it >= cur && !dd &&
                               !(h.x > me.mxx + eps || me.mnx > h.z + eps || h.y > me.mxy + eps || me.mny > h.w + eps);
                        if (cand) {
                            // the candidate lies inside its hull: IoU > thresh needs more intersection than this
                            // kept box has with the hull (measured along the kept box's own axes)
                            const float hc[8] = {h.x, h.y, h.x, h.w, h.z, h.w, h.z, h.y};
                            const float asum = me.area + ar;
                            cand = !(rect_inter_bound(me.c, hc) * 1.002f + 1e-6f * asum < need_frac * asum);
                        }
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, cand);
                    if (cand) q[qn + __popc(bal & lanemask_lt())] = t;
                    qn += __popc(bal);
                    __syncwarp();
                    if (qn >= 32) {
                        clip(q[qn - 32 + lane]);
                        qn -= 32;
                        __syncwarp();
                    }
                } else {
                    if (t < t1) cand = items[t] >= cur && !df[t];
                    if (cand && standup_iou(mef, hull[t]) > th) df[t] = 1;
                }
            }
        }
    }
    if constexpr (ROTATED) {
        if (lane < qn) clip(q[lane]);
    }
}

__global__ void nms_stripe_finish_kernel(const int* __restrict__ kept_cnt, int limit, int64_t keep_stride, int B,
                                         int* __restrict__ keep_count) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) keep_count[b] = (int)min((int64_t)min(kept_cnt[b], limit), keep_stride);
}

// ---------------------------------------------------------------------------------------------
// Whole NMS of one frame in one launch when at most kSmallN boxes survive pre_max_size (the live
// path: pre_max_size = 100, configs/train.yaml:176): top-k, per-box prep, all pairs with the mask in
// shared memory, sweep by one thread.  One launch instead of four and no global intermediates.
// A frame is a thread-block CLUSTER of `csize` CTAs (1, 2, 4 or 8; as many as fit one wave of the batch):
// every CTA selects and prepares the same <= 128 boxes (identical results, no exchange needed), takes every
// csize-th pair of the i < j triangle, and ORs its suppression bits into the mask of CTA 0 through
// distributed shared memory; CTA 0 sweeps.  A single frame on 8 SMs tests one pair per thread instead of five.
constexpr int kSmallN = 128;
template <bool ROTATED>
__global__ void __launch_bounds__(kSortThreads)
nms_small_kernel(BoxSrc bs, const float* __restrict__ scores,
                 const int* __restrict__ n_valid, int64_t N, int k, int post_max, float thresh,
                 int* __restrict__ keep, int64_t keep_stride, int* __restrict__ keep_count, float* __restrict__ dets, int K,
                 const int* __restrict__ order, int64_t order_stride, const int* __restrict__ n_sorted) {
    namespace cg = cooperative_groups;
    using BoxG = typename std::conditional<ROTATED, RBoxG, float4>::type;
    __shared__ unsigned long long skey[kSelectMaxK];
    __shared__ BoxG s_box[kSmallN];
    __shared__ unsigned long long s_mask[kSmallN][2];
    __shared__ int s_kept[kSmallN];
    __shared__ int s_nk;
    cg::cluster_group cluster = cg::this_cluster();
    const int csize = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int b = blockIdx.x / csize;
    const float* sc = scores + (int64_t)b * N;
    const int nv = n_valid ? max(0, min(n_valid[b], (int)N)) : (int)N;
    int n;
    if (order) {
        // long score lists (KITTI: 107 k anchors) are selected by the 8-CTA cluster top-k beforehand: same order, same tie rule
        n = min(n_sorted[b], kSmallN);
        if ((int)threadIdx.x < n) {
            const int a = order[(int64_t)b * order_stride + threadIdx.x];
            skey[threadIdx.x] = ((unsigned long long)score_key(sc[a]) << 32) | (unsigned)a;
        }
        __syncthreads();
    } else {
        n = block_topk(sc, nv, min(k, kSmallN), skey);
    }
    if (n == 0) {  // the same in every CTA of the cluster
        if (rank == 0) {
            if (threadIdx.x == 0) keep_count[b] = 0;
            if (dets)
                for (int i = threadIdx.x; i < K * 2; i += kSortThreads)
                    reinterpret_cast<float4*>(dets + (int64_t)b * K * 8)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        return;
    }
    if (threadIdx.x < n) {
        const int64_t row = (int64_t)b * N + (int)(skey[threadIdx.x] & 0xffffffffu);
        if constexpr (ROTATED) {
            float r[5];
            src_bev(bs, row, r);
            RBox rb;
            rbox_prepare(r, rb);
            RBoxG& d = s_box[threadIdx.x];
#pragma unroll
            for (int i = 0; i < 8; ++i) d.c[i] = rb.c[i];
            d.area = rb.area; d.mnx = rb.mnx; d.mny = rb.mny; d.mxx = rb.mxx; d.mxy = rb.mxy;
        } else {
            s_box[threadIdx.x] = src_standup(bs, row);
        }
        s_mask[threadIdx.x][0] = 0ull;
        s_mask[threadIdx.x][1] = 0ull;
    }
    if (csize > 1) cluster.sync(); else __syncthreads();
    // 32-bit ORs: native shared-memory atomics, also from the other CTAs of the cluster (a 64-bit OR is a
    // compare-and-swap loop on local shared memory and loses bits against remote updates)
    unsigned* own = reinterpret_cast<unsigned*>(&s_mask[0][0]);
    unsigned* mask0 = csize > 1 ? cluster.map_shared_rank(own, 0) : own;
    const double th = (double)thresh;
    const int npairs = n * (n - 1) / 2;
    for (int p = rank + csize * (int)threadIdx.x; p < npairs; p += csize * kSortThreads) {
        // pair p of the triangle, column-major: p = j (j - 1) / 2 + i with i < j
        int j = (int)((1.f + sqrtf(1.f + 8.f * (float)p)) * 0.5f);
        while (j * (j - 1) / 2 > p) --j;
        while ((j + 1) * j / 2 <= p) ++j;
        const int i = p - j * (j - 1) / 2;
        bool sup;
        if constexpr (ROTATED) {
            RBox a, c;
            load_rbox(&s_box[i], a);
            load_rbox(&s_box[j], c);
            sup = rbox_iou(a, c, -1) > th;  // devRotateIoU(row, col), nms_gpu.py:445-449
        } else {
            sup = standup_iou(s_box[i], s_box[j]) > th;
        }
        if (sup) atomicOr(&mask0[4 * i + (j >> 5)], 1u << (j & 31));
    }
    if (csize > 1) cluster.sync(); else __syncthreads();
    if (rank != 0) return;
    if (threadIdx.x == 0) {
        unsigned long long rm0 = 0ull, rm1 = 0ull;
        const int limit = (int)min((int64_t)(post_max > 0 ? post_max : n), keep_stride);
        int nk = 0;
        int* kp = keep + (int64_t)b * keep_stride;
        for (int i = 0; i < n && nk < limit; ++i) {
            const unsigned long long rm = i < 64 ? rm0 : rm1;
            if (!((rm >> (i & 63)) & 1ull)) {
                const int a = (int)(skey[i] & 0xffffffffu);
                s_kept[nk] = a;
                kp[nk++] = a;
                rm0 |= s_mask[i][0];
                rm1 |= s_mask[i][1];
            }
        }
        keep_count[b] = nk;
        s_nk = nk;
    }
    if (!dets) return;
    // final detections of the frame (what gather_dets does for the longer lists): decoded box + score, zero padded
    __syncthreads();
    const int nk = min(s_nk, K);
    for (int kq = threadIdx.x; kq < K; kq += kSortThreads) {
        float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (kq < nk) {
            const int64_t row = (int64_t)b * N + s_kept[kq];
            src_decoded(bs, row, o);
            o[7] = scores[row];
        }
        float4* dst = reinterpret_cast<float4*>(dets + ((int64_t)b * K + kq) * 8);
        dst[0] = make_float4(o[0], o[1], o[2], o[3]);
        dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
}

// Cluster size of nms_small for a batch of B frames: the largest of 8, 4, 2, 1 that keeps the batch in one wave of
// one CTA per SM.  A cluster lives inside one GPC (18-20 SMs on B200), so clusters of 8 pack two per GPC: 16 frames
// as 16 x 8 CTAs measured 62 us against 40 us for 32 frames as 32 x 4.
static int nms_small_cluster(int B) {
#ifdef PP_NMS_SMALL_CLUSTER
    return PP_NMS_SMALL_CLUSTER;
#endif
    const int sms = num_sms();
    if (B * 16 <= sms) return 8;
    if (B * 4 <= sms) return 4;
    if (B * 2 <= sms) return 2;
    return 1;
}

template <typename... Args>
static cudaError_t launch_clustered(void (*kernel)(Args...), unsigned grid, unsigned block, int csize, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid);
    lc.blockDim = dim3(block);
    lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)csize;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    lc.attrs = at;
    lc.numAttrs = 1;
    return cudaLaunchKernelEx(&lc, kernel, args...);
}

// final detections: out[b,k,:] = (boxes[b, keep[b,k], 0:box_dim], scores[b, keep[b,k]]), zero padded
__global__ void __launch_bounds__(256)
gather_dets_kernel(const float* __restrict__ boxes, int box_dim, const float* __restrict__ scores, int64_t N,
                   const int* __restrict__ keep, int64_t keep_stride, const int* __restrict__ keep_count,
                   int K, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int od = box_dim + 1;
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= K * od) return;
    const int k = e / od, d = e - k * od;
    float v = 0.f;
    if (k < min(keep_count[b], K)) {
        const int64_t i = (int64_t)b * N + keep[(int64_t)b * keep_stride + k];
        v = d < box_dim ? boxes[i * box_dim + d] : scores[i];
    }
    out[((int64_t)b * K + k) * od + d] = v;
}

// same with the boxes decoded on the fly (pp_decode_nms_dev)
__global__ void __launch_bounds__(256)
decode_gather_dets_kernel(BoxSrc bs, const float* __restrict__ scores, int64_t N, const int* __restrict__ keep,
                          int64_t keep_stride, const int* __restrict__ keep_count, int K, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= K) return;
    float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (k < min(keep_count[b], K)) {
        const int64_t row = (int64_t)b * N + keep[(int64_t)b * keep_stride + k];
        src_decoded(bs, row, o);
        o[7] = scores[row];
    }
    float4* dst = reinterpret_cast<float4*>(out + ((int64_t)b * K + k) * 8);
    dst[0] = make_float4(o[0], o[1], o[2], o[3]);
    dst[1] = make_float4(o[4], o[5], o[6], o[7]);
}

// d3_box_overlap (second/utils/eval.py:131-163): BEV intersection of camera boxes (columns 0,2,3,5,6, float32,
// criterion 2) x height overlap and the chosen ratio in float64, stored as float32.  "Next" row N4.
__global__ void __launch_bounds__(256)
d3_overlap_kernel(const double* __restrict__ boxes, int64_t N, const double* __restrict__ qboxes, int64_t K,
                  int criterion, float* __restrict__ out) {
    __shared__ RBoxG s_q[64];
    __shared__ RBoxG s_b[64];
    __shared__ double s_qh[64][3], s_bh[64][3];  // y, height, volume
    const int64_t n0 = (int64_t)blockIdx.y * 64, k0 = (int64_t)blockIdx.x * 64;
    if (threadIdx.x < 128) {
        const bool isq = threadIdx.x < 64;
        const int t = threadIdx.x & 63;
        const int64_t g = (isq ? k0 : n0) + t;
        if (g < (isq ? K : N)) {
            const double* src = (isq ? qboxes : boxes) + g * 7;
            const float r[5] = {(float)src[0], (float)src[2], (float)src[3], (float)src[5], (float)src[6]};
            RBox rb;
            rbox_prepare(r, rb);
            RBoxG& d = isq ? s_q[t] : s_b[t];
#pragma unroll
            for (int i = 0; i < 8; ++i) d.c[i] = rb.c[i];
            d.area = rb.area; d.mnx = rb.mnx; d.mny = rb.mny; d.mxx = rb.mxx; d.mxy = rb.mxy;
            double* h = isq ? s_qh[t] : s_bh[t];
            h[0] = src[1]; h[1] = src[4]; h[2] = __dmul_rn(__dmul_rn(src[3], src[4]), src[5]);
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 64 * 64; e += 256) {
        const int bn = e >> 6, qk = e & 63;
        if (n0 + bn >= N || k0 + qk >= K) continue;
        RBox q, bx;
        load_rbox(&s_q[qk], q);
        load_rbox(&s_b[bn], bx);
        float rinc = (float)rbox_iou(q, bx, 2);
        if (rinc > 0.f) {
            const double by = s_bh[bn][0], qy = s_qh[qk][0];
            const double iw = __dsub_rn(fmin(by, qy), fmax(__dsub_rn(by, s_bh[bn][1]), __dsub_rn(qy, s_qh[qk][1])));
            if (iw > 0.0) {
                const double inc = __dmul_rn(iw, (double)rinc);
                const double a1 = s_bh[bn][2], a2 = s_qh[qk][2];
                const double ua = criterion == -1 ? __dsub_rn(__dadd_rn(a1, a2), inc) : criterion == 0 ? a1 : criterion == 1 ? a2 : 1.0;
                rinc = (float)__ddiv_rn(inc, ua);
            } else {
                rinc = 0.f;
            }
        }
        out[(n0 + bn) * K + k0 + qk] = rinc;
    }
}

}  // namespace pp

using namespace pp;

extern "C" int pp_d3_box_overlap_dev(const double* boxes, int64_t N, const double* query_boxes, int64_t K,
                                     int criterion, float* out, void* stream) {
    PP_CHECK_ARG(N >= 0 && K >= 0 && criterion >= -1 && criterion <= 2, "pp_d3_box_overlap_dev: bad arguments");
    if (N == 0 || K == 0) return PP_OK;
    PP_CHECK_ARG(boxes && query_boxes && out, "pp_d3_box_overlap_dev: null argument");
    PP_CHECK_ARG(ceil_div(N, 64) <= 65535, "pp_d3_box_overlap_dev: N too large (chunk the boxes)");
    const dim3 g((unsigned)ceil_div(K, 64), (unsigned)ceil_div(N, 64));
    PP_TIMED("d3_overlap", static_cast<cudaStream_t>(stream));
    d3_overlap_kernel<<<g, 256, 0, static_cast<cudaStream_t>(stream)>>>(boxes, N, query_boxes, K, criterion, out);
    PP_LAUNCHED();
    return PP_OK;
}

extern "C" int pp_gather_dets_dev(const float* boxes, int box_dim, const float* scores, int B, int64_t N,
                                  const int32_t* keep, int64_t keep_stride, const int32_t* keep_count, int K,
                                  float* out, void* stream) {
    PP_CHECK_ARG(B > 0 && B <= 65535 && K > 0 && box_dim > 0 && N >= 0, "pp_gather_dets_dev: bad arguments");
    PP_CHECK_ARG(boxes && scores && keep && keep_count && out, "pp_gather_dets_dev: null argument");
    const dim3 g((unsigned)ceil_div((int64_t)K * (box_dim + 1), 256), B);
    PP_TIMED("gather_dets", static_cast<cudaStream_t>(stream));
    gather_dets_kernel<<<g, 256, 0, static_cast<cudaStream_t>(stream)>>>(boxes, box_dim, scores, N, keep, keep_stride,
                                                                         keep_count, K, out);
    PP_LAUNCHED();
    return PP_OK;
}

namespace {
struct NmsWs {
    int* order; int* n_sorted; void* sorted; unsigned long long* mask; unsigned* kbuf; int* ibuf;
    int* kept_cnt; unsigned char* dead; void* grid; int* bin_start; int* bin_items; float4* bin_hull; float* bin_area; int* slot_of;
    int* cursor; int* s_idx; int* s_n; int* new_kept; int* n_new;
    int64_t n_cap, cb_cap; size_t total; bool full_sort, stripes;
};
NmsWs nms_carve(void* ws, int kind, int B, int64_t N, int pre_max) {
    NmsWs w;
    Carver c(ws);
    w.n_cap = (pre_max > 0 && pre_max < N) ? pre_max : N;
    w.cb_cap = (w.n_cap + 63) / 64;
    w.full_sort = !(pre_max > 0 && pre_max <= kSelectMaxK) && !(N <= kSelectMaxK);
    w.order = c.take<int>((size_t)B * w.n_cap + 1);
    w.n_sorted = c.take<int>(B);
    if (kind == PP_NMS_ROTATED) w.sorted = c.take<RBoxG>((size_t)B * w.n_cap + 1);
    else w.sorted = c.take<float4>((size_t)B * w.n_cap + 1);
    w.stripes = w.n_cap > kStripeMin;
    w.kept_cnt = nullptr; w.dead = nullptr; w.grid = nullptr; w.bin_start = nullptr; w.bin_items = nullptr; w.bin_hull = nullptr; w.bin_area = nullptr; w.slot_of = nullptr;
    w.cursor = nullptr; w.s_idx = nullptr; w.s_n = nullptr; w.new_kept = nullptr; w.n_new = nullptr;
    if (w.stripes) {
        w.mask = c.take<unsigned long long>((size_t)B * kStripe * kMaskGroup + 1);
        w.kept_cnt = c.take<int>(B);
        w.cursor = c.take<int>(B);
        w.s_n = c.take<int>(B);
        w.n_new = c.take<int>(B);
        w.dead = c.take<unsigned char>((size_t)B * w.n_cap + 1);
        w.grid = c.take<StripeGrid>(B);
        w.bin_start = c.take<int>((size_t)B * (kGridBins + 1));
        w.bin_items = c.take<int>((size_t)B * w.n_cap + 1);
        w.bin_hull = c.take<float4>((size_t)B * w.n_cap + 1);
        w.bin_area = c.take<float>((size_t)B * w.n_cap + 1);
        w.slot_of = c.take<int>((size_t)B * w.n_cap + 1);
        w.s_idx = c.take<int>((size_t)B * kStripe);
        w.new_kept = c.take<int>((size_t)B * kStripe);
    } else {
        w.mask = c.take<unsigned long long>((size_t)B * w.n_cap * w.cb_cap + 1);
    }
    if (w.full_sort) {
        w.kbuf = c.take<unsigned>((size_t)B * 2 * N);
        w.ibuf = c.take<int>((size_t)B * 2 * N);
    } else { w.kbuf = nullptr; w.ibuf = nullptr; }
    w.total = c.used();
    return w;
}
}  // namespace

// Top-k of long score lists for other translation units (predict.cu): order [B][order_stride], n_sorted [B].
int nms_topk_long_dev(const float* scores, int B, int64_t N, int k, int* order, int64_t order_stride, int* n_sorted, cudaStream_t st) {
    PP_TIMED("nms_topk", st);
    return pp::nms_topk_cluster_launch(scores, nullptr, B, N, k, order, order_stride, n_sorted, st);
}

extern "C" size_t pp_nms_workspace_bytes(int kind, int B, int64_t N, int pre_max_size) {
    if (B <= 0 || N < 0) return 0;
    return nms_carve(nullptr, kind, B, N > 0 ? N : 1, pre_max_size).total + 256;
}

static int nms_run(int kind, const float* boxes, int box_stride, const float* anchors, int64_t anchor_period,
                   const float* scores, const int32_t* n_valid, int B, int64_t N, int pre_max_size, int post_max_size,
                   float thresh, int32_t* keep, int64_t keep_stride, int32_t* keep_count, void* workspace,
                   size_t workspace_bytes, void* stream, float* dets = nullptr, int K = 0, bool* dets_done = nullptr) {
    const BoxSrc bsrc{boxes, box_stride, anchors, anchor_period};
    PP_CHECK_ARG(kind == PP_NMS_STANDUP || kind == PP_NMS_ROTATED, "pp_nms_dev: bad kind");
    PP_CHECK_ARG(B > 0 && B <= 65535 && N >= 0 && N < ((int64_t)1 << 31), "pp_nms_dev: bad B/N");
    PP_CHECK_ARG(keep && keep_count && workspace && keep_stride > 0, "pp_nms_dev: null argument");
    PP_CHECK_ARG(box_stride >= (kind == PP_NMS_ROTATED ? 5 : 4), "pp_nms_dev: box_stride too small");
    PP_CHECK_ARG(thresh >= 0.f, "pp_nms_dev: the IoU threshold must be >= 0 (disjoint boxes are never tested)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (N == 0) {
        PP_CUDA(cudaMemsetAsync(keep_count, 0, sizeof(int) * B, st));
        return PP_OK;
    }
    PP_CHECK_ARG(boxes && scores, "pp_nms_dev: null boxes/scores");
    const NmsWs w = nms_carve(workspace, kind, B, N, pre_max_size);
    if (w.total > workspace_bytes) {
        set_error("pp_nms_dev: workspace %zu < required %zu", workspace_bytes, w.total);
        return PP_E_WORKSPACE;
    }
    PP_CHECK_ARG(w.cb_cap * 8 <= 200 * 1024, "pp_nms_dev: more than 1.6M boxes per frame after pre_max_size");
    if (w.n_cap <= kSmallN) {
        const int* order = nullptr;
        if (N >= kLongScores) {  // one CTA walking 100 k scores five times costs more than the rest of the frame
            PP_TIMED("nms_topk", st);
            PP_TRY_RC(nms_topk_cluster_launch(scores, n_valid, B, N, (int)w.n_cap, w.order, w.n_cap, w.n_sorted, st));
            order = w.order;
        }
        PP_TIMED("nms_small", st);
        const int csize = nms_small_cluster(B);
        PP_CUDA(launch_clustered(kind == PP_NMS_ROTATED ? nms_small_kernel<true> : nms_small_kernel<false>, (unsigned)(B * csize),
                                 (unsigned)kSortThreads, csize, st, bsrc, scores, n_valid, N, (int)w.n_cap, post_max_size, thresh,
                                 keep, keep_stride, keep_count, dets, K, order, (int64_t)w.n_cap, (const int*)w.n_sorted));
        PP_LAUNCHED();
        if (dets_done) *dets_done = true;
        return PP_OK;
    }
    if (w.full_sort) {
        PP_TIMED("nms_sort", st);
        nms_sort_kernel<<<B, kSortThreads, 0, st>>>(scores, n_valid, N, pre_max_size, w.kbuf, w.ibuf, w.order,
                                                   w.n_cap, w.n_sorted);
    } else {
        const int k = (int)w.n_cap;
        PP_TIMED("nms_topk", st);
        if (N >= kLongScores) {  // long score lists: one 8-CTA cluster per frame (DSMEM histograms)
            PP_TRY_RC(nms_topk_cluster_launch(scores, n_valid, B, N, k, w.order, w.n_cap, w.n_sorted, st));
        } else {
            nms_topk_kernel<<<B, kSortThreads, 0, st>>>(scores, n_valid, N, k, w.order, w.n_cap, w.n_sorted);
            PP_LAUNCHED();
        }
    }
    if (w.full_sort) PP_LAUNCHED();
    {
        const dim3 g((unsigned)ceil_div(w.n_cap, 256), B);
        PP_TIMED("nms_prep", st);
        if (kind == PP_NMS_ROTATED)
            nms_prep_kernel<true><<<g, 256, 0, st>>>(bsrc, N, w.order, w.n_cap, w.n_sorted, w.sorted, w.n_cap);
        else
            nms_prep_kernel<false><<<g, 256, 0, st>>>(bsrc, N, w.order, w.n_cap, w.n_sorted, w.sorted, w.n_cap);
        PP_LAUNCHED();
    }
    if (w.stripes) {
        const int limit = post_max_size > 0 ? post_max_size : 0x7fffffff;
        PP_CUDA(cudaMemsetAsync(w.kept_cnt, 0, sizeof(int) * B, st));
        PP_CUDA(cudaMemsetAsync(w.cursor, 0, sizeof(int) * B, st));
        PP_CUDA(cudaMemsetAsync(w.dead, 0, (size_t)B * w.n_cap, st));
        const bool rot = kind == PP_NMS_ROTATED;
        StripeGrid* sgrid = static_cast<StripeGrid*>(w.grid);
        {
            PP_TIMED("nms_bin_build", st);
            if (rot) nms_bin_build_kernel<true><<<B, 1024, 0, st>>>(w.sorted, w.n_cap, w.n_sorted, sgrid, w.bin_start, w.bin_items, w.bin_hull, w.bin_area, w.slot_of);
            else nms_bin_build_kernel<false><<<B, 1024, 0, st>>>(w.sorted, w.n_cap, w.n_sorted, sgrid, w.bin_start, w.bin_items, w.bin_hull, w.bin_area, w.slot_of);
            PP_LAUNCHED();
        }
        // at most ceil(n/kAlive) stripes: every stripe moves the cursor past kAlive alive boxes or to the end;
        // frames that are done (or reached post_max_size) make their CTAs exit at once
        for (int64_t base = 0; base < w.n_cap; base += kAlive) {
            {
                PP_TIMED("nms_select", st);
                nms_select_kernel<<<B, 1024, 0, st>>>(w.n_sorted, w.dead, w.slot_of, w.n_cap, w.kept_cnt, limit, w.cursor, w.s_idx, w.s_n);
                PP_LAUNCHED();
            }
            {
                const dim3 g(kAlive / 64, B);
                PP_TIMED("nms_stripe_mask", st);
                if (rot) nms_stripe_mask_kernel<true><<<g, 64 * kMaskQ, 0, st>>>(w.sorted, w.n_cap, w.s_idx, w.s_n, thresh, w.mask);
                else nms_stripe_mask_kernel<false><<<g, 64 * kMaskQ, 0, st>>>(w.sorted, w.n_cap, w.s_idx, w.s_n, thresh, w.mask);
                PP_LAUNCHED();
            }
            {
                PP_TIMED("nms_stripe_sweep", st);
                nms_stripe_sweep_kernel<<<B, 256, 0, st>>>(w.s_idx, w.s_n, w.n_cap, w.mask, w.order, limit, w.kept_cnt, w.new_kept,
                                                           w.n_new, keep, keep_stride);
                PP_LAUNCHED();
            }
            if (base + kAlive < w.n_cap) {
                const dim3 g(kAlive / kPushWarps, B);
                PP_TIMED("nms_push", st);
                if (rot) nms_push_kernel<true><<<g, kPushWarps * 32, 0, st>>>(w.sorted, w.n_cap, w.new_kept, w.n_new, w.cursor, w.n_sorted, w.kept_cnt, limit, thresh, sgrid, w.bin_start, w.bin_items, w.bin_hull, w.bin_area, w.dead);
                else nms_push_kernel<false><<<g, kPushWarps * 32, 0, st>>>(w.sorted, w.n_cap, w.new_kept, w.n_new, w.cursor, w.n_sorted, w.kept_cnt, limit, thresh, sgrid, w.bin_start, w.bin_items, w.bin_hull, w.bin_area, w.dead);
                PP_LAUNCHED();
            }
        }
        nms_stripe_finish_kernel<<<(unsigned)ceil_div(B, 128), 128, 0, st>>>(w.kept_cnt, limit, keep_stride, B, keep_count);
        PP_LAUNCHED();
        return PP_OK;
    }
    {
        const dim3 g((unsigned)ceil_div(w.cb_cap, kMaskGroup), (unsigned)w.cb_cap, B);
        PP_CHECK_ARG(w.cb_cap <= 65535, "pp_nms_dev: too many boxes per frame");
        PP_TIMED("nms_mask", st);
        if (kind == PP_NMS_ROTATED)
            nms_mask_kernel<true><<<g, 64 * kMaskQ, 0, st>>>(w.sorted, w.n_cap, w.n_sorted, thresh, w.mask, w.n_cap * w.cb_cap);
        else
            nms_mask_kernel<false><<<g, 64 * kMaskQ, 0, st>>>(w.sorted, w.n_cap, w.n_sorted, thresh, w.mask, w.n_cap * w.cb_cap);
        PP_LAUNCHED();
    }
    if (w.cb_cap <= kSweepSmemWords) {
        const size_t smem = (size_t)kSweepSmemWords * 64 * w.cb_cap * 8 + (size_t)kSweepSmemWords * 64 * 4;
        if (smem > 48 * 1024)
            PP_CUDA(cudaFuncSetAttribute(nms_sweep_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PP_TIMED("nms_sweep", st);
        nms_sweep_smem_kernel<<<B, 512, smem, st>>>(w.mask, w.n_cap * w.cb_cap, (int)w.cb_cap, w.n_sorted, w.order, w.n_cap,
                                                    post_max_size, keep, keep_stride, keep_count);
        PP_LAUNCHED();
    } else {
        const size_t smem = (size_t)w.cb_cap * 8 + 8;
        if (smem > 48 * 1024)
            PP_CUDA(cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PP_TIMED("nms_sweep", st);
        nms_sweep_kernel<<<B, kSweepThreads, smem, st>>>(w.mask, w.n_cap * w.cb_cap, w.n_sorted, w.order, w.n_cap,
                                                        post_max_size, keep, keep_stride, keep_count);
        PP_LAUNCHED();
    }
    return PP_OK;
}

extern "C" int pp_nms_dev(int kind, const float* boxes, int box_stride, const float* scores,
                          const int32_t* n_valid, int B, int64_t N, int pre_max_size, int post_max_size,
                          float thresh, int32_t* keep, int64_t keep_stride, int32_t* keep_count,
                          void* workspace, size_t workspace_bytes, void* stream) {
    return nms_run(kind, boxes, box_stride, nullptr, 0, scores, n_valid, B, N, pre_max_size, post_max_size, thresh, keep,
                   keep_stride, keep_count, workspace, workspace_bytes, stream);
}

extern "C" int pp_decode_nms_dev(int kind, const float* box_encodings, const float* anchors, int64_t anchor_period,
                                 const float* scores, const int32_t* n_valid, int B, int64_t N, int pre_max_size,
                                 int post_max_size, float thresh, int32_t* keep, int64_t keep_stride,
                                 int32_t* keep_count, float* dets, int K, void* workspace, size_t workspace_bytes,
                                 void* stream) {
    PP_CHECK_ARG(N == 0 || anchors, "pp_decode_nms_dev: null anchors");
    PP_CHECK_ARG(anchor_period >= 0, "pp_decode_nms_dev: bad anchor_period");
    PP_CHECK_ARG(!dets || (K > 0 && (reinterpret_cast<uintptr_t>(dets) & 15) == 0), "pp_decode_nms_dev: dets needs K > 0 and 16-byte alignment");
    bool dets_done = false;  // the one-launch path for <= 128 boxes writes the detections itself
    const int rc = nms_run(kind, box_encodings, 7, anchors, anchor_period, scores, n_valid, B, N, pre_max_size, post_max_size,
                           thresh, keep, keep_stride, keep_count, workspace, workspace_bytes, stream, dets, K, &dets_done);
    if (rc || !dets || dets_done) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (N == 0) {
        PP_CUDA(cudaMemsetAsync(dets, 0, (size_t)B * K * 8 * sizeof(float), st));
        return PP_OK;
    }
    const BoxSrc bsrc{box_encodings, 7, anchors, anchor_period};
    const dim3 g((unsigned)ceil_div(K, 256), B);
    PP_TIMED("gather_dets", st);
    decode_gather_dets_kernel<<<g, 256, 0, st>>>(bsrc, scores, N, keep, keep_stride, keep_count, K, dets);
    PP_LAUNCHED();
    return PP_OK;
}

extern "C" int pp_rotate_iou_dev(const float* boxes, int64_t N, const float* query_boxes, int64_t K,
                                 int criterion, float* out, void* stream) {
    PP_CHECK_ARG(N >= 0 && K >= 0 && criterion >= -1 && criterion <= 2, "pp_rotate_iou_dev: bad arguments");
    if (N == 0 || K == 0) return PP_OK;
    PP_CHECK_ARG(boxes && query_boxes && out, "pp_rotate_iou_dev: null argument");
    PP_CHECK_ARG(ceil_div(N, 64) <= 65535, "pp_rotate_iou_dev: N too large (chunk the boxes)");
    const dim3 g((unsigned)ceil_div(K, 64), (unsigned)ceil_div(N, 64));
    PP_TIMED("rotate_iou_matrix", static_cast<cudaStream_t>(stream));
    rotate_iou_matrix_kernel<<<g, 256, 0, static_cast<cudaStream_t>(stream)>>>(boxes, N, query_boxes, K, criterion, out);
    PP_LAUNCHED();
    return PP_OK;
}
