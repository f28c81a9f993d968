// Pieces shared by nms.cu and predict.cu: score ordering (radix-select top-k + bitonic sort), the
// 64-byte prepared rotated box, and the axis-aligned "+1" IoU of the live path.
#pragma once
#include "pp_common.cuh"
#include "rotated_iou.cuh"

namespace pp {

constexpr int kSortThreads = 1024;
constexpr int kSelectMaxK = 1024;

__device__ __forceinline__ unsigned score_key(float s) {
    if (s != s) return 0xFFFFFFFFu;  // numpy sorts NaN last, i.e. first after the [::-1]
    const unsigned u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// ---------------------------------------------------------------------------------------------
// Top-k (k <= 1024) per frame: radix select of the k-th largest key, compaction, bitonic sort.
// Boxes whose score is -inf are absent (pp_anchor_mask_dev writes -inf for masked-out anchors: the
// reference gathers `box_preds[a_mask]` before scoring, model/voxelnet.py:1119-1137).  Block-wide count
// of present scores; all threads of the block must call it.
__device__ __forceinline__ int block_count_present(const float* __restrict__ sc, int nv) {
    int total = 0;
    for (int i0 = 0; i0 < nv; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        total += __syncthreads_count(i < nv && sc[i] != -INFINITY);
    }
    return total;
}

// Warp 0: find the digit d (255..0) where the count of keys with a larger digit is < rem <= that
// count + hist[d]; returns d and the number still needed inside digit d.  8 bins per lane.
__device__ __forceinline__ void select_digit(const unsigned* hist, unsigned rem, unsigned prefix, int shift,
                                             unsigned* prefix_out, unsigned* rem_out, unsigned* cnt_out) {
    const int lane = lane_id();
    // lane l owns bins 255-8l .. 248-8l (descending)
    unsigned h[8], tot = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) { h[q] = hist[255 - 8 * lane - q]; tot += h[q]; }
    unsigned inc = tot;  // inclusive prefix over lanes (descending bins)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    const unsigned before = inc - tot;  // keys in bins above this lane's range
    const bool mine = before < rem && rem <= inc;
    const unsigned bal = __ballot_sync(0xffffffffu, mine);
    if (bal == 0) {  // fewer than rem keys in total: take digit 0
        const unsigned total = __shfl_sync(0xffffffffu, inc, 31);
        if (lane == 0) { *prefix_out = prefix; *rem_out = rem - total + hist[0]; *cnt_out = hist[0]; }
        return;
    }
    if (mine) {
        unsigned acc = before;
        int q = 0;
        for (; q < 7; ++q) {
            if (acc + h[q] >= rem) break;
            acc += h[q];
        }
        *prefix_out = prefix | ((unsigned)(255 - 8 * lane - q) << shift);
        *rem_out = rem - acc;
        *cnt_out = h[q];
    }
}

// One histogram pass of the radix select over `count` elements: element e has key key_of(e) and takes part when
// in(key); bins are the byte at `shift`.  Warp-aggregated: scores cluster in a few top-byte bins, a plain atomicAdd
// would serialise the whole block on one shared-memory word.  All threads of the block must call it.
template <typename KeyOf, typename In>
__device__ __forceinline__ void topk_hist_pass(unsigned* hist, int count, int shift, KeyOf key_of, In in) {
    for (int e0 = 0; e0 < count; e0 += kSortThreads) {
        const int e = e0 + threadIdx.x;
        int d = -1;
        if (e < count) {
            const unsigned key = key_of(e);
            if (in(key)) d = (int)((key >> shift) & 255u);
        }
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (d >= 0 && (int)lane_id() == __ffs(peers) - 1) atomicAdd(&hist[d], (unsigned)__popc(peers));
    }
}

// Block-wide: the n = min(kmax, present scores) best of sc[0..nv) by (score desc, index desc), sorted, as
// (key<<32 | index) in skey[0..n); returns n (kmax <= kSelectMaxK).  All kSortThreads threads of the block must call it.
//   pass 1   byte histogram of all keys + count of absent (-inf) scores
//   as soon as the keys that can still be selected (larger prefix, or inside the chosen bin) number <= kTopkCand, their
//   indices are compacted into a shared-memory list and the remaining passes, the tie pass and the collection walk
//   that list instead of the whole frame (d435i: 10 240 scores -> ~1 900 after the first byte)
constexpr int kTopkCand = 4096;
__device__ __forceinline__ int block_topk(const float* __restrict__ sc, int nv, int kmax,
                                          unsigned long long* skey /*[kSelectMaxK]*/) {
    __shared__ unsigned hist[256];
    __shared__ unsigned cand[kTopkCand];
    __shared__ unsigned s_prefix, s_remaining, s_count, s_scratch, s_absent, s_ncand;

    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) { s_absent = 0; s_ncand = 0; s_prefix = 0; }
    __syncthreads();
    // ---- pass 1 over the frame: top byte + absent count
    for (int i0 = 0; i0 < nv; i0 += kSortThreads) {
        const int i = i0 + threadIdx.x;
        int d = -1;
        bool absent = false;
        if (i < nv) {
            const float v = sc[i];
            absent = v == -INFINITY;
            d = (int)(score_key(v) >> 24);
        }
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (d >= 0 && (int)lane_id() == __ffs(peers) - 1) atomicAdd(&hist[d], (unsigned)__popc(peers));
        const unsigned ab = __ballot_sync(0xffffffffu, absent);
        if (ab && lane_id() == 0) atomicAdd(&s_absent, (unsigned)__popc(ab));
    }
    __syncthreads();
    const int kk = min(kmax, nv - (int)s_absent);
    if (kk <= 0) return 0;
    unsigned T = 0, Tidx = 0;
    int ncand = -1;  // < 0: the frame itself is the element list
    auto key_at = [&](int e) { return score_key(sc[ncand < 0 ? e : (int)cand[e]]); };
    auto idx_at = [&](int e) { return ncand < 0 ? (unsigned)e : cand[e]; };
    if (kk < nv) {
        // ---- k-th largest key
        for (int shift = 24; shift >= 0; shift -= 8) {
            const unsigned prefix = s_prefix;
            if (shift != 24) {
                if (threadIdx.x < 256) hist[threadIdx.x] = 0;
                __syncthreads();
                topk_hist_pass(hist, ncand < 0 ? nv : ncand, shift, key_at,
                               [&](unsigned key) { return ((key ^ prefix) >> (shift + 8)) == 0; });
                __syncthreads();
            }
            if (threadIdx.x < 32) {
                const unsigned rem = shift == 24 ? (unsigned)kk : s_remaining;
                __syncwarp();
                select_digit(hist, rem, prefix, shift, &s_prefix, &s_remaining, &s_count);
            }
            __syncthreads();
            // keys that can still be selected: larger prefix (kk - remaining of them) or inside the chosen bin
            if (ncand < 0 && shift > 0 && (unsigned)kk - s_remaining + s_count <= (unsigned)kTopkCand) {
                const unsigned floor_key = s_prefix;
                for (int i0 = 0; i0 < nv; i0 += kSortThreads) {
                    const int i = i0 + threadIdx.x;
                    const bool take = i < nv && (score_key(sc[i]) >> shift) >= (floor_key >> shift);
                    const unsigned bal = __ballot_sync(0xffffffffu, take);
                    unsigned base = 0;
                    if (lane_id() == 0 && bal) base = atomicAdd(&s_ncand, (unsigned)__popc(bal));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (take) cand[base + __popc(bal & ((1u << lane_id()) - 1u))] = (unsigned)i;
                }
                __syncthreads();
                ncand = (int)s_ncand;
            }
        }
        T = s_prefix;
        const unsigned need_eq = s_remaining, have_eq = s_count;
        __syncthreads();
        if (have_eq > need_eq) {
            // ---- among keys == T keep the need_eq largest indices (tie rule: descending index)
            if (threadIdx.x == 0) { s_prefix = 0; s_remaining = need_eq; }
            for (int shift = 24; shift >= 0; shift -= 8) {
                if (threadIdx.x < 256) hist[threadIdx.x] = 0;
                __syncthreads();
                const unsigned prefix = s_prefix;
                const int count = ncand < 0 ? nv : ncand;
                for (int e = threadIdx.x; e < count; e += kSortThreads) {
                    if (key_at(e) != T) continue;
                    const unsigned key = idx_at(e);
                    if (shift == 24 || ((key ^ prefix) >> (shift + 8)) == 0) atomicAdd(&hist[(key >> shift) & 255u], 1u);
                }
                __syncthreads();
                if (threadIdx.x < 32) {
                    const unsigned rem = s_remaining;
                    __syncwarp();
                    select_digit(hist, rem, prefix, shift, &s_prefix, &s_remaining, &s_scratch);
                }
                __syncthreads();
            }
            Tidx = s_prefix;
            __syncthreads();
        }
    }
    // ---- collection (arbitrary order) + bitonic sort, descending on (key, index)
    if (threadIdx.x == 0) s_count = 0;
    int np2 = 1;
    while (np2 < kk) np2 <<= 1;
    for (int i = threadIdx.x; i < np2; i += kSortThreads) skey[i] = 0ull;
    __syncthreads();
    {
        const int count = ncand < 0 ? nv : ncand;
        for (int e = threadIdx.x; e < count; e += kSortThreads) {
            const unsigned key = key_at(e), i = idx_at(e);
            if (kk == nv || key > T || (key == T && i >= Tidx)) {
                const unsigned pos = atomicAdd(&s_count, 1u);
                if (pos < (unsigned)kSelectMaxK) skey[pos] = ((unsigned long long)key << 32) | i;
            }
        }
    }
    __syncthreads();
    for (int size = 2; size <= np2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < (np2 >> 1); t += kSortThreads) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const unsigned long long a = skey[lo], c = skey[hi];
                if ((a < c) == desc) { skey[lo] = c; skey[hi] = a; }
            }
            __syncthreads();
        }
    }
    return kk;
}

// ---------------------------------------------------------------------------------------------
struct __align__(16) RBoxG {  // 64 bytes in global / shared memory
    float c[8];
    float area, mnx, mny, mxx, mxy, pad0, pad1, pad2;
};

__device__ __forceinline__ void load_rbox(const RBoxG* g, RBox& r) {
    const float4* p = reinterpret_cast<const float4*>(g);
    const float4 a = p[0], b = p[1], c = p[2], d = p[3];
    r.c[0] = a.x; r.c[1] = a.y; r.c[2] = a.z; r.c[3] = a.w;
    r.c[4] = b.x; r.c[5] = b.y; r.c[6] = b.z; r.c[7] = b.w;
    r.area = c.x; r.mnx = c.y; r.mny = c.z; r.mxx = c.w; r.mxy = d.x;
}

// iou_device, eval_helper_functions.py:553-564: float32 differences, then "+ 1" onwards in float64
__device__ __forceinline__ double standup_iou(const float4& a, const float4& b) {
    const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
    const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
    const double width = fmax((double)__fsub_rn(right, left) + 1.0, 0.0);
    const double height = fmax((double)__fsub_rn(bottom, top) + 1.0, 0.0);
    const double interS = width * height;
    const double Sa = ((double)__fsub_rn(a.z, a.x) + 1.0) * ((double)__fsub_rn(a.w, a.y) + 1.0);
    const double Sb = ((double)__fsub_rn(b.z, b.x) + 1.0) * ((double)__fsub_rn(b.w, b.y) + 1.0);
    return interS / (Sa + Sb - interS);
}

}  // namespace pp

// nms.cu: top-k of long score lists (>= 16 384 per frame) by one 8-CTA cluster per frame; order [B][order_stride]
// (descending score, then descending index), n_sorted [B].
int nms_topk_long_dev(const float* scores, int B, int64_t N, int k, int* order, int64_t order_stride, int* n_sorted, cudaStream_t st);
constexpr int64_t kLongScoreList = 16384;
