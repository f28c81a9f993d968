// Device-side predict glue ("next" row N2): the per-frame body of VoxelNet.predict,
// model/voxelnet.py:1105-1326 of the reference, for a batch of frames without leaving the device.
//
// The reference pulls nine tensors to the host (.numpy(), 1067-1084) and runs, per frame, in numpy:
//   anchor-mask gather (1119-1137), dir argmax (1143), sigmoid (1150), optional score threshold
//   (1190-1198), top-100 by argpartition (1207), second_box_decode on those (1227), rotated ->
//   standup boxes (1233-1249), nms (1259-1265, itself a numba.cuda launch with H2D/D2H),
//   direction flip (1301-1306) and box_lidar_to_camera (1319).
// Here: one streaming pass turns cls_preds (+ mask, + threshold) into scores with -inf for absent
// anchors, then ONE CTA per frame does the rest: radix-select top-k, decode of the <= 128 selected
// boxes, box prep, all-pairs suppression mask in shared memory, greedy sweep, flip, camera transform.
// Selections larger than 128 boxes (KITTI-style top_k / nms_pre_max_size of 1000 and more; the reference's
// code has no limit there) take the general decode + NMS path of nms.cu on the same scores, followed by
// a small kernel for the per-detection tail (label, direction flip, camera box).
#include <type_traits>

#include "box_math.cuh"
#include "nms_common.cuh"

namespace pp {

constexpr int kPredictMaxSel = 128;  // boxes entering NMS (reference: 100)

// scores[b,a] = max_c sigmoid(cls[b,a,c]) if the anchor is present, else -inf.
// sigmoid_array, model/voxelnet.py:722-723: 1 / (1 + np.exp(-x)) in float32.
__device__ __forceinline__ float sigmoid_f32(float x) {
    return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x)));
}

__global__ void __launch_bounds__(256)
predict_score_kernel(const float* __restrict__ cls, const unsigned char* __restrict__ mask, int64_t total,
                     int NC, float score_thr, float* __restrict__ scores) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    float s = -INFINITY;
    if (!mask || mask[i] == 1) {  // np.where(a_mask == 1), model/voxelnet.py:1112
        const float* c = cls + i * NC;
        s = sigmoid_f32(c[0]);
        for (int k = 1; k < NC; ++k) s = fmaxf(s, sigmoid_f32(c[k]));
        if (score_thr > 0.f && !(s >= score_thr)) s = -INFINITY;  // top_scores >= thresh, 1192
    }
    scores[i] = s;
}

// (r_rect @ velo2cam) in float32, rows 0..2 (load_data.py:1515); sequential k, no FMA
__device__ __forceinline__ void camera_matrix(const float* __restrict__ R, const float* __restrict__ T, float* M) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) {
            float acc = __fmul_rn(R[i * 4], T[j]);
            for (int k = 1; k < 4; ++k) acc = __fadd_rn(acc, __fmul_rn(R[i * 4 + k], T[k * 4 + j]));
            M[i * 4 + j] = acc;
        }
}

// Output row k of frame b: anchor a with decoded box `dec` (a < 0: zero padding).  Label (1176-1184), direction flip
// (1301-1306), camera box (load_data.py:1511-1523).
__device__ __forceinline__ void predict_emit(int b, int k, int a, const float* dec, const float* __restrict__ cls,
                                             const float* __restrict__ dir, const float* __restrict__ sc, const float* M,
                                             bool has_cam, int64_t A, int NC, int K, float* __restrict__ box3d_lidar,
                                             double* __restrict__ box3d_camera, float* __restrict__ out_scores,
                                             int* __restrict__ out_labels, int* __restrict__ out_index) {
    float o[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    double cam[7] = {0., 0., 0., 0., 0., 0., 0.};
    float score = 0.f;
    int label = 0;
    if (a >= 0) {
        score = sc[a];
#pragma unroll
        for (int d = 0; d < 7; ++d) o[d] = dec[d];
        if (NC > 1) {  // argmax over classes (first maximum)
            const float* c = cls + ((int64_t)b * A + a) * NC;
            float best = sigmoid_f32(c[0]);
            for (int q = 1; q < NC; ++q) {
                const float v = sigmoid_f32(c[q]);
                if (v > best) { best = v; label = q; }
            }
        }
        if (dir) {
            // np.argmax(dir_preds, -1): 1 only when the second logit is strictly larger
            const float* dp = dir + ((int64_t)b * A + a) * 2;
            const bool dl = dp[1] > dp[0];
            const bool opp = (o[6] > 0.f) != dl;
            // float32 += float64 array: the sum is formed in float64 and rounded once
            o[6] = (float)((double)o[6] + (opp ? 3.141592653589793 : 0.0));
        }
        if (has_cam && box3d_camera) {
            const double x = (double)o[0], y = (double)o[1], z = (double)o[2];
#pragma unroll
            for (int i = 0; i < 3; ++i)
                cam[i] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, (double)M[i * 4]), __dmul_rn(y, (double)M[i * 4 + 1])),
                                             __dmul_rn(z, (double)M[i * 4 + 2])),
                                   (double)M[i * 4 + 3]);
            cam[3] = (double)o[4]; cam[4] = (double)o[5]; cam[5] = (double)o[3]; cam[6] = (double)o[6];  // l, h, w, r
        }
    }
    float* po = box3d_lidar + ((int64_t)b * K + k) * 7;
#pragma unroll
    for (int d = 0; d < 7; ++d) po[d] = o[d];
    if (box3d_camera) {
        double* pc = box3d_camera + ((int64_t)b * K + k) * 7;
#pragma unroll
        for (int d = 0; d < 7; ++d) pc[d] = cam[d];
    }
    if (out_scores) out_scores[(int64_t)b * K + k] = score;
    if (out_labels) out_labels[(int64_t)b * K + k] = label;
    if (out_index) out_index[(int64_t)b * K + k] = a;
}

template <bool ROTATED>
__global__ void __launch_bounds__(kSortThreads)
predict_frame_kernel(const float* __restrict__ box_preds, const float* __restrict__ cls, const float* __restrict__ dir,
                     const float* __restrict__ anchors, int64_t anchor_frame_stride, const float* __restrict__ rect,
                     const float* __restrict__ trv2c, const float* __restrict__ scores, int64_t A, int NC, int n_max,
                     int post_max, float thresh, int K, float* __restrict__ box3d_lidar,
                     double* __restrict__ box3d_camera, float* __restrict__ out_scores, int* __restrict__ out_labels,
                     int* __restrict__ out_index, int* __restrict__ out_count, const int* __restrict__ order,
                     const int* __restrict__ n_sorted) {
    using BoxG = typename std::conditional<ROTATED, RBoxG, float4>::type;
    __shared__ unsigned long long skey[kSelectMaxK];
    __shared__ BoxG s_box[kPredictMaxSel];
    __shared__ unsigned long long s_mask[kPredictMaxSel][2];
    __shared__ float s_dec[kPredictMaxSel][7];
    __shared__ int s_keep[kPredictMaxSel];
    __shared__ int s_nk;
    __shared__ float s_M[12];
    const int b = blockIdx.x;
    const float* sc = scores + (int64_t)b * A;
    int n;
    if (order) {  // long score lists: selected beforehand by the cluster top-k of nms.cu (same order, same tie rule)
        n = min(n_sorted[b], n_max);
        if ((int)threadIdx.x < n) {
            const int a = order[(int64_t)b * kPredictMaxSel + threadIdx.x];
            skey[threadIdx.x] = ((unsigned long long)score_key(sc[a]) << 32) | (unsigned)a;
        }
        __syncthreads();
    } else {
        n = block_topk(sc, (int)A, n_max, skey);
    }
    if (n > 0) {
        if (threadIdx.x < n) {
            const int a = (int)(skey[threadIdx.x] & 0xffffffffu);
            float* d = s_dec[threadIdx.x];
            box_decode_one(box_preds + ((int64_t)b * A + a) * 7, anchors + (int64_t)b * anchor_frame_stride + (int64_t)a * 7, d);
            if constexpr (ROTATED) {
                const float r[5] = {d[0], d[1], d[3], d[4], d[6]};
                RBox rb;
                rbox_prepare(r, rb);
                RBoxG& g = s_box[threadIdx.x];
#pragma unroll
                for (int i = 0; i < 8; ++i) g.c[i] = rb.c[i];
                g.area = rb.area; g.mnx = rb.mnx; g.mny = rb.mny; g.mxx = rb.mxx; g.mxy = rb.mxy;
            } else {
                s_box[threadIdx.x] = rbox_standup_one(d[0], d[1], d[3], d[4], d[6]);
            }
            s_mask[threadIdx.x][0] = 0ull;
            s_mask[threadIdx.x][1] = 0ull;
        }
        if (threadIdx.x == kSortThreads - 1 && rect && trv2c) camera_matrix(rect + (int64_t)b * 16, trv2c + (int64_t)b * 16, s_M);
        __syncthreads();
        const double th = (double)thresh;
        for (int idx = threadIdx.x; idx < n * n; idx += kSortThreads) {
            const int i = idx / n, j = idx - i * n;
            if (j <= i) continue;
            bool sup;
            if constexpr (ROTATED) {
                RBox x, y;
                load_rbox(&s_box[i], x);
                load_rbox(&s_box[j], y);
                sup = rbox_iou(x, y, -1) > th;
            } else {
                sup = standup_iou(s_box[i], s_box[j]) > th;
            }
            if (sup) atomicOr(&s_mask[i][j >> 6], 1ull << (j & 63));
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long rm0 = 0ull, rm1 = 0ull;
            const int limit = min(post_max > 0 ? post_max : n, K);
            int nk = 0;
            for (int i = 0; i < n && nk < limit; ++i) {
                const unsigned long long rm = i < 64 ? rm0 : rm1;
                if (!((rm >> (i & 63)) & 1ull)) {
                    s_keep[nk++] = i;
                    rm0 |= s_mask[i][0];
                    rm1 |= s_mask[i][1];
                }
            }
            s_nk = nk;
            out_count[b] = nk;
        }
        __syncthreads();
    } else {
        if (threadIdx.x == 0) { s_nk = 0; out_count[b] = 0; }
        __syncthreads();
    }
    const int nk = s_nk;
    // ---- selected boxes: direction flip (1301-1306), camera boxes (load_data.py:1511-1523), zero padding
    for (int k = threadIdx.x; k < K; k += kSortThreads) {
        const bool live = k < nk;
        const int t = live ? s_keep[k] : 0;
        const int a = live ? (int)(skey[t] & 0xffffffffu) : -1;
        predict_emit(b, k, a, live ? s_dec[t] : nullptr, cls, dir, sc, s_M, rect && trv2c, A, NC, K, box3d_lidar, box3d_camera,
                     out_scores, out_labels, out_index);
    }
}

// Tail of the large-selection path: keep[b, 0..count) are anchor indices in NMS order.
__global__ void __launch_bounds__(128)
predict_tail_kernel(const float* __restrict__ box_preds, const float* __restrict__ cls, const float* __restrict__ dir,
                    const float* __restrict__ anchors, int64_t anchor_frame_stride, const float* __restrict__ rect,
                    const float* __restrict__ trv2c, const float* __restrict__ scores, int64_t A, int NC, int K,
                    const int* __restrict__ keep, const int* __restrict__ keep_count, float* __restrict__ box3d_lidar,
                    double* __restrict__ box3d_camera, float* __restrict__ out_scores, int* __restrict__ out_labels,
                    int* __restrict__ out_index, int* __restrict__ out_count) {
    __shared__ float s_M[12];
    const int b = blockIdx.y;
    const bool has_cam = rect && trv2c;
    if (threadIdx.x == 0 && has_cam) camera_matrix(rect + (int64_t)b * 16, trv2c + (int64_t)b * 16, s_M);
    __syncthreads();
    const int nk = min(keep_count[b], K);
    if (blockIdx.x == 0 && threadIdx.x == 0) out_count[b] = nk;
    const int k = blockIdx.x * 128 + threadIdx.x;
    if (k >= K) return;
    int a = -1;
    float dec[7];
    if (k < nk) {
        a = keep[(int64_t)b * K + k];
        box_decode_one(box_preds + ((int64_t)b * A + a) * 7, anchors + (int64_t)b * anchor_frame_stride + (int64_t)a * 7, dec);
    }
    predict_emit(b, k, a, dec, cls, dir, scores + (int64_t)b * A, s_M, has_cam, A, NC, K, box3d_lidar, box3d_camera, out_scores,
                 out_labels, out_index);
}

}  // namespace pp

using namespace pp;

static int predict_n_max(const pp_predict_cfg* cfg) {
    int n_max = cfg->top_k;
    if (cfg->nms_pre_max_size > 0 && cfg->nms_pre_max_size < n_max) n_max = cfg->nms_pre_max_size;
    return n_max;
}

extern "C" size_t pp_predict_workspace_bytes(const pp_predict_cfg* cfg, int B, int64_t A, int K) {
    if (!cfg || B <= 0 || A < 0 || K <= 0) return 0;
    Carver c(nullptr);
    c.take<float>((size_t)B * A + 1);
    const int n_max = predict_n_max(cfg);
    if (n_max <= kPredictMaxSel && A >= kLongScoreList) {
        c.take<int32_t>((size_t)B * kPredictMaxSel);
        c.take<int32_t>((size_t)B);
    }
    if (n_max > kPredictMaxSel) {
        c.take<int32_t>((size_t)B * K);
        c.take<int32_t>((size_t)B);
        c.take<unsigned char>(pp_nms_workspace_bytes(cfg->nms_kind, B, A, n_max));
    }
    return c.used() + 256;
}

extern "C" int pp_predict_dev(const pp_predict_cfg* cfg, const float* box_preds, const float* cls_preds,
                              const float* dir_preds, const float* anchors, const uint8_t* anchors_mask,
                              const float* rect, const float* Trv2c, int B, int64_t A, int K, float* box3d_lidar,
                              double* box3d_camera, float* scores, int32_t* label_preds, int32_t* anchor_index,
                              int32_t* count, void* workspace, size_t workspace_bytes, void* stream) {
    PP_CHECK_ARG(cfg && B > 0 && B <= 65535 && A >= 0 && A < ((int64_t)1 << 31) && K > 0, "pp_predict_dev: bad B/A/K");
    PP_CHECK_ARG(box3d_lidar && count && workspace, "pp_predict_dev: null output");
    PP_CHECK_ARG(cfg->num_class >= 1 && cfg->top_k >= 1, "pp_predict_dev: bad num_class/top_k");
    PP_CHECK_ARG(cfg->nms_kind == PP_NMS_STANDUP || cfg->nms_kind == PP_NMS_ROTATED, "pp_predict_dev: bad nms_kind");
    PP_CHECK_ARG(cfg->nms_iou_threshold >= 0.f, "pp_predict_dev: the IoU threshold must be >= 0");
    const int n_max = predict_n_max(cfg);
    PP_CHECK_ARG((rect == nullptr) == (Trv2c == nullptr), "pp_predict_dev: rect and Trv2c go together");
    if (workspace_bytes < pp_predict_workspace_bytes(cfg, B, A, K)) {
        set_error("pp_predict_dev: workspace %zu < %zu", workspace_bytes, pp_predict_workspace_bytes(cfg, B, A, K));
        return PP_E_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Carver c(workspace);
    float* sc = c.take<float>((size_t)B * A + 1);
    if (A > 0) {
        PP_CHECK_ARG(box_preds && cls_preds && anchors, "pp_predict_dev: null input");
        PP_TIMED("predict_score", st);
        predict_score_kernel<<<(unsigned)ceil_div((int64_t)B * A, 256), 256, 0, st>>>(
            cls_preds, anchors_mask, (int64_t)B * A, cfg->num_class, cfg->nms_score_threshold, sc);
        PP_LAUNCHED();
    }
    const float* dir = cfg->use_direction_classifier ? dir_preds : nullptr;
    PP_CHECK_ARG(!cfg->use_direction_classifier || dir_preds || A == 0, "pp_predict_dev: dir_preds missing");
    const int64_t astride = cfg->anchors_per_frame ? A * 7 : 0;
    if (n_max > kPredictMaxSel) {
        // the general path: top-k + decode + NMS of nms.cu on the masked scores (absent anchors are -inf there too)
        int32_t* keep = c.take<int32_t>((size_t)B * K);
        int32_t* kcnt = c.take<int32_t>((size_t)B);
        const size_t nms_bytes = pp_nms_workspace_bytes(cfg->nms_kind, B, A, n_max);
        void* nms_ws = c.take<unsigned char>(nms_bytes);
        PP_TRY_RC(pp_decode_nms_dev(cfg->nms_kind, box_preds, anchors, cfg->anchors_per_frame ? 0 : A, sc, nullptr, B, A, n_max,
                                 cfg->nms_post_max_size, cfg->nms_iou_threshold, keep, K, kcnt, nullptr, 0, nms_ws, nms_bytes,
                                 stream));
        PP_TIMED("predict_tail", st);
        predict_tail_kernel<<<dim3((unsigned)ceil_div(K, 128), B), 128, 0, st>>>(
            box_preds, cls_preds, dir, anchors, astride, rect, Trv2c, sc, A, cfg->num_class, K, keep, kcnt, box3d_lidar,
            box3d_camera, scores, label_preds, anchor_index, count);
        PP_LAUNCHED();
        return PP_OK;
    }
    int32_t *order = nullptr, *n_sorted = nullptr;
    if (A >= kLongScoreList) {  // KITTI heads (107 k anchors): one CTA walking the scores five times was 3/4 of the call
        order = c.take<int32_t>((size_t)B * kPredictMaxSel);
        n_sorted = c.take<int32_t>((size_t)B);
        PP_TRY_RC(nms_topk_long_dev(sc, B, A, n_max, order, kPredictMaxSel, n_sorted, st));
    }
    {
        PP_TIMED("predict_frame", st);
        if (cfg->nms_kind == PP_NMS_ROTATED)
            predict_frame_kernel<true><<<B, kSortThreads, 0, st>>>(box_preds, cls_preds, dir, anchors, astride, rect, Trv2c, sc, A,
                                                                   cfg->num_class, n_max, cfg->nms_post_max_size,
                                                                   cfg->nms_iou_threshold, K, box3d_lidar, box3d_camera,
                                                                   scores, label_preds, anchor_index, count, order, n_sorted);
        else
            predict_frame_kernel<false><<<B, kSortThreads, 0, st>>>(box_preds, cls_preds, dir, anchors, astride, rect, Trv2c, sc, A,
                                                                    cfg->num_class, n_max, cfg->nms_post_max_size,
                                                                    cfg->nms_iou_threshold, K, box3d_lidar, box3d_camera,
                                                                    scores, label_preds, anchor_index, count, order, n_sorted);
        PP_LAUNCHED();
    }
    return PP_OK;
}
