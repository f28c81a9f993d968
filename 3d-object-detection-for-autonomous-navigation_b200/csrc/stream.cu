// pp_stream: batches of frames from HOST memory through the whole path, inside the library.
//
// The reference hands numpy clouds to points_to_voxel on the tf.data thread (load_data.py:2966), merges the
// samples of a batch (merge_second_batch, load_data.py:2164-2224) and gets numpy detections back from
// VoxelNet.predict (model/voxelnet.py:1259-1326).  This object is that boundary for a batch: one call takes the
// batch's clouds in host memory and returns its detections in host memory, with everything between them
// (voxelize + decorate, scatter, decode + NMS) chained on the device.  The copy of batch k+1 into the second
// device staging buffer runs on a copy stream while batch k is processed, so in steady state a batch costs
// max(transfer, compute).  The layers of the host framework that sit between the stages (PFN dense / RPN) are not
// part of this library: their outputs are bound as device tensors (pp_stream_bind).
#include <new>
#include <string.h>

#include "pp_common.cuh"

using namespace pp;

namespace {
constexpr int kSlots = 2;
}

struct pp_stream {
    pp_stream_cfg cfg;
    int device = 0;
    int64_t A = 0;
    int nx = 0, ny = 0, nz = 0;
    int64_t cap_rows = 0, cap_points = 0;
    cudaStream_t compute = nullptr, copy = nullptr;
    cudaEvent_t ev_copied[kSlots] = {}, ev_consumed[kSlots] = {}, ev_done[kSlots] = {};
    // device
    void* d_points[kSlots] = {};
    int64_t* d_off[kSlots] = {};
    float *d_anchors = nullptr, *d_voxels = nullptr, *d_decorated = nullptr, *d_canvas = nullptr, *d_dets = nullptr;
    int32_t *d_coors = nullptr, *d_num = nullptr, *d_vnum = nullptr, *d_vbase = nullptr, *d_keep = nullptr, *d_kcnt = nullptr;
    int32_t* d_cellrow = nullptr;  // the voxelizer's cell -> row map, read by the scatter (grids of at most 4 z slabs)
    void *ws_vox = nullptr, *ws_sc = nullptr, *ws_nms = nullptr;
    size_t ws_vox_bytes = 0, ws_sc_bytes = 0, ws_nms_bytes = 0;
    // host (pinned)
    int64_t* h_off[kSlots] = {};
    void* h_stage[kSlots] = {};       // pageable clouds go through these
    float* h_dets[kSlots] = {};       // pageable result buffers are filled from these
    int32_t* h_kcnt[kSlots] = {};
    float* user_dets[kSlots] = {};    // where wait() has to copy to (nullptr: the DMA wrote the caller's memory)
    int32_t* user_kcnt[kSlots] = {};
    int n_frames_of[kSlots] = {};
    bool busy[kSlots] = {};
    // bound tensors of the host framework
    const float *pfn_feats = nullptr, *box_enc = nullptr, *scores = nullptr;
    int64_t seq = 0;
    size_t esz() const { return cfg.point_dtype == PP_F64 ? 8 : 4; }
};

static bool page_locked(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

extern "C" void pp_stream_destroy(pp_stream* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->compute) cudaStreamSynchronize(s->compute);
    if (s->copy) cudaStreamSynchronize(s->copy);
    for (int k = 0; k < kSlots; ++k) {
        cudaFree(s->d_points[k]); cudaFree(s->d_off[k]);
        cudaFreeHost(s->h_off[k]); cudaFreeHost(s->h_stage[k]); cudaFreeHost(s->h_dets[k]); cudaFreeHost(s->h_kcnt[k]);
        if (s->ev_copied[k]) cudaEventDestroy(s->ev_copied[k]);
        if (s->ev_consumed[k]) cudaEventDestroy(s->ev_consumed[k]);
        if (s->ev_done[k]) cudaEventDestroy(s->ev_done[k]);
    }
    cudaFree(s->d_anchors); cudaFree(s->d_voxels); cudaFree(s->d_decorated); cudaFree(s->d_canvas); cudaFree(s->d_dets);
    cudaFree(s->d_cellrow); cudaFree(s->d_coors); cudaFree(s->d_num); cudaFree(s->d_vnum); cudaFree(s->d_vbase); cudaFree(s->d_keep); cudaFree(s->d_kcnt);
    cudaFree(s->ws_vox); cudaFree(s->ws_sc); cudaFree(s->ws_nms);
    if (s->compute) cudaStreamDestroy(s->compute);
    if (s->copy) cudaStreamDestroy(s->copy);
    cudaGetLastError();
    delete s;
}

static int stream_init(pp_stream* s, const float* anchors_host) {
    const pp_stream_cfg& c = s->cfg;
    int32_t g[3];
    PP_TRY_RC(pp_grid_size(c.vox.voxel_size, c.vox.coors_range, c.vox.arith_f32, g));
    s->nx = g[0]; s->ny = g[1]; s->nz = g[2];
    const int64_t ncell = (int64_t)g[0] * g[1] * g[2];
    const int B = c.max_frames, P = c.vox.max_points, D = c.D;
    s->cap_rows = (int64_t)B * (c.vox.max_voxels < ncell ? c.vox.max_voxels : ncell);
    s->cap_points = (int64_t)B * c.max_frame_points;
    PP_CUDA(cudaStreamCreateWithFlags(&s->compute, cudaStreamNonBlocking));
    PP_CUDA(cudaStreamCreateWithFlags(&s->copy, cudaStreamNonBlocking));
    const size_t pts_bytes = (size_t)s->cap_points * D * s->esz();
    for (int k = 0; k < kSlots; ++k) {
        PP_CUDA(cudaEventCreateWithFlags(&s->ev_copied[k], cudaEventDisableTiming));
        PP_CUDA(cudaEventCreateWithFlags(&s->ev_consumed[k], cudaEventDisableTiming));
        PP_CUDA(cudaEventCreateWithFlags(&s->ev_done[k], cudaEventDisableTiming));
        PP_CUDA(cudaMalloc(&s->d_points[k], pts_bytes ? pts_bytes : 16));
        PP_CUDA(cudaMalloc(&s->d_off[k], (size_t)(B + 1) * 8));
        PP_CUDA(cudaHostAlloc(&s->h_off[k], (size_t)(B + 1) * 8, cudaHostAllocDefault));
        PP_CUDA(cudaHostAlloc(&s->h_dets[k], (size_t)B * c.post_max * 8 * 4 + 16, cudaHostAllocDefault));
        PP_CUDA(cudaHostAlloc(&s->h_kcnt[k], (size_t)B * 4 + 16, cudaHostAllocDefault));
    }
    PP_CUDA(cudaMalloc(&s->d_anchors, (size_t)s->A * 28));
    PP_CUDA(cudaMemcpy(s->d_anchors, anchors_host, (size_t)s->A * 28, cudaMemcpyHostToDevice));
    if (c.keep_voxels) PP_CUDA(cudaMalloc(&s->d_voxels, (size_t)s->cap_rows * P * D * 4));
    PP_CUDA(cudaMalloc(&s->d_decorated, (size_t)s->cap_rows * P * (D + 5) * 4));
    PP_CUDA(cudaMalloc(&s->d_coors, (size_t)s->cap_rows * 16));
    PP_CUDA(cudaMalloc(&s->d_num, (size_t)s->cap_rows * 4));
    PP_CUDA(cudaMalloc(&s->d_vnum, (size_t)B * 4));
    PP_CUDA(cudaMalloc(&s->d_vbase, (size_t)(B + 1) * 4));
    if (s->nz <= 4) PP_CUDA(cudaMalloc(&s->d_cellrow, (size_t)B * s->nx * s->ny * s->nz * 4));
    PP_CUDA(cudaMalloc(&s->d_canvas, (size_t)B * c.C * s->ny * s->nx * 4));
    PP_CUDA(cudaMalloc(&s->d_keep, (size_t)B * c.post_max * 4));
    PP_CUDA(cudaMalloc(&s->d_kcnt, (size_t)B * 4));
    PP_CUDA(cudaMalloc(&s->d_dets, (size_t)B * c.post_max * 8 * 4));
    s->ws_vox_bytes = pp_voxelize_workspace_bytes(&c.vox, s->cap_points, B, c.max_frame_points, D, PP_F32);
    s->ws_sc_bytes = pp_scatter_workspace_bytes(B, s->ny, s->nx, s->cap_rows);
    s->ws_nms_bytes = pp_nms_workspace_bytes(c.nms_kind, B, s->A, c.pre_max);
    PP_CHECK_ARG(s->ws_vox_bytes > 0, "pp_stream_create: bad voxelizer configuration");
    PP_CUDA(cudaMalloc(&s->ws_vox, s->ws_vox_bytes));
    PP_CUDA(cudaMalloc(&s->ws_sc, s->ws_sc_bytes ? s->ws_sc_bytes : 16));
    PP_CUDA(cudaMalloc(&s->ws_nms, s->ws_nms_bytes ? s->ws_nms_bytes : 16));
    return PP_OK;
}

extern "C" int pp_stream_create(int device, const pp_stream_cfg* cfg, const float* anchors, int64_t A, pp_stream** out) {
    PP_CHECK_ARG(cfg && anchors && out && A > 0, "pp_stream_create: null argument");
    PP_CHECK_ARG(cfg->max_frames > 0 && cfg->max_frame_points > 0 && cfg->C > 0 && cfg->post_max > 0 &&
                     (cfg->D == 3 || cfg->D == 4) && (cfg->point_dtype == PP_F32 || cfg->point_dtype == PP_F64),
                 "pp_stream_create: bad configuration");
    int ndev = 0;
    PP_CUDA(cudaGetDeviceCount(&ndev));
    PP_CHECK_ARG(device >= 0 && device < ndev, "pp_stream_create: device %d of %d", device, ndev);
    PP_CUDA(cudaSetDevice(device));
    pp_stream* s = new (std::nothrow) pp_stream();
    if (!s) { set_error("out of host memory"); return PP_E_NOMEM; }
    s->cfg = *cfg;
    s->device = device;
    s->A = A;
    const int rc = stream_init(s, anchors);
    if (rc != PP_OK) { pp_stream_destroy(s); return rc; }
    *out = s;
    return PP_OK;
}

extern "C" int pp_stream_bind(pp_stream* s, const float* pfn_feats, const float* box_encodings, const float* scores) {
    PP_CHECK_ARG(s && pfn_feats && box_encodings && scores, "pp_stream_bind: null argument");
    s->pfn_feats = pfn_feats; s->box_enc = box_encodings; s->scores = scores;
    return PP_OK;
}

extern "C" int64_t pp_stream_cap_rows(pp_stream* s) { return s ? s->cap_rows : 0; }
extern "C" int64_t pp_stream_anchor_count(pp_stream* s) { return s ? s->A : 0; }

extern "C" int pp_stream_wait(pp_stream* s, int64_t ticket) {
    PP_CHECK_ARG(s, "pp_stream_wait: null stream");
    PP_CUDA(cudaSetDevice(s->device));
    for (int k = 0; k < kSlots; ++k) {
        if (!s->busy[k]) continue;
        if (ticket >= 0 && (ticket & 1) != k) continue;
        PP_CUDA(cudaEventSynchronize(s->ev_done[k]));
        const int n = s->n_frames_of[k];
        if (s->user_dets[k]) memcpy(s->user_dets[k], s->h_dets[k], (size_t)n * s->cfg.post_max * 8 * 4);
        if (s->user_kcnt[k]) memcpy(s->user_kcnt[k], s->h_kcnt[k], (size_t)n * 4);
        s->user_dets[k] = nullptr; s->user_kcnt[k] = nullptr;
        s->busy[k] = false;
    }
    return PP_OK;
}

extern "C" int pp_stream_submit(pp_stream* s, const void* points, const int64_t* frame_offsets, int n_frames,
                                float* dets, int32_t* counts, int64_t* ticket) {
    PP_CHECK_ARG(s && points && frame_offsets && dets && counts, "pp_stream_submit: null argument");
    PP_CHECK_ARG(s->pfn_feats, "pp_stream_submit: pp_stream_bind first");
    const pp_stream_cfg& c = s->cfg;
    PP_CHECK_ARG(n_frames > 0 && n_frames <= c.max_frames, "pp_stream_submit: n_frames %d of %d", n_frames, c.max_frames);
    const int64_t total = frame_offsets[n_frames] - frame_offsets[0];
    PP_CHECK_ARG(frame_offsets[0] == 0 && total >= 0 && total <= s->cap_points, "pp_stream_submit: %lld points, capacity %lld",
                 (long long)total, (long long)s->cap_points);
    int64_t maxf = 0;
    for (int b = 0; b < n_frames; ++b) {
        const int64_t nb = frame_offsets[b + 1] - frame_offsets[b];
        PP_CHECK_ARG(nb >= 0 && nb <= c.max_frame_points, "pp_stream_submit: frame %d has %lld points, capacity %lld", b,
                     (long long)nb, (long long)c.max_frame_points);
        maxf = nb > maxf ? nb : maxf;
    }
    PP_CUDA(cudaSetDevice(s->device));
    const int k = (int)(s->seq & 1);
    if (s->busy[k]) PP_TRY_RC(pp_stream_wait(s, k));  // at most two batches in flight
    const size_t bytes = (size_t)total * c.D * s->esz();
    // ---- copy stream: clouds + offsets into staging buffer k (free once the batch that used it was voxelized)
    PP_CUDA(cudaStreamWaitEvent(s->copy, s->ev_consumed[k], 0));
    memcpy(s->h_off[k], frame_offsets, (size_t)(n_frames + 1) * 8);
    PP_CUDA(cudaMemcpyAsync(s->d_off[k], s->h_off[k], (size_t)(n_frames + 1) * 8, cudaMemcpyHostToDevice, s->copy));
    if (bytes) {
        const void* src = points;
        if (!page_locked(points)) {
            // pageable memory: one host copy into the slot's pinned buffer (grown on demand), then one DMA
            if (!s->h_stage[k]) PP_CUDA(cudaHostAlloc(&s->h_stage[k], (size_t)s->cap_points * c.D * s->esz(), cudaHostAllocDefault));
            memcpy(s->h_stage[k], points, bytes);
            src = s->h_stage[k];
        }
        PP_CUDA(cudaMemcpyAsync(s->d_points[k], src, bytes, cudaMemcpyHostToDevice, s->copy));
    }
    PP_CUDA(cudaEventRecord(s->ev_copied[k], s->copy));
    // ---- compute stream
    cudaStream_t st = s->compute;
    PP_CUDA(cudaStreamWaitEvent(st, s->ev_copied[k], 0));
    PP_TRY_RC(pp_voxelize_dev(&c.vox, s->d_points[k], c.point_dtype, c.D, s->d_off[k], n_frames, total, maxf, PP_F32,
                              s->d_voxels, s->d_decorated, s->d_coors, 4, s->d_num, s->cap_rows, s->d_vnum, s->d_vbase,
                              nullptr, s->d_cellrow, s->ws_vox, s->ws_vox_bytes, st));
    PP_CUDA(cudaEventRecord(s->ev_consumed[k], st));
    if (s->d_cellrow)
        PP_TRY_RC(pp_scatter_cells_dev(s->pfn_feats, s->d_cellrow, s->nz, c.C, n_frames, s->ny, s->nx, c.layout, s->d_canvas, st));
    else
        PP_TRY_RC(pp_scatter_dev(s->pfn_feats, s->d_coors, s->cap_rows, s->d_vbase + n_frames, c.C, n_frames, s->ny, s->nx,
                             c.layout, s->d_canvas, s->ws_sc, s->ws_sc_bytes, st));
    PP_TRY_RC(pp_decode_nms_dev(c.nms_kind, s->box_enc, s->d_anchors, s->A, s->scores, nullptr, n_frames, s->A, c.pre_max,
                                c.post_max, c.iou_threshold, s->d_keep, c.post_max, s->d_kcnt, s->d_dets, c.post_max,
                                s->ws_nms, s->ws_nms_bytes, st));
    // ---- detections back
    const size_t dbytes = (size_t)n_frames * c.post_max * 8 * 4;
    const bool direct = page_locked(dets) && page_locked(counts);
    PP_CUDA(cudaMemcpyAsync(direct ? dets : s->h_dets[k], s->d_dets, dbytes, cudaMemcpyDeviceToHost, st));
    PP_CUDA(cudaMemcpyAsync(direct ? counts : s->h_kcnt[k], s->d_kcnt, (size_t)n_frames * 4, cudaMemcpyDeviceToHost, st));
    PP_CUDA(cudaEventRecord(s->ev_done[k], st));
    s->user_dets[k] = direct ? nullptr : dets;
    s->user_kcnt[k] = direct ? nullptr : counts;
    s->n_frames_of[k] = n_frames;
    s->busy[k] = true;
    if (ticket) *ticket = s->seq;
    ++s->seq;
    return PP_OK;
}

extern "C" int pp_stream_view(pp_stream* s, pp_stream_tensors* out) {
    PP_CHECK_ARG(s && out, "pp_stream_view: null argument");
    out->voxels = s->d_voxels; out->decorated = s->d_decorated; out->coors = s->d_coors; out->num_points = s->d_num;
    out->voxel_num = s->d_vnum; out->voxel_base = s->d_vbase; out->canvas = s->d_canvas; out->dets = s->d_dets;
    out->keep_count = s->d_kcnt; out->compute_stream = s->compute;
    return PP_OK;
}
