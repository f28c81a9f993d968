/*
 * pp_b200.h -- C ABI of the B200-native PointPillars pre/post-processing hot path.
 *
 * One shared library (libpp_b200.so, sm_100a) replaces the data-parallel stages of
 * krullgit/3D-Object-Detection-for-autonomous-navigation.  Each entry point names the
 * reference interface it stands in for (paths relative to the reference checkout).
 *
 * Three layers, all plain C (pointers + sizes, no framework types):
 *
 *   *_dev   device pointers in, device pointers out, asynchronous on the caller's stream,
 *           caller-provided workspace.  This is what a DLPack / __cuda_array_interface__
 *           consumer (TensorFlow via tf.experimental.dlpack, torch, cupy) binds.
 *   *_host  host (numpy) buffers in and out through a pp_ctx that owns a stream, a grow-only device
 *           arena and grow-only pinned staging buffers (pageable caller memory is copied through them
 *           in pieces so that the host copy overlaps the transfer).  Synchronous, like the
 *           reference's numpy functions.
 *           This is what the reference's call sites bind through ctypes.
 *   pp_stream  a batch of host clouds in, host detections out per call; the copy of the next batch
 *           overlaps the kernels of the current one inside the library (end of this header).
 *
 * Conventions
 *   - Every function returns 0 on success, <0 on error (PP_E_*); pp_last_error_string() gives the
 *     thread-local message.  Nothing throws or exits across this boundary.
 *   - The library never frees or retains caller memory.  State lives in pp_ctx / pp_stream objects; process-wide are
 *     only the kernel launch-configuration cache and pp_voxelize_set_small_path_min_points.
 *   - A pp_ctx must not be used from two threads at once; create one per thread (the reference
 *     calls the voxelizer on the tf.data thread and NMS on the main thread).
 *   - `stream` is a cudaStream_t passed as void*.  The legacy default stream is never used
 *     implicitly.
 *   - There is no CPU fallback: without a CUDA device every compute entry returns PP_E_CUDA.
 */
#ifndef PP_B200_H
#define PP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PP_API __attribute__((visibility("default")))

#define PP_OK 0
#define PP_E_INVALID (-1)   /* bad argument */
#define PP_E_CUDA (-2)      /* CUDA runtime error (message has the cudaError string) */
#define PP_E_WORKSPACE (-3) /* workspace / output capacity too small */
#define PP_E_NOMEM (-4)

#define PP_F32 0
#define PP_F64 1

#define PP_LAYOUT_NCHW 0 /* reference output of PointPillarsScatter.call */
#define PP_LAYOUT_NHWC 1 /* what RPN.call transposes to (model/voxelnet.py:697) */

PP_API const char* pp_last_error_string(void);
PP_API int pp_version(void);
/* Number of kernel launches issued by this thread through the library since the last reset
 * (bench.py's gpu_launches). */
PP_API int64_t pp_launch_count(int reset);

/* Per-launch device timing (the reference's measure_time_extended timers, model/voxelnet.py:
 * 753-760, done with CUDA events on the launching stream).  pp_profile_start() arms this thread;
 * pp_profile_stop() synchronises and returns the number of records: names as '\n'-separated
 * text in launch order, ms[i] the device time of launch i. */
PP_API int pp_profile_start(void);
PP_API int pp_profile_stop(char* names, size_t names_cap, float* ms, int max_records);

/* ---- grid -------------------------------------------------------------------------------
 * np.round((range[3:]-range[:3])/voxel_size).astype(int32), round-half-to-even:
 * load_data.py:612-615 and 730-731.  arith_f32: the wrapper cast python lists to float32
 * because points are float32 (load_data.py:726-729). */
PP_API int pp_grid_size(const double voxel_size[3], const double coors_range[6], int arith_f32,
                 int32_t grid_xyz[3]);

/* ---- voxelizer --------------------------------------------------------------------------
 * Replaces points_to_voxel / _points_to_voxel_reverse_kernel / _points_to_voxel_kernel,
 * load_data.py:593-771 (call site load_data.py:2966), for a batch of independent frames.
 * Bit-exact first-come semantics: voxel id = order of first touch, slot = order of arrival,
 * max_points cap per voxel, and the scan BREAKS at the first point that would open voxel
 * number max_voxels (load_data.py:630-634).  NaN coordinates drop the point (UB in the
 * reference). */
typedef struct pp_voxel_cfg {
    double voxel_size[3];
    double coors_range[6];
    int32_t max_points;
    int32_t max_voxels;
    int32_t reverse_index; /* 1: coors are (z,y,x) (the reference's call), 0: (x,y,z) */
    int32_t arith_f32;     /* 1 only when points AND voxel_size/coors_range were float32 */
} pp_voxel_cfg;

/* Workspace of pp_voxelize_dev for a batch of n_frames frames, total_points points in all, at most
 * max_frame_points in one frame (the batch total is a valid bound when the largest frame is not known),
 * D values per point, voxels of out_dtype.  A call with smaller counts fits in the same workspace. */
PP_API size_t pp_voxelize_workspace_bytes(const pp_voxel_cfg* cfg, int64_t total_points, int n_frames,
                                   int64_t max_frame_points, int D, int out_dtype);

/*  points         [total_points, D] point_dtype, frames concatenated
 *  frame_offsets  device int64 [n_frames+1], frame b = rows [off[b], off[b+1])
 *  max_frame_points  max over b of off[b+1]-off[b] (host knows it; sizes the grid)
 *  voxels         [cap_rows, max_points, D] out_dtype (PP_F32, or PP_F64 when points are f64),
 *                 frames packed back to back (merge_second_batch layout, load_data.py:2203-2214);
 *                 every written row is zero-padded.  May be NULL when `decorated` is given.
 *  decorated      optional [cap_rows, max_points, D+5] float32: PillarFeatureNet decoration
 *                 (model/pointpillars.py:143-203) fused into the same pass
 *  coors          [cap_rows, coors_cols] int32; coors_cols 3 = reference's per-frame coors,
 *                 4 = (frame, c0, c1, c2) as merge_second_batch pads it
 *  num_points     [cap_rows] int32
 *  voxel_num      [n_frames] int32; voxel_base [n_frames+1] int32 exclusive scan (row of frame b's
 *                 first voxel); cap_rows must be >= min(sum of per-frame voxel counts,
 *                 n_frames*max_voxels) -- n_frames*max_voxels is always enough
 *  point_slot     optional [total_points] int32: voxel_in_frame*max_points+slot, -1 if dropped
 *  cell_voxel     optional [n_frames*nz*ny*nx] int32: packed row of the cell's voxel or -1
 *                 (the reference's coor_to_voxelidx map) */
PP_API int pp_voxelize_dev(const pp_voxel_cfg* cfg, const void* points, int point_dtype, int D,
                    const int64_t* frame_offsets, int n_frames, int64_t total_points,
                    int64_t max_frame_points, int out_dtype, void* voxels, float* decorated,
                    int32_t* coors, int coors_cols, int32_t* num_points, int64_t cap_rows,
                    int32_t* voxel_num, int32_t* voxel_base, int32_t* point_slot,
                    int32_t* cell_voxel, void* workspace, size_t workspace_bytes, void* stream);

/* Two bit-identical implementations sit behind pp_voxelize_dev: grids of at most 16 384 cells (the d435i grid) have a
 * path that keeps its per-cell tables in shared memory (chunks of 4 096 points for a few frames, of 16 384 for batches);
 * below ~150 000 points it has no advantage over the any-grid path (measured: profiles/r02_notes.md).  Batches of fewer
 * than `n` points take the any-grid path; default 150 000, 0 = always the shared-memory path when the grid allows it,
 * n < 0 = back to the default.  Process-wide; size workspaces after setting it. */
PP_API int pp_voxelize_set_small_path_min_points(int64_t n);

/* ---- pillar decoration ------------------------------------------------------------------
 * Replaces lines 143-203 of PillarFeatureNet.call (model/pointpillars.py), constants 121-124,
 * mask 23-49.  voxels [M,P,D] f32, num_points [M], coors [M,4] (batch,z,y,x) ->
 * out [M,P,D+5] f32.  vx, vy, x_offset, y_offset as the reference computes them in python
 * doubles (x_offset = vx/2 + range[0]); they are rounded to float32 like TF constants. */
PP_API int pp_decorate_dev(const float* voxels, const int32_t* num_points, const int32_t* coors, int64_t M,
                    int P, int D, double vx, double vy, double x_offset, double y_offset, float* out,
                    void* stream);

/* ---- scatter ----------------------------------------------------------------------------
 * Replaces PointPillarsScatter.call, model/pointpillars.py:285-341: out[b,:,y,x] = SUM of the
 * feature rows with coords (b,*,y,x) (z ignored, duplicates added, lines 302,317); everything
 * else zero.  `out` is fully written (no memset needed).  Rows with b outside [0,B) or (y,x)
 * outside the canvas are ignored.  `coords` and `out` must be 16-byte aligned.  M may be read from device (`M_dev`, e.g. voxel_base+n_frames)
 * when the host does not know it; then M is the capacity. */
PP_API size_t pp_scatter_workspace_bytes(int B, int ny, int nx, int64_t M);
PP_API int pp_scatter_dev(const float* feats, const int32_t* coords, int64_t M, const int32_t* M_dev, int C,
                   int B, int ny, int nx, int layout, float* out, void* workspace,
                   size_t workspace_bytes, void* stream);
/* The same canvas straight from the voxelizer's cell -> row map (pp_voxelize_dev's cell_voxel, [B][nz][ny][nx], nz <= 4):
 * for coords that come from pp_voxelize_dev the result is identical to pp_scatter_dev (rows of the z slabs of one
 * (y, x) are added in ascending row order, as the reference's scatter_nd does, model/pointpillars.py:302-317), without
 * the link pass and its workspace.  What the chained pipeline (pp_stream, FramePipeline) uses between its own stages. */
PP_API int pp_scatter_cells_dev(const float* features, const int32_t* cell_voxel, int nz, int C, int B, int ny, int nx,
                         int layout, float* out, void* stream);


/* ---- box decode -------------------------------------------------------------------------
 * Replaces second_box_decode (default flags), libraries/eval_helper_functions.py:388-461
 * (duplicate second/core/box_np_ops.py:69-104).  [N,7] f32 each, order x,y,z,w,l,h,r.
 * `anchor_period` > 0: anchors holds anchor_period rows reused cyclically (one anchor set for a
 * batch of frames); 0: one anchor row per encoding row. */
PP_API int pp_box_decode_dev(const float* box_encodings, const float* anchors, int64_t N,
                      int64_t anchor_period, float* out, void* stream);

/* Rotated BEV box (x,y,w,l,r) [N,5] -> standup box (xmin,ymin,xmax,ymax) [N,4]:
 * center_to_corner_box2d + corner_to_standup_nd_jit, load_data.py:1525-1594, 1330-1341, as
 * called at model/voxelnet.py:1233-1249.  in_stride = floats per input row (5, or 7 to read
 * x,y,w,l,r straight out of decoded [N,7] boxes at columns 0,1,3,4,6). */
PP_API int pp_rbox_to_standup_dev(const float* boxes, int in_stride, int64_t N, float* out, void* stream);

/* ---- NMS --------------------------------------------------------------------------------
 * kind PP_NMS_STANDUP replaces nms / nms_gpu / nms_kernel / nms_postprocess,
 *   libraries/eval_helper_functions.py:463-598 (axis-aligned IoU with the "+1" convention);
 *   boxes [B,N,4] (xmin,ymin,xmax,ymax).
 * kind PP_NMS_ROTATED replaces rotate_nms_gpu / rotate_nms_kernel / devRotateIoU,
 *   second/core/non_max_suppression/nms_gpu.py:180-490; boxes [B,N,5] (x,y,w,l,angle), or
 *   box_stride 7: decoded boxes (x,y,z,w,l,h,r), BEV columns 0,1,3,4,6 read in place.
 * Both: scores [B,N]; order = descending score, ties by descending index; optional top
 * pre_max_size (<=0: all), greedy suppression of IoU > thresh (strict, thresh rounded to
 * float32), first post_max_size kept (<=0: all).  Boxes whose score is -inf are treated as absent
 * (see pp_anchor_mask_dev).
 *   n_valid   optional device [B]: only the first n_valid[b] boxes of frame b are used
 *   keep      [B, keep_stride] int32 indices into the frame's boxes, in keep order
 *   keep_count[B] int32 (count is clipped to keep_stride) */
#define PP_NMS_STANDUP 0
#define PP_NMS_ROTATED 1
PP_API size_t pp_nms_workspace_bytes(int kind, int B, int64_t N, int pre_max_size);
PP_API int pp_nms_dev(int kind, const float* boxes, int box_stride, const float* scores,
               const int32_t* n_valid, int B, int64_t N, int pre_max_size, int post_max_size,
               float thresh, int32_t* keep, int64_t keep_stride, int32_t* keep_count,
               void* workspace, size_t workspace_bytes, void* stream);

/* Decode + NMS in one call: box_encodings [B,N,7] and anchors ([anchor_period,7] reused cyclically, or [B*N,7]
 * with anchor_period 0) instead of boxes.  second_box_decode (eval_helper_functions.py:388-461) is applied only to
 * the boxes NMS looks at -- the order of the reference's live path: top-k (model/voxelnet.py:1207), decode (1227),
 * standup boxes (1233-1249, kind PP_NMS_STANDUP) or BEV boxes (kind PP_NMS_ROTATED), NMS (1259) -- so a batch never
 * materialises all decoded anchors.  keep / keep_count as pp_nms_dev; dets optional [B,K,8] (16-byte aligned):
 * decoded box + score of each kept box, zero padded (what pp_gather_dets_dev returns).  Workspace:
 * pp_nms_workspace_bytes. */
PP_API int pp_decode_nms_dev(int kind, const float* box_encodings, const float* anchors, int64_t anchor_period,
                      const float* scores, const int32_t* n_valid, int B, int64_t N, int pre_max_size,
                      int post_max_size, float thresh, int32_t* keep, int64_t keep_stride, int32_t* keep_count,
                      float* dets, int K, void* workspace, size_t workspace_bytes, void* stream);

/* Final detections of a batch (the only tensor the pipeline copies back to the host):
 * out[b,k,:] = (boxes[b,keep[b,k],0:box_dim], scores[b,keep[b,k]]) for k < keep_count[b], zeros
 * after.  Mirrors `box_preds[selected]`, `top_scores[selected]`, model/voxelnet.py:1281-1287. */
PP_API int pp_gather_dets_dev(const float* boxes, int box_dim, const float* scores, int B, int64_t N,
                       const int32_t* keep, int64_t keep_stride, const int32_t* keep_count, int K,
                       float* out, void* stream);

/* ---- rotated IoU matrix -----------------------------------------------------------------
 * Replaces rotate_iou_gpu and rotate_iou_gpu_eval, nms_gpu.py:526-561 and 618-653 (kernels
 * 493-523, 579-615).  boxes [N,5], query_boxes [K,5] f32 -> out [N,K] f32,
 * out[n,k] = devRotateIoUEval(query_boxes[k], boxes[n], criterion); criterion -1 IoU,
 * 0 inter/area(query), 1 inter/area(box), 2 intersection area. */
PP_API int pp_rotate_iou_dev(const float* boxes, int64_t N, const float* query_boxes, int64_t K,
                      int criterion, float* out, void* stream);

/* "Next" row N4: d3_box_overlap, second/utils/eval.py:131-163 (bev_box_overlap, 126-128, is pp_rotate_iou_dev).
 * boxes [N,7], query_boxes [K,7] float64 camera boxes (x,y,z,l,h,w,ry) -> out [N,K] float32. */
PP_API int pp_d3_box_overlap_dev(const double* boxes, int64_t N, const double* query_boxes, int64_t K,
                          int criterion, float* out, void* stream);

/* ---- anchor mask ("next" row N1) ---------------------------------------------------------------
 * Replaces the per-sample anchor mask of the data loader, load_data.py:3043-3072:
 * sparse_sum_for_anchors_mask (586-591) + cumsum(0).cumsum(1) + fused_get_anchors_area (558-584) on
 * rbbox2d_to_near_bbox(anchors[:, [0,1,3,4,6]]) (534-548), mask = area > anchor_area_threshold.
 *   pp_anchor_cells_dev   anchors [A,7] f32 -> cells [A,4] int32 (x0,y0,x1,y1), once per anchor set
 *   pp_anchor_mask_dev    coors [M,coors_cols] (3: one frame (z,y,x); 4: (frame,z,y,x)), M or *M_dev rows ->
 *                         area [B,A] f32 (optional), mask [B,A] uint8 (optional),
 *                         masked_scores [B,A] = mask ? scores : -inf (optional; pp_nms_dev ignores -inf
 *                         scores, which is how the reference's `box_preds[a_mask]` gather is expressed) */
PP_API int pp_anchor_cells_dev(const float* anchors, int64_t A, const double voxel_size[3],
                        const double coors_range[6], int32_t* cells, void* stream);
PP_API size_t pp_anchor_mask_workspace_bytes(int B, int ny, int nx);
PP_API int pp_anchor_mask_dev(const int32_t* coors, int coors_cols, int64_t M, const int32_t* M_dev, int B, int ny,
                       int nx, const int32_t* cells, int64_t A, float threshold, const float* scores,
                       float* area, uint8_t* mask, float* masked_scores, void* workspace,
                       size_t workspace_bytes, void* stream);

/* ---- predict glue ("next" row N2) ---------------------------------------------------------------
 * Replaces the per-frame body of VoxelNet.predict, model/voxelnet.py:1105-1326, for a batch of frames:
 * anchor-mask gather (1119-1137), dir argmax (1143), sigmoid_array (722-723, 1150), optional score
 * threshold (1190-1198), top-k by score (hard-coded 100, 1207), second_box_decode of the selected
 * boxes (1227), standup boxes (1233-1249) and nms (1259-1265) -- or rotated NMS when nms_kind is
 * PP_NMS_ROTATED --, `box_preds[selected]` (1281-1287), direction flip (1301-1306) and
 * box_lidar_to_camera (1319, load_data.py:1511-1523).  One launch for the scores, then one CTA per frame for
 * the rest when min(top_k, nms_pre_max_size) <= 128 (the reference's 100); larger selections (the reference's
 * code has no limit) run pp_decode_nms_dev on the same scores plus one launch for the per-detection tail.
 * Workspace: pp_predict_workspace_bytes(cfg, B, A, K).
 *   box_preds [B,A,7], cls_preds [B,A,num_class], dir_preds [B,A,2] (NULL without direction classifier),
 *   anchors [A,7] shared (anchors_per_frame 0) or [B,A,7], anchors_mask [B,A] uint8 or NULL,
 *   rect, Trv2c [B,4,4] float32 or both NULL (then box3d_camera is not written)
 *   outputs, K rows per frame, zero padded after count[b]:
 *   box3d_lidar [B,K,7] f32, box3d_camera [B,K,7] f64 (x,y,z,l,h,w,r: float64 like the reference's
 *   concatenate of float64 xyz with float32 columns), scores [B,K] f32, label_preds [B,K] int32,
 *   anchor_index [B,K] int32 (index into the frame's A anchors, -1 padded), count [B] (0 = the reference's None). */
typedef struct pp_predict_cfg {
    int32_t num_class;                /* columns of cls_preds (encode_background_as_zeros) */
    int32_t use_direction_classifier;
    int32_t top_k;                    /* 100 in the reference */
    int32_t nms_pre_max_size;         /* <= 0: none */
    int32_t nms_post_max_size;        /* <= 0: none */
    int32_t nms_kind;                 /* PP_NMS_STANDUP (the live path) or PP_NMS_ROTATED */
    float nms_iou_threshold;
    float nms_score_threshold;        /* <= 0: off */
    int32_t anchors_per_frame;
} pp_predict_cfg;
PP_API size_t pp_predict_workspace_bytes(const pp_predict_cfg* cfg, int B, int64_t A, int K);
PP_API int pp_predict_dev(const pp_predict_cfg* cfg, const float* box_preds, const float* cls_preds,
                   const float* dir_preds, const float* anchors, const uint8_t* anchors_mask, const float* rect,
                   const float* Trv2c, int B, int64_t A, int K, float* box3d_lidar, double* box3d_camera,
                   float* scores, int32_t* label_preds, int32_t* anchor_index, int32_t* count, void* workspace,
                   size_t workspace_bytes, void* stream);

/* ---- sensor ingest ("next" row N3) ---------------------------------------------------------------
 * Replaces the production branch of dataLoader.__getitem__, load_data.py:2434-2443 (same sequence in
 * scripts/realsense_make_dataset.py:382-414):
 *     ros_numpy.point_cloud2.pointcloud2_to_xyz_array(pc)   rows with finite x,y,z only, as float64
 *     [start::step]                                          ([1::4] in the reference)
 *     np.dot(points, r), np.dot(points, r2)                  n_rot float64 3x3 matrices, row vectors
 *     points + [0, 0, 1]                                     translation
 * cloud  [B, n_in] records of point_step bytes (sensor_msgs/PointCloud2 data), float32 fields at byte
 *        offsets off_x, off_y, off_z; a plain [N,3] float32 array is point_step 12, offsets 0,4,8
 * rotations  HOST pointer, n_rot (<= 4) row-major 3x3 float64; translation HOST pointer [3] or NULL
 * points_out [B, cap, 3] float64; rows at and after n_out[b] are NaN (the voxelizer drops NaN points), so
 *        the batch feeds pp_voxelize_dev with frame_offsets b*cap.  cap >= ceil((n_in - start)/step) never
 *        truncates.  n_out [B] int32.
 * Bit-identical to numpy for matrices whose entries are 0, +-1, +-2^k (the reference's from_euler(+-90 deg)
 * matrices); within 1 ulp per product-sum otherwise (BLAS summation order is unspecified). */
PP_API size_t pp_ingest_workspace_bytes(int B, int64_t n_in);
PP_API int pp_ingest_dev(const void* cloud, int B, int64_t n_in, int point_step, int off_x, int off_y, int off_z,
                  int start, int step, const double* rotations, int n_rot, const double* translation,
                  double* points_out, int64_t cap, int32_t* n_out, void* workspace, size_t workspace_bytes,
                  void* stream);

/* ---- context + host-buffer layer ----------------------------------------------------------- */
typedef struct pp_ctx pp_ctx;
PP_API int pp_ctx_create(int device, pp_ctx** out);
PP_API void pp_ctx_destroy(pp_ctx* ctx);
PP_API void* pp_ctx_stream(pp_ctx* ctx);
PP_API int pp_ctx_device(pp_ctx* ctx);
PP_API int pp_ctx_sync(pp_ctx* ctx);
/* Page-locked host memory.  The *_host functions hand caller buffers that are page-locked (these, cudaHostAlloc /
 * cudaHostRegister memory, torch pinned tensors) straight to the copy engine; pageable buffers go through the
 * context's pinned ring. */
PP_API int pp_host_alloc(size_t bytes, void** out);
PP_API void pp_host_free(void* p);

/* points_to_voxel(points, voxel_size, coors_range, max_points, reverse_index, max_voxels),
 * load_data.py:695-771: host points [N,D] -> host voxels [max_voxels,max_points,D] in the
 * points' dtype (only the first *voxel_num_out rows are written, zero padded), coors
 * [max_voxels,3], num_points [max_voxels]; point_slot optional [N]. */
PP_API int pp_points_to_voxel_host(pp_ctx* ctx, const pp_voxel_cfg* cfg, const void* points,
                            int point_dtype, int64_t N, int D, void* voxels, int32_t* coors,
                            int32_t* num_points, int32_t* voxel_num_out, int32_t* point_slot);
PP_API int pp_decorate_host(pp_ctx* ctx, const float* voxels, const int32_t* num_points,
                     const int32_t* coors, int64_t M, int P, int D, double vx, double vy,
                     double x_offset, double y_offset, float* out);
PP_API int pp_scatter_host(pp_ctx* ctx, const float* feats, const int32_t* coords, int64_t M, int C, int B,
                    int ny, int nx, int layout, float* out);
PP_API int pp_box_decode_host(pp_ctx* ctx, const float* box_encodings, const float* anchors, int64_t N,
                       float* out);
PP_API int pp_rbox_to_standup_host(pp_ctx* ctx, const float* boxes, int64_t N, float* out);
/* keep: int64 [min(N, pre, post)]; *keep_count_out == 0 is the reference's `None`. */
PP_API int pp_nms_host(pp_ctx* ctx, int kind, const float* boxes, const float* scores, int64_t N,
                int pre_max_size, int post_max_size, float thresh, int64_t* keep,
                int32_t* keep_count_out);
/* anchors_area / anchors_mask of one frame, load_data.py:3043-3072: coors [M,3] (z,y,x), anchors [A,7]. */
PP_API int pp_anchors_mask_host(pp_ctx* ctx, const int32_t* coors, int64_t M, const float* anchors, int64_t A,
                         const double voxel_size[3], const double coors_range[6], float threshold,
                         float* area_out, uint8_t* mask_out);
PP_API int pp_d3_box_overlap_host(pp_ctx* ctx, const double* boxes, int64_t N, const double* query_boxes, int64_t K,
                           int criterion, float* out);
PP_API int pp_rotate_iou_host(pp_ctx* ctx, const float* boxes, int64_t N, const float* query_boxes,
                       int64_t K, int criterion, float* out);
/* One host cloud in, host float64 points [*n_out, 3] out (points_out holds cap rows). */
PP_API int pp_ingest_host(pp_ctx* ctx, const void* cloud, int64_t n_in, int point_step, int off_x, int off_y, int off_z,
                   int start, int step, const double* rotations, int n_rot, const double* translation,
                   double* points_out, int64_t cap, int32_t* n_out);
/* VoxelNet.predict(example, preds_dict) per-frame body on host arrays (see pp_predict_dev). */
PP_API int pp_predict_host(pp_ctx* ctx, const pp_predict_cfg* cfg, const float* box_preds, const float* cls_preds,
                    const float* dir_preds, const float* anchors, const uint8_t* anchors_mask, const float* rect,
                    const float* Trv2c, int B, int64_t A, int K, float* box3d_lidar, double* box3d_camera,
                    float* scores, int32_t* label_preds, int32_t* anchor_index, int32_t* count);

/* ---- batches of frames from host memory -------------------------------------------------------------------
 * The reference hands numpy clouds to points_to_voxel on the tf.data thread (load_data.py:2966), merges the samples
 * of a batch (merge_second_batch, load_data.py:2164-2224) and takes numpy detections out of VoxelNet.predict
 * (model/voxelnet.py:1259-1326).  pp_stream is that boundary for a batch: clouds in host memory in, detections in
 * host memory out, with voxelize + decorate -> scatter -> decode + NMS chained on the device in between.  The clouds
 * of batch k+1 are copied into a second device staging buffer on a copy stream while batch k is processed.
 * Page-locked caller memory (pp_host_alloc, cudaHostAlloc / cudaHostRegister, torch pinned tensors) is handed to the
 * copy engine directly; pageable memory is staged through pinned buffers of the stream (one extra host copy).
 * The host framework's layers between the stages stay outside: their output tensors are bound as device pointers.
 * One producer thread per pp_stream; different pp_streams are independent. */
typedef struct pp_stream pp_stream;
typedef struct pp_stream_cfg {
    pp_voxel_cfg vox;
    int32_t D;                /* values per point (3 or 4) */
    int32_t point_dtype;      /* PP_F32 / PP_F64: dtype of the host clouds */
    int32_t C;                /* PFN output channels = canvas channels */
    int32_t layout;           /* PP_LAYOUT_NCHW (the reference's canvas) / PP_LAYOUT_NHWC */
    int32_t nms_kind;         /* PP_NMS_ROTATED / PP_NMS_STANDUP */
    int32_t pre_max, post_max;
    float iou_threshold;
    int32_t max_frames;       /* frames per batch */
    int32_t keep_voxels;      /* 1: also materialise the raw voxel rows [rows, max_points, D] */
    int64_t max_frame_points; /* points per frame */
} pp_stream_cfg;
typedef struct pp_stream_tensors {   /* device tensors of the last batch (pp_stream_view) */
    const float* voxels;      /* [rows, max_points, D] or NULL */
    const float* decorated;   /* [rows, max_points, D+5]: the PFN's input */
    const int32_t* coors;     /* [rows, 4] (frame, z, y, x) */
    const int32_t* num_points;
    const int32_t* voxel_num; /* [max_frames] */
    const int32_t* voxel_base;/* [max_frames + 1] */
    const float* canvas;      /* [max_frames, C, ny, nx] or NHWC: the RPN's input */
    const float* dets;        /* [max_frames, post_max, 8] box7 + score */
    const int32_t* keep_count;
    void* compute_stream;     /* the stream these tensors are produced on */
} pp_stream_tensors;
/* anchors: HOST [A,7] float32 (one set for every frame). */
PP_API int pp_stream_create(int device, const pp_stream_cfg* cfg, const float* anchors, int64_t A, pp_stream** out);
PP_API void pp_stream_destroy(pp_stream* s);
PP_API int64_t pp_stream_cap_rows(pp_stream* s);     /* rows of pfn_feats the scatter stage may read */
PP_API int64_t pp_stream_anchor_count(pp_stream* s);
/* DEVICE tensors of the host framework: PFN output [cap_rows, C], RPN box encodings [max_frames, A, 7] and scores
 * [max_frames, A]; they must stay valid while batches are in flight. */
PP_API int pp_stream_bind(pp_stream* s, const float* pfn_feats, const float* box_encodings, const float* scores);
/* One batch.  points: HOST [frame_offsets[n_frames], D] of cfg.point_dtype, frames concatenated; frame_offsets: HOST
 * int64 [n_frames + 1] starting at 0; dets: HOST [n_frames, post_max, 8] float32 (box7 + score, zero padded); counts:
 * HOST [n_frames] int32.  Returns after the work has been enqueued (at most two batches are in flight: the call
 * first waits for the batch before the previous one); `points` may be reused as soon as pp_stream_wait returns for
 * this ticket (pageable memory: as soon as this call returns); dets / counts are valid after pp_stream_wait. */
PP_API int pp_stream_submit(pp_stream* s, const void* points, const int64_t* frame_offsets, int n_frames,
                     float* dets, int32_t* counts, int64_t* ticket);
/* ticket < 0: every batch in flight. */
PP_API int pp_stream_wait(pp_stream* s, int64_t ticket);
PP_API int pp_stream_view(pp_stream* s, pp_stream_tensors* out);

#ifdef __cplusplus
}
#endif
#endif /* PP_B200_H */
