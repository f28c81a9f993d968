// Internal interface between voxelize.cu (dispatch, any-grid path) and voxelize_small.cu (grids whose
// per-cell tables fit in shared memory).  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pp_b200.h"

namespace pp {

struct VoxParams;

bool vox_small_eligible(const pp_voxel_cfg* cfg, int64_t ncell, int n_frames, int64_t total_points,
                        int64_t max_frame_points, int D);
size_t vox_small_workspace_bytes(const pp_voxel_cfg* cfg, int64_t ncell, int n_frames, int64_t total_points,
                                 int64_t max_frame_points, int D, int out_dtype);
int vox_small_run(const pp_voxel_cfg* cfg, const VoxParams& p, const void* points, int point_dtype,
                  const int64_t* frame_offsets, int n_frames, int64_t total_points, int64_t max_frame_points,
                  int out_dtype, void* voxels, float* decorated, int32_t* coors, int coors_cols, int32_t* num_points,
                  int64_t cap_rows, int32_t* voxel_num, int32_t* voxel_base, int32_t* point_slot, int32_t* cell_voxel,
                  void* workspace, size_t workspace_bytes, cudaStream_t st);

void vox_small_set_min_points(int64_t n);

}  // namespace pp
