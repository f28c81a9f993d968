// Deterministic first-come pillarization for grids whose per-cell tables fit in shared memory
// (the d435i grid of configs/train.yaml: 80 x 64 x 2 = 10 240 cells).  Same results, bit for bit, as the
// reference's sequential loop (load_data.py:593-692) and as the any-grid path in voxelize.cu, but built
// as a two-level counting sort: no pass issues a global atomic per point, none sorts, none ranks a bitmap.
//
//   scan    (frame, chunk of 16 384 points)  rows staged by TMA bulk copies, four points per thread; cell id in
//           the reference's arithmetic; in-range points are compacted IN INDEX ORDER into 16/32-byte records
//           (coordinates in the output type) and 4-byte tags {cell | index in chunk}; the chunk's per-cell counts
//           live in SHARED memory (ATOMS) and are written out once per chunk, plus a saturated snapshot of them at
//           the three quarter boundaries of the chunk.
//   prefix  (frame, 256 cells; short batches: 32 cells x 8 ranges of the frame's chunks)  per cell: exclusive prefix of the chunk counts = the slot base of every chunk
//           (uint8, saturated: a base >= max_points means "full"); min(points, max_points) goes to the tail word
//           of the cell's row of the slot table; the chunk in which a cell first appears is counted, which gives
//           every chunk the number of voxels opened before it.
//   place   (frame, quarter chunk)  ONE WARP walks the quarter's record tags in order, 32 per step, with the
//           running per-cell counts (chunk base + quarter snapshot) in shared memory: slot = count[cell] + rank
//           among the step's earlier records of the cell.  The step's peers are found through a shared-memory
//           atomicOr table keyed by the cell (match_any costs 60 cycles per step and SM on scattered cells, the table
//           4.4: tools/micro/match_tput.cu).  A record that finds count 0 opens its cell: because the walk is in index
//           order, the running number of such records IS the voxel id (order of first touch), and the record that
//           would open voxel number max_voxels is the reference's `break` position (load_data.py:630-634).  The
//           record's position goes to entry `slot` of the cell's row of the slot table (max_points + 1 words per
//           cell, L2 resident; scattering the 16-byte records themselves was measured DRAM-random-write bound).
//   finish  (pillar)  one warp per two pillars, lane = slot: cell -> slot row -> records (before the break position);
//           every lane owns its slot of the output rows whether or not a point sits in it, so the zero-padded voxel
//           rows and the fused PillarFeatureNet decoration (model/pointpillars.py:143-203) leave as dense coalesced
//           stores straight from registers; num_points, coors, point->slot map.  (Staging the rows in shared memory
//           for TMA bulk stores was measured slower: profiles/r02_notes.md.)
//
// The kernels live in voxelize_small_impl.cuh and are compiled for two chunk sizes: 16 384 points (batches: fewest
// per-chunk tables per point) and 4 096 points (a frame or a few: four times the CTAs and a quarter of the
// dependent chain per CTA; one d435i frame 92 -> 59 us).
#include "pp_common.cuh"
#include "vox_common.cuh"
#include "vox_internal.h"

#define PP_VS_NS vs14
#define PP_VS_SHIFT 14
#define PP_VS_SUB 4
#define PP_VS_STAGES 2
#define PP_VS_KB 5
#define PP_VS_PPT 4
#ifndef PP_BATCH_PLACE_THREADS
#define PP_BATCH_PLACE_THREADS 128
#endif
#define PP_VS_PLACE_THREADS PP_BATCH_PLACE_THREADS
#include "voxelize_small_impl.cuh"
#undef PP_VS_PLACE_THREADS
#undef PP_VS_PPT
#undef PP_VS_NS
#undef PP_VS_SHIFT
#undef PP_VS_SUB
#undef PP_VS_STAGES
#undef PP_VS_KB
#ifndef PP_SHORT_SHIFT
#define PP_SHORT_SHIFT 12
#endif
#define PP_VS_NS vs12
#define PP_VS_SHIFT PP_SHORT_SHIFT
#ifdef PP_SHORT_SUB
#define PP_VS_SUB PP_SHORT_SUB
#elif PP_SHORT_SHIFT >= 12
#define PP_VS_SUB 4
#elif PP_SHORT_SHIFT == 11
#define PP_VS_SUB 2
#else
#define PP_VS_SUB 1
#endif
#ifndef PP_SHORT_STAGES
#define PP_SHORT_STAGES 2
#endif
#ifndef PP_SHORT_KB
#define PP_SHORT_KB 5
#endif
#ifndef PP_SHORT_PPT
#define PP_SHORT_PPT 4
#endif
#define PP_VS_STAGES PP_SHORT_STAGES
#define PP_VS_KB PP_SHORT_KB
#ifndef PP_SHORT_PLACE_THREADS
#define PP_SHORT_PLACE_THREADS 128
#endif
#define PP_VS_PPT PP_SHORT_PPT
#define PP_VS_PLACE_THREADS PP_SHORT_PLACE_THREADS
#include "voxelize_small_impl.cuh"
#undef PP_VS_PLACE_THREADS
#undef PP_VS_PPT
#undef PP_VS_NS
#undef PP_VS_SHIFT
#undef PP_VS_SUB
#undef PP_VS_STAGES
#undef PP_VS_KB

namespace pp {

#ifndef PP_SMALL_MIN_POINTS
#define PP_SMALL_MIN_POINTS 150000
#endif
#ifndef PP_SHORT_BATCH_CHUNKS
#define PP_SHORT_BATCH_CHUNKS 640
#endif
// batches below this many points take the any-grid path (process-wide, pp_voxelize_set_small_path_min_points)
static int64_t g_small_min_points = PP_SMALL_MIN_POINTS;
// batches of at most this many 4 096-point chunks (frames x chunks of the largest frame) use the small chunk size
constexpr int64_t kShortBatchChunks = PP_SHORT_BATCH_CHUNKS;

void vox_small_set_min_points(int64_t n) { g_small_min_points = n < 0 ? PP_SMALL_MIN_POINTS : n; }

static int64_t short_chunks(int n_frames, int64_t max_frame_points) {
    return (int64_t)n_frames * (max_frame_points > 0 ? ceil_div(max_frame_points, (int64_t)vs12::kChunk) : 1);
}
static bool short_batch(int n_frames, int64_t max_frame_points) { return short_chunks(n_frames, max_frame_points) <= kShortBatchChunks; }

bool vox_small_eligible(const pp_voxel_cfg* cfg, int64_t ncell, int n_frames, int64_t total_points,
                        int64_t max_frame_points, int D) {
    if (total_points < g_small_min_points) return false;  // small clouds: the any-grid path has less fixed work per frame
    return short_batch(n_frames, max_frame_points) ? vs12::eligible(cfg, ncell, n_frames, total_points, max_frame_points, D)
                                                   : vs14::eligible(cfg, ncell, n_frames, total_points, max_frame_points, D);
}

// Bytes for this batch AND for any smaller one (a workspace sized for the largest batch also serves a single frame,
// which takes the small chunk size: its per-chunk tables are bounded by kShortBatchChunks).
size_t vox_small_workspace_bytes(const pp_voxel_cfg* cfg, int64_t ncell, int n_frames, int64_t total_points,
                                 int64_t max_frame_points, int D, int out_dtype) {
    const int64_t sc = short_chunks(n_frames, max_frame_points);
    const size_t b12 = vs12::workspace_bytes_of(cfg, ncell, n_frames, sc < kShortBatchChunks ? sc : kShortBatchChunks, total_points, D, out_dtype);
    if (short_batch(n_frames, max_frame_points)) return b12;
    const size_t b14 = vs14::workspace_bytes_of(cfg, ncell, n_frames, (int64_t)n_frames * vs14::chunks_of(max_frame_points), total_points, D, out_dtype);
    return b14 > b12 ? b14 : b12;
}

int vox_small_run(const pp_voxel_cfg* cfg, const VoxParams& p, const void* points, int point_dtype,
                  const int64_t* frame_offsets, int n_frames, int64_t total_points, int64_t max_frame_points,
                  int out_dtype, void* voxels, float* decorated, int32_t* coors, int coors_cols, int32_t* num_points,
                  int64_t cap_rows, int32_t* voxel_num, int32_t* voxel_base, int32_t* point_slot, int32_t* cell_voxel,
                  void* workspace, size_t workspace_bytes, cudaStream_t st) {
    return (short_batch(n_frames, max_frame_points) ? vs12::run : vs14::run)(
        cfg, p, points, point_dtype, frame_offsets, n_frames, total_points, max_frame_points, out_dtype, voxels, decorated, coors,
        coors_cols, num_points, cap_rows, voxel_num, voxel_base, point_slot, cell_voxel, workspace, workspace_bytes, st);
}

}  // namespace pp
