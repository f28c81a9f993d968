// C-ABI glue: thread-local error string, launch counter, pp_ctx (stream + device arena) and
// the host-buffer entry points the reference's numpy call sites bind (include/pp_b200.h).
#include <map>
#include <mutex>
#include <new>
#include <tuple>
#include <vector>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "pp_common.cuh"

namespace pp {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches += n; }

struct ProfRec { const char* name; cudaEvent_t e0, e1; };
static thread_local bool g_prof_on = false;
static thread_local std::vector<ProfRec>* g_prof = nullptr;

LaunchTimer::LaunchTimer(const char* name, cudaStream_t st) : st_(st), slot_(-1) {
    if (!g_prof_on) return;
    ProfRec r{name, nullptr, nullptr};
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    cudaEventRecord(r.e0, st);
    g_prof->push_back(r);
    slot_ = (int)g_prof->size() - 1;
}
LaunchTimer::~LaunchTimer() {
    if (slot_ >= 0) cudaEventRecord((*g_prof)[slot_].e1, st_);
}

// ---- launch-configuration cache ------------------------------------------------------------------
static std::mutex g_cfg_mu;
static int g_sm_count[64] = {0};

int num_sms() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int n = g_sm_count[dev];
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        g_sm_count[dev] = n;
    }
    return n;
}

struct KernelKey {
    const void* k; int dev, threads; size_t smem;
    bool operator<(const KernelKey& o) const { return std::tie(k, dev, threads, smem) < std::tie(o.k, o.dev, o.threads, o.smem); }
};
static std::map<KernelKey, int> g_kernel_cfg;
static std::map<std::pair<const void*, int>, size_t> g_kernel_smem;  // largest dynamic smem opted into so far

int kernel_config(const void* kernel, int threads, size_t smem, int* ctas_per_sm) {
    int dev = 0;
    PP_CUDA(cudaGetDevice(&dev));
    const KernelKey key{kernel, dev, threads, smem};
    std::lock_guard<std::mutex> lock(g_cfg_mu);
    auto it = g_kernel_cfg.find(key);
    if (it != g_kernel_cfg.end()) { *ctas_per_sm = it->second; return PP_OK; }
    size_t& opted = g_kernel_smem[{kernel, dev}];
    if (smem > 48 * 1024 && smem > opted) {
        PP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        opted = smem;
    }
    int per_sm = 0;
    PP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    g_kernel_cfg[key] = per_sm;
    *ctas_per_sm = per_sm;
    return PP_OK;
}

}  // namespace pp

using namespace pp;

extern "C" const char* pp_last_error_string(void) { return g_err; }
extern "C" int pp_version(void) { return 100; }
extern "C" int64_t pp_launch_count(int reset) {
    const int64_t v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

extern "C" int pp_profile_start(void) {
    if (!g_prof) g_prof = new std::vector<ProfRec>();
    for (auto& r : *g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    g_prof->clear();
    g_prof_on = true;
    return PP_OK;
}

extern "C" int pp_profile_stop(char* names, size_t names_cap, float* ms, int max_records) {
    g_prof_on = false;
    if (!g_prof) return 0;
    int n = 0;
    size_t off = 0;
    if (names && names_cap) names[0] = 0;
    for (auto& r : *g_prof) {
        if (cudaEventSynchronize(r.e1) == cudaSuccess && n < max_records) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) {
                const size_t len = strlen(r.name);
                if (names && off + len + 2 <= names_cap) {
                    memcpy(names + off, r.name, len);
                    names[off + len] = '\n';
                    names[off + len + 1] = 0;
                    off += len + 1;
                    if (ms) ms[n] = t;
                    ++n;
                }
            }
        }
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    g_prof->clear();
    return n;
}

// ---- context ----------------------------------------------------------------------------------
// Host <-> device transfers of the *_host layer.  Caller memory that is already page-locked (cudaHostAlloc /
// cudaHostRegister / torch pinned tensors / pp_host_alloc) is handed to the copy engine directly.  Pageable memory
// (a fresh numpy array, the reference's case) is moved through the context's own pinned ring in pieces: a few host
// threads copy piece k+1 into the ring while the DMA of piece k is in flight, so the host copy overlaps the transfer
// instead of preceding it (a pageable cudaMemcpyAsync does the same inside the driver with one thread).
constexpr size_t kPiece = (size_t)2 << 20;   // bytes per staged piece
constexpr int kRing = 4;                      // pieces in flight
#ifndef PP_COPY_THREADS
#define PP_COPY_THREADS 12
#endif
constexpr int kCopyThreads = PP_COPY_THREADS;                // 407 040 f64 points in / 9.9 MB out on the 32-vCPU host of the GPU box: 2 threads 1.04 ms, 4: 1.02, 8: 0.75, 12: 0.69, 16: 0.69

static void parallel_memcpy(void* dst, const void* src, size_t bytes) {
    if (bytes < ((size_t)256 << 10)) { memcpy(dst, src, bytes); return; }
    const size_t per = (bytes / kCopyThreads + 4095) & ~(size_t)4095;
#pragma omp parallel for num_threads(kCopyThreads) schedule(static, 1)
    for (int t = 0; t < kCopyThreads; ++t) {
        const size_t o = (size_t)t * per;
        if (o < bytes) memcpy(static_cast<char*>(dst) + o, static_cast<const char*>(src) + o, bytes - o < per ? bytes - o : per);
    }
}

static bool is_page_locked(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

struct pp_ctx {
    int device;
    cudaStream_t stream;
    struct Buf { void* p = nullptr; size_t cap = 0; };
    Buf dev[12];
    void* ring = nullptr;            // kRing pinned pieces of kPiece bytes
    cudaEvent_t ring_ev[kRing] = {};  // the DMA that last used the piece
    bool ring_busy[kRing] = {};
    void* small = nullptr;           // pinned scratch for scalars read back
    // grow-only device buffers, one per slot
    int get(int slot, size_t bytes, void** out) {
        Buf& b = dev[slot];
        if (bytes < 256) bytes = 256;
        if (b.cap < bytes) {
            if (b.p) PP_CUDA(cudaFree(b.p));
            b.p = nullptr; b.cap = 0;
            const size_t want = bytes + bytes / 4;
            PP_CUDA(cudaMalloc(&b.p, want));
            b.cap = want;
        }
        *out = b.p;
        return PP_OK;
    }
    int ensure_ring() {
        if (ring) return PP_OK;
        PP_CUDA(cudaHostAlloc(&ring, kRing * kPiece, cudaHostAllocDefault));
        PP_CUDA(cudaHostAlloc(&small, 4096, cudaHostAllocDefault));
        for (int k = 0; k < kRing; ++k) PP_CUDA(cudaEventCreateWithFlags(&ring_ev[k], cudaEventDisableTiming));
        return PP_OK;
    }
    char* piece(int k) { return static_cast<char*>(ring) + (size_t)k * kPiece; }
    int wait_piece(int k) {
        if (ring_busy[k]) { PP_CUDA(cudaEventSynchronize(ring_ev[k])); ring_busy[k] = false; }
        return PP_OK;
    }
    // host -> device, asynchronous on the context stream; `src` may be reused as soon as this returns
    int h2d(void* dst_dev, const void* src, size_t bytes) {
        if (bytes == 0) return PP_OK;
        if (is_page_locked(src)) {
            PP_CUDA(cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyHostToDevice, stream));
            return PP_OK;
        }
        PP_TRY_RC(ensure_ring());
        int k = 0;
        for (size_t o = 0; o < bytes; o += kPiece, k = (k + 1) % kRing) {
            const size_t nb = bytes - o < kPiece ? bytes - o : kPiece;
            PP_TRY_RC(wait_piece(k));
            parallel_memcpy(piece(k), static_cast<const char*>(src) + o, nb);
            PP_CUDA(cudaMemcpyAsync(static_cast<char*>(dst_dev) + o, piece(k), nb, cudaMemcpyHostToDevice, stream));
            PP_CUDA(cudaEventRecord(ring_ev[k], stream));
            ring_busy[k] = true;
        }
        return PP_OK;
    }
    // device -> host; `dst` is complete when this returns (the stream has been drained up to the copy)
    int d2h(void* dst, const void* src_dev, size_t bytes) {
        if (bytes == 0) return PP_OK;
        if (is_page_locked(dst)) {
            PP_CUDA(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, stream));
            PP_CUDA(cudaStreamSynchronize(stream));
            return PP_OK;
        }
        PP_TRY_RC(ensure_ring());
        const size_t np = (bytes + kPiece - 1) / kPiece;
        for (int k = 0; k < kRing; ++k) PP_TRY_RC(wait_piece(k));
        size_t issued = 0;
        for (size_t done = 0; done < np; ++done) {
            for (; issued < np && issued < done + kRing; ++issued) {  // keep kRing DMAs in flight
                const int k = (int)(issued % kRing);
                const size_t o = issued * kPiece, nb = bytes - o < kPiece ? bytes - o : kPiece;
                PP_CUDA(cudaMemcpyAsync(piece(k), static_cast<const char*>(src_dev) + o, nb, cudaMemcpyDeviceToHost, stream));
                PP_CUDA(cudaEventRecord(ring_ev[k], stream));
            }
            const int k = (int)(done % kRing);
            const size_t o = done * kPiece, nb = bytes - o < kPiece ? bytes - o : kPiece;
            PP_CUDA(cudaEventSynchronize(ring_ev[k]));
            parallel_memcpy(static_cast<char*>(dst) + o, piece(k), nb);
        }
        return PP_OK;
    }
    // one small value back (voxel counts): pinned scratch + event wait, no full-stream drain semantics needed
    int read_back(void* dst, const void* src_dev, size_t bytes) {
        PP_TRY_RC(ensure_ring());
        PP_CUDA(cudaMemcpyAsync(small, src_dev, bytes, cudaMemcpyDeviceToHost, stream));
        PP_CUDA(cudaStreamSynchronize(stream));
        memcpy(dst, small, bytes);
        return PP_OK;
    }
};

extern "C" int pp_ctx_create(int device, pp_ctx** out) {
    PP_CHECK_ARG(out, "pp_ctx_create: null out");
    int ndev = 0;
    PP_CUDA(cudaGetDeviceCount(&ndev));
    PP_CHECK_ARG(device >= 0 && device < ndev, "pp_ctx_create: device %d of %d", device, ndev);
    PP_CUDA(cudaSetDevice(device));
    pp_ctx* c = new (std::nothrow) pp_ctx();
    if (!c) { set_error("out of host memory"); return PP_E_NOMEM; }
    c->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e));
        delete c;
        return PP_E_CUDA;
    }
    *out = c;
    return PP_OK;
}

extern "C" void pp_ctx_destroy(pp_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto& b : c->dev)
        if (b.p) cudaFree(b.p);
    if (c->ring) {
        cudaFreeHost(c->ring);
        cudaFreeHost(c->small);
        for (int k = 0; k < kRing; ++k) cudaEventDestroy(c->ring_ev[k]);
    }
    cudaStreamDestroy(c->stream);
    delete c;
}
extern "C" void* pp_ctx_stream(pp_ctx* c) { return c ? static_cast<void*>(c->stream) : nullptr; }
extern "C" int pp_ctx_device(pp_ctx* c) { return c ? c->device : -1; }
extern "C" int pp_ctx_sync(pp_ctx* c) {
    PP_CHECK_ARG(c, "null ctx");
    PP_CUDA(cudaStreamSynchronize(c->stream));
    return PP_OK;
}

// Page-locked host memory for callers that want the direct-DMA path of the *_host functions
// (numpy: np.frombuffer over the returned block).
extern "C" int pp_host_alloc(size_t bytes, void** out) {
    PP_CHECK_ARG(out, "pp_host_alloc: null out");
    PP_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return PP_OK;
}
extern "C" void pp_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

#define PP_TRY(expr)          \
    do {                      \
        int rc__ = (expr);    \
        if (rc__) return rc__; \
    } while (0)

#define PP_ENTER(c)                                   \
    PP_CHECK_ARG(c, "null ctx");                      \
    PP_CUDA(cudaSetDevice((c)->device));              \
    cudaStream_t st = (c)->stream;                    \
    (void)st

// ---- host layer ---------------------------------------------------------------------------------
extern "C" int pp_points_to_voxel_host(pp_ctx* c, const pp_voxel_cfg* cfg, const void* points,
                                       int point_dtype, int64_t N, int D, void* voxels, int32_t* coors,
                                       int32_t* num_points, int32_t* voxel_num_out, int32_t* point_slot) {
    PP_ENTER(c);
    PP_CHECK_ARG(cfg && voxels && coors && num_points && voxel_num_out && N >= 0, "pp_points_to_voxel_host: bad argument");
    PP_CHECK_ARG(point_dtype == PP_F32 || point_dtype == PP_F64, "bad point_dtype");
    const size_t esz = point_dtype == PP_F64 ? 8 : 4;
    const int P = cfg->max_points, MV = cfg->max_voxels;
    const size_t ws_bytes = pp_voxelize_workspace_bytes(cfg, N, 1, N, D, point_dtype);
    PP_CHECK_ARG(ws_bytes > 0, "pp_points_to_voxel_host: bad config");
    void *d_pts, *d_vox, *d_coors, *d_num, *d_misc, *d_ws, *d_slot = nullptr;
    PP_TRY(c->get(0, (size_t)N * D * esz, &d_pts));
    PP_TRY(c->get(1, (size_t)MV * P * D * esz, &d_vox));
    PP_TRY(c->get(2, (size_t)MV * 3 * 4, &d_coors));
    PP_TRY(c->get(3, (size_t)MV * 4, &d_num));
    PP_TRY(c->get(4, 256, &d_misc));  // frame_offsets[2] (int64) | voxel_num | voxel_base[2]
    PP_TRY(c->get(5, ws_bytes, &d_ws));
    if (point_slot) PP_TRY(c->get(6, (size_t)N * 4, &d_slot));
    int64_t* d_off = static_cast<int64_t*>(d_misc);
    int32_t* d_vnum = reinterpret_cast<int32_t*>(d_off + 2);
    int32_t* d_vbase = d_vnum + 1;
    const int64_t off[2] = {0, N};
    PP_CUDA(cudaMemcpyAsync(d_off, off, sizeof(off), cudaMemcpyHostToDevice, st));  // 16 bytes: copied at enqueue time
    PP_TRY(c->h2d(d_pts, points, (size_t)N * D * esz));
    PP_TRY(pp_voxelize_dev(cfg, d_pts, point_dtype, D, d_off, 1, N, N, point_dtype, d_vox, nullptr,
                           static_cast<int32_t*>(d_coors), 3, static_cast<int32_t*>(d_num), MV, d_vnum,
                           d_vbase, static_cast<int32_t*>(d_slot), nullptr, d_ws, ws_bytes, st));
    int32_t m = 0;
    PP_TRY(c->read_back(&m, d_vnum, 4));
    *voxel_num_out = m;
    // the three result arrays leave in one stream; the last copy drains it
    if (m > 0) {
        if (point_slot && N > 0) PP_CUDA(cudaMemcpyAsync(point_slot, d_slot, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
        PP_CUDA(cudaMemcpyAsync(coors, d_coors, (size_t)m * 12, cudaMemcpyDeviceToHost, st));
        PP_CUDA(cudaMemcpyAsync(num_points, d_num, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
        PP_TRY(c->d2h(voxels, d_vox, (size_t)m * P * D * esz));
    } else if (point_slot && N > 0) {
        PP_TRY(c->d2h(point_slot, d_slot, (size_t)N * 4));
    }
    PP_CUDA(cudaStreamSynchronize(st));
    return PP_OK;
}

extern "C" int pp_decorate_host(pp_ctx* c, const float* voxels, const int32_t* num_points, const int32_t* coors,
                                int64_t M, int P, int D, double vx, double vy, double x_offset, double y_offset,
                                float* out) {
    PP_ENTER(c);
    PP_CHECK_ARG(M >= 0 && P >= 1 && D >= 3, "pp_decorate_host: bad shape");
    if (M == 0) return PP_OK;
    void *d_v, *d_n, *d_c, *d_o;
    PP_TRY(c->get(0, (size_t)M * P * D * 4, &d_v));
    PP_TRY(c->get(1, (size_t)M * 4, &d_n));
    PP_TRY(c->get(2, (size_t)M * 16, &d_c));
    PP_TRY(c->get(3, (size_t)M * P * (D + 5) * 4, &d_o));
    PP_CUDA(cudaMemcpyAsync(d_v, voxels, (size_t)M * P * D * 4, cudaMemcpyHostToDevice, st));
    PP_CUDA(cudaMemcpyAsync(d_n, num_points, (size_t)M * 4, cudaMemcpyHostToDevice, st));
    PP_CUDA(cudaMemcpyAsync(d_c, coors, (size_t)M * 16, cudaMemcpyHostToDevice, st));
    PP_TRY(pp_decorate_dev(static_cast<float*>(d_v), static_cast<int32_t*>(d_n), static_cast<int32_t*>(d_c), M, P, D,
                           vx, vy, x_offset, y_offset, static_cast<float*>(d_o), st));
    PP_CUDA(cudaMemcpyAsync(out, d_o, (size_t)M * P * (D + 5) * 4, cudaMemcpyDeviceToHost, st));
    PP_CUDA(cudaStreamSynchronize(st));
    return PP_OK;
}

extern "C" int pp_scatter_host(pp_ctx* c, const float* feats, const int32_t* coords, int64_t M, int C, int B,
                               int ny, int nx, int layout, float* out) {
    PP_ENTER(c);
    PP_CHECK_ARG(M >= 0 && C > 0 && B > 0 && ny > 0 && nx > 0 && out, "pp_scatter_host: bad argument");
    const size_t ws_bytes = pp_scatter_workspace_bytes(B, ny, nx, M);
    const size_t out_bytes = (size_t)B * C * ny * nx * 4;
    void *d_f, *d_c, *d_o, *d_ws;
    PP_TRY(c->get(0, (size_t)M * C * 4, &d_f));
    PP_TRY(c->get(1, (size_t)M * 16, &d_c));
    PP_TRY(c->get(2, out_bytes, &d_o));
    PP_TRY(c->get(3, ws_bytes, &d_ws));
    if (M > 0) {
        PP_CUDA(cudaMemcpyAsync(d_f, feats, (size_t)M * C * 4, cudaMemcpyHostToDevice, st));
        PP_CUDA(cudaMemcpyAsync(d_c, coords, (size_t)M * 16, cudaMemcpyHostToDevice, st));
    }
    PP_TRY(pp_scatter_dev(static_cast<float*>(d_f), static_cast<int32_t*>(d_c), M, nullptr, C, B, ny, nx, layout,
                          static_cast<float*>(d_o), d_ws, ws_bytes, st));
    PP_CUDA(cudaMemcpyAsync(out, d_o, out_bytes, cudaMemcpyDeviceToHost, st));
    PP_CUDA(cudaStreamSynchronize(st));
    return PP_OK;
}

extern "C" int pp_box_decode_host(pp_ctx* c, const float* box_encodings, const float* anchors, int64_t N, float* out) {
    PP_ENTER(c);
    PP_CHECK_ARG(N >= 0, "pp_box_decode_host: N < 0");
    if (N == 0) return PP_OK;
    void *d_e, *d_a, *d_o;
    PP_TRY(c->get(0, (size_t)N * 28, &d_e));
    PP_TRY(c->get(1, (size_t)N * 28, &d_a));
    PP_TRY(c->get(2, (size_t)N * 28, &d_o));
    PP_CUDA(cudaMemcpyAsync(d_e, box_encodings, (size_t)N * 28, cudaMemcpyHostToDevice, st));
    PP_CUDA(cudaMemcpyAsync(d_a, anchors, (size_t)N * 28, cudaMemcpyHostToDevice, st));
    PP_TRY(pp_box_decode_dev(static_cast<float*>(d_e), static_cast<float*>(d_a), N, 0, static_cast<float*>(d_o), st));
    PP_CUDA(cudaMemcpyAsync(out, d_o, (size_t)N * 28, cudaMemcpyDeviceToHost, st));
    PP_CUDA(cudaStreamSynchronize(st));
    return PP_OK;
}

extern "C" int pp_rbox_to_standup_host(pp_ctx* c, const float* boxes, int64_t N, float* out) {
    PP_ENTER(c);
    PP_CHECK_ARG(N >= 0, "pp_rbox_to_standup_host: N < 0");
    if (N == 0) return PP_OK;
    void *d_b, *d_o;
    PP_TRY(c->get(0, (size_t)N * 20, &d_b));
    PP_TRY(c->get(1, (size_t)N * 16, &d_o));
    PP_CUDA(cudaMemcpyAsync(d_b, boxes, (size_t)N * 20, cudaMemcpyHostToDevice, st));
    PP_TRY(pp_rbox_to_standup_dev(static_cast<float*>(d_b), 5, N, static_cast<float*>(d_o), st));
    PP_CUDA(cudaMemcpyAsync(out, d_o, (size_t)N * 16, cudaMemcpyDeviceToHost, st));
    PP_CUDA(cudaStreamSynchronize(st));
    return PP_OK;
}

extern "C" int pp_nms_host(pp_ctx* c, int kind, const float* boxes, const float* scores, int64_t N,
                           int pre_max_size, int post_max_size, float thresh, int64_t* keep,
                           int32_t* keep_count_out) {
    PP_ENTER(c);
    PP_CHECK_ARG(keep_count_out && N >= 0, "pp_nms_host: bad argument");
    PP_CHECK_ARG(kind == PP_NMS_STANDUP || kind == PP_NMS_ROTATED, "pp_nms_host: bad kind");
    *keep_count_out = 0;
    if (N == 0) return PP_OK;
    const int bs = kind == PP_NMS_ROTATED ? 5 : 4;
    int64_t cap = N;
    if (pre_max_size > 0 && pre_max_size < cap) cap = pre_max_size;
    if (post_max_size > 0 && post_max_size < cap) cap = post_max_size;
    const size_t ws_bytes = pp_nms_workspace_bytes(kind, 1, N, pre_max_size);
    void *d_b, *d_s, *d_k, *d_ws;
    PP_TRY(c->get(0, (size_t)N * bs * 4, &d_b));
    PP_TRY(c->get(1, (size_t)N * 4, &d_s));
    PP_TRY(c->get(2, (size_t)(cap + 1) * 4, &d_k));
    PP_TRY(c->get(3, ws_bytes, &d_ws));
    PP_CUDA(cudaMemcpyAsync(d_b, boxes, (size_t)N * bs * 4, cudaMemcpyHostToDevice, st));
    PP_CUDA(cudaMemcpyAsync(d_s, scores, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    int32_t* d_keep = static_cast<int32_t*>(d_k);
    int32_t* d_cnt = d_keep + cap;
    PP_TRY(pp_nms_dev(kind, static_cast<float*>(d_b), bs, static_cast<float*>(d_s), nullptr, 1, N, pre_max_size,
                      post_max_size, thresh, d_keep, cap, d_cnt, d_ws, ws_bytes, st));
    // keep indices + count in one copy
    int32_t* h = static_cast<int32_t*>(malloc((size_t)(cap + 1) * 4));
    if (!h) { set_error("out of host memory"); return PP_E_NOMEM; }
    cudaError_t e = cudaMemcpyAsync(h, d_keep, (size_t)(cap + 1) * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        free(h);
        set_error("pp_nms_host: %s", cudaGetErrorString(e));
        return PP_E_CUDA;
    }
    const int32_t k = h[cap];
    for (int32_t i = 0; i < k; ++i) keep[i] = h[i];
    free(h);
    *keep_count_out = k;
    return PP_OK;
}

extern "C" int pp_anchors_mask_host(pp_ctx* c, const int32_t* coors, int64_t M, const float* anchors, int64_t A,
                                    const double voxel_size[3], const double coors_range[6], float threshold,
                                    float* area_out, uint8_t* mask_out) {
    PP_ENTER(c);
    PP_CHECK_ARG(M >= 0 && A >= 0 && voxel_size && coors_range, "pp_anchors_mask_host: bad argument");
    if (A == 0) return PP_OK;
    int32_t grid[3];
    pp_grid_size(voxel_size, coors_range, 0, grid);
    const size_t ws_bytes = pp_anchor_mask_workspace_bytes(1, grid[1], grid[0]);
    void *d_co, *d_an, *d_cells, *d_area, *d_mask, *d_ws;
    PP_TRY(c->get(0, (size_t)M * 12, &d_co));
    PP_TRY(c->get(1, (size_t)A * 28, &d_an));
    PP_TRY(c->get(2, (size_t)A * 16, &d_cells));
    PP_TRY(c->get(3, (size_t)A * 4, &d_area));
    PP_TRY(c->get(4, (size_t)A, &d_mask));
    PP_TRY(c->get(5, ws_bytes, &d_ws));
    if (M > 0) PP_CUDA(cudaMemcpyAsync(d_co, coors, (size_t)M * 12, cudaMemcpyHostToDevice, st));
    PP_CUDA(cudaMemcpyAsync(d_an, anchors, (size_t)A * 28, cudaMemcpyHostToDevice, st));
    PP_TRY(pp_anchor_cells_dev(static_cast<float*>(d_an), A, voxel_size, coors_range, static_cast<int32_t*>(d_cells), st));
    PP_TRY(pp_anchor_mask_dev(static_cast<int32_t*>(d_co), 3, M, nullptr, 1, grid[1], grid[0], static_cast<int32_t*>(d_cells),
                              A, threshold, nullptr, static_cast<float*>(d_area), static_cast<uint8_t*>(d_mask), nullptr,
                              d_ws, ws_bytes, st));
    if (area_out) PP_CUDA(cudaMemcpyAsync(area_out, d_area, (size_t)A * 4, cudaMemcpyDeviceToHost, st));
    if (mask_out) PP_CUDA(cudaMemcpyAsync(mask_out, d_mask, (size_t)A, cudaMemcpyDeviceToHost, st));
    PP_CUDA(cudaStreamSynchronize(st));
    return PP_OK;
}

extern "C" int pp_d3_box_overlap_host(pp_ctx* c, const double* boxes, int64_t N, const double* query_boxes, int64_t K,
                                      int criterion, float* out) {
    PP_ENTER(c);
    PP_CHECK_ARG(N >= 0 && K >= 0, "pp_d3_box_overlap_host: bad argument");
    if (N == 0 || K == 0) return PP_OK;
    void *d_b, *d_q, *d_o;
    PP_TRY(c->get(0, (size_t)N * 56, &d_b));
    PP_TRY(c->get(1, (size_t)K * 56, &d_q));
    PP_TRY(c->get(2, (size_t)N * K * 4, &d_o));
    PP_CUDA(cudaMemcpyAsync(d_b, boxes, (size_t)N * 56, cudaMemcpyHostToDevice, st));
    PP_CUDA(cudaMemcpyAsync(d_q, query_boxes, (size_t)K * 56, cudaMemcpyHostToDevice, st));
    PP_TRY(pp_d3_box_overlap_dev(static_cast<double*>(d_b), N, static_cast<double*>(d_q), K, criterion,
                                 static_cast<float*>(d_o), st));
    PP_CUDA(cudaMemcpyAsync(out, d_o, (size_t)N * K * 4, cudaMemcpyDeviceToHost, st));
    PP_CUDA(cudaStreamSynchronize(st));
    return PP_OK;
}

extern "C" int pp_rotate_iou_host(pp_ctx* c, const float* boxes, int64_t N, const float* query_boxes, int64_t K,
                                  int criterion, float* out) {
    PP_ENTER(c);
    PP_CHECK_ARG(N >= 0 && K >= 0, "pp_rotate_iou_host: bad argument");
    if (N == 0 || K == 0) return PP_OK;
    void *d_b, *d_q, *d_o;
    PP_TRY(c->get(0, (size_t)N * 20, &d_b));
    PP_TRY(c->get(1, (size_t)K * 20, &d_q));
    PP_TRY(c->get(2, (size_t)N * K * 4, &d_o));
    PP_CUDA(cudaMemcpyAsync(d_b, boxes, (size_t)N * 20, cudaMemcpyHostToDevice, st));
    PP_CUDA(cudaMemcpyAsync(d_q, query_boxes, (size_t)K * 20, cudaMemcpyHostToDevice, st));
    PP_TRY(pp_rotate_iou_dev(static_cast<float*>(d_b), N, static_cast<float*>(d_q), K, criterion,
                             static_cast<float*>(d_o), st));
    PP_CUDA(cudaMemcpyAsync(out, d_o, (size_t)N * K * 4, cudaMemcpyDeviceToHost, st));
    PP_CUDA(cudaStreamSynchronize(st));
    return PP_OK;
}

extern "C" int pp_predict_host(pp_ctx* c, const pp_predict_cfg* cfg, const float* box_preds, const float* cls_preds,
                               const float* dir_preds, const float* anchors, const uint8_t* anchors_mask,
                               const float* rect, const float* Trv2c, int B, int64_t A, int K, float* box3d_lidar,
                               double* box3d_camera, float* scores, int32_t* label_preds, int32_t* anchor_index,
                               int32_t* count) {
    PP_ENTER(c);
    PP_CHECK_ARG(cfg && B > 0 && A >= 0 && K > 0 && box3d_lidar && count, "pp_predict_host: bad argument");
    const size_t nbox = (size_t)B * A;
    const size_t n_anch = cfg->anchors_per_frame ? nbox : (size_t)A;
    const size_t ws_bytes = pp_predict_workspace_bytes(cfg, B, A, K);
    void *d_bp, *d_cls, *d_dir = nullptr, *d_an, *d_mask = nullptr, *d_rect = nullptr, *d_trv = nullptr, *d_ws, *d_out;
    PP_TRY(c->get(0, nbox * 28, &d_bp));
    PP_TRY(c->get(1, nbox * 4 * cfg->num_class, &d_cls));
    PP_TRY(c->get(3, n_anch * 28, &d_an));
    PP_TRY(c->get(7, ws_bytes, &d_ws));
    if (A > 0) {
        PP_CHECK_ARG(box_preds && cls_preds && anchors, "pp_predict_host: null input");
        PP_CUDA(cudaMemcpyAsync(d_bp, box_preds, nbox * 28, cudaMemcpyHostToDevice, st));
        PP_CUDA(cudaMemcpyAsync(d_cls, cls_preds, nbox * 4 * cfg->num_class, cudaMemcpyHostToDevice, st));
        PP_CUDA(cudaMemcpyAsync(d_an, anchors, n_anch * 28, cudaMemcpyHostToDevice, st));
        if (dir_preds && cfg->use_direction_classifier) {
            PP_TRY(c->get(2, nbox * 8, &d_dir));
            PP_CUDA(cudaMemcpyAsync(d_dir, dir_preds, nbox * 8, cudaMemcpyHostToDevice, st));
        }
        if (anchors_mask) {
            PP_TRY(c->get(4, nbox, &d_mask));
            PP_CUDA(cudaMemcpyAsync(d_mask, anchors_mask, nbox, cudaMemcpyHostToDevice, st));
        }
    }
    if (rect && Trv2c) {
        PP_TRY(c->get(5, (size_t)B * 64, &d_rect));
        PP_TRY(c->get(6, (size_t)B * 64, &d_trv));
        PP_CUDA(cudaMemcpyAsync(d_rect, rect, (size_t)B * 64, cudaMemcpyHostToDevice, st));
        PP_CUDA(cudaMemcpyAsync(d_trv, Trv2c, (size_t)B * 64, cudaMemcpyHostToDevice, st));
    }
    const size_t rows = (size_t)B * K;
    Carver sz(nullptr);
    sz.take<double>(rows * 7); sz.take<float>(rows * 7); sz.take<float>(rows); sz.take<int32_t>(rows); sz.take<int32_t>(rows);
    sz.take<int32_t>(B);
    PP_TRY(c->get(8, sz.used(), &d_out));
    Carver o(d_out);
    double* o_cam = o.take<double>(rows * 7);
    float* o_lid = o.take<float>(rows * 7);
    float* o_sc = o.take<float>(rows);
    int32_t* o_lab = o.take<int32_t>(rows);
    int32_t* o_idx = o.take<int32_t>(rows);
    int32_t* o_cnt = o.take<int32_t>(B);
    PP_TRY(pp_predict_dev(cfg, static_cast<float*>(d_bp), static_cast<float*>(d_cls), static_cast<float*>(d_dir),
                          static_cast<float*>(d_an), static_cast<uint8_t*>(d_mask), static_cast<float*>(d_rect),
                          static_cast<float*>(d_trv), B, A, K, o_lid, box3d_camera ? o_cam : nullptr, o_sc, o_lab, o_idx, o_cnt,
                          d_ws, ws_bytes, st));
    PP_CUDA(cudaMemcpyAsync(box3d_lidar, o_lid, rows * 28, cudaMemcpyDeviceToHost, st));
    if (box3d_camera) PP_CUDA(cudaMemcpyAsync(box3d_camera, o_cam, rows * 56, cudaMemcpyDeviceToHost, st));
    if (scores) PP_CUDA(cudaMemcpyAsync(scores, o_sc, rows * 4, cudaMemcpyDeviceToHost, st));
    if (label_preds) PP_CUDA(cudaMemcpyAsync(label_preds, o_lab, rows * 4, cudaMemcpyDeviceToHost, st));
    if (anchor_index) PP_CUDA(cudaMemcpyAsync(anchor_index, o_idx, rows * 4, cudaMemcpyDeviceToHost, st));
    PP_CUDA(cudaMemcpyAsync(count, o_cnt, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    PP_CUDA(cudaStreamSynchronize(st));
    return PP_OK;
}

extern "C" int pp_ingest_host(pp_ctx* c, const void* cloud, int64_t n_in, int point_step, int off_x, int off_y, int off_z,
                              int start, int step, const double* rotations, int n_rot, const double* translation,
                              double* points_out, int64_t cap, int32_t* n_out) {
    PP_ENTER(c);
    PP_CHECK_ARG(n_in >= 0 && point_step > 0 && cap >= 0 && n_out && (points_out || cap == 0), "pp_ingest_host: bad argument");
    const size_t ws_bytes = pp_ingest_workspace_bytes(1, n_in);
    void *d_c, *d_o, *d_n, *d_ws;
    PP_TRY(c->get(0, (size_t)n_in * point_step, &d_c));
    PP_TRY(c->get(1, (size_t)cap * 24, &d_o));
    PP_TRY(c->get(2, 256, &d_n));
    PP_TRY(c->get(3, ws_bytes, &d_ws));
    if (n_in > 0) PP_CUDA(cudaMemcpyAsync(d_c, cloud, (size_t)n_in * point_step, cudaMemcpyHostToDevice, st));
    PP_TRY(pp_ingest_dev(d_c, 1, n_in, point_step, off_x, off_y, off_z, start, step, rotations, n_rot, translation,
                         static_cast<double*>(d_o), cap, static_cast<int32_t*>(d_n), d_ws, ws_bytes, st));
    int32_t n = 0;
    PP_CUDA(cudaMemcpyAsync(&n, d_n, 4, cudaMemcpyDeviceToHost, st));
    PP_CUDA(cudaStreamSynchronize(st));
    *n_out = n;
    if (n > 0) {
        PP_CUDA(cudaMemcpyAsync(points_out, d_o, (size_t)n * 24, cudaMemcpyDeviceToHost, st));
        PP_CUDA(cudaStreamSynchronize(st));
    }
    return PP_OK;
}
