// Shared host/device helpers for libpp_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "pp_b200.h"

namespace pp {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// Per-launch CUDA-event timing (pp_profile_start/stop): when profiling is on, every kernel launch
// is bracketed by two events on the launching stream.  Off by default: one thread-local branch.
struct LaunchTimer {
    LaunchTimer(const char* name, cudaStream_t st);
    ~LaunchTimer();
    cudaStream_t st_;
    int slot_;
};
#define PP_TIMED(name, st) pp::LaunchTimer pp_timer__(name, st)

#define PP_CHECK_ARG(cond, ...)                \
    do {                                       \
        if (!(cond)) {                         \
            pp::set_error(__VA_ARGS__);        \
            return PP_E_INVALID;               \
        }                                      \
    } while (0)

#define PP_CUDA(expr)                                                                      \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            pp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                          __LINE__);                                                       \
            return PP_E_CUDA;                                                              \
        }                                                                                  \
    } while (0)

// after a kernel launch
#define PP_LAUNCHED()                                                                        \
    do {                                                                                     \
        pp::count_launch();                                                                  \
        cudaError_t e__ = cudaGetLastError();                                                \
        if (e__ != cudaSuccess) {                                                            \
            pp::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, \
                          __LINE__);                                                         \
            return PP_E_CUDA;                                                                \
        }                                                                                    \
    } while (0)

// Multiprocessor count of the current device (148 on B200), queried once per device.
int num_sms();
// Per-process cache of a kernel's launch configuration: raises the dynamic shared-memory limit when `smem`
// needs it and returns the resident CTAs per SM for (threads, smem).  cudaFuncSetAttribute and the occupancy
// query cost microseconds each, which matters on the single-frame path; they run once per (kernel, smem, device).
int kernel_config(const void* kernel, int threads, size_t smem, int* ctas_per_sm);

#define PP_TRY_RC(expr)        \
    do {                       \
        int rc__ = (expr);     \
        if (rc__) return rc__; \
    } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace.
struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(void* p) : base(static_cast<char*>(p)) {}
    template <typename T>
    T* take(size_t n) {
        off = align_up(off, 256);
        T* r = reinterpret_cast<T*>(base + off);
        off += n * sizeof(T);
        return r;
    }
    size_t used() const { return align_up(off, 256); }
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
// ---- TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP / SYNCS) --------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    // dst, src 16-byte aligned, bytes a multiple of 16
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// ---- TMA 1-D bulk copy shared -> global (bulk async-group completion) ---------------------------------
__device__ __forceinline__ void tma_bulk_s2g(void* dst_gmem, const void* src_smem, unsigned bytes) {
    // dst, src 16-byte aligned, bytes a multiple of 16
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (TMA) that reads them
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// block-wide exclusive scan of one int per thread (blockDim.x multiple of 32, <= 1024)
__device__ __forceinline__ int block_excl_scan(int v, int* total, int* sm /*[33]*/) {
    const int lane = lane_id(), w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) sm[w] = inc;
    __syncthreads();
    if (w == 0) {
        const int nw = blockDim.x >> 5;
        int x = lane < nw ? sm[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += u;
        }
        if (lane < nw) sm[lane] = x;  // inclusive over warps
        if (lane == 31) sm[32] = x;
    }
    __syncthreads();
    const int woff = w ? sm[w - 1] : 0;
    *total = sm[32];
    const int r = woff + inc - v;
    __syncthreads();
    return r;
}
#endif

}  // namespace pp
