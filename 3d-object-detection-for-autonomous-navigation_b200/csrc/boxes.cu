// Anchor box decoding and rotated->standup conversion (float32 elementwise passes).
//   decode:  second_box_decode, libraries/eval_helper_functions.py:388-461 of the reference
//   standup: center_to_corner_box2d + corner_to_standup_nd_jit, load_data.py:1525-1594, 1330-1341
// Rows are 7 floats (28 bytes), so a block stages 256 rows through shared memory with 16-byte
// loads/stores; arithmetic uses explicit _rn intrinsics so the compiler cannot contract the
// numpy two-step multiply/add into an FMA.
#include "box_math.cuh"

namespace pp {

constexpr int kBoxThreads = 256;

// cooperative copy of `nfloat` floats global->shared (vectorised when 16-byte aligned)
__device__ __forceinline__ void stage_in(float* sm, const float* g, int nfloat) {
    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
        const int n4 = nfloat >> 2;
        for (int k = threadIdx.x; k < n4; k += blockDim.x)
            reinterpret_cast<float4*>(sm)[k] = __ldg(reinterpret_cast<const float4*>(g) + k);
        for (int k = (n4 << 2) + threadIdx.x; k < nfloat; k += blockDim.x) sm[k] = g[k];
    } else {
        for (int k = threadIdx.x; k < nfloat; k += blockDim.x) sm[k] = g[k];
    }
}
__device__ __forceinline__ void stage_out(float* g, const float* sm, int nfloat) {
    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
        const int n4 = nfloat >> 2;
        for (int k = threadIdx.x; k < n4; k += blockDim.x)
            reinterpret_cast<float4*>(g)[k] = reinterpret_cast<const float4*>(sm)[k];
        for (int k = (n4 << 2) + threadIdx.x; k < nfloat; k += blockDim.x) g[k] = sm[k];
    } else {
        for (int k = threadIdx.x; k < nfloat; k += blockDim.x) g[k] = sm[k];
    }
}

__global__ void __launch_bounds__(kBoxThreads)
box_decode_kernel(const float* __restrict__ enc, const float* __restrict__ anchors, int64_t N,
                  int64_t period, float* __restrict__ out) {
    __shared__ __align__(16) float s_t[kBoxThreads * 7];
    __shared__ __align__(16) float s_a[kBoxThreads * 7];
    const int64_t base = (int64_t)blockIdx.x * kBoxThreads;
    const int m = (int)min((int64_t)kBoxThreads, N - base);
    stage_in(s_t, enc + base * 7, m * 7);
    if (period <= 0) {
        stage_in(s_a, anchors + base * 7, m * 7);
    } else {
        for (int k = threadIdx.x; k < m * 7; k += kBoxThreads) {
            const int r = k / 7;
            s_a[k] = __ldg(anchors + ((base + r) % period) * 7 + (k - r * 7));
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < m) {
        float* t = s_t + threadIdx.x * 7;
        const float* a = s_a + threadIdx.x * 7;
        box_decode_one(t, a, t);
    }
    __syncthreads();
    stage_out(out + base * 7, s_t, m * 7);
}

__global__ void __launch_bounds__(kBoxThreads)
rbox_to_standup_kernel(const float* __restrict__ boxes, int stride, int64_t N, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * kBoxThreads + threadIdx.x;
    if (i >= N) return;
    const float* b = boxes + i * stride;
    float cx, cy, w, l, r;
    if (stride == 7) { cx = b[0]; cy = b[1]; w = b[3]; l = b[4]; r = b[6]; }
    else { cx = b[0]; cy = b[1]; w = b[2]; l = b[3]; r = b[4]; }
    reinterpret_cast<float4*>(out)[i] = rbox_standup_one(cx, cy, w, l, r);
}

}  // namespace pp

using namespace pp;

extern "C" int pp_box_decode_dev(const float* box_encodings, const float* anchors, int64_t N,
                                 int64_t anchor_period, float* out, void* stream) {
    PP_CHECK_ARG(N >= 0, "pp_box_decode_dev: N < 0");
    if (N == 0) return PP_OK;
    PP_CHECK_ARG(box_encodings && anchors && out, "pp_box_decode_dev: null argument");
    PP_TIMED("box_decode", static_cast<cudaStream_t>(stream));
    box_decode_kernel<<<(unsigned)ceil_div(N, kBoxThreads), kBoxThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        box_encodings, anchors, N, anchor_period, out);
    PP_LAUNCHED();
    return PP_OK;
}

extern "C" int pp_rbox_to_standup_dev(const float* boxes, int in_stride, int64_t N, float* out, void* stream) {
    PP_CHECK_ARG(N >= 0 && (in_stride == 5 || in_stride == 7), "pp_rbox_to_standup_dev: bad arguments");
    if (N == 0) return PP_OK;
    PP_CHECK_ARG(boxes && out && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                 "pp_rbox_to_standup_dev: null or unaligned argument");
    PP_TIMED("rbox_to_standup", static_cast<cudaStream_t>(stream));
    rbox_to_standup_kernel<<<(unsigned)ceil_div(N, kBoxThreads), kBoxThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        boxes, in_stride, N, out);
    PP_LAUNCHED();
    return PP_OK;
}
