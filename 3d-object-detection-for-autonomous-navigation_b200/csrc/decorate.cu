// PillarFeatureNet point decoration (model/pointpillars.py:143-203 of the reference) as one
// pass: a warp stages one pillar row [P,D] in shared memory, reduces the xyz sums, and streams
// the decorated row [P,D+5] out with consecutive lanes on consecutive addresses.
// The reference runs ~15 TensorFlow kernels and materialises six [M,P,*] intermediates here.
#include "pp_common.cuh"

namespace pp {

constexpr int kDecWarps = 8;

__global__ void __launch_bounds__(kDecWarps * 32)
decorate_kernel(const float* __restrict__ voxels, const int* __restrict__ num_points,
                const int* __restrict__ coors, int64_t M, int P, int D, float vx, float vy,
                float x_off, float y_off, float* __restrict__ out) {
    extern __shared__ float dsm[];
    const int lane = lane_id(), w = threadIdx.x >> 5;
    float* row = dsm + (size_t)w * P * D;
    const int nin = P * D, Do = D + 5, nout = P * Do;
    for (int64_t m = (int64_t)blockIdx.x * kDecWarps + w; m < M; m += (int64_t)gridDim.x * kDecWarps) {
        const float* src = voxels + m * nin;
        float sx = 0.f, sy = 0.f, sz = 0.f;
        for (int k = lane; k < nin; k += 32) {
            const float v = src[k];
            row[k] = v;
            const int d = k % D;
            // the reference sums ALL P slots (padding is zero), line 143
            sx += d == 0 ? v : 0.f;
            sy += d == 1 ? v : 0.f;
            sz += d == 2 ? v : 0.f;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            sx += __shfl_xor_sync(0xffffffffu, sx, o);
            sy += __shfl_xor_sync(0xffffffffu, sy, o);
            sz += __shfl_xor_sync(0xffffffffu, sz, o);
        }
        const int np = num_points[m];
        const float nf = (float)np;
        const float mx = __fdiv_rn(sx, nf), my = __fdiv_rn(sy, nf), mz = __fdiv_rn(sz, nf);
        // separate multiply and add, as two TF ops (lines 160-161, 169-170)
        const float ex = __fadd_rn(__fmul_rn((float)coors[4 * m + 3], vx), x_off);
        const float ey = __fadd_rn(__fmul_rn((float)coors[4 * m + 2], vy), y_off);
        __syncwarp();
        float* dst = out + m * nout;
        for (int k = lane; k < nout; k += 32) {
            const int s = k / Do, d = k - s * Do;
            const float* q = row + s * D;
            float v;
            if (d < D) v = q[d];
            else if (d == D) v = q[0] - mx;
            else if (d == D + 1) v = q[1] - my;
            else if (d == D + 2) v = q[2] - mz;
            else if (d == D + 3) v = q[0] - ex;
            else v = q[1] - ey;
            // mask multiply (lines 199-203) keeps the reference's NaN behaviour for num_points==0
            dst[k] = __fmul_rn(v, s < np ? 1.f : 0.f);
        }
        __syncwarp();
    }
}

}  // namespace pp

using namespace pp;

extern "C" int pp_decorate_dev(const float* voxels, const int32_t* num_points, const int32_t* coors,
                               int64_t M, int P, int D, double vx, double vy, double x_offset,
                               double y_offset, float* out, void* stream) {
    PP_CHECK_ARG(M >= 0 && P >= 1 && D >= 3 && D <= 16, "pp_decorate_dev: bad shape M=%lld P=%d D=%d",
                 (long long)M, P, D);
    if (M == 0) return PP_OK;
    PP_CHECK_ARG(voxels && num_points && coors && out, "pp_decorate_dev: null argument");
    const size_t smem = (size_t)kDecWarps * P * D * sizeof(float);
    PP_CHECK_ARG(smem <= 200 * 1024, "pp_decorate_dev: P*D too large");
    if (smem > 48 * 1024)
        PP_CUDA(cudaFuncSetAttribute(decorate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = ceil_div(M, kDecWarps);
    if (blocks > (int64_t)num_sms() * 8) blocks = (int64_t)num_sms() * 8;
    PP_TIMED("decorate", static_cast<cudaStream_t>(stream));
    decorate_kernel<<<(unsigned)blocks, kDecWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
        voxels, num_points, coors, M, P, D, (float)vx, (float)vy, (float)x_offset, (float)y_offset, out);
    PP_LAUNCHED();
    return PP_OK;
}
