// Latency of warp-level primitives the voxelizer leans on (one warp, dependent chains, clock64).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int mode, int* out, long long* cyc, const int* vals) {
    __shared__ unsigned char tbl[16384];
    for (int i = threadIdx.x; i < 16384; i += 32) tbl[i] = 0;
    __syncwarp();
    int v = vals[threadIdx.x];
    int acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < 256; ++it) {
        if (mode == 0) {  // match_any, result feeds the next value (dependent)
            unsigned m = __match_any_sync(0xffffffffu, v);
            acc += __popc(m);
            v += (int)(m & 0);
        } else if (mode == 1) {  // ballot chain
            unsigned m = __ballot_sync(0xffffffffu, v & 1);
            acc += __popc(m);
            v += (int)(m & 0);
        } else if (mode == 2) {  // LDS.U8 -> STS.U8 -> syncwarp chain on the table
            int c = (v * 37 + it) & 16383;
            int cnt = tbl[c];
            tbl[c] = (unsigned char)(cnt + 1);
            __syncwarp();
            acc += cnt;
        } else if (mode == 3) {  // shfl chain
            v = __shfl_sync(0xffffffffu, v, (threadIdx.x + 1) & 31);
            acc += v;
        } else if (mode == 4) {  // 4 independent match_any per iteration
            unsigned m0 = __match_any_sync(0xffffffffu, v);
            unsigned m1 = __match_any_sync(0xffffffffu, v ^ 1);
            unsigned m2 = __match_any_sync(0xffffffffu, v ^ 2);
            unsigned m3 = __match_any_sync(0xffffffffu, v ^ 3);
            acc += __popc(m0) + __popc(m1) + __popc(m2) + __popc(m3);
            v += (int)((m0 ^ m1 ^ m2 ^ m3) & 0);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) *cyc = t1 - t0;
    out[threadIdx.x] = acc + v;
}
int main() {
    int *out, *vals; long long* cyc;
    cudaMalloc(&out, 128); cudaMalloc(&vals, 128); cudaMalloc(&cyc, 8);
    const char* names[] = {"match_any", "ballot", "lds_sts_syncwarp", "shfl", "match_any x4 independent"};
    for (int pat = 0; pat < 4; ++pat) {
        int h[32];
        for (int i = 0; i < 32; ++i) h[i] = pat == 0 ? i * 977 : pat == 1 ? 5 : pat == 2 ? (i / 4) * 131 : (i & 1) * 999;
        cudaMemcpy(vals, h, 128, cudaMemcpyHostToDevice);
        const char* pn[] = {"32 distinct", "all equal", "8 groups of 4", "2 groups"};
        for (int mode = 0; mode < 5; ++mode) {
            if (pat > 0 && mode != 0 && mode != 4) continue;
            long long c = 0;
            k<<<1, 32>>>(mode, out, cyc, vals); cudaDeviceSynchronize();
            k<<<1, 32>>>(mode, out, cyc, vals); cudaDeviceSynchronize();
            cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-28s %-14s %.1f cycles/iteration\n", names[mode], pn[pat], c / 256.0);
        }
    }
    return 0;
}
