// Throughput (not latency) of the ways to find, inside a warp, the lanes that hold the same key:
// many warps per SM, independent iterations.  Prints cycles per warp-step per SM.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned lanemask_lt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

template <int MODE>
__global__ void k(const int* vals, int iters, int* out, long long* cyc) {
    const int lane = threadIdx.x & 31;
    int v = vals[lane] + (threadIdx.x >> 5) * 7;
    int acc = 0;
    __shared__ unsigned scr[32][33];
    const int w = threadIdx.x >> 5;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const int c = (v + it * 3) & 0x3fff;
        if (MODE == 0) {
            const unsigned m = __match_any_sync(0xffffffffu, c);
            acc += __popc(m & lanemask_lt()) + __popc(m);
        } else if (MODE == 1) {  // one ballot per key bit
            unsigned peers = 0xffffffffu;
#pragma unroll
            for (int bit = 0; bit < 14; ++bit) {
                const unsigned bal = __ballot_sync(0xffffffffu, (c >> bit) & 1);
                peers &= ((c >> bit) & 1) ? bal : ~bal;
            }
            acc += __popc(peers & lanemask_lt()) + __popc(peers);
        } else if (MODE == 2) {  // bitonic sort of (key << 5 | lane), run heads by ballot, results back through shared memory
            unsigned x = ((unsigned)c << 5) | lane;
#pragma unroll
            for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
                for (int j = kk >> 1; j >= 1; j >>= 1) {
                    const unsigned o = __shfl_xor_sync(0xffffffffu, x, j);
                    const bool asc = (lane & kk) == 0, low = (lane & j) == 0;
                    x = (asc == low) ? min(x, o) : max(x, o);
                }
            }
            const unsigned prev = __shfl_up_sync(0xffffffffu, x, 1);
            const bool head = lane == 0 || (prev >> 5) != (x >> 5);
            const unsigned heads = __ballot_sync(0xffffffffu, head);
            const int start = 31 - __clz(heads & (lanemask_lt() | (1u << lane)));
            const unsigned after = heads & ~((2u << lane) - 1u);
            const int end = after ? __ffs(after) - 1 : 32;
            scr[w][x & 31] = (unsigned)(lane - start) | ((unsigned)(end - start) << 8);
            __syncwarp();
            const unsigned r = scr[w][lane];
            __syncwarp();
            acc += (r & 255) + (r >> 8);
        } else if (MODE == 3) {  // shared-memory bit masks keyed by a folded key (atomicOr), no verification
            const int hsh = c & 31;  // worst case: tiny table, every step verifies
            atomicOr(&scr[w][hsh], 1u << lane);
            __syncwarp();
            const unsigned m = scr[w][hsh];
            __syncwarp();
            scr[w][hsh] = 0u;
            __syncwarp();
            acc += __popc(m & lanemask_lt()) + __popc(m);
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
    int *out, *vals; long long* cyc;
    cudaMalloc(&out, 4 * 148 * 1024 * 4); cudaMalloc(&vals, 128); cudaMalloc(&cyc, 8);
    const char* names[] = {"match_any", "14 ballots", "bitonic sort + heads", "smem atomicOr masks"};
    const int iters = 2000;
    for (int pat = 0; pat < 3; ++pat) {
        int h[32];
        for (int i = 0; i < 32; ++i) h[i] = pat == 0 ? i * 977 : pat == 1 ? (i % 28) * 131 : (i / 8) * 131;
        cudaMemcpy(vals, h, 128, cudaMemcpyHostToDevice);
        const char* pn[] = {"32 distinct", "28 distinct", "4 distinct"};
        for (int warps = 4; warps <= 32; warps *= 2) {
            for (int mode = 0; mode < 4; ++mode) {
                long long c = 0;
                auto launch = [&]() {
                    if (mode == 0) k<0><<<148, warps * 32>>>(vals, iters, out, cyc);
                    if (mode == 1) k<1><<<148, warps * 32>>>(vals, iters, out, cyc);
                    if (mode == 2) k<2><<<148, warps * 32>>>(vals, iters, out, cyc);
                    if (mode == 3) k<3><<<148, warps * 32>>>(vals, iters, out, cyc);
                };
                launch(); cudaDeviceSynchronize();
                launch(); cudaDeviceSynchronize();
                cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
                printf("%-22s %-12s %2d warps/SM: %7.1f cycles per step per warp, %6.1f cycles per step per SM\n", names[mode], pn[pat],
                       warps, c / (double)iters, c / (double)iters / warps);
            }
        }
    }
    return 0;
}
