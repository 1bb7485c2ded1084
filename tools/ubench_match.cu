// micro-benchmark: cost of __match_any_sync on sm_100a (distinct vs few distinct values), 1..8 warps per CTA
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int mode, int iters, unsigned* out, long long* cyc) {
    int lane = threadIdx.x & 31;
    unsigned acc = 0;
    int v = mode == 0 ? lane : (mode == 1 ? (lane & 3) : (lane * 7919 + blockIdx.x) % 645);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        unsigned m = __match_any_sync(0xffffffffu, v + (acc & 1));
        acc += __popc(m);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
    unsigned* out; long long* cyc; cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
    for (int warps : {1, 4, 8, 16}) for (int mode : {0, 1, 2}) for (int blocks : {1, 444}) {
        k<<<blocks, warps * 32>>>(mode, 2000, out, cyc);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("warps/CTA %2d blocks %3d mode %d : %.1f cycles per match (per warp)\n", warps, blocks, mode, c / 2000.0);
    }
    return 0;
}
