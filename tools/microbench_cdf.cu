// Cycle cost of the sampler's serial CDF (build_cdf_warp) on one warp, by section, and of alternatives.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -I odam_b200/csrc -I include -o /tmp/mb_cdf tools/microbench_cdf.cu && /tmp/mb_cdf
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "odam_sq.h"
#include "sq_math.cuh"
#include "sq_device.cuh"
using namespace odam;

__global__ void k(long long *out, float *sink, int reps)
{
    __shared__ GridTab g;
    __shared__ __align__(16) float cdf[kGPad];
    const int lane = threadIdx.x;
    for (int i = lane; i < kGPad; i += 32) g.slot[i] = make_float4(0.f, 0.001f * (i % 7 + 1), 0.f, 0.f);
    __syncwarp();
    long long t[6] = {0, 0, 0, 0, 0, 0};
    for (int r = 0; r < reps; r++) {
        long long c0 = clock64();
        build_cdf_warp(g, cdf, 1.3f, lane);
        long long c1 = clock64();
        t[0] += c1 - c0;
        // sections
        c0 = clock64();
        for (int i = lane; i < kGPad; i += 32) cdf[i] = i < kG ? __fmul_rn(1.3f, g.slot[i].y) : 0.f;
        __syncwarp();
        c1 = clock64();
        t[1] += c1 - c0;
        if (lane == 0) {
            float c = 0.001f;
            for (int i = 1; i < kG; i++) { c = __fadd_rn(__fadd_rn(c, 0.001f), cdf[i]); cdf[i] = c; }
        }
        __syncwarp();
        long long c2 = clock64();
        t[2] += c2 - c1;
        // pure register chain of 402 dependent FADDs
        float c = cdf[3];
#pragma unroll 8
        for (int i = 0; i < 402; i++) c = __fadd_rn(c, 0.001f);
        long long c3 = clock64();
        t[3] += c3 - c2;
        if (c == 1234.5f) sink[0] = c;
        float s = cdf[kG - 1];
        float mine[7];
#pragma unroll
        for (int k2 = 0; k2 < 7; k2++) { int i = lane + 32 * k2; mine[k2] = i < kG ? __fdiv_rn(cdf[i], s) : 0.f; }
        __syncwarp();
#pragma unroll
        for (int k2 = 0; k2 < 7; k2++) { int i = lane + 32 * k2; if (i < kG) cdf[i] = mine[k2]; }
        __syncwarp();
        long long c4 = clock64();
        t[4] += c4 - c3;
    }
    if (lane == 0) for (int i = 0; i < 6; i++) out[i] = t[i] / reps;
    if (cdf[lane] == 1234.5f) sink[1] = cdf[lane];
}
int main()
{
    long long *o; float *s; cudaMalloc(&o, 64); cudaMalloc(&s, 8);
    k<<<1, 32>>>(o, s, 50); cudaDeviceSynchronize();
    long long h[6]; cudaMemcpy(h, o, 48, cudaMemcpyDeviceToHost);
    printf("build_cdf_warp %lld cycles | fill %lld | naive serial scan %lld | 402 dependent FADD in regs %lld | normalise %lld\n",
           h[0], h[1], h[2], h[3], h[4]);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
