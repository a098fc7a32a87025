// Latency calibration on sm_100a: cycles per dependent op for the instruction classes the sampler chain uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o /tmp/microbench tools/microbench.cu && /tmp/microbench
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template <int OP>
__global__ void k(double *out, long long *cyc, double seed, float fseed)
{
    __shared__ float sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (float)((i * 7 + 1) & 1023);
    __syncthreads();
    double x = seed; float f = fseed; int idx = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        if (OP == 0) x = __fma_rn(x, 0.999999, 1e-9);
        if (OP == 1) f = __fmaf_rn(f, 0.999999f, 1e-9f);
        if (OP == 2) { idx = (int)sm[idx & 1023]; }
        if (OP == 3) f = __fsqrt_rn(f + 1.0f);
        if (OP == 4) f = __fdiv_rn(f, 1.000001f) + 0.5f;
        if (OP == 5) f = roundf(f * 1.0001f + 0.3f);
        if (OP == 6) x = (double)(float)x * 1.0000001;
        if (OP == 7) x = rint(x * 1.0000001) + 0.25;
        if (OP == 8) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f)); f += 1.0f; }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[OP] = t1 - t0; }
    out[threadIdx.x] = x + f + idx;
}
int main()
{
    double *o; long long *c;
    cudaMalloc(&o, 8 * 1024); cudaMallocManaged(&c, 8 * 16);
    const char *names[] = {"DFMA", "FFMA", "LDS+I2F chain", "FADD+sqrt_rn", "fdiv_rn+FADD", "FFMA+roundf", "F2F.f32.f64+f64.f32+DMUL", "DMUL+rint+DADD", "MUFU.RCP+FADD"};
    for (int warps = 1; warps <= 16; warps *= 4) {
        k<0><<<1, 32 * warps>>>(o, c, 1.0, 1.0f); k<1><<<1, 32 * warps>>>(o, c, 1.0, 1.0f); k<2><<<1, 32 * warps>>>(o, c, 1.0, 1.0f);
        k<3><<<1, 32 * warps>>>(o, c, 1.0, 1.0f); k<4><<<1, 32 * warps>>>(o, c, 1.0, 1.0f); k<5><<<1, 32 * warps>>>(o, c, 1.0, 1.0f);
        k<6><<<1, 32 * warps>>>(o, c, 1.0, 1.0f); k<7><<<1, 32 * warps>>>(o, c, 1.0, 1.0f); k<8><<<1, 32 * warps>>>(o, c, 1.0, 1.0f);
        cudaDeviceSynchronize();
        for (int i = 0; i < 9; i++) printf("warps=%2d %-28s %.1f cycles/iter\n", warps, names[i], (double)c[i] / N);
    }
    return 0;
}
