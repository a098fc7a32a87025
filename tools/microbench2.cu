// FFMA throughput by operand form on sm_100a (full chip): reg*const+const vs reg*reg+reg.
#include <cstdio>
#include <cuda_runtime.h>
template <int FORM>
__global__ void __launch_bounds__(1024) k(float *sink, int iters, float a, float b)
{
    float x[8], y[8], z[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i; y[i] = 1.0f + 1e-7f * (threadIdx.x + i); z[i] = 1e-9f * i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (FORM == 0) x[i] = __fmaf_rn(x[i], a, b);
                if (FORM == 1) x[i] = __fmaf_rn(x[i], y[i], z[i]);
                if (FORM == 2) x[i] = __fmaf_rn(y[i], z[(i + 1) & 7], x[i]);
                if (FORM == 3) x[i] = fminf(fminf(x[i], y[i]), z[i]) + 0.f * a;
            }
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r += x[i];
    if (r == 123.456f) sink[0] = r;
}
int main()
{
    float *s; cudaMalloc(&s, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *names[] = {"x=fma(x,const,const)", "x=fma(x,y,z) 3 regs", "x=fma(y,z',x) 3 regs", "min3-ish"};
    for (int form = 0; form < 3; form++) {
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            int iters = 2048, blocks = 148 * 2;
            if (form == 0) k<0><<<blocks, 1024>>>(s, iters, 0.999f, 0.001f);
            if (form == 1) k<1><<<blocks, 1024>>>(s, iters, 0.999f, 0.001f);
            if (form == 2) k<2><<<blocks, 1024>>>(s, iters, 0.999f, 0.001f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fl = 2.0 * 64 * iters * blocks * 1024.0;
            if (rep == 2) printf("%-26s %.1f TFLOP/s\n", names[form], fl / ms / 1e9);
        }
    }
    return 0;
}
