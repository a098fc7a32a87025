// Pipe throughput on sm_100a for the instruction mix of phase E: scalar FFMA vs packed FFMA2, alone and mixed with
// the other pipes the scan loop uses (FMNMX3 on the ALU pipe, MUFU.RCP on the XU pipe).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o /tmp/mb2 tools/microbench_ffma2.cu && /tmp/mb2
// Prints, per variant, SM cycles per loop iteration per SM sub-partition at full occupancy (8 warps per scheduler),
// i.e. the reciprocal issue/pipe throughput of the mix, and the equivalent FP32 TFLOP/s for the FMA-only variants.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpk2(uint64_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fmin3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float rcpa(float d) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d)); return r; }

constexpr int kIters = 4096;

// V: 0 = 16 scalar FFMA; 1 = 8 FFMA2 (same flops); 2 = 16 FFMA2; 3 = 8 FFMA2 + 8 FMNMX3; 4 = 8 FFMA2 + 8 FMNMX3 + 2 MUFU
//    5 = 16 FFMA + 8 FMNMX3; 6 = the E mix per 4 points: 18 FFMA2 + 4 MUFU + 4 FMUL2(as FFMA2) + 8 FMNMX3
template <int V>
__global__ void __launch_bounds__(1024) k(float *sink, float a, float b)
{
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x + i;
    uint64_t p[8];
#pragma unroll
    for (int i = 0; i < 8; i++) p[i] = pk2(x[2 * i], x[2 * i + 1]);
    float m0 = a, m1 = b, m2 = a + b, m3 = a - b;
    const uint64_t A = pk2(a, a), B = pk2(b, b);
    for (int it = 0; it < kIters; it++) {
        if (V == 0 || V == 5) {
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = __fmaf_rn(x[i], a, b);
        }
        if (V == 1 || V == 3 || V == 4) {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = ffma2(p[i], A, B);
        }
        if (V == 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = ffma2(p[i], A, B);
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = ffma2(p[i], B, A);
        }
        if (V == 6) {
#pragma unroll
            for (int r = 0; r < 2; r++) {
#pragma unroll
                for (int i = 0; i < 8; i++) p[i] = ffma2(p[i], A, B);
            }
            p[0] = ffma2(p[0], B, A); p[1] = ffma2(p[1], B, A);
            p[2] = ffma2(p[2], B, A); p[3] = ffma2(p[3], B, A); p[4] = ffma2(p[4], B, A); p[5] = ffma2(p[5], B, A);
        }
        if (V == 3 || V == 4 || V == 5 || V == 6) {
            float l0, h0, l1, h1;
            if (V == 5) { l0 = x[0]; h0 = x[1]; l1 = x[2]; h1 = x[3]; }
            else { unpk2(p[0], l0, h0); unpk2(p[1], l1, h1); }
            m0 = fmin3(m0, l0, h0); m1 = fmin3(m1, l1, h1); m2 = fmin3(m2, l0, h1); m3 = fmin3(m3, l1, h0);
            m0 = fmin3(m0, l1, h1); m1 = fmin3(m1, l0, h0); m2 = fmin3(m2, l1, h0); m3 = fmin3(m3, l0, h1);
        }
        if (V == 4 || V == 6) {
            float l, h;
            unpk2(p[2], l, h);
            l = rcpa(l); h = rcpa(h);
            p[2] = pk2(l, h);
            if (V == 6) { unpk2(p[3], l, h); l = rcpa(l); h = rcpa(h); p[3] = pk2(l, h); }
        }
    }
    float r = m0 + m1 + m2 + m3;
#pragma unroll
    for (int i = 0; i < 16; i++) r += x[i];
#pragma unroll
    for (int i = 0; i < 8; i++) { float l, h; unpk2(p[i], l, h); r += l + h; }
    if (r == 123.456f) sink[0] = r;
}

template <int V>
static void run(const char *name, double flop_per_iter, int sms, double ghz, float *sink)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 2, threads = 1024;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k<V><<<blocks, threads>>>(sink, 0.999f, 0.001f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    // 64 warps per SM = 16 per scheduler; cycles per loop iteration per scheduler = time * clock / (iters * 16)
    const double cyc = best * 1e-3 * ghz * 1e9 / ((double)kIters * 16);
    printf("%-52s %8.3f ms  %6.2f cycles/iter/SMSP", name, best, cyc);
    if (flop_per_iter > 0) printf("  %7.2f TFLOP/s", flop_per_iter * kIters * (double)blocks * threads / (best * 1e-3) / 1e12);
    printf("\n");
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz / 1e6;
    float *sink;
    cudaMalloc(&sink, 4096);
    printf("%s, %d SMs, clock attr %.3f GHz (cycles assume that clock)\n", prop.name, prop.multiProcessorCount, ghz);
    run<0>("16 FFMA", 32, prop.multiProcessorCount, ghz, sink);
    run<1>("8 FFMA2 (same flops as 16 FFMA)", 32, prop.multiProcessorCount, ghz, sink);
    run<2>("16 FFMA2", 64, prop.multiProcessorCount, ghz, sink);
    run<5>("16 FFMA + 8 FMNMX3", 0, prop.multiProcessorCount, ghz, sink);
    run<3>("8 FFMA2 + 8 FMNMX3", 0, prop.multiProcessorCount, ghz, sink);
    run<4>("8 FFMA2 + 8 FMNMX3 + 2 MUFU.RCP", 0, prop.multiProcessorCount, ghz, sink);
    run<6>("E mix / 4 points: 22 FFMA2 + 8 FMNMX3 + 4 MUFU.RCP", 0, prop.multiProcessorCount, ghz, sink);
    return 0;
}
