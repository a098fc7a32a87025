// Issue cost of the instruction classes phase E and the sampler use, alone and interleaved 1:1 with FFMA, on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o build/mbp tools/microbench_pipes.cu && build/mbp
// Full occupancy (16 warps per scheduler), 8 independent chains per thread; prints SM cycles per instruction per SM
// sub-partition: 1.0 = full rate, 2.0 = half rate; "mixed" < sum of the two alone means the pipes overlap.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 2048;
constexpr int kChains = 8;

enum Op { FFMA, FADD, FMUL, FMNMX, FMNMX3, IADD3, LOP3, IMNMX, FSETSEL, MUFU_RCP, DFMA, SHFL, LDS32, F2I, NOP_ };

template <int OP>
__device__ __forceinline__ void op(float &x, float a, float b, int &i, double &d, const float *sm)
{
    if (OP == FFMA) x = __fmaf_rn(x, a, b);
    if (OP == FADD) x = __fadd_rn(x, a);
    if (OP == FMUL) x = __fmul_rn(x, a);
    if (OP == FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(x) : "f"(a));
    if (OP == FMNMX3) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(a), "f"(b));
    if (OP == IADD3) asm volatile("add.s32 %0, %0, %1;" : "+r"(i) : "r"(__float_as_int(a)));
    if (OP == LOP3) asm volatile("xor.b32 %0, %0, %1;" : "+r"(i) : "r"(__float_as_int(a)));
    if (OP == IMNMX) asm volatile("min.s32 %0, %0, %1;" : "+r"(i) : "r"(__float_as_int(a)));
    if (OP == FSETSEL) x = x > a ? x : b;
    if (OP == MUFU_RCP) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x));
    if (OP == DFMA) d = __fma_rn(d, 0.999, 0.001);
    if (OP == SHFL) x = __shfl_xor_sync(0xffffffffu, x, 1);
    if (OP == LDS32) x = sm[(__float_as_int(x) & 1023)];
    if (OP == F2I) i = __float2int_rn(x) + i;
}

template <int A, int B>
__global__ void __launch_bounds__(1024) k(float *sink, float a, float b)
{
    __shared__ float sm[1024];
    sm[threadIdx.x] = (float)threadIdx.x;
    __syncthreads();
    float x[kChains], y[kChains];
    int ii[kChains];
    double dd[kChains];
#pragma unroll
    for (int c = 0; c < kChains; c++) { x[c] = threadIdx.x + c; y[c] = x[c] * 0.5f; ii[c] = c; dd[c] = c; }
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int c = 0; c < kChains; c++) {
            op<A>(x[c], a, b, ii[c], dd[c], sm);
            if (B != NOP_) op<B>(y[c], a, b, ii[c], dd[c], sm);
        }
    }
    float r = 0.f;
#pragma unroll
    for (int c = 0; c < kChains; c++) r += x[c] + y[c] + (float)ii[c] + (float)dd[c];
    if (r == 123.456f) sink[0] = r;
}

template <int A, int B>
static double run(int sms, double ghz, float *sink)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        k<A, B><<<sms * 2, 1024>>>(sink, 0.999f, 0.001f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const int per_iter = kChains * (B == NOP_ ? 1 : 2);
    return best * 1e-3 * ghz * 1e9 / ((double)kIters * 16 * per_iter);   // cycles per instruction per scheduler
}

#define ALONE(NAME, OPC) printf("%-10s alone %5.2f   with FFMA 1:1: %5.2f cycles per pair\n", NAME, run<OPC, NOP_>(sms, ghz, sink), 2 * run<FFMA, OPC>(sms, ghz, sink))

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz / 1e6;
    const int sms = prop.multiProcessorCount;
    float *sink;
    cudaMalloc(&sink, 4096);
    printf("%s: cycles per warp instruction per scheduler (clock attr %.3f GHz; loop overhead included)\n", prop.name, ghz);
    ALONE("FFMA", FFMA);
    ALONE("FADD", FADD);
    ALONE("FMUL", FMUL);
    ALONE("FMNMX", FMNMX);
    ALONE("FMNMX3", FMNMX3);
    ALONE("IADD", IADD3);
    ALONE("XOR", LOP3);
    ALONE("IMNMX", IMNMX);
    ALONE("FSETP+SEL", FSETSEL);
    ALONE("MUFU.RCP", MUFU_RCP);
    ALONE("DFMA", DFMA);
    ALONE("SHFL", SHFL);
    ALONE("LDS", LDS32);
    ALONE("F2I+IADD", F2I);
    return 0;
}
