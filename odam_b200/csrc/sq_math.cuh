// sq_math.cuh -- the transcendentals of the path, in double precision, host+device.
//
// (1) libm-faithful cosf / sinf / powf for the SAMPLER: the reference's C++ sampler (fast_sampler/sampling.cpp:59-67)
//     calls glibc's float routines on theta in [-pi, pi] and x in [0,1], p in [0.2,1.6]; its discrete decisions
//     (roundf split, CDF bucket) depend on the last bit of those results, so the device runs glibc's own
//     double-precision algorithm, operation for operation (sq_glibc_*; bit-identical to libm, tests/test_sq_math.py).
// (2) correctly rounded helpers for everything that mirrors torch ops instead (sigmoid, yaw sin/cos, logs for the
//     backward pass): lean fp64 with relative error < 2^-45, rounded once to float (sq_sincos_pi, sq_exp_neg, sq_log01).
//
// Host+device so that tests/ can check the very same code on the CPU against extended precision
// (tests/test_sq_math.py through csrc/sq_math_host.cpp).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "sq_glibc_data.h"

#ifdef __CUDACC__
#define SQ_HD __host__ __device__ __forceinline__
#else
#define SQ_HD static inline
#endif

namespace odam {

SQ_HD double sq_fma(double a, double b, double c)
{
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}

SQ_HD double sq_bits_to_double(uint64_t u)
{
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, 8);
    return d;
#endif
}
SQ_HD uint64_t sq_double_to_bits(double d)
{
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u;
    memcpy(&u, &d, 8);
    return u;
#endif
}

// sin and cos of a float angle, |theta| <= ~3.2 (two-term Cody-Waite reduction by pi/2, then the classic
// degree-13/14 minimax kernels on |r| <= pi/4).  Results in double, |error| < 1 ulp(double).
SQ_HD void sq_sincos_pi(float theta, double &s, double &c)
{
    const double x = (double)theta;
    const double two_over_pi = 0.63661977236758134308;
    const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17;
    const double kf = rint(x * two_over_pi);
    const int k = (int)kf;
    double r = sq_fma(-kf, pio2_hi, x);
    r = sq_fma(-kf, pio2_lo, r);
    const double z = r * r;
    // sin(r) = r + r^3 * S(z)
    double ps = 1.58969099521155010221e-10;
    ps = sq_fma(ps, z, -2.50507602534068634195e-08);
    ps = sq_fma(ps, z, 2.75573137070700676789e-06);
    ps = sq_fma(ps, z, -1.98412698298579493134e-04);
    ps = sq_fma(ps, z, 8.33333333332248946124e-03);
    ps = sq_fma(ps, z, -1.66666666666666324348e-01);
    const double sr = sq_fma(r * z, ps, r);
    // cos(r) = 1 - z/2 + z^2 * C(z)
    double pc = -1.13596475577881948265e-11;
    pc = sq_fma(pc, z, 2.08757232129817482790e-09);
    pc = sq_fma(pc, z, -2.75573143513906633035e-07);
    pc = sq_fma(pc, z, 2.48015872894767294178e-05);
    pc = sq_fma(pc, z, -1.38888888888741095749e-03);
    pc = sq_fma(pc, z, 4.16666666666666019037e-02);
    const double cr = sq_fma(z * z, pc, sq_fma(-0.5, z, 1.0));
    switch (k & 3) {
        case 0: s = sr; c = cr; break;
        case 1: s = cr; c = -sr; break;
        case 2: s = -sr; c = -cr; break;
        default: s = -cr; c = sr; break;
    }
}

// Taylor coefficients 1/13! .. 1/2! of exp, then ln2 split and log2(e).  On the device they sit in constant
// memory so that every DFMA takes its coefficient as a c[bank][offset] operand (an immediate double costs two
// extra issue slots per use).
#define SQ_EXP_COEFS                                                                                          \
    {1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0,   \
     1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5,                                       \
     6.93147180369123816490e-01, 1.90821492927058770002e-10, 1.44269504088896338700e+00, 0.0}
#ifdef __CUDACC__
__constant__ double kSqExpC[16] = SQ_EXP_COEFS;
#endif
static const double kSqExpH[16] = SQ_EXP_COEFS;
#ifdef __CUDA_ARCH__
#define SQ_EC(i) kSqExpC[i]
#else
#define SQ_EC(i) kSqExpH[i]
#endif

// exp(y) for -700 < y <= 0:  y = k ln2 + r, |r| <= ln2/2, Taylor to r^13, scaled by 2^k through the exponent field
SQ_HD double sq_exp_neg(double y)
{
    const double kf = rint(y * SQ_EC(14));
    double r = sq_fma(-kf, SQ_EC(12), y);
    r = sq_fma(-kf, SQ_EC(13), r);
    double e = SQ_EC(0);
    e = sq_fma(e, r, SQ_EC(1));
    e = sq_fma(e, r, SQ_EC(2));
    e = sq_fma(e, r, SQ_EC(3));
    e = sq_fma(e, r, SQ_EC(4));
    e = sq_fma(e, r, SQ_EC(5));
    e = sq_fma(e, r, SQ_EC(6));
    e = sq_fma(e, r, SQ_EC(7));
    e = sq_fma(e, r, SQ_EC(8));
    e = sq_fma(e, r, SQ_EC(9));
    e = sq_fma(e, r, SQ_EC(10));
    e = sq_fma(e, r, SQ_EC(11));
    e = sq_fma(e, r, 1.0);
    e = sq_fma(e, r, 1.0);
    const int k = (int)kf;  // never subnormal on the caller's domain
    return sq_bits_to_double(sq_double_to_bits(e) + ((uint64_t)(int64_t)k << 52));
}

// log(x) for a float 0 < x <= 1, in double:  x = 2^E * m, m in [sqrt(.5), sqrt(2));
// log m = 2 atanh(s), s = (m-1)/(m+1), odd series to s^17
SQ_HD double sq_log01(float xf)
{
    const double x = (double)xf;
    uint64_t ux = sq_double_to_bits(x);
    int E = (int)((ux >> 52) & 0x7ff) - 1023;
    uint64_t um = (ux & 0x000fffffffffffffull) | 0x3ff0000000000000ull;
    double m = sq_bits_to_double(um);
    if (m > 1.41421356237309514547) { m *= 0.5; E += 1; }
    const double f = m - 1.0;
    const double s = f / (m + 1.0);
    const double z = s * s;
    double q = 1.0 / 17.0;
    q = sq_fma(q, z, 1.0 / 15.0);
    q = sq_fma(q, z, 1.0 / 13.0);
    q = sq_fma(q, z, 1.0 / 11.0);
    q = sq_fma(q, z, 1.0 / 9.0);
    q = sq_fma(q, z, 1.0 / 7.0);
    q = sq_fma(q, z, 1.0 / 5.0);
    q = sq_fma(q, z, 1.0 / 3.0);
    const double lm = sq_fma(2.0 * s * z, q, 2.0 * s);          // log(m)
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double dE = (double)E;
    return sq_fma(dE, ln2_hi, sq_fma(dE, ln2_lo, lm));
}

// x^p for 0 <= x <= 1 (float), 0 < p < 2 (float): exp(p * log(x)) in double
SQ_HD double sq_pow01(float xf, float pf)
{
    if (xf == 0.0f) return 0.0;
    return sq_exp_neg((double)pf * sq_log01(xf));
}


// ================================================================================================================
// glibc-faithful cosf / sinf / powf (x86-64, glibc >= 2.28, the FMA ifunc variants __cosf_fma, __sinf_fma, __powf_fma).
// Same double-precision operation sequence as the compiled library -- every multiply, add and fused multiply-add
// in the same place (transcribed from the disassembly of glibc 2.39) -- and the same tables, so the float results are
// bit-identical to what the reference's C++ sampler gets from libm on an FMA-capable x86-64 host.  Domains: the ones
// the sampler uses (|theta| < 120 for cos/sin; x a positive normal float, 0.2 <= y <= 1.6 for pow).
// ================================================================================================================
static const uint64_t kGlibcSincosH[28] = SQ_GLIBC_SINCOS;
static const uint64_t kGlibcLog2TabH[32] = SQ_GLIBC_LOG2TAB;
static const uint64_t kGlibcPolyH[9] = SQ_GLIBC_POLY;
static const uint64_t kGlibcExp2TabH[32] = SQ_GLIBC_EXP2TAB;
#ifdef __CUDACC__
__constant__ uint64_t kGlibcSincosD[28] = SQ_GLIBC_SINCOS;
__device__ const uint64_t kGlibcLog2TabD[32] = SQ_GLIBC_LOG2TAB;   // indexed per lane -> global/L1, not constant
__constant__ uint64_t kGlibcPolyD[9] = SQ_GLIBC_POLY;
__device__ const uint64_t kGlibcExp2TabD[32] = SQ_GLIBC_EXP2TAB;
#endif
#ifdef __CUDA_ARCH__
#define SQ_GT(name, i) sq_bits_to_double(name##D[i])
#define SQ_GTU(name, i) name##D[i]
#else
#define SQ_GT(name, i) sq_bits_to_double(name##H[i])
#define SQ_GTU(name, i) name##H[i]
#endif

SQ_HD double sq_mul(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
SQ_HD double sq_add(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
SQ_HD uint32_t sq_float_bits(float f)
{
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
SQ_HD float sq_bits_float(uint32_t u)
{
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

// table index: 6 c0, 7 c1, 8 s1, 9 c2, 10 s2, 11 c3, 12 s3, 13 c4 (+14 for the second, negated, table)
SQ_HD float sq_glibc_cos_poly(double x2, int b)
{
    const double x4 = sq_mul(x2, x2);
    const double A = sq_fma(x2, SQ_GT(kGlibcSincos, b + 7), SQ_GT(kGlibcSincos, b + 6));
    const double B = sq_fma(x2, SQ_GT(kGlibcSincos, b + 13), SQ_GT(kGlibcSincos, b + 11));
    const double x6 = sq_mul(x2, x4);
    const double C = sq_fma(x4, SQ_GT(kGlibcSincos, b + 9), A);
    return (float)sq_fma(B, x6, C);
}
SQ_HD float sq_glibc_sin_poly(double xs, double x2, int b)
{
    const double S = sq_fma(x2, SQ_GT(kGlibcSincos, b + 12), SQ_GT(kGlibcSincos, b + 10));
    const double x3 = sq_mul(x2, xs);
    const double x5 = sq_mul(x3, x2);
    const double L = sq_fma(x3, SQ_GT(kGlibcSincos, b + 8), xs);
    return (float)sq_fma(S, x5, L);
}
// reduce_fast: n = round(x * 2/pi) via the 2^24-scaled constant, x - n * pi/2 in one fused step
SQ_HD double sq_glibc_reduce(double x, int &n)
{
    const double r = sq_mul(x, SQ_GT(kGlibcSincos, 4));
    n = ((int)r + 0x800000) >> 24;
    return sq_fma(-(double)n, SQ_GT(kGlibcSincos, 5), x);
}
SQ_HD float sq_glibc_cosf(float xf)
{
    const uint32_t top = (sq_float_bits(xf) >> 20) & 0x7ff;
    const double x = (double)xf;
    if (top <= 0x3f3) {
        if (top <= 0x397) return 1.0f;
        return sq_glibc_cos_poly(sq_mul(x, x), 0);
    }
    int n;
    const double xr = sq_glibc_reduce(x, n);   // |x| < 120 assumed (top <= 0x42e)
    const int b = (n & 2) ? 14 : 0;
    const double x2 = sq_mul(xr, xr);
    if ((n & 1) == 0) return sq_glibc_cos_poly(x2, b);
    return sq_glibc_sin_poly(sq_mul(xr, SQ_GT(kGlibcSincos, n & 3)), x2, b);
}
SQ_HD float sq_glibc_sinf(float xf)
{
    const uint32_t top = (sq_float_bits(xf) >> 20) & 0x7ff;
    const double x = (double)xf;
    if (top <= 0x3f3) {
        if (top <= 0x397) return xf;
        return sq_glibc_sin_poly(x, sq_mul(x, x), 0);
    }
    int n;
    const double xr = sq_glibc_reduce(x, n);
    const int b = (n & 2) ? 14 : 0;
    const double x2 = sq_mul(xr, xr);
    if (n & 1) return sq_glibc_cos_poly(x2, b);
    return sq_glibc_sin_poly(sq_mul(xr, SQ_GT(kGlibcSincos, n & 3)), x2, b);
}
// powf's log2_inline: log2 of a positive normal float, as the double glibc multiplies by y
SQ_HD double sq_glibc_log2(float xf)
{
    const uint32_t ix = sq_float_bits(xf);
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) & 15u);
    const uint32_t top = tmp & 0xff800000u;
    const uint32_t iz = ix - top;
    const int k = (int)top >> 23;
    const double z = (double)sq_bits_float(iz);
    const double r = sq_fma(z, SQ_GT(kGlibcLog2Tab, 2 * i), -1.0);
    const double y0 = sq_add((double)k, SQ_GT(kGlibcLog2Tab, 2 * i + 1));
    const double p01 = sq_fma(r, SQ_GT(kGlibcPoly, 0), SQ_GT(kGlibcPoly, 1));
    const double p23 = sq_fma(r, SQ_GT(kGlibcPoly, 2), SQ_GT(kGlibcPoly, 3));
    const double r2 = sq_mul(r, r);
    const double q = sq_fma(r, SQ_GT(kGlibcPoly, 4), y0);
    const double r4 = sq_mul(r2, r2);
    const double q2 = sq_fma(r2, p23, q);
    return sq_fma(p01, r4, q2);
}
// powf's exp2_inline (sign_bias = 0) rounded to float: 2^(ylogx) for |ylogx| < 126
SQ_HD float sq_glibc_exp2(double ylogx)
{
    const double shift = SQ_GT(kGlibcPoly, 5);
    double kd = sq_add(ylogx, shift);
    const uint64_t ki = sq_double_to_bits(kd);
    kd = sq_add(kd, -shift);
    const double r = sq_add(ylogx, -kd);
    const uint64_t t = SQ_GTU(kGlibcExp2Tab, (int)(ki & 31)) + (ki << 47);
    const double s = sq_bits_to_double(t);
    const double z = sq_fma(SQ_GT(kGlibcPoly, 6), r, SQ_GT(kGlibcPoly, 7));
    const double r2 = sq_mul(r, r);
    double y = sq_fma(r, SQ_GT(kGlibcPoly, 8), 1.0);
    y = sq_fma(z, r2, y);
    return (float)sq_mul(y, s);
}
// powf(x, y) for 0 <= x (normal or zero), y in the sampler's range
SQ_HD float sq_glibc_powf(float x, float y)
{
    if (x == 0.0f) return 0.0f;
    return sq_glibc_exp2(sq_mul((double)y, sq_glibc_log2(x)));
}

// sign(c) * |c|^p as the reference's fexp (sampling.cpp:59-61), float in / float out, libm-faithful
SQ_HD float sq_signed_pow(float c, float p)
{
    return copysignf(sq_glibc_powf(fabsf(c), p), c);
}

// float angle -> (sign(cos)|cos|^e, sign(sin)|sin|^e) exactly as sampling.cpp:64-67 gets them from libm:
// powf(fabsf(cosf(theta)), e) with cosf/sinf rounded to float in between
SQ_HD void sq_grid_node(float theta, float e, float &fc, float &fs)
{
    fc = sq_signed_pow(sq_glibc_cosf(theta), e);
    fs = sq_signed_pow(sq_glibc_sinf(theta), e);
}

}  // namespace odam
