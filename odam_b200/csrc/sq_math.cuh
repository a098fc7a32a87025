// sq_math.cuh -- lean fp64 evaluation of the three transcendentals the sampler needs, rounded once to fp32.
//
// The reference's sampler (fast_sampler/sampling.cpp:59-67) calls glibc's float cosf/sinf/powf on
//   theta in [-pi, pi]            (grid angles),
//   x in [0, 1], p in [0.2, 1.6]  (|cos|^e, |sin|^e).
// Those routines are within 0.56 ulp of the exact value; here each function is evaluated in double precision
// with a relative error < 2^-45 on exactly that domain and rounded once to float, i.e. the correctly rounded
// float in all but ~2^-20 of the cases.  (CUDA's generic double pow()/sincos() do the same job in ~4x the
// instructions and ~2x the registers because they carry every special case; the restricted domain needs none.)
//
// Host+device so that tests/ can check the very same code on the CPU against extended precision
// (tests/test_sq_math.py through csrc/sq_math_host.cpp).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

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

// sign(c) * |c|^p as the reference's fexp (sampling.cpp:59-61), float in / float out
SQ_HD float sq_signed_pow(float c, float p)
{
    float r = (float)sq_pow01(fabsf(c), p);
    return copysignf(r, c);
}

// float angle -> (sign(cos)|cos|^e, sign(sin)|sin|^e) with cosf/sinf rounded to float in between, as the
// reference does (powf(fabsf(cosf(theta)), e))
SQ_HD void sq_grid_node(float theta, float e, float &fc, float &fs)
{
    double s, c;
    sq_sincos_pi(theta, s, c);
    fc = sq_signed_pow((float)c, e);
    fs = sq_signed_pow((float)s, e);
}

}  // namespace odam
