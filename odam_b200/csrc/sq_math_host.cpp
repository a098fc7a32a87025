// Host build of sq_math.cuh for CPU-side unit tests (tests/test_sq_math.py).  Not part of the product path.
#include "sq_math.cuh"

extern "C" {
void sq_math_sincos(const float *theta, int n, double *s, double *c)
{
    for (int i = 0; i < n; i++) odam::sq_sincos_pi(theta[i], s[i], c[i]);
}
void sq_math_pow01(const float *x, const float *p, int n, double *out)
{
    for (int i = 0; i < n; i++) out[i] = odam::sq_pow01(x[i], p[i]);
}
void sq_math_glibc(const float *x, const float *p, int n, float *c, float *s, float *pw)
{
    for (int i = 0; i < n; i++) {
        c[i] = odam::sq_glibc_cosf(x[i]);
        s[i] = odam::sq_glibc_sinf(x[i]);
        pw[i] = odam::sq_glibc_powf(x[i] < 0 ? -x[i] : x[i], p[i]);
    }
}
void sq_math_grid_node(const float *theta, const float *e, int n, float *fc, float *fs)
{
    for (int i = 0; i < n; i++) odam::sq_grid_node(theta[i], e[i], fc[i], fs[i]);
}
}
