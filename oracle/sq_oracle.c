/*
 * oracle/sq_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, strict fp32, no FMA contraction) of the reference's
 * multi-view superquadric optimisation step.  It is the checker for the CUDA
 * path in odam_b200/csrc; nothing under odam_b200/ may link, import or call it.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against
 * outputs of the reference itself (tests/golden/*.npz, produced by
 * tests/golden/make_golden.py importing /root/reference in the build
 * container) and, when oracle/_ref/libref_sampler.so exists, against the
 * reference's own C++ sampler compiled from the sources where they lie.
 *
 * What each function follows (paths relative to /root/reference):
 *   sq_uniform_stream      src/super_quadric/learnable_primitives/fast_sampler/sampling.cpp:18-28
 *                          (libstdc++ mt19937 + uniform_real_distribution<float>)
 *   sq_dc_grid             sampling.cpp:76-125  (divide & conquer equal-arc-length grid)
 *   sq_oracle_sample       sampling.cpp:128-215 (CDF sampling of eta, uniform index sampling of omega)
 *                          with B=1, M=1, N=1000, buffer_size=201, seed=0
 *                          (fast_sampler/_sampler.pyx:413-441)
 *   forward in sq_step     src/super_quadric/sq_libs.py:26-27,556-595 (squashing, rotz, points),
 *                          learnable_primitives/sampling.py:508-509,586-615 (signed power, clamp),
 *                          sq_libs.py:395-430 (projection, extrema, masked L1), :463-466 (prior)
 *   backward in sq_step    the autograd graph of the above, written out analytically
 *   adam in sq_step        torch/optim/adam.py _single_tensor_adam (non-capturable branch)
 *                          as configured at sq_libs.py:373-387
 *
 * Build:  see oracle/Makefile  (gcc -O2 -ffp-contract=off ...; -DSQ_F64 builds the
 * float64 "ideal arithmetic" variant used to put the fp32 differences in context).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SQ_G 201   /* _sampler.pyx:423 buffer_size */
#define SQ_N 1000  /* sq_libs.py:545 EqualDistanceSamplerSQ(1000) */

#ifdef SQ_F64
typedef double real;
#define R_FMA fma
#define R_COS cos
#define R_SIN sin
#define R_POW pow
#define R_EXP exp
#define R_LOG log
#define R_SQRT sqrt
#define R_ABS fabs
#else
typedef float real;
#define R_FMA fmaf
#define R_COS cosf
#define R_SIN sinf
#define R_POW powf
#define R_EXP expf
#define R_LOG logf
#define R_SQRT sqrtf
#define R_ABS fabsf
#endif

/* ------------------------------------------------------------------ */
/* uniform stream: sampling.cpp:18-28, seed 0 re-applied every call    */
/* ------------------------------------------------------------------ */
void sq_uniform_stream(uint32_t seed, int n, float *out)
{
    uint32_t mt[624];
    int idx = 624;
    mt[0] = seed;
    for (int i = 1; i < 624; i++)
        mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    for (int k = 0; k < n; k++) {
        if (idx == 624) {
            for (int i = 0; i < 624; i++) {
                uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
                mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        /* generate_canonical<float,24>: one 32-bit draw, float(y)/2^32, clipped below 1 */
        float r = (float)y / 4294967296.0f;
        if (r >= 1.0f) r = nextafterf(1.0f, 0.0f);
        out[k] = r;
    }
}

/* ------------------------------------------------------------------ */
/* sampler (always fp32: it IS fp32 in the reference)                  */
/* ------------------------------------------------------------------ */
static inline float sgn_powf(float x, float p) /* sampling.cpp:59-61 */
{
    return copysignf(powf(fabsf(x), p), x);
}

static inline void ellipse_pt(float th, float a1, float a2, float e, float *c) /* :64-67 */
{
    c[0] = a1 * sgn_powf(cosf(th), e);
    c[1] = a2 * sgn_powf(sinf(th), e);
}

static inline float chord(const float *p, const float *q) /* :69-73 */
{
    float d1 = p[0] - q[0];
    float d2 = p[1] - q[1];
    return sqrtf(d1 * d1 + d2 * d2);
}

typedef struct { float A[2], B[2], ta, tb; int n, off; } sq_node;

/* Every node with n>0 writes exactly one fixed slot, so the traversal order is
 * free (sampling.cpp uses a LIFO stack; here a FIFO queue = level order, which
 * is what the CUDA path does).  Returns the depth of the tree. */
int sq_dc_grid(float a1, float a2, float e, float ta, float tb, float *grid)
{
    sq_node q[SQ_G + 2];
    int depth_of[SQ_G + 2];
    int head = 0, tail = 0, max_depth = 0;
    ellipse_pt(ta, a1, a2, e, q[0].A);
    ellipse_pt(tb, a1, a2, e, q[0].B);
    q[0].ta = ta; q[0].tb = tb; q[0].n = SQ_G - 2; q[0].off = 1;
    depth_of[0] = 1;
    tail = 1;
    grid[0] = ta;
    while (head < tail) {
        sq_node nd = q[head];
        int d = depth_of[head];
        head++;
        if (d > max_depth) max_depth = d;
        float th = (nd.ta + nd.tb) / 2;
        float C[2];
        ellipse_pt(th, a1, a2, e, C);
        float dA = chord(nd.A, C);
        float dB = chord(C, nd.B);
        int nA = (int)roundf((dA / (dA + dB)) * (float)(nd.n - 1));
        int nB = nd.n - nA - 1;
        if (nA + nd.off < 0 || nA + nd.off >= SQ_G) return -1; /* NaN geometry */
        grid[nA + nd.off] = th;
        if (nA > 0) {
            sq_node *c = &q[tail]; depth_of[tail] = d + 1; tail++;
            c->A[0] = nd.A[0]; c->A[1] = nd.A[1]; c->B[0] = C[0]; c->B[1] = C[1];
            c->ta = nd.ta; c->tb = th; c->n = nA; c->off = nd.off;
        }
        if (nB > 0) {
            sq_node *c = &q[tail]; depth_of[tail] = d + 1; tail++;
            c->A[0] = C[0]; c->A[1] = C[1]; c->B[0] = nd.B[0]; c->B[1] = nd.B[1];
            c->ta = th; c->tb = nd.tb; c->n = nB; c->off = nd.off + nA + 1;
        }
    }
    grid[SQ_G - 1] = tb;
    return max_depth;
}

/* libstdc++ std::lower_bound bisection (kept verbatim in behaviour because the
 * CDF is not guaranteed sorted at its last entry). */
static int lower_bound_201(const float *cdf, float val)
{
    int first = 0, len = SQ_G;
    while (len > 0) {
        int half = len >> 1;
        int mid = first + half;
        if (cdf[mid] < val) { first = mid + 1; len = len - half - 1; }
        else len = half;
    }
    return first;
}

static float g_uniform[2 * SQ_N];
static int g_uniform_ready = 0;

/* a[3], e[2] fp32 -> grids [201] each, indices [1000] each, angles [1000] each.
 * Any output pointer may be NULL.  Returns max tree depth, or -1 on NaN geometry. */
int sq_oracle_sample(const float *a, const float *e,
                     float *eta_grid, float *omega_grid,
                     int *eta_idx, int *omega_idx,
                     float *etas, float *omegas, float *cdf_out)
{
    const float pi = (float)acos(-1.0); /* sampling.cpp:14 */
    const float pi_2 = pi / 2;          /* :15 */
    float ge[SQ_G], go[SQ_G], cdf[SQ_G];
    if (!g_uniform_ready) { sq_uniform_stream(0u, 2 * SQ_N, g_uniform); g_uniform_ready = 1; }

    int d1 = sq_dc_grid(a[0], a[2], e[0], pi_2, -pi_2, ge); /* :183-190 */
    if (d1 < 0) return -1;
    /* sample_etas, :128-155 */
    const float smoothing = 0.001f;
    const float a1a2 = a[0] + a[1];
    cdf[0] = smoothing;
    for (int i = 1; i < SQ_G; i++)
        cdf[i] = cdf[i - 1] + smoothing + a1a2 * sgn_powf(cosf(ge[i]), e[0]);
    float s = cdf[SQ_G - 1];
    for (int i = 0; i < SQ_G; i++) cdf[i] /= s;
    for (int i = 0; i < SQ_N; i++) {
        int j = lower_bound_201(cdf, g_uniform[i]);
        if (j >= SQ_G) j = SQ_G - 1; /* unreachable: cdf[200]==1 > u */
        if (eta_idx) eta_idx[i] = j;
        if (etas) etas[i] = ge[j];
    }
    int d2 = sq_dc_grid(a[0], a[1], e[1], pi, -pi, go); /* :202-209 */
    if (d2 < 0) return -1;
    for (int i = 0; i < SQ_N; i++) { /* :210-212 */
        int k = (int)(g_uniform[SQ_N + i] * (float)SQ_G);
        if (omega_idx) omega_idx[i] = k;
        if (omegas) omegas[i] = go[k];
    }
    if (eta_grid) memcpy(eta_grid, ge, sizeof ge);
    if (omega_grid) memcpy(omega_grid, go, sizeof go);
    if (cdf_out) memcpy(cdf_out, cdf, sizeof cdf);
    return d1 > d2 ? d1 : d2;
}

/* ------------------------------------------------------------------ */
/* one optimisation trajectory for one object                          */
/* ------------------------------------------------------------------ */
/* parameter vector layout (9): t[0..2], angle, s[0..2], h[0..1]  (include/odam_sq.h) */

static inline real clamp_eps(real v) /* sampling.py:613-615 */
{
    real m = R_ABS(v) > (real)1e-6f ? R_ABS(v) : (real)1e-6f;
    return (v > 0 ? (real)1 : (real)-1) * m;
}
static inline real clamp_grad(real v) /* d/dv of the above: torch.max(|v|,c) splits ties 0.5/0.5 */
{
    real av = R_ABS(v);
    if (av > (real)1e-6f) return 1;
    if (av == (real)1e-6f) return (real)0.5;
    return 0;
}
static inline real sgn(real v) { return (v > 0) - (v < 0); }

typedef struct {
    /* per-sample cached quantities for the backward pass */
    real fce, fse, fco, fso;       /* signed powers */
    real lce, lse, lco, lso;       /* log|cos|, log|sin| */
    real x0, y0, z0;               /* before clamp */
    real x, y, z;                  /* local, after clamp */
    real X, Y, Z;                  /* world */
} sq_pt;

/*
 * Runs n_iters Adam steps.  Optional teacher-forcing inputs (NULL = fresh run):
 *   m0,v0 [9] Adam moments, step0 = number of steps already taken, s0[3] = prior anchor.
 * History outputs (any may be NULL): params after each step [n_iters*9], loss before each
 * step [n_iters], grad [n_iters*9], arg-extreme sample index per view/side
 * [n_iters*V*4] (-1 when no valid point), pred box [n_iters*V*4], eta/omega grid index of
 * every sample [n_iters*1000] each.
 * Returns 0, or -(iter+1) if the sampler hit NaN geometry at that iteration.
 */
int sq_oracle_run(const float *init9, int V, const float *Ms, const float *box,
                  const unsigned char *mask, const float *prior9, int n_iters,
                  int optimize_shapes, double lr, double lr_shape,
                  const float *m0, const float *v0, int step0, const float *s0,
                  float *hist_params, float *hist_loss, float *hist_grad,
                  int *hist_arg, float *hist_pred, int *hist_eta_idx, int *hist_omega_idx,
                  float *m_out, float *v_out)
{
    real p[9], m[9], v[9], sprior[3];
    for (int i = 0; i < 9; i++) { p[i] = init9[i]; m[i] = m0 ? m0[i] : 0; v[i] = v0 ? v0[i] : 0; }
    for (int i = 0; i < 3; i++) sprior[i] = s0 ? s0[i] : init9[4 + i]; /* sq_libs.py:454 */
    sq_pt *pt = (sq_pt *)malloc(sizeof(sq_pt) * SQ_N);
    int *ej = (int *)malloc(sizeof(int) * SQ_N), *ok = (int *)malloc(sizeof(int) * SQ_N);
    float ge[SQ_G], go[SQ_G];
    const double beta1 = 0.9, beta2 = 0.999, eps = 1e-8;
    int rc = 0;

    for (int it = 0; it < n_iters; it++) {
        real *t = p, ang = p[3], *s = p + 4, *h = p + 7;
        /* ---- forward: sq_libs.py:577-595 ---- */
        real a[3], e[2], sig[2];
        for (int k = 0; k < 3; k++) a[k] = s[k] * s[k];
        for (int k = 0; k < 2; k++) {
            sig[k] = (real)1 / ((real)1 + R_EXP(-h[k]));
            e[k] = sig[k] * (real)1.4f + (real)0.2f; /* squashing, :26-27 */
        }
        float af[3] = {(float)a[0], (float)a[1], (float)a[2]}, ef[2] = {(float)e[0], (float)e[1]};
        if (sq_oracle_sample(af, ef, ge, go, ej, ok, NULL, NULL, NULL) < 0) { rc = -(it + 1); break; }
        if (hist_eta_idx) memcpy(hist_eta_idx + (size_t)it * SQ_N, ej, sizeof(int) * SQ_N);
        if (hist_omega_idx) memcpy(hist_omega_idx + (size_t)it * SQ_N, ok, sizeof(int) * SQ_N);
        real cz = R_COS(ang), sz = R_SIN(ang);
        for (int i = 0; i < SQ_N; i++) {
            real eta = ge[ej[i]], om = go[ok[i]];
            if (eta == 0) eta += (real)1e-6f; /* sampling.py:591-592 */
            if (om == 0) om += (real)1e-6f;
            real ce = R_COS(eta), se = R_SIN(eta), co = R_COS(om), so = R_SIN(om);
            sq_pt *q = &pt[i];
            q->fce = sgn(ce) * R_POW(R_ABS(ce), e[0]); /* sampling.py:508-509 */
            q->fse = sgn(se) * R_POW(R_ABS(se), e[0]);
            q->fco = sgn(co) * R_POW(R_ABS(co), e[1]);
            q->fso = sgn(so) * R_POW(R_ABS(so), e[1]);
            q->lce = R_LOG(R_ABS(ce)); q->lse = R_LOG(R_ABS(se));
            q->lco = R_LOG(R_ABS(co)); q->lso = R_LOG(R_ABS(so));
            q->x0 = a[0] * q->fce * q->fco; /* :605-607, left-assoc */
            q->y0 = a[1] * q->fce * q->fso;
            q->z0 = a[2] * q->fse;
            q->x = clamp_eps(q->x0); q->y = clamp_eps(q->y0); q->z = clamp_eps(q->z0);
            /* pts @ R.T + t with R=[[c,-s,0],[s,c,0],[0,0,1]] (sq_libs.py:556-575,590-592);
             * torch's CPU mm accumulates k in order with FMA (probed), then the add rounds */
            real X = R_FMA(q->z, (real)0, R_FMA(q->y, -sz, q->x * cz));
            real Y = R_FMA(q->z, (real)0, R_FMA(q->y, cz, q->x * sz));
            real Z = R_FMA(q->z, (real)1, R_FMA(q->y, (real)0, q->x * (real)0));
            q->X = X + t[0]; q->Y = Y + t[1]; q->Z = Z + t[2];
        }
        /* ---- constraint_2d: sq_libs.py:395-430 ---- */
        real g[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        real side_sum[4] = {0, 0, 0, 0};
        for (int vi = 0; vi < V; vi++) {
            const float *M = Ms + 12 * vi;
            real best[4] = {(real)1000000, (real)-1000000, (real)1000000, (real)-1000000};
            int arg[4] = {-1, -1, -1, -1};
            for (int i = 0; i < SQ_N; i++) {
                const sq_pt *q = &pt[i];
                real qx = R_FMA(1, M[3], R_FMA(q->Z, M[2], R_FMA(q->Y, M[1], q->X * M[0])));
                real qy = R_FMA(1, M[7], R_FMA(q->Z, M[6], R_FMA(q->Y, M[5], q->X * M[4])));
                real qz = R_FMA(1, M[11], R_FMA(q->Z, M[10], R_FMA(q->Y, M[9], q->X * M[8])));
                int valid = qz > (real)0.5f;      /* :399 */
                real d = R_ABS(qz) + (real)1e-6f;  /* :400 */
                real u = qx / d, w = qy / d;
                /* :402-413 where(valid, coord, +-1e6) then min/max, first index wins ties;
                 * a winning sentinel carries no gradient (torch.where routes it to the
                 * constant branch) -> arg = -1 */
                real c4[4] = {valid ? u : (real)1000000, valid ? u : (real)-1000000,
                              valid ? w : (real)1000000, valid ? w : (real)-1000000};
                for (int sd = 0; sd < 4; sd++) {
                    int better = (sd & 1) ? (c4[sd] > best[sd]) : (c4[sd] < best[sd]);
                    if (i == 0 || better) { best[sd] = c4[sd]; arg[sd] = valid ? i : -1; }
                }
            }
            for (int sd = 0; sd < 4; sd++) {
                if (hist_arg) hist_arg[((size_t)it * V + vi) * 4 + sd] = arg[sd];
                if (hist_pred) hist_pred[((size_t)it * V + vi) * 4 + sd] = (float)best[sd];
                real target = box[4 * vi + sd]; /* = -line[-1], sq_libs.py:424,449 */
                real diff = best[sd] - target;
                real l = R_ABS(diff);
                real mk = mask[4 * vi + sd] ? 1 : 0;
                if (l != l) l = 0;                /* :426-427 nan -> 0 */
                side_sum[sd] += l * mk;
                if (arg[sd] < 0 || !mk || diff != diff) continue;
                /* ---- backward through this arg-extreme ---- */
                real c = sgn(diff) * mk / (real)V;
                const sq_pt *q = &pt[arg[sd]];
                real qx = R_FMA(1, M[3], R_FMA(q->Z, M[2], R_FMA(q->Y, M[1], q->X * M[0])));
                real qy = R_FMA(1, M[7], R_FMA(q->Z, M[6], R_FMA(q->Y, M[5], q->X * M[4])));
                real qz = R_FMA(1, M[11], R_FMA(q->Z, M[10], R_FMA(q->Y, M[9], q->X * M[8])));
                real d = R_ABS(qz) + (real)1e-6f;
                real gq[3] = {0, 0, 0};
                real num = sd < 2 ? qx : qy;
                gq[sd < 2 ? 0 : 1] = c / d;
                gq[2] = -c * num / (d * d) * sgn(qz);
                real gp[3];
                for (int k = 0; k < 3; k++) gp[k] = M[k] * gq[0] + M[4 + k] * gq[1] + M[8 + k] * gq[2];
                g[0] += gp[0]; g[1] += gp[1]; g[2] += gp[2];
                g[3] += gp[0] * (-q->x * sz - q->y * cz) + gp[1] * (q->x * cz - q->y * sz);
                real gl[3] = {cz * gp[0] + sz * gp[1], -sz * gp[0] + cz * gp[1], gp[2]};
                gl[0] *= clamp_grad(q->x0); gl[1] *= clamp_grad(q->y0); gl[2] *= clamp_grad(q->z0);
                real ga[3] = {gl[0] * q->fce * q->fco, gl[1] * q->fce * q->fso, gl[2] * q->fse};
                for (int k = 0; k < 3; k++) g[4 + k] += (real)2 * s[k] * ga[k];
                real ge1 = (gl[0] * q->x0 + gl[1] * q->y0) * q->lce + gl[2] * q->z0 * q->lse;
                real ge2 = gl[0] * q->x0 * q->lco + gl[1] * q->y0 * q->lso;
                g[7] += ge1 * (real)1.4f * sig[0] * ((real)1 - sig[0]);
                g[8] += ge2 * (real)1.4f * sig[1] * ((real)1 - sig[1]);
            }
        }
        real loss = 0;
        for (int sd = 0; sd < 4; sd++) loss += side_sum[sd] / (real)V; /* torch.mean over ALL V */
        if (prior9) { /* sq_libs.py:463-466 */
            real dd[3] = {sprior[0] - s[0], sprior[1] - s[1], sprior[2] - s[2]};
            real q3 = 0;
            for (int r = 0; r < 3; r++) {
                real row = 0;
                for (int cidx = 0; cidx < 3; cidx++) row += (real)prior9[3 * r + cidx] * dd[cidx];
                q3 += dd[r] * row;
            }
            loss += q3 * (real)20;
            for (int r = 0; r < 3; r++) {
                real acc = 0;
                for (int cidx = 0; cidx < 3; cidx++)
                    acc += ((real)prior9[3 * r + cidx] + (real)prior9[3 * cidx + r]) * dd[cidx];
                g[4 + r] += -(real)20 * acc;
            }
        }
        if (!optimize_shapes) { g[7] = 0; g[8] = 0; }
        if (hist_loss) hist_loss[it] = (float)loss;
        if (hist_grad) for (int i = 0; i < 9; i++) hist_grad[(size_t)it * 9 + i] = (float)g[i];
        /* ---- Adam: torch/optim/adam.py _single_tensor_adam, non-capturable branch ---- */
        double step = (double)(step0 + it + 1);
        double bc1 = 1.0 - pow(beta1, step), bc2 = 1.0 - pow(beta2, step);
        double bc2_sqrt = pow(bc2, 0.5);
        int np_ = optimize_shapes ? 9 : 7;
        for (int i = 0; i < np_; i++) {
            double step_size = (i < 7 ? lr : lr_shape) / bc1;
            /* roundings probed bit-for-bit against torch 2.11 CPU kernels (AVX2+FMA):
             * lerp_ and addcmul_ fuse their last multiply-add, addcdiv_ does not */
            m[i] = R_FMA((real)(1.0 - beta1), g[i] - m[i], m[i]);       /* lerp_, weight<0.5 */
            v[i] = v[i] * (real)beta2;                                   /* mul_ */
            v[i] = R_FMA((real)(1.0 - beta2) * g[i], g[i], v[i]);        /* addcmul_ */
            real denom = R_SQRT(v[i]) / (real)bc2_sqrt + (real)eps;
            p[i] = p[i] + ((real)(-step_size) * m[i]) / denom;           /* addcdiv_ */
        }
        if (hist_params) for (int i = 0; i < 9; i++) hist_params[(size_t)it * 9 + i] = (float)p[i];
    }
    if (m_out) for (int i = 0; i < 9; i++) m_out[i] = (float)m[i];
    if (v_out) for (int i = 0; i < 9; i++) v_out[i] = (float)v[i];
    free(pt); free(ej); free(ok);
    return rc;
}

/* forward only: params -> 1000 world points [1000*3] (compute_ellipsoid_points(use_numpy=True)) */
int sq_oracle_points(const float *p9, float *out_xyz)
{
    float a[3], e[2];
    float ge[SQ_G], go[SQ_G];
    int *ej = (int *)malloc(sizeof(int) * SQ_N), *ok = (int *)malloc(sizeof(int) * SQ_N);
    for (int k = 0; k < 3; k++) a[k] = p9[4 + k] * p9[4 + k];
    for (int k = 0; k < 2; k++) e[k] = 1.0f / (1.0f + expf(-p9[7 + k])) * 1.4f + 0.2f;
    int rc = sq_oracle_sample(a, e, ge, go, ej, ok, NULL, NULL, NULL);
    if (rc >= 0) {
        float cz = cosf(p9[3]), sz = sinf(p9[3]);
        for (int i = 0; i < SQ_N; i++) {
            float eta = ge[ej[i]], om = go[ok[i]];
            if (eta == 0) eta += 1e-6f;
            if (om == 0) om += 1e-6f;
            float ce = cosf(eta), se = sinf(eta), co = cosf(om), so = sinf(om);
            float fce = copysignf(powf(fabsf(ce), e[0]), ce) * (ce != 0);
            float fse = copysignf(powf(fabsf(se), e[0]), se) * (se != 0);
            float fco = copysignf(powf(fabsf(co), e[1]), co) * (co != 0);
            float fso = copysignf(powf(fabsf(so), e[1]), so) * (so != 0);
            float x = (float)clamp_eps((real)(a[0] * fce * fco));
            float y = (float)clamp_eps((real)(a[1] * fce * fso));
            float z = (float)clamp_eps((real)(a[2] * fse));
            out_xyz[3 * i + 0] = fmaf(y, -sz, x * cz) + p9[0];
            out_xyz[3 * i + 1] = fmaf(y, cz, x * sz) + p9[1];
            out_xyz[3 * i + 2] = z + p9[2];
        }
    }
    free(ej); free(ok);
    return rc < 0 ? -1 : 0;
}
