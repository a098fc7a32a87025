// sq_postproc.cuh -- the two geometry steps either side of the optimiser in the reference's call chain, on the device:
//
//   obb_of_points     compute_oriented_bbox  (reference src/utils/box_utils.py:319-410), called right after the
//                     optimiser on the 1000 final surface points of every object (src/scripts/run_multi_view.py:66-67)
//   box3d_iou_pair    box3d_iou              (box_utils.py:97-120 with polygon_clip :23-67, poly_area :70-73,
//                     box3d_vol :89-95), the pair cost of merge_process (src/scripts/run_merge.py:79-122)
//
// compute_oriented_bbox is: xy convex hull (scipy/Qhull) -> subtract the float32 mean of the hull vertices -> for the
// direction of every hull edge EXCEPT the closing one (last vertex -> first vertex of Qhull's vertex list) the
// axis-aligned extent of the rotated hull -> smallest area wins (ties: smallest folded angle) -> 4 corners at z_max,
// then the same 4 at z_min.  Which edge is "the closing one" is decided by where Qhull's vertex list starts; that is
// reproduced here (qhull_head_facet below) instead of being defined away, because in ~1 % of the objects the skipped
// edge is the best one and the reference then returns the runner-up rectangle.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>

namespace odam {

constexpr int kHullMax = 1024;

constexpr int kChains = 4;     // gift-wrapping chains, one per extreme point of the set

struct HullScratch {
    uint16_t hv[kHullMax];      // hull vertices (sample indices), counter-clockwise
    // centred hull coordinates (float32, as the reference's in-place subtraction); before that: the four chains'
    // vertex lists (uint16_t[kChains][kHullMax]), then the facet queue of qhull_head_facet
    float cx[kHullMax], cy[kHullMax];
    double rarea[32];
    double rang[32];
    int ridx[32];
    float wx[kChains][32], wy[kChains][32];
    int widx[kChains][32];
    int h, start_pos, flag;
    int ccur[kChains], clen[kChains], cactive;         // per chain: current vertex, list length; bit mask of live chains
    float ccx[kChains], ccy[kChains], cex[kChains], cey[kChains];   // current vertex and end vertex coordinates
    float zmin, zmax, meanx, meany;
    double best_ang;
};
static_assert(offsetof(HullScratch, cy) == offsetof(HullScratch, cx) + sizeof(float) * kHullMax, "cx and cy are one block");

__device__ __forceinline__ double cross2(double ax, double ay, double bx, double by) { return ax * by - ay * bx; }

// gift-wrapping comparator: is candidate b a better "next counter-clockwise hull vertex after cur" than a?
// (b strictly clockwise of a as seen from cur; collinear -> the farther one; equal points -> the lower sample index)
__device__ __forceinline__ bool wrap_better(float cx, float cy, float ax, float ay, int ai, float bx, float by, int bi)
{
    if (bi < 0) return false;
    if (ai < 0) return true;
    {
        // float32 filter: the sign of the orientation determinant is certain when it exceeds the rounding error of
        // its own evaluation (four subtractions, two products, one difference: < 4 * 2^-24 of |p1| + |p2|; the
        // threshold leaves a factor 4).  Coincident points and collinear triples give 0 here and take the exact path.
        const float fax = ax - cx, fay = ay - cy, fbx = bx - cx, fby = by - cy;
        const float p1 = fax * fby, p2 = fay * fbx;
        const float cr32 = p1 - p2;
        const float mag = fabsf(p1) + fabsf(p2);
        if (mag > 1e-30f && fabsf(cr32) > 1e-6f * mag) return cr32 < 0.f;   // (no verdict from denormal products)
    }
    const double dax = (double)ax - cx, day = (double)ay - cy, dbx = (double)bx - cx, dby = (double)by - cy;
    const double da = dax * dax + day * day, db = dbx * dbx + dby * dby;
    if (db == 0.0) return false;            // b coincides with cur
    if (da == 0.0) return true;
    const double cr = cross2(dax, day, dbx, dby);
    if (cr != 0.0) return cr < 0.0;
    if (db != da) return db > da;
    return bi < ai;
}

// Where does scipy's ConvexHull(points).vertices start?  scipy walks Qhull's facet list from its head (the oldest
// surviving facet) counter-clockwise, so vertices[0] is the first vertex of that facet.  For points in convex position
// Qhull's build is a breadth-first quickhull: initial simplex = (min-x point, max-x point, and whichever of the min-y /
// max-y points is farther from their line), facets in the order [(maxx,minx), (third,minx), (third,maxx)]; a facet with
// outside points is replaced by two new facets appended to the list -- first the one that keeps the facet's OLDER
// vertex, then the one with the younger -- split at its furthest point.  The first facet in that order with no
// outside points is the head.  (Interior points never change this: they are never furthest.)  Verified against scipy
// on 10^4 hulls of superquadric samples (tests/test_postproc.py).  Returns the position (in the counter-clockwise
// list hv) of the head facet's first vertex; *flag is set when the initial simplex is so thin that Qhull may have
// searched beyond the extreme points (not observed on superquadrics).
__device__ int qhull_head_facet(HullScratch &H, const float *px, const float *py, int *flag)
{
    const int h = H.h;
    if (h < 3) return 0;
    // extreme points in Qhull's point order (= sample index order): strict comparisons keep the first occurrence
    int imin_x = 0, imax_x = 0, imin_y = 0, imax_y = 0;
    {
        // "first occurrence" is by sample index, not by position in the hull list
        auto less_idx = [&](int a, int b) { return H.hv[a] < H.hv[b]; };
        for (int k = 1; k < h; k++) {
            const float x = px[H.hv[k]], y = py[H.hv[k]];
            const float xn = px[H.hv[imin_x]], xx = px[H.hv[imax_x]], yn = py[H.hv[imin_y]], yx = py[H.hv[imax_y]];
            if (x < xn || (x == xn && less_idx(k, imin_x))) imin_x = k;
            if (x > xx || (x == xx && less_idx(k, imax_x))) imax_x = k;
            if (y < yn || (y == yn && less_idx(k, imin_y))) imin_y = k;
            if (y > yx || (y == yx && less_idx(k, imax_y))) imax_y = k;
        }
    }
    auto X = [&](int k) { return (double)px[H.hv[k]]; };
    auto Y = [&](int k) { return (double)py[H.hv[k]]; };
    auto dist = [&](int u, int v, int p) { return fabs(cross2(X(v) - X(u), Y(v) - Y(u), X(p) - X(u), Y(p) - Y(u))); };
    const int a = imin_x, b = imax_x;
    int third = -1;
    double bd = -1.0;
    const int cand[2] = {imin_y, imax_y};
    for (int c = 0; c < 2; c++) {
        if (cand[c] == a || cand[c] == b) continue;
        const double d = dist(a, b, cand[c]);
        if (d > bd) { bd = d; third = cand[c]; }
    }
    const double len2 = (X(b) - X(a)) * (X(b) - X(a)) + (Y(b) - Y(a)) * (Y(b) - Y(a));
    if (third < 0 || bd < 1e-2 * len2) {   // degenerate / very thin: Qhull would look at all points
        *flag = 1;
        third = -1; bd = -1.0;
        for (int k = 0; k < h; k++) {
            if (k == a || k == b) continue;
            const double d = dist(a, b, k);
            if (d > bd) { bd = d; third = k; }
        }
        if (third < 0) return 0;
    }
    // vertex ages (insertion order) by hull position; facets as counter-clockwise arcs u -> v of the current polygon
    // (the outside points of facet (u, v) are the hull positions strictly between u and v)
    // the facet queue lives in the (not yet used) centred-coordinate arrays: 1024 entries of {u | v << 16} and
    // {age_u | age_v << 16}.  On superquadric hulls the head is found within the first ~80 entries (queue length
    // ~170); a full queue is flagged.
    constexpr int kQ = kHullMax;
    uint32_t *quv = reinterpret_cast<uint32_t *>(H.cx), *qage = reinterpret_cast<uint32_t *>(H.cy);
    int qn = 0;
    bool dropped = false;
    auto push = [&](int u, int v, int au, int av) {
        if (qn < kQ) { quv[qn] = (uint32_t)u | ((uint32_t)v << 16); qage[qn] = (uint32_t)au | ((uint32_t)av << 16); qn++; }
        else dropped = true;
    };
    auto ccw_after = [&](int u, int v, int w) {   // is v before w when walking counter-clockwise from u?
        const int dv = (v - u + h) % h, dw = (w - u + h) % h;
        return dv < dw;
    };
    // orient each initial facet so that it is an arc of the triangle's counter-clockwise boundary
    auto push_edge = [&](int p, int q, int ap, int aq, int other) {
        // the arc p -> q (counter-clockwise) must not contain `other`
        if (ccw_after(p, q, other)) push(p, q, ap, aq); else push(q, p, aq, ap);
    };
    push_edge(b, a, 1, 0, third);       // facet 0 omits `third`
    push_edge(third, a, 2, 0, b);       // facet 1 omits max-x
    push_edge(third, b, 2, 1, a);       // facet 2 omits min-x
    int next_age = 3;
    for (int qi = 0; qi < qn; qi++) {
        const int u = (int)(quv[qi] & 0xffffu), v = (int)(quv[qi] >> 16);
        const int age_u_q = (int)(qage[qi] & 0xffffu), age_v_q = (int)(qage[qi] >> 16);
        const int len = (v - u + h) % h;
        if (len == 1) { if (dropped) *flag = 1; return u; }   // no outside points: the head of Qhull's facet list
        int p = -1;
        double best = -1.0;
        for (int s = 1; s < len; s++) {
            const int k = (u + s) % h;
            const double d = dist(u, v, k);
            if (d > best) { best = d; p = k; }
        }
        const int ap = next_age++;
        // two new facets: first the one that keeps the facet's older vertex, then the one with the younger
        if (age_u_q < age_v_q) { push(u, p, age_u_q, ap); push(p, v, ap, age_v_q); }
        else { push(p, v, ap, age_v_q); push(u, p, age_u_q, ap); }
    }
    *flag = 1;
    return 0;
}

// Python's float modulo for a positive divisor
__device__ __forceinline__ double py_mod_pos(double a, double b)
{
    double r = fmod(a, b);
    if (r != 0.0 && r < 0.0) r += b;
    return r;
}

// compute_oriented_bbox for the n_pts points in px/py/pz (shared memory, float32).  Whole CTA; out = 8 corners x 3.
__device__ void obb_of_points(HullScratch &H, const float *px, const float *py, const float *pz, int n_pts,
                              double *out /*[24]*/, int *out_flag)
{
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    // ---- z range and the four gift-wrapping starts ----
    // The hull is wrapped counter-clockwise by FOUR chains at once, each from one extreme point of the set to the next:
    //   S0 = min x (then min y), S1 = min y (then min x), S2 = max x (then min y), S3 = max y (then max x),
    // lowest index among duplicates.  Each is the END of an edge of the counter-clockwise walk (the bottom end of the
    // left side, the left end of the bottom side, ...), i.e. a vertex the one-chain walk from S0 emits too, and the
    // step function is the same, so the concatenated lists are that walk's list -- in a quarter of the rounds (a round
    // costs two CTA barriers and a butterfly whatever the number of chains; measured 0.28 -> 0.21 ms for 50 objects).
    float zmin = INFINITY, zmax = -INFINITY;
    float e0[kChains], e1[kChains];   // lexicographic keys: (x, y), (y, x), (-x, y), (-y, -x)
    int ei[kChains];
#pragma unroll
    for (int k = 0; k < kChains; k++) { e0[k] = e1[k] = INFINITY; ei[k] = -1; }
    auto ext_take = [](float &a0, float &a1, int &ai, float b0, float b1, int bi) {
        if (bi >= 0 && (ai < 0 || b0 < a0 || (b0 == a0 && (b1 < a1 || (b1 == a1 && bi < ai))))) { a0 = b0; a1 = b1; ai = bi; }
    };
    for (int i = tid; i < n_pts; i += T) {
        const float x = px[i], y = py[i], z = pz[i];
        zmin = fminf(zmin, z); zmax = fmaxf(zmax, z);
        ext_take(e0[0], e1[0], ei[0], x, y, i);
        ext_take(e0[1], e1[1], ei[1], y, x, i);
        ext_take(e0[2], e1[2], ei[2], -x, y, i);
        ext_take(e0[3], e1[3], ei[3], -y, -x, i);
    }
    for (int o = 16; o > 0; o >>= 1) {
        zmin = fminf(zmin, __shfl_xor_sync(0xffffffffu, zmin, o));
        zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
#pragma unroll
        for (int k = 0; k < kChains; k++) {
            const float o0 = __shfl_xor_sync(0xffffffffu, e0[k], o), o1 = __shfl_xor_sync(0xffffffffu, e1[k], o);
            const int oi = __shfl_xor_sync(0xffffffffu, ei[k], o);
            ext_take(e0[k], e1[k], ei[k], o0, o1, oi);
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < kChains; k++) { H.wx[k][warp] = e0[k]; H.wy[k][warp] = e1[k]; H.widx[k][warp] = ei[k]; }
        H.rarea[warp] = zmin; H.rang[warp] = zmax;
    }
    __syncthreads();
    uint16_t *const chain_list = reinterpret_cast<uint16_t *>(H.cx);   // [kChains][kHullMax], see HullScratch
    if (tid == 0) {
        for (int w = 1; w < nwarps; w++) {
#pragma unroll
            for (int k = 0; k < kChains; k++) ext_take(e0[k], e1[k], ei[k], H.wx[k][w], H.wy[k][w], H.widx[k][w]);
            zmin = fminf(zmin, (float)H.rarea[w]); zmax = fmaxf(zmax, (float)H.rang[w]);
        }
        H.zmin = zmin; H.zmax = zmax;
        H.flag = 0;
        int active = 0;
        for (int k = 0; k < kChains; k++) {
            const int a = ei[k], b = ei[(k + 1) % kChains];
            H.clen[k] = 0; H.ccur[k] = a;
            if (a < 0 || b < 0) continue;   // no points at all
            H.ccx[k] = px[a]; H.ccy[k] = py[a];
            H.cex[k] = px[b]; H.cey[k] = py[b];
            if (px[a] != px[b] || py[a] != py[b]) {   // a chain whose ends coincide is empty
                chain_list[k * kHullMax] = (uint16_t)a;
                H.clen[k] = 1;
                active |= 1 << k;
            }
        }
        if (!active && ei[0] >= 0) { chain_list[0] = (uint16_t)ei[0]; H.clen[0] = 1; }   // a single distinct point
        H.cactive = active;
    }
    __syncthreads();
    // ---- gift wrapping, one hull vertex per chain and round.  Each chain has a quarter of the warps to itself: every
    // thread its best candidate among its share of the points, warp butterfly, then the group's first lane combines
    // the group's warps and advances the chain ----
    const int max_rounds = min(n_pts, kHullMax - 1);
    const bool grouped = nwarps >= kChains;                  // else (tiny CTAs) every warp works on every chain in turn
    const int gw = grouped ? nwarps / kChains : nwarps;      // warps per chain
    const int kc = grouped ? min(warp / gw, kChains - 1) : 0;   // this warp's chain
    const int wloc = warp - kc * gw;                         // its position in the group
    const int gthreads = grouped ? (kc == kChains - 1 ? nwarps - gw * (kChains - 1) : gw) * 32 : T;
    const int tloc = tid - kc * gw * 32;
    for (int round = 0; round <= max_rounds; round++) {
        const int active = H.cactive;
        if (!active) break;
        for (int k = kc; k < kChains; k += (grouped ? kChains : 1)) {   // one pass when every chain has its warps
            if (!((active >> k) & 1)) continue;
            const float cx = H.ccx[k], cy = H.ccy[k];
            float bx = 0.f, by = 0.f;
            int bi = -1;
            for (int i = tloc; i < n_pts; i += gthreads) {
                const float x = px[i], y = py[i];
                if (wrap_better(cx, cy, bx, by, bi, x, y, i)) { bx = x; by = y; bi = i; }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const float ox = __shfl_xor_sync(0xffffffffu, bx, o), oy = __shfl_xor_sync(0xffffffffu, by, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (wrap_better(cx, cy, bx, by, bi, ox, oy, oi)) { bx = ox; by = oy; bi = oi; }
            }
            if (lane == 0) { H.wx[k][wloc] = bx; H.wy[k][wloc] = by; H.widx[k][wloc] = bi; }
        }
        __syncthreads();
        if (tid < kChains && ((active >> tid) & 1)) {   // one thread per chain: combine its group's warps, advance
            const int k = tid;
            const int nw = grouped ? (k == kChains - 1 ? nwarps - gw * (kChains - 1) : gw) : nwarps;
            const float cx = H.ccx[k], cy = H.ccy[k];
            float bx = H.wx[k][0], by = H.wy[k][0];
            int bi = H.widx[k][0];
            for (int w = 1; w < nw; w++)
                if (wrap_better(cx, cy, bx, by, bi, H.wx[k][w], H.wy[k][w], H.widx[k][w])) { bx = H.wx[k][w]; by = H.wy[k][w]; bi = H.widx[k][w]; }
            if (bi < 0 || (bx == H.cex[k] && by == H.cey[k])) {
                atomicAnd(&H.cactive, ~(1 << k));                       // reached the next chain's start
            } else if (round == max_rounds || H.clen[k] >= kHullMax) {
                atomicAnd(&H.cactive, ~(1 << k)); H.flag = 2;           // did not close (cannot happen with exact orientation tests)
            } else {
                chain_list[k * kHullMax + H.clen[k]] = (uint16_t)bi;
                H.clen[k] = H.clen[k] + 1;
                H.ccur[k] = bi; H.ccx[k] = bx; H.ccy[k] = by;
            }
        }
        __syncthreads();
    }
    // the chains' lists, one after the other, are the counter-clockwise hull from S0
    {
        int off = 0;
        for (int k = 0; k < kChains; k++) {
            const int len = H.clen[k];
            for (int j = tid; j < len; j += T) H.hv[off + j] = chain_list[k * kHullMax + j];
            off += len;
        }
        __syncthreads();
        if (tid == 0) H.h = min(off, kHullMax);
        __syncthreads();
    }
    // the start vertex must carry the lowest index among its duplicates too (it was chosen that way above)
    const int h = H.h;
    // ---- Qhull's vertex order: rotate so that the list starts where scipy's hull.vertices starts ----
    if (tid == 0) {
        int flag = 0;
        H.start_pos = qhull_head_facet(H, px, py, &flag);
        H.flag |= flag | (h < 3 ? 2 : 0);
        // mean of the hull vertices in that order: sequential float32 sum, float32 division (np.mean(axis=0) of a
        // float32 [h, 2] array), then the in-place float32 subtraction
        float sx32 = 0.f, sy32 = 0.f;
        for (int k = 0; k < h; k++) {
            const int p = H.hv[(H.start_pos + k) % h];
            sx32 = __fadd_rn(sx32, px[p]);
            sy32 = __fadd_rn(sy32, py[p]);
        }
        H.meanx = __fdiv_rn(sx32, (float)h);
        H.meany = __fdiv_rn(sy32, (float)h);
    }
    __syncthreads();
    const int s0 = H.start_pos;
    for (int k = tid; k < h; k += T) {
        const int p = H.hv[(s0 + k) % h];
        H.cx[k] = __fsub_rn(px[p], H.meanx);
        H.cy[k] = __fsub_rn(py[p], H.meany);
    }
    __syncthreads();
    // ---- one candidate per hull edge k -> k+1, k = 0..h-2 (the closing edge h-1 -> 0 is not considered) ----
    const double half_pi = 1.5707963267948966;   // math.pi / 2
    double my_area = INFINITY, my_ang = INFINITY;
    int my_k = -1;
    double mnx = 0, mxx = 0, mny = 0, mxy = 0;
    for (int k = tid; k < h - 1; k += T) {
        const float ex = __fsub_rn(H.cx[k + 1], H.cx[k]), ey = __fsub_rn(H.cy[k + 1], H.cy[k]);
        const double ang = fabs(py_mod_pos(atan2((double)ey, (double)ex), half_pi));
        const double c = cos(ang), s1 = cos(ang - half_pi), s2 = cos(ang + half_pi);
        double a0 = INFINITY, a1 = -INFINITY, b0 = INFINITY, b1 = -INFINITY;
        for (int j = 0; j < h; j++) {
            const double x = (double)H.cx[j], y = (double)H.cy[j];
            const double rx = c * x + s1 * y, ry = s2 * x + c * y;
            a0 = fmin(a0, rx); a1 = fmax(a1, rx); b0 = fmin(b0, ry); b1 = fmax(b1, ry);
        }
        const double area = (a1 - a0) * (b1 - b0);
        // np.unique sorts the angles ascending; the first strictly smaller area wins -> ties go to the smaller angle
        if (area < my_area || (area == my_area && ang < my_ang)) {
            my_area = area; my_ang = ang; my_k = k; mnx = a0; mxx = a1; mny = b0; mxy = b1;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double oa = __shfl_xor_sync(0xffffffffu, my_area, o), og = __shfl_xor_sync(0xffffffffu, my_ang, o);
        const int ok = __shfl_xor_sync(0xffffffffu, my_k, o);
        const double t0 = __shfl_xor_sync(0xffffffffu, mnx, o), t1 = __shfl_xor_sync(0xffffffffu, mxx, o);
        const double t2 = __shfl_xor_sync(0xffffffffu, mny, o), t3 = __shfl_xor_sync(0xffffffffu, mxy, o);
        if (ok >= 0 && (my_k < 0 || oa < my_area || (oa == my_area && og < my_ang))) {
            my_area = oa; my_ang = og; my_k = ok; mnx = t0; mxx = t1; mny = t2; mxy = t3;
        }
    }
    if (lane == 0) { H.rarea[warp] = my_area; H.rang[warp] = my_ang; H.ridx[warp] = my_k; }
    __syncthreads();
    // the winning warp's lane 0 writes the result
    bool mine = lane == 0 && my_k >= 0;
    if (mine)
        for (int w = 0; w < nwarps; w++) {
            if (w == warp || H.ridx[w] < 0) continue;
            const double oa = H.rarea[w], og = H.rang[w];
            if (oa < my_area || (oa == my_area && (og < my_ang || (og == my_ang && w < warp)))) mine = false;
        }
    if (h < 3 && tid == 0) {   // degenerate input: the reference would raise inside Qhull; return the axis-aligned box
        mine = true; my_ang = 0.0; my_area = 0.0;
        mnx = mxx = mny = mxy = 0.0;
    }
    if (mine) {
        if (!(my_area < 1e10)) { my_ang = 0.0; mnx = mxx = mny = mxy = 0.0; }   // the reference's initial min_bbox survives
        const double c = cos(my_ang), s1 = cos(my_ang - half_pi), s2 = cos(my_ang + half_pi);
        const double cxs[4] = {mxx, mxx, mnx, mnx}, cys[4] = {mxy, mny, mny, mxy};
        for (int q = 0; q < 4; q++) {
            // np.dot([x, y], R) with R = [[c, s1], [s2, c]]
            const double X = cxs[q] * c + cys[q] * s2 + (double)H.meanx;
            const double Y = cxs[q] * s1 + cys[q] * c + (double)H.meany;
            out[q * 3 + 0] = X; out[q * 3 + 1] = Y; out[q * 3 + 2] = (double)H.zmax;
            out[(q + 4) * 3 + 0] = X; out[(q + 4) * 3 + 1] = Y; out[(q + 4) * 3 + 2] = (double)H.zmin;
        }
        if (out_flag) *out_flag = H.flag;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------
// box3d_iou (box_utils.py:97-120): corners [8][3] double, upper four first; returns (iou3d, iou2d)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double poly_area2(const double *x, const double *y, int n)
{
    // 0.5 * |dot(x, roll(y, 1)) - dot(y, roll(x, 1))|
    double a = 0.0, b = 0.0;
    for (int i = 0; i < n; i++) {
        const int j = (i + n - 1) % n;
        a += x[i] * y[j];
        b += y[i] * x[j];
    }
    return 0.5 * fabs(a - b);
}

// Sutherland-Hodgman, subject and clip given counter-clockwise (box_utils.py:23-67, same predicates and the same
// intersection formula); returns the vertex count (0 = empty)
__device__ int polygon_clip4(const double *sx, const double *sy, const double *cxp, const double *cyp, double *ox, double *oy)
{
    double ax[12], ay[12], bx[12], by[12];
    int na = 4;
    for (int i = 0; i < 4; i++) { ax[i] = sx[i]; ay[i] = sy[i]; }
    double c1x = cxp[3], c1y = cyp[3];
    for (int ci = 0; ci < 4; ci++) {
        const double c2x = cxp[ci], c2y = cyp[ci];
        int nb = 0;
        double px_ = ax[na - 1], py_ = ay[na - 1];
        auto inside = [&](double x, double y) { return (c2x - c1x) * (y - c1y) > (c2y - c1y) * (x - c1x); };
        auto inter = [&](double s0, double s1, double e0, double e1, double &ix, double &iy) {
            const double dcx = c1x - c2x, dcy = c1y - c2y, dpx = s0 - e0, dpy = s1 - e1;
            const double n1 = c1x * c2y - c1y * c2x, n2 = s0 * e1 - s1 * e0;
            const double n3 = 1.0 / (dcx * dpy - dcy * dpx);
            ix = (n1 * dpx - n2 * dcx) * n3;
            iy = (n1 * dpy - n2 * dcy) * n3;
        };
        for (int i = 0; i < na; i++) {
            const double ex = ax[i], ey = ay[i];
            if (inside(ex, ey)) {
                if (!inside(px_, py_)) { inter(px_, py_, ex, ey, bx[nb], by[nb]); nb++; }
                bx[nb] = ex; by[nb] = ey; nb++;
            } else if (inside(px_, py_)) {
                inter(px_, py_, ex, ey, bx[nb], by[nb]); nb++;
            }
            px_ = ex; py_ = ey;
        }
        c1x = c2x; c1y = c2y;
        if (nb == 0) return 0;
        na = nb;
        for (int i = 0; i < na; i++) { ax[i] = bx[i]; ay[i] = by[i]; }
    }
    for (int i = 0; i < na; i++) { ox[i] = ax[i]; oy[i] = ay[i]; }
    return na;
}

__device__ void box3d_iou_pair(const double *A, const double *B, double *iou3d, double *iou2d)
{
    double ax[4], ay[4], bx[4], by[4];
    for (int i = 0; i < 4; i++) {   // corners 3, 2, 1, 0: counter-clockwise
        ax[i] = A[(3 - i) * 3 + 0]; ay[i] = A[(3 - i) * 3 + 1];
        bx[i] = B[(3 - i) * 3 + 0]; by[i] = B[(3 - i) * 3 + 1];
    }
    const double area1 = poly_area2(ax, ay, 4), area2 = poly_area2(bx, by, 4);
    double ix[12], iy[12];
    const int ni = polygon_clip4(ax, ay, bx, by, ix, iy);
    // the reference takes ConvexHull(inter).volume; the clip of two convex polygons is convex, so that is its area
    const double inter_area = ni >= 3 ? poly_area2(ix, iy, ni) : 0.0;
    *iou2d = inter_area / (area1 + area2 - inter_area);
    const double zmax = fmin(A[2], B[2]), zmin = fmax(A[4 * 3 + 2], B[4 * 3 + 2]);
    const double inter_vol = inter_area * fmax(0.0, zmax - zmin);
    auto vol = [](const double *C) {
        auto d = [&](int i, int j) {
            const double dx = C[i * 3] - C[j * 3], dy = C[i * 3 + 1] - C[j * 3 + 1], dz = C[i * 3 + 2] - C[j * 3 + 2];
            return sqrt(dx * dx + dy * dy + dz * dz);
        };
        return d(0, 1) * d(1, 2) * d(0, 4);
    };
    *iou3d = inter_vol / (vol(A) + vol(B) - inter_vol);
}

}  // namespace odam
