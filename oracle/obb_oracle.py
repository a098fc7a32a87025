"""CPU restatement of the device-side oriented-box step (odam_b200/csrc/sq_postproc.cuh) -- TEST INFRASTRUCTURE ONLY.

The reference's compute_oriented_bbox (/root/reference/src/utils/box_utils.py:319-410) gets its hull from
scipy.spatial.ConvexHull (Qhull) and skips the hull's closing edge, so its result depends on where Qhull's vertex
list starts.  The kernel cannot call Qhull; it builds the hull itself and PREDICTS Qhull's first vertex.  This module
restates both pieces in plain Python so that they can be checked against scipy on thousands of hulls without a GPU
(tests/test_postproc.py):
  hull_ccw(xy)            strict convex hull, counter-clockwise, lowest sample index among duplicates
  qhull_first_vertex(..)  position of scipy's hull.vertices[0] in that list (breadth-first quickhull order, see the
                          comment on qhull_head_facet in sq_postproc.cuh)
  oriented_bbox(pts)      the whole step with the reference's dtype behaviour (float32 mean / centring for float32 input)
Parity status: pinned against scipy's own vertex order and against outputs of the reference's compute_oriented_bbox
(tests/golden/call_site.npz, optim_process.npz, intermediate.npz).
"""
import math

import numpy as np


def _cross(o, a, b):
    return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])


def hull_ccw(xy):
    """Indices of the strict convex hull of xy [n, 2] in counter-clockwise order (Andrew's monotone chain on float64;
    collinear points are not vertices, duplicates are represented by their lowest index)."""
    xy = np.asarray(xy, np.float64)
    order = np.lexsort((np.arange(len(xy)), xy[:, 1], xy[:, 0]))
    pts = []
    for i in order:       # drop exact duplicates, keeping the lowest index
        if pts and xy[pts[-1], 0] == xy[i, 0] and xy[pts[-1], 1] == xy[i, 1]:
            continue
        pts.append(int(i))
    if len(pts) < 3:
        return pts
    lower, upper = [], []
    for i in pts:
        while len(lower) >= 2 and _cross(xy[lower[-2]], xy[lower[-1]], xy[i]) <= 0:
            lower.pop()
        lower.append(i)
    for i in reversed(pts):
        while len(upper) >= 2 and _cross(xy[upper[-2]], xy[upper[-1]], xy[i]) <= 0:
            upper.pop()
        upper.append(i)
    return lower[:-1] + upper[:-1]


def qhull_first_vertex(xy, hv):
    """Position in hv (counter-clockwise hull, sample indices) of scipy's ConvexHull(xy).vertices[0]."""
    h = len(hv)
    if h < 3:
        return 0
    P = np.asarray(xy, np.float64)[hv]
    idx = np.asarray(hv)

    def first_extreme(vals, sign):
        best = 0
        for k in range(1, h):
            if sign * vals[k] < sign * vals[best] or (vals[k] == vals[best] and idx[k] < idx[best]):
                best = k
        return best
    a, b = first_extreme(P[:, 0], 1), first_extreme(P[:, 0], -1)
    cands = [first_extreme(P[:, 1], 1), first_extreme(P[:, 1], -1)]
    dist = lambda u, v, p: abs(_cross(P[u], P[v], P[p]))
    third, bd = -1, -1.0
    for c in cands:
        if c in (a, b):
            continue
        d = dist(a, b, c)
        if d > bd:
            bd, third = d, c
    if third < 0 or bd < 1e-2 * float(((P[b] - P[a]) ** 2).sum()):
        third, bd = -1, -1.0
        for k in range(h):
            if k in (a, b):
                continue
            d = dist(a, b, k)
            if d > bd:
                bd, third = d, k
    queue = []

    def push_edge(p, q, ap, aq, other):
        if (q - p) % h < (other - p) % h:
            queue.append((p, q, ap, aq))
        else:
            queue.append((q, p, aq, ap))
    push_edge(b, a, 1, 0, third)
    push_edge(third, a, 2, 0, b)
    push_edge(third, b, 2, 1, a)
    age, qi = 3, 0
    while qi < len(queue):
        u, v, au, av = queue[qi]
        qi += 1
        n = (v - u) % h
        if n == 1:
            return u
        best, p = -1.0, -1
        for s in range(1, n):
            k = (u + s) % h
            d = dist(u, v, k)
            if d > best:
                best, p = d, k
        ap, age = age, age + 1
        if au < av:
            queue += [(u, p, au, ap), (p, v, ap, av)]
        else:
            queue += [(p, v, ap, av), (u, p, au, ap)]
    return 0


def oriented_bbox(pts):
    """compute_oriented_bbox (box_utils.py:319-410) with the hull and Qhull's vertex order restated as above."""
    pts = np.asarray(pts)
    z_min, z_max = pts[:, 2].min(), pts[:, 2].max()
    xy = pts[:, :2]
    hv = hull_ccw(xy)
    s = qhull_first_vertex(xy, hv)
    hv = hv[s:] + hv[:s]
    contour = xy[hv]
    mean = np.zeros(2, xy.dtype)
    for row in contour:                      # np.mean(axis=0): sequential sum in the array's dtype
        mean = (mean + row).astype(xy.dtype)
    mean = (mean / xy.dtype.type(len(hv))).astype(xy.dtype)
    contour = (contour - mean).astype(xy.dtype)
    best = (0.0, 1e10, 0.0, 0.0, 0.0, 0.0)
    cands = []
    for i in range(len(hv) - 1):
        ex, ey = contour[i + 1, 0] - contour[i, 0], contour[i + 1, 1] - contour[i, 1]
        cands.append(abs(math.atan2(float(ey), float(ex)) % (math.pi / 2)))
    C = contour.astype(np.float64)
    for ang in sorted(set(cands)):
        R = np.array([[math.cos(ang), math.cos(ang - math.pi / 2)], [math.cos(ang + math.pi / 2), math.cos(ang)]])
        rot = R @ C.T
        mnx, mxx, mny, mxy = rot[0].min(), rot[0].max(), rot[1].min(), rot[1].max()
        area = (mxx - mnx) * (mxy - mny)
        if area < best[1]:
            best = (ang, area, mnx, mxx, mny, mxy)
    ang, _, mnx, mxx, mny, mxy = best
    R = np.array([[math.cos(ang), math.cos(ang - math.pi / 2)], [math.cos(ang + math.pi / 2), math.cos(ang)]])
    rect = np.array([[mxx, mxy], [mxx, mny], [mnx, mny], [mnx, mxy]]) @ R + mean.astype(np.float64)[None]
    up = np.concatenate([rect, np.full((4, 1), float(z_max))], 1)
    lo = np.concatenate([rect, np.full((4, 1), float(z_min))], 1)
    return np.concatenate([up, lo], 0)
