"""CPU tests of the post-optimisation geometry: the host mirror of compute_oriented_bbox against the reference's
recorded outputs, and the restatement of the device algorithm (own hull + predicted Qhull vertex order,
oracle/obb_oracle.py) against scipy itself on thousands of superquadric hulls."""
import numpy as np
import pytest
from scipy.spatial import ConvexHull

from conftest import golden
from oracle import c_oracle, obb_oracle


def _random_params(rng, n):
    p = np.zeros((n, 9), np.float32)
    p[:, 0:3] = rng.uniform(-3, 3, (n, 3))
    p[:, 3] = rng.uniform(-np.pi, np.pi, n)
    p[:, 4:7] = np.sqrt(rng.uniform(0.3, 1.5, (n, 3)) / 2)
    p[:, 7:9] = rng.uniform(-2.5, 2.5, (n, 2))
    return p


def test_hull_and_qhull_vertex_order_match_scipy():
    """ConvexHull(xy).vertices -- the set, the counter-clockwise order AND the first vertex (which decides the edge
    compute_oriented_bbox skips) -- predicted without Qhull, on 1500 superquadric surfaces incl. cube-like and
    pinched shapes, axis-aligned ones (ties in the extreme points) and tiny ones."""
    rng = np.random.default_rng(5)
    P = _random_params(rng, 1500)
    P[:100, 7:9] = -10000.0                      # cubes (e = 0.2)
    P[100:200, 3] = rng.choice([0.0, np.pi / 2, -np.pi / 2, np.pi], 100)   # axis-aligned: equal extreme coordinates
    P[200:250, 4:7] *= 0.05
    bad_set, bad_start = [], []
    for k, p in enumerate(P):
        xy = c_oracle.points(p)[:, :2]
        want = xy[ConvexHull(xy).vertices]
        hv = obb_oracle.hull_ccw(xy)
        s = obb_oracle.qhull_first_vertex(xy, hv)
        got = xy[hv[s:] + hv[:s]]
        if got.shape != want.shape or set(map(tuple, got)) != set(map(tuple, want)):
            bad_set.append(k)
        elif not np.array_equal(got, want):
            bad_start.append(k)
    print(f"hull vertex set differs from Qhull's: {len(bad_set)} of {len(P)}; first vertex mispredicted: {len(bad_start)} "
          f"of {len(P)} {bad_start[:8]}")
    assert not bad_set, bad_set[:10]
    # Qhull's facet order is predicted, not computed by Qhull: a residual of ~1 in 1000 hulls starts elsewhere (the
    # skipped edge then differs, which changes the box only when that edge is the best one: ~1 % of those)
    assert len(bad_start) <= 3, bad_start


def test_obb_restatement_matches_reference_outputs():
    """The restated step against what the reference's compute_oriented_bbox returned (float64 inputs in
    call_site.npz, the call site's float32 points in intermediate.npz)."""
    G = golden("call_site.npz")
    for k in range(6):
        got = obb_oracle.oriented_bbox(G[f"obb{k}_pts"].astype(np.float64))
        assert np.abs(got - G[f"obb{k}_box"]).max() < 1e-9, k
    R = golden("intermediate.npz")
    for k in range(int(R["iters"])):
        got = obb_oracle.oriented_bbox(R["surface_points"][k])
        assert np.abs(got - R["bbox_qc"][k]).max() < 1e-9, k


def test_host_mirror_matches_reference_outputs_float32():
    from odam_b200.postprocess import compute_oriented_bbox
    R = golden("intermediate.npz")
    for k in range(int(R["iters"])):
        assert np.abs(compute_oriented_bbox(R["surface_points"][k]) - R["bbox_qc"][k]).max() < 1e-9, k


def test_obb_restatement_matches_host_mirror_on_random_surfaces():
    from odam_b200.postprocess import compute_oriented_bbox
    rng = np.random.default_rng(6)
    for p in _random_params(rng, 300):
        pts = c_oracle.points(p)
        assert np.abs(obb_oracle.oriented_bbox(pts) - compute_oriented_bbox(pts)).max() < 1e-9
