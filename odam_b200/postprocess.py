"""Post-optimisation geometry of the call site (reference src/scripts/run_multi_view.py:49,66-67):
the detector's 3-D box and the oriented box of the optimised surface points.  Host-side numpy; same
conventions and corner order as reference src/utils/box_utils.py:286-308 (get_3d_box) and :319-410
(compute_oriented_bbox), re-implemented vectorised (no per-edge Python loops, no matplotlib import).
"""
import numpy as np
from scipy.spatial import ConvexHull


def rotz(t):
    c, s = np.cos(t), np.sin(t)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def get_3d_box(box_size, rot_mat, center):
    """8 corners [8,3] of a box of size (l, w, h) rotated by rot_mat about its centre (box_utils.py:286-308)."""
    l, w, h = box_size
    sx = np.array([1, 1, -1, -1, 1, 1, -1, -1]) * (l / 2)
    sy = np.array([1, -1, -1, 1, 1, -1, -1, 1]) * (w / 2)
    sz = np.array([1, 1, 1, 1, -1, -1, -1, -1]) * (h / 2)
    return (np.asarray(rot_mat) @ np.vstack([sx, sy, sz])).T + np.asarray(center)[None, :]


def compute_oriented_bbox(pts):
    """Oriented 3-D box (8 corners) of surface points, +z up: min-area rectangle of the xy convex hull over the
    hull-edge directions, extruded from z_min to z_max (box_utils.py:319-410, including its conventions: hull
    centred on the mean of its vertices, the closing hull edge not considered, angles folded into [0, pi/2),
    first smallest area wins, upper four corners first)."""
    pts = np.asarray(pts)                      # the reference keeps the caller's dtype (float32 from the call site) ...
    z_min, z_max = pts[:, 2].min(), pts[:, 2].max()
    xy = pts[:, :2]
    hull = xy[ConvexHull(xy).vertices]
    centre = np.mean(hull, axis=0)             # ... so the mean and the centring happen in that dtype
    hull = hull - centre
    edges = np.diff(hull, axis=0).astype(np.float64)
    hull = hull.astype(np.float64)
    centre = centre.astype(np.float64)
    angles = np.unique(np.abs(np.arctan2(edges[:, 1], edges[:, 0]) % (np.pi / 2)))
    c, s = np.cos(angles), np.cos(angles - np.pi / 2)
    s2 = np.cos(angles + np.pi / 2)
    rx = c[:, None] * hull[None, :, 0] + s[:, None] * hull[None, :, 1]
    ry = s2[:, None] * hull[None, :, 0] + c[:, None] * hull[None, :, 1]
    min_x, max_x, min_y, max_y = rx.min(1), rx.max(1), ry.min(1), ry.max(1)
    area = (max_x - min_x) * (max_y - min_y)
    k = int(np.argmax(area == area[area < 1e10].min())) if (area < 1e10).any() else 0
    R = np.array([[c[k], s[k]], [s2[k], c[k]]])
    rect = np.array([[max_x[k], max_y[k]], [max_x[k], min_y[k]], [min_x[k], min_y[k]], [min_x[k], max_y[k]]]) @ R
    rect += centre[None, :]
    upper = np.concatenate([rect, np.full((4, 1), z_max)], axis=1)
    lower = np.concatenate([rect, np.full((4, 1), z_min)], axis=1)
    return np.concatenate([upper, lower], axis=0)
