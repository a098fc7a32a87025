"""Where does the drop-in's optim_process (the reference's call site, run_multi_view.py:22-76) spend its time?
50 tracks x 50 frames x 200 iterations (BASELINE config 2 shape) through the reference-facing API."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from odam_b200 import synthetic  # noqa: E402
from odam_b200.run_multi_view import optim_process  # noqa: E402

n_obj, n_frames, iters = 50, 50, 200
scene = synthetic.make_scene(n_obj, n_frames, seed=2)
frame_ids = np.arange(n_frames)
tracks = []
for i in range(n_obj):
    t = -np.ones((n_frames, 82))
    t[:, 0], t[:, 1] = frame_ids, scene.cls[i]
    b = np.where(scene.mask[i].astype(bool), scene.box[i], np.array([5.0, 1290.0, 5.0, 960.0])[None])
    t[:, 2:6] = np.stack([b[:, 0], b[:, 2], b[:, 1], b[:, 3]], 1)
    t[:, 6:9], t[:, 9:12], t[:, 12] = scene.dims[i], scene.translate[i], scene.angle[i]
    tracks.append(t)
P_cws = [scene.P_cws[0][f] for f in range(n_frames)]
args = (tracks, frame_ids, [np.eye(4)] * n_frames, P_cws, 968, 1296, synthetic.K, "super_quadric", True, iters, 10)
optim_process(*args)  # warm-up (library init, workspace)
ts = []
for _ in range(5):
    t0 = time.perf_counter(); optim_process(*args); ts.append(time.perf_counter() - t0)
print(f"optim_process, {n_obj} tracks x {n_frames} frames x {iters} iterations: median {np.median(ts) * 1e3:.1f} ms "
      f"({n_obj * n_frames * iters / np.median(ts) / 1e6:.1f} M unit/s through the reference call site)")
pr = cProfile.Profile(); pr.enable(); optim_process(*args); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
