#!/bin/bash
# compute-sanitizer over the round-2 kernels (oriented boxes, merge costs, sample_on_batch with several primitives) and
# a small optimiser launch of both builds
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
from odam_b200 import api, synthetic
rng = np.random.default_rng(0)
P = np.zeros((6, 9), np.float32); P[:, 0:3] = rng.uniform(-2, 2, (6, 3)); P[:, 3] = rng.uniform(-3, 3, 6)
P[:, 4:7] = np.sqrt(rng.uniform(0.3, 1.5, (6, 3)) / 2); P[:, 7:9] = rng.uniform(-2, 2, (6, 2)); P[0, 7:9] = -10000
boxes, flags, pts = api.oriented_boxes_host(P, want_points=True)
b2, f2 = api.oriented_boxes_of_points_host(pts[:, :500])
cost = api.merge_cost_host(boxes, np.array([0, 0, 4, 5, 1, 1]))
api.sample_on_batch(rng.uniform(0.1, 0.8, (2, 2, 3)).astype(np.float32), rng.uniform(0.2, 1.6, (2, 2, 2)).astype(np.float32))
tracks = api.pack_scene(synthetic.make_scene(3, 12, seed=3))
for lay in (1, 2):
    api.optimize_host(tracks, prior=api.prior_table(), n_iters=3, threads=256, code_layout=lay)
api.optimize_host(tracks, prior=api.prior_table(), n_iters=3, cluster=2)
api.optimize_host(tracks, prior=api.prior_table(), n_iters=3, cluster=3, extras=("out_corners", "out_box_flag"))   # 128-register build, fused boxes
print("ok", boxes.shape, cost.shape, flags.tolist())
PY
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python /tmp/san.py > gpurun_out/sanitize_$tool.log 2>&1
  tail -4 gpurun_out/sanitize_$tool.log
done
timeout 900 compute-sanitizer --tool racecheck python /tmp/san.py > gpurun_out/sanitize_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|hazard" gpurun_out/sanitize_racecheck.log | sort | uniq -c | head -20
