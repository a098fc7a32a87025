"""Round-2 golden vectors, again produced by RUNNING THE REFERENCE ITSELF (build container only):

    python tests/golden/make_golden_wide.py        # needs /root/reference; writes tests/golden/

Adds, without touching the round-1 fixtures (tests/golden/make_golden.py):
  ref_runs_wide.npz   reference trajectories at the BASELINE shapes -- V=50 (config 2), V=30 (config 3), V=300
                      (config 4) -- plus the edge cases the reference's loop has: a view behind the camera (the
                      z <= 0.5 sentinel path of sq_libs.py:399-413), a track whose views are partly fully masked, and
                      an all-masked track (only the prior acts);
  optim_process.npz   the return dict of the reference's own call site ``optim_process``
                      (src/scripts/run_multi_view.py:22-76) on synthetic 82-column tracks (SURVEY 8c stubs),
                      10 iterations, incl. one track with too few views (keeps its initial quadric);
  intermediate.npz    ``SuperQuadricOptimizer.run_with_intermediate`` (sq_libs.py:478-527): per-step surface points and
                      oriented boxes;
  sampler_batch.npz   ONE ``fast_sample_on_batch`` call with B*M > 1 (the generator keeps drawing across primitives,
                      sampling.cpp:169-214).
Every trajectory is asserted bit-identical with oracle/torch_oracle.py here, as in round 1.
"""
import os
import sys
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402

OUT = mg.OUT


def record_case(sq, torch_oracle, c_oracle, fast_sample_on_batch, prior_by_class, scene, i, rep, pr, iters, V, Ms=None,
                box=None, mask=None):
    """One reference run + the discrete decisions of every step (from the bit-identical torch restatement)."""
    Ms = scene.P_cws[i][:V] if Ms is None else Ms
    box = scene.box[i][:V] if box is None else box
    mask = scene.mask[i][:V] if mask is None else mask
    lines = []
    for v in range(V):
        d = {}
        for s, name in enumerate(("x_min", "x_max", "y_min", "y_max")):
            if mask[v, s]:
                d[name] = np.array([1, 0, -box[v, s]]) if name[0] == "x" else np.array([0, 1, -box[v, s]])
        lines.append(d)

    class _S:   # record_reference_run reads the views through these two members
        translate, angle, dims, cls = scene.translate, scene.angle, scene.dims, scene.cls
        P_cws = {i: Ms}
        V_ = V

        @staticmethod
        def gt_lines(_i):
            return lines
    _S.V = V
    rec = mg.record_reference_run(sq, _S, i, rep, pr, iters, V)
    t = torch_oracle.run(scene.translate[i], scene.angle[i], scene.dims[i], Ms, box, mask,
                         prior_by_class[scene.cls[i]] if pr else None, iters, rep, sampler=fast_sample_on_batch,
                         anomaly=False)
    for x in ("params", "grad", "m", "v", "loss"):
        assert np.array_equal(t[x], rec[x], equal_nan=True), (x, np.abs(t[x] - rec[x]).max())
    rec["arg"] = t["arg"].astype(np.int16)
    rec["resid_sign"] = np.sign(t["pred"] - box.astype(np.float32)[None]).astype(np.int8)
    eta_idx = np.zeros((iters, 1000), np.uint8)
    for it in range(iters):
        eta_idx[it] = c_oracle.sample(t["ae"][it, :3], t["ae"][it, 3:])["eta_idx"]
    rec["eta_idx"] = eta_idx
    rec.update(Ms=np.asarray(Ms, np.float64), box=np.asarray(box, np.float64), mask=np.asarray(mask, np.uint8),
               cls=int(scene.cls[i]), translate=scene.translate[i], angle=scene.angle[i], dims=scene.dims[i])
    return rec


def main():
    b = mg.build_reference_extension()
    sys.path[:0] = [mg.REF, b]
    os.chdir(mg.REF)
    import torch
    torch.set_num_threads(1)
    import src.super_quadric.sq_libs as sq
    from learnable_primitives.fast_sampler import fast_sample_on_batch
    from oracle import c_oracle, torch_oracle
    from odam_b200 import synthetic
    import pickle
    c_oracle.build()
    with open(os.path.join(mg.REF, "src", "super_quadric", "scale_prior"), "rb") as f:
        prior = pickle.load(f)
    prior_by_class = np.stack([np.asarray(prior[sq.CLASS_MAPPER[c]], np.float64) for c in range(8)])

    # ---- sampler: one call with several primitives ----
    rng = np.random.default_rng(4242)
    A = rng.uniform(0.1, 0.8, (3, 2, 3)).astype(np.float32)
    E = rng.uniform(0.2, 1.6, (3, 2, 2)).astype(np.float32)
    et, om = fast_sample_on_batch(A, E, 1000)
    np.savez_compressed(os.path.join(OUT, "sampler_batch.npz"), a=A, e=E, etas=np.asarray(et), omegas=np.asarray(om))
    print("sampler_batch: B=3, M=2")

    # ---- trajectories at the BASELINE shapes and the loop's edge cases ----
    out = dict(prior_by_class=prior_by_class)
    cases = []
    s2 = synthetic.make_scene(4, 50, seed=2)      # config 2 draws (first objects of the bench scene)
    s3 = synthetic.make_scene(4, 30, seed=3)      # config 3
    s4 = synthetic.make_scene(2, 300, seed=4)     # config 4
    s1 = synthetic.make_scene(10, 20, seed=1)     # config 1
    cases.append(("V50 config-2 object 0", s2, 0, "super_quadric", True, 40, 50, {}))
    cases.append(("V50 config-2 object 1", s2, 1, "super_quadric", True, 40, 50, {}))
    cases.append(("V30 config-3 object 0", s3, 0, "super_quadric", True, 40, 30, {}))
    cases.append(("V300 config-4 object 0", s4, 0, "super_quadric", True, 30, 300, {}))
    # a camera looking away in view 0 (every point has z <= 0.5 there: +-1e6 sentinels, no gradient from that view)
    Ms = s1.P_cws[2].copy()
    Ms[0, 2, :] = -Ms[0, 2, :]
    cases.append(("V20 view 0 behind the camera", s1, 2, "super_quadric", True, 30, 20, dict(Ms=Ms)))
    # views 3..8 fully masked (run() accepts empty line dicts; the call site normally drops such frames)
    mk = s1.mask[3].copy()
    mk[3:9] = 0
    cases.append(("V20 views 3-8 fully masked", s1, 3, "super_quadric", True, 30, 20, dict(mask=mk)))
    # nothing to fit at all: only the prior term (zero gradient at the anchor)
    cases.append(("V12 all masked", s1, 4, "super_quadric", True, 12, 12, dict(mask=np.zeros((12, 4), np.uint8))))
    # a camera so close that the object straddles the z = 0.5 plane in view 1 (some points valid, some not)
    Ms = s1.P_cws[5].copy()
    c = np.append(s1.translate[5], 1.0)
    zc = Ms[1, 2] @ c
    Ms[1, 2, 3] -= zc - 0.5
    cases.append(("V20 object straddles z=0.5 in view 1", s1, 5, "super_quadric", True, 30, 20, dict(Ms=Ms)))
    out["names"] = np.array([c[0] for c in cases])
    out["case_repr"] = np.array([c[3] for c in cases])
    out["case_prior"] = np.array([c[4] for c in cases])
    out["case_iters"] = np.array([c[5] for c in cases])
    out["case_views"] = np.array([c[6] for c in cases])
    for k, (name, scene, i, rep, pr, iters, V, kw) in enumerate(cases):
        rec = record_case(sq, torch_oracle, c_oracle, fast_sample_on_batch, prior_by_class, scene, i, rep, pr, iters, V,
                          **kw)
        for x, val in rec.items():
            out[f"w{k}_{x}"] = val
        print(f"wide case {k}: {name}: loss {rec['loss'][0]:.4f} -> {rec['loss'][-1]:.4f}; torch_oracle bit-identical")
    np.savez_compressed(os.path.join(OUT, "ref_runs_wide.npz"), **out)

    # ---- the reference's own call site ----
    for name in ("quaternion", "easydict", "open3d", "matplotlib", "matplotlib.pyplot", "matplotlib.patches",
                 "matplotlib._color_data", "plyfile", "trimesh"):
        sys.modules.setdefault(name, types.ModuleType(name))   # test-only shims for imports the path never calls
    sys.modules["easydict"].EasyDict = dict
    sys.modules["plyfile"].PlyData = sys.modules["plyfile"].PlyElement = None
    import src.scripts.run_multi_view as rmv
    import src.utils.box_utils as box_utils
    scene = synthetic.make_scene(6, 14, seed=31)
    rows = [14, 14, 12, 14, 6, 11]          # object 4: too few views -> keeps its initial quadric
    seq = synthetic.scene_to_tracks(scene, rows, seed=5)
    n_iters, n_views = 10, 10
    ref = rmv.optim_process(seq["tracks"], seq["img_names"], list(seq["T_wcs"]), list(seq["P_cws"]), seq["img_h"],
                            seq["img_w"], seq["K"], "super_quadric", True, n_iters, n_views)
    flat = lambda Q: np.concatenate([Q.translate.detach().numpy().ravel(), np.atleast_1d(Q.angle.detach().numpy()),
                                     Q.scales.detach().numpy().ravel(), Q.shapes.detach().numpy().ravel()]).astype(np.float32)
    cs = dict(rows=np.array(rows), n_iters=n_iters, n_views=n_views, img_names=seq["img_names"], T_wcs=seq["T_wcs"],
              P_cws=seq["P_cws"], K=seq["K"], img_h=seq["img_h"], img_w=seq["img_w"],
              quadrics=np.stack([flat(Q) for Q in ref["quadrics"]]), bboxes_qc=np.stack(ref["bboxes_qc"]),
              bboxes_dl=np.stack(ref["bboxes_dl"]), obj_class=np.array([Q.obj_class for Q in ref["quadrics"]]))
    for i, t in enumerate(seq["tracks"]):
        cs[f"track{i}"] = t
    np.savez_compressed(os.path.join(OUT, "optim_process.npz"), **cs)
    print("optim_process: 6 tracks, 10 iterations; quadrics", cs["quadrics"].shape)

    # ---- run_with_intermediate ----
    i, V, iters = 1, 20, 6
    opt = sq.SuperQuadricOptimizer(s1.translate[i], s1.angle[i], s1.dims[i].copy(), int(s1.cls[i]), "super_quadric", True)
    Q, steps = opt.run_with_intermediate(s1.gt_lines(i)[:V], None, s1.P_cws[i][:V], iters)
    np.savez_compressed(os.path.join(OUT, "intermediate.npz"), obj=i, V=V, iters=iters, final=flat(Q),
                        loss=np.array([float(l[0]) for l in opt.loss_log], np.float32),
                        surface_points=np.stack([s["surface_points"] for s in steps]).astype(np.float32),
                        bbox_qc=np.stack([s["bbox_qc"] for s in steps]))
    print("intermediate: 6 steps")

    # ---- merge cost matrix (N4): the reference's box3d_iou on pairs of oriented boxes ----
    rng = np.random.default_rng(99)
    nb = 24
    boxes = []
    for k in range(nb):
        ctr = rng.uniform(-1.0, 1.0, 3) * np.array([1, 1, 0.3])
        if k % 5 == 1:
            ctr = boxes[-1].mean(0) + rng.normal(0, 0.05, 3)      # near-duplicates (what merging is for)
        boxes.append(box_utils.get_3d_box(rng.uniform(0.3, 1.5, 3), box_utils.rotz(rng.uniform(-np.pi, np.pi)), ctr))
    boxes.append(boxes[0] + np.array([1e-3, -2e-3, 5e-4]))        # an almost exact duplicate (an exact one makes the
                                                                   # reference's clipper divide by zero and raise)
    boxes.append(boxes[3] + np.array([10.0, 0, 0]))                # far away: IoU 0
    for k in range(6):                                             # oriented boxes as the call site produces them
        boxes.append(box_utils.compute_oriented_bbox(out[f"w{k}_final_points"].astype(np.float64)
                                                     - out[f"w{k}_final_points"].mean(0) * (k % 2)))
    boxes = np.stack(boxes)
    iou3d = np.zeros((len(boxes), len(boxes)))
    iou2d = np.zeros_like(iou3d)
    for a in range(len(boxes)):                                    # merge_process evaluates the pairs i < j only
        for b_ in range(a + 1, len(boxes)):
            iou3d[a, b_], iou2d[a, b_] = box_utils.box3d_iou(boxes[a], boxes[b_])
    np.savez_compressed(os.path.join(OUT, "box_iou.npz"), boxes=boxes, iou3d=iou3d, iou2d=iou2d)
    print("box_iou:", boxes.shape, "mean IoU", iou3d.mean())
    print("wrote", OUT)


if __name__ == "__main__":
    main()
