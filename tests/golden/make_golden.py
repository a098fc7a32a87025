"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (build container only).

The reference is Python + one Cython/C++ extension; it cannot travel to the GPU box, so its
outputs are committed as small fixtures and every oracle in oracle/ is pinned against them.

    python tests/golden/make_golden.py            # needs /root/reference; writes tests/golden/

What it does
  1. builds the reference's sampler extension in a scratch dir (nothing is
     written under /root/reference, no reference source is copied into this repo),
  2. imports src.super_quadric.sq_libs from /root/reference (cwd = /root/reference because
     sq_libs.py:388 opens ./src/super_quadric/scale_prior),
  3. records sampler known-answer vectors (fast_sample_on_batch) and full optimiser trajectories
     (SuperQuadricOptimizer.run with a hook on optimizer.step) on seeded synthetic scenes,
  4. asserts that oracle/torch_oracle.py reproduces those trajectories BIT FOR BIT here, and that
     oracle/sq_oracle.c reproduces the sampler bit for bit,
  5. exports the scale-prior matrices (a data file of the reference) to odam_b200/data/.
"""
import json
import os
import pickle
import shutil
import subprocess
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("ODAM_REFERENCE", "/root/reference")
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)


def build_reference_extension():
    b = os.path.join(tempfile.gettempdir(), "odam_ref_build")
    so = [f for f in os.listdir(os.path.join(b, "learnable_primitives", "fast_sampler"))
          if f.startswith("_sampler") and f.endswith(".so")] if os.path.isdir(b) else []
    if not so:
        shutil.rmtree(b, ignore_errors=True)
        os.makedirs(b)
        sq = os.path.join(REF, "src", "super_quadric")
        shutil.copytree(os.path.join(sq, "learnable_primitives"), os.path.join(b, "learnable_primitives"))
        shutil.copy(os.path.join(sq, "setup.py"), b)
        os.remove(os.path.join(b, "learnable_primitives", "fast_sampler", "_sampler.c"))
        subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=b, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return b


def record_reference_run(sq, scene, i, representation, prior, n_iters, V=None):
    """SuperQuadricOptimizer(...).run(...) on object i with a per-step hook."""
    import torch
    V = V or scene.V
    opt = sq.SuperQuadricOptimizer(scene.translate[i], scene.angle[i], scene.dims[i].copy(),
                                   int(scene.cls[i]), representation, prior)
    Q = opt.Q_init
    leaves = (Q.translate, Q.angle, Q.scales, Q.shapes)
    flat = lambda xs: np.concatenate([np.atleast_1d(x.detach().numpy().astype(np.float32)).ravel() for x in xs])
    rec = dict(init=flat(leaves).copy(), params=[], grad=[], m=[], v=[])
    adam = opt.optimizer
    orig_step = adam.step

    def step(*a, **k):
        rec["grad"].append(flat([x.grad if x.grad is not None else torch.zeros_like(x) for x in leaves]))
        r = orig_step(*a, **k)
        rec["params"].append(flat(leaves))
        for key, name in (("m", "exp_avg"), ("v", "exp_avg_sq")):
            rec[key].append(flat([adam.state[x][name] if x in adam.state else torch.zeros_like(x) for x in leaves]))
        return r

    adam.step = step
    opt.run(scene.gt_lines(i)[:V], None, scene.P_cws[i][:V], n_iters)
    rec["loss"] = np.array([float(l[0]) for l in opt.loss_log], np.float32)
    for k in ("params", "grad", "m", "v"):
        rec[k] = np.stack(rec[k]).astype(np.float32)
    rec["final_points"] = Q.compute_ellipsoid_points(use_numpy=True)[0].astype(np.float32)
    return rec


def prepare_tracks_fixture(sq):
    """The body of OdamProcess._prepare_tracks' loop (reference src/processor.py:188-205) run with the reference's own
    SuperQuadric / geometry helpers on synthetic tracks: per track the projected box of its mean-pose quadric in one
    camera.  (The method itself needs the detector/associator state of OdamProcess; the loop body is what the
    drop-in's odam_b200.processor.prepare_track_boxes mirrors.)"""
    import src.utils.geometry_utils as geo_utils
    from odam_b200 import synthetic
    rng = np.random.default_rng(88)
    K = synthetic.K
    # a camera 3 m back, 1.4 m up, looking at the origin region (+z up world)
    eye, at = np.array([3.0, 0.4, 1.4]), np.array([0.0, 0.0, 0.4])
    f = (at - eye) / np.linalg.norm(at - eye); r = np.cross(f, [0, 0, 1.0]); r /= np.linalg.norm(r); d = np.cross(f, r)
    T_wc = np.eye(4); T_wc[:3, 0], T_wc[:3, 1], T_wc[:3, 2], T_wc[:3, 3] = r, d, f, eye
    out = dict(T_wc=T_wc, K=K)
    for t in range(6):
        nrow = int(rng.integers(1, 12))
        track = -np.ones((nrow, 82))
        track[:, 6:9] = rng.uniform(0.02 if t == 5 else 0.3, 1.5, 3)[None] + rng.normal(0, 0.02, (nrow, 3))   # t=5: clipped dims
        track[:, 9:12] = rng.uniform(-0.8, 0.8, 3)[None] * np.array([1, 1, 0.3]) + rng.normal(0, 0.05, (nrow, 3))
        track[:, 12] = rng.uniform(-np.pi, np.pi) + rng.normal(0, 0.1, nrow)
        azi_wo = np.mean(track[:, 12], axis=0)
        t_wo = np.mean(track[:, 9: 12], axis=0)
        dimensions = np.clip(np.mean(track[:, 6: 9], axis=0), a_min=0.05, a_max=np.inf)
        Q = sq.SuperQuadric(t_wo, azi_wo, np.sqrt(dimensions / 2), shapes=np.array([-0., -0.]))
        pts, _ = Q.compute_ellipsoid_points(use_numpy=True)
        box_3d_c = (geo_utils.get_homogeneous(pts) @ np.linalg.inv(T_wc).T)[:, :3]
        pixels = geo_utils.projection(box_3d_c, K)
        x_min, y_min, _ = np.min(pixels, axis=0)
        x_max, y_max, _ = np.max(pixels, axis=0)
        out.update({f"t{t}_track": track, f"t{t}_pts": pts, f"t{t}_box": np.array([x_min, y_min, x_max, y_max])})
    np.savez_compressed(os.path.join(OUT, "prepare_tracks.npz"), **out)
    print("wrote prepare_tracks.npz")


def main():
    b = build_reference_extension()
    sys.path[:0] = [REF, b]
    os.chdir(REF)
    import torch
    torch.set_num_threads(1)
    import src.super_quadric.sq_libs as sq
    if "--only-prepare-tracks" in sys.argv:   # add this fixture without regenerating the others
        prepare_tracks_fixture(sq)
        return
    from learnable_primitives.fast_sampler import fast_sample_on_batch
    from oracle import c_oracle, torch_oracle
    from odam_b200 import synthetic

    c_oracle.build()

    # ---- 5. prior matrices (data) ----
    with open(os.path.join(REF, "src", "super_quadric", "scale_prior"), "rb") as f:
        prior = pickle.load(f)
    os.makedirs(os.path.join(REPO, "odam_b200", "data"), exist_ok=True)
    with open(os.path.join(REPO, "odam_b200", "data", "scale_prior.json"), "w") as f:
        json.dump({"source": "reference data file src/super_quadric/scale_prior (ShapeNet synset id -> 3x3 inverse "
                             "covariance of sqrt(dim/2), float64); exported by tests/golden/make_golden.py",
                   "class_mapper": {str(k): v for k, v in sq.CLASS_MAPPER.items()},
                   "matrices": {k: np.asarray(v, np.float64).tolist() for k, v in prior.items()}}, f, indent=1)
    prior_by_class = np.stack([np.asarray(prior[sq.CLASS_MAPPER[c]], np.float64) for c in range(8)])

    # ---- 3a. sampler known-answer vectors ----
    rng = np.random.default_rng(2024)
    A = [np.array([.5, .4, .3]), np.array([.25, .25, .25]), np.array([0.15, 0.15, 0.35])]
    E = [np.array([.9, .9]), np.array([.2, .2]), np.array([0.20715195, 1.3855394])]
    for _ in range(29):
        A.append(rng.uniform(0.1, 0.8, 3))
        E.append(rng.uniform(0.2, 1.6, 2))
    for e1 in (0.2, 0.200001, 1.6):  # cube-like / extreme exponents (non-monotone CDF tail)
        A.append(rng.uniform(0.1, 0.8, 3))
        E.append(np.array([e1, rng.uniform(0.2, 1.6)]))
    A = np.asarray(A, np.float32)
    E = np.asarray(E, np.float32)
    etas = np.zeros((len(A), 1000), np.float32)
    omegas = np.zeros((len(A), 1000), np.float32)
    for k in range(len(A)):
        et, om = fast_sample_on_batch(A[k].reshape(1, 1, 3), E[k].reshape(1, 1, 2), 1000)
        etas[k], omegas[k] = et.ravel(), om.ravel()
        et2, om2 = fast_sample_on_batch(A[k].reshape(1, 1, 3), E[k].reshape(1, 1, 2), 1000)
        assert np.array_equal(et, et2) and np.array_equal(om, om2)            # stateless
        o = c_oracle.sample(A[k], E[k])
        assert np.array_equal(o["etas"], etas[k]) and np.array_equal(o["omegas"], omegas[k]), k
        et3, om3 = c_oracle.ref_sample_on_batch(A[k].reshape(1, 1, 3), E[k].reshape(1, 1, 2))
        assert np.array_equal(et3.ravel(), etas[k]) and np.array_equal(om3.ravel(), omegas[k]), k
    np.savez_compressed(os.path.join(OUT, "sampler_kat.npz"), a=A, e=E, etas=etas, omegas=omegas,
                        uniforms_head=c_oracle.uniform_stream()[:8])
    print(f"sampler: {len(A)} parameter sets; C restatement and oracle/_ref agree bit for bit")

    # ---- 3b. optimiser trajectories ----
    scene = synthetic.make_scene(10, 20, seed=1)
    cases = [(i, "super_quadric", True, 200, 20) for i in range(6)]
    cases += [(6, "super_quadric", False, 60, 20), (7, "cube", True, 60, 20), (8, "quadric", True, 60, 20),
              (9, "super_quadric", True, 60, 11)]
    out = dict(translate=scene.translate, angle=scene.angle, dims=scene.dims, cls=scene.cls, P_cws=scene.P_cws,
               box=scene.box, mask=scene.mask, prior_by_class=prior_by_class,
               case_obj=np.array([c[0] for c in cases]), case_repr=np.array([c[1] for c in cases]),
               case_prior=np.array([c[2] for c in cases]), case_iters=np.array([c[3] for c in cases]),
               case_views=np.array([c[4] for c in cases]))
    for k, (i, rep, pr, iters, V) in enumerate(cases):
        rec = record_reference_run(sq, scene, i, rep, pr, iters, V)
        rec2 = record_reference_run(sq, scene, i, rep, pr, iters, V)
        assert all(np.array_equal(rec[x], rec2[x]) for x in rec), "reference is not run-to-run deterministic?"
        t = torch_oracle.run(scene.translate[i], scene.angle[i], scene.dims[i], scene.P_cws[i][:V], scene.box[i][:V],
                             scene.mask[i][:V], prior_by_class[scene.cls[i]] if pr else None, iters, rep,
                             sampler=fast_sample_on_batch, anomaly=False)
        for x in ("params", "grad", "m", "v", "loss"):
            assert np.array_equal(t[x], rec[x]), (k, x, np.abs(t[x] - rec[x]).max())
        assert np.array_equal(torch_oracle.points(rec["params"][-1], fast_sample_on_batch), rec["final_points"])
        for x, val in rec.items():
            out[f"c{k}_{x}"] = val
        # the reference's discrete decisions per iteration (from the bit-identical torch restatement):
        # arg-extreme sample per view/side, and the eta-grid bucket of every sample
        out[f"c{k}_arg"] = t["arg"].astype(np.int16)
        # ... and the sign of every residual pred - target (the L1 loss is non-smooth there: a residual within
        # 1 ulp of zero flips the sign of that side's whole gradient contribution)
        out[f"c{k}_resid_sign"] = np.sign(t["pred"] - scene.box[i][:V].astype(np.float32)[None]).astype(np.int8)
        eta_idx = np.zeros((iters, 1000), np.uint8)
        for it in range(iters):
            o = c_oracle.sample(t["ae"][it, :3], t["ae"][it, 3:])
            et = o["etas"].copy()
            et[et == 0] += np.float32(1e-6)
            assert np.array_equal(et, t["etas"][it])
            eta_idx[it] = o["eta_idx"]
        out[f"c{k}_eta_idx"] = eta_idx
        print(f"case {k}: obj {i} {rep} prior={pr} iters={iters} V={V}: loss {rec['loss'][0]:.4f} -> "
              f"{rec['loss'][-1]:.4f}; torch_oracle bit-identical")
    np.savez_compressed(os.path.join(OUT, "ref_runs.npz"), **out)

    # ---- 3c. call-site helpers (staging before, oriented box after the optimiser) ----
    import types
    for name in ("quaternion", "easydict", "open3d", "matplotlib", "matplotlib.pyplot", "matplotlib.patches",
                 "matplotlib._color_data", "plyfile", "trimesh"):
        sys.modules.setdefault(name, types.ModuleType(name))   # test-only shims for imports the path never calls
    sys.modules["easydict"].EasyDict = dict
    sys.modules["plyfile"].PlyData = sys.modules["plyfile"].PlyElement = None
    import src.utils.box_utils as box_utils
    import src.utils.tracking_gt_utils as tgu
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(77)
    n_frames = 30
    frame_ids = np.arange(100, 100 + n_frames)
    T_wcs = [np.eye(4) for _ in range(n_frames)]
    K = synthetic.K
    cs = dict(frame_ids=frame_ids)
    for t in range(4):
        nrow = int(rng.integers(5, 25))
        frames = np.sort(rng.choice(frame_ids, nrow, replace=False))
        track = -np.ones((nrow, 82))
        track[:, 0] = frames
        track[:, 1] = rng.integers(0, 8, nrow)
        x0, y0 = rng.uniform(-30, 900, nrow), rng.uniform(-30, 600, nrow)
        track[:, 2:6] = np.stack([x0, y0, x0 + rng.uniform(40, 500, nrow), y0 + rng.uniform(40, 420, nrow)], 1)
        track[:, 6:9] = rng.uniform(0.3, 1.5, (nrow, 3))
        track[:, 9:12] = rng.normal(0, 1, 3)[None] + rng.normal(0, 0.05, (nrow, 3))
        track[:, 12] = rng.uniform(-0.3, 0.3) + rng.normal(0, 0.1, nrow)
        track[:, 13] = rng.uniform(0.5, 1, nrow)
        lines, bl, pv, oc, T_wos, scales, dp = tgu.load_pred_object(track, frame_ids, T_wcs, 968, 1296, K)
        T_wo = tgu.averaging_T_wos(T_wos)
        sc = np.mean(np.asarray([x for x in scales if len(x) > 0]), axis=0)
        valid = [i for i in range(n_frames) if len(bl[i]) > 0]
        boxv = np.zeros((len(valid), 4)); maskv = np.zeros((len(valid), 4), np.uint8)
        for a, i in enumerate(valid):
            for b, nm in enumerate(("x_min", "x_max", "y_min", "y_max")):
                if nm in bl[i]:
                    boxv[a, b] = -bl[i][nm][-1]; maskv[a, b] = 1
        cs.update({f"t{t}_track": track, f"t{t}_class": oc, f"t{t}_T_wo": T_wo, f"t{t}_dims": sc,
                   f"t{t}_valid": np.array(valid), f"t{t}_box": boxv, f"t{t}_mask": maskv,
                   f"t{t}_yaw": Rotation.from_matrix(T_wo[:3, :3]).as_euler("zxy")[0],
                   f"t{t}_bbox_dl": box_utils.get_3d_box(sc, T_wo[:3, :3], T_wo[:3, 3])})
    for k in range(6):
        cs[f"obb{k}_pts"] = out[f"c{k}_final_points"]
        cs[f"obb{k}_box"] = box_utils.compute_oriented_bbox(out[f"c{k}_final_points"].astype(np.float64))
    np.savez_compressed(os.path.join(OUT, "call_site.npz"), **cs)
    prepare_tracks_fixture(sq)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
