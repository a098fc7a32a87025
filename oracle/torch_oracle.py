"""PyTorch/autograd restatement of the reference's optimiser loop -- TEST INFRASTRUCTURE ONLY.

Same tensor ops, in the same order and with the same shapes, as
  /root/reference/src/super_quadric/sq_libs.py:353-393 (construction, Adam groups),
  :395-430 (projection, masked extrema, L1), :432-475 (run loop, prior term),
  :556-595 (rotz, squashing, points) and
  /root/reference/src/super_quadric/learnable_primitives/sampling.py:586-615 (angles -> surface),
so that on one host it reproduces the reference bit for bit
(tests/golden/make_golden.py asserts exactly that against the imported reference;
tests/test_oracle_golden.py re-checks it against the committed vectors).

Parity status: PINNED against outputs of the reference itself (tests/golden/).

It is also the ``cpu_baseline`` / ``--impl reference`` arm of bench.py: the
reference's Python cannot travel to the GPU box, this port of it can.
"""
import numpy as np
import torch

from . import c_oracle

SIDES = ("x_min", "x_max", "y_min", "y_max")


def default_sampler():
    """The reference's own compiled sampler when oracle/_ref exists, else the C restatement."""
    return c_oracle.ref_sample_on_batch if c_oracle.have_ref_sampler() else c_oracle.sample_on_batch


class _NoAnomaly:
    def __init__(self, *_a, **_k):
        pass


def surface_points(translate, angle, scales, shapes, sampler):
    """sq_libs.py:577-595 + sampling.py:586-615.  Returns (pts[1000,3], etas, omegas)."""
    cosz = torch.cos(angle)
    sinz = torch.sin(angle)
    zeros = angle.detach() * 0
    ones = zeros.detach() + 1
    R = torch.stack([cosz, -sinz, zeros, sinz, cosz, zeros, zeros, zeros, ones], dim=0).reshape(3, 3)
    a = (scales ** 2).unsqueeze(0).unsqueeze(0)
    e = (torch.sigmoid(shapes) * (1.6 - 0.2) + 0.2).unsqueeze(0).unsqueeze(0)
    etas, omegas = sampler(a.detach().cpu().numpy(), e.detach().cpu().numpy(), 1000)
    etas[etas == 0] += 1e-6
    omegas[omegas == 0] += 1e-6
    etas = a.new_tensor(etas)
    omegas = a.new_tensor(omegas)
    a1 = a[:, :, 0].unsqueeze(-1)
    a2 = a[:, :, 1].unsqueeze(-1)
    a3 = a[:, :, 2].unsqueeze(-1)
    e1 = e[:, :, 0].unsqueeze(-1)
    e2 = e[:, :, 1].unsqueeze(-1)

    def sp(x, p):
        return torch.sign(x) * (torch.abs(x) ** p)

    x = a1 * sp(torch.cos(etas), e1) * sp(torch.cos(omegas), e2)
    y = a2 * sp(torch.cos(etas), e1) * sp(torch.sin(omegas), e2)
    z = a3 * sp(torch.sin(etas), e1)
    lim = x.new_tensor(1e-6)
    x = ((x > 0).float() * 2 - 1) * torch.max(torch.abs(x), lim)
    y = ((y > 0).float() * 2 - 1) * torch.max(torch.abs(y), lim)
    z = ((z > 0).float() * 2 - 1) * torch.max(torch.abs(z), lim)
    pts = torch.stack([x, y, z], -1)[0, 0]
    pts = pts @ R.T
    pts = pts + translate.unsqueeze(0)
    surface_points.last_ae = (a.detach().numpy().reshape(3).copy(), e.detach().numpy().reshape(2).copy())
    return pts, etas, omegas


def box_loss(pts_w, Ms, target, mask, want_arg=False):
    """sq_libs.py:395-430.  target/mask: [V,4] in SIDES order; returns (loss, pred[V,4], arg[V,4])."""
    V = Ms.shape[0]
    homo = torch.cat([pts_w, torch.ones_like(pts_w[:, 2:])], dim=1)[None, :, :]
    pix = homo @ Ms.permute(0, 2, 1)
    valid = pix[:, :, 2] > 0.5
    pix = pix[:, :, :2] / (torch.abs(pix[:, :, 2:]) + 1e-6)
    big = torch.ones_like(pix[:, :, 0]) * 1000000
    ext = (
        torch.min(torch.where(valid, pix[:, :, 0], big), dim=1),
        torch.max(torch.where(valid, pix[:, :, 0], -big), dim=1),
        torch.min(torch.where(valid, pix[:, :, 1], big), dim=1),
        torch.max(torch.where(valid, pix[:, :, 1], -big), dim=1),
    )
    loss = 0
    for s in range(4):
        l = torch.nn.functional.l1_loss(ext[s].values, target[:, s], reduction="none")
        l = torch.where(torch.isnan(l), torch.zeros_like(l), l)
        l = l * mask[:, s]
        loss += torch.mean(l)
    if not want_arg:
        return loss, None, None
    pred = torch.stack([x.values for x in ext], 1).detach()
    arg = torch.stack([x.indices for x in ext], 1)
    rows = torch.arange(V)
    for s in range(4):  # a winning sentinel carries no gradient
        arg[:, s] = torch.where(valid[rows, arg[:, s]], arg[:, s], torch.full_like(arg[:, s], -1))
    return loss, pred, arg


def run(translate, angle, dims, Ms, box, mask, prior33=None, n_iters=200,
        representation="super_quadric", sampler=None, anomaly=True, record=True, params_are_scales=False):
    """One object, as SuperQuadricOptimizer(...).run(...) does it.

    dims are the 3-D box dimensions (the constructor takes sqrt(dims/2), sq_libs.py:361);
    box/mask [V,4] in SIDES order, box in pixels (= -line[-1]); prior33 = 3x3 matrix or None.
    Returns dict(final[9], loss[n_iters], and when record: params/grad/m/v [n_iters,9] after each
    step, arg/pred [n_iters,V,4], etas/omegas [n_iters,1000]).
    """
    sampler = sampler or default_sampler()
    f32 = torch.float32
    scales0 = np.asarray(dims, np.float64) if params_are_scales else np.sqrt(np.asarray(dims, np.float64) / 2)
    shapes0 = np.array([-10000., -10000.]) if representation == "cube" else np.array([-0., -0.])
    shapes = torch.tensor(shapes0, dtype=f32, requires_grad=True)
    t = torch.tensor(np.asarray(translate, np.float64), dtype=f32, requires_grad=True)
    ang = torch.tensor(float(angle), dtype=f32, requires_grad=True)
    scales = torch.tensor(scales0, dtype=f32, requires_grad=True)
    groups = [{"params": [t, ang, scales]}]
    if representation == "super_quadric":
        groups.append({"params": [shapes], "lr": 0.1})
    opt = torch.optim.Adam(groups, lr=0.01)
    Ms = torch.tensor(np.asarray(Ms).reshape(-1, 3, 4)).float()
    target = torch.tensor(np.asarray(box, np.float64)).float()
    maskt = torch.tensor(np.asarray(mask)).float()
    A = None if prior33 is None else torch.tensor(np.asarray(prior33, np.float64).reshape(3, 3)).float()
    scales_init = scales.detach().clone()
    leaves = (t, ang, scales, shapes)
    flat = lambda xs: np.concatenate([np.atleast_1d(x.detach().numpy().astype(np.float32)).ravel() for x in xs])
    out = dict(loss=np.zeros(n_iters, np.float32), init=flat(leaves).copy())
    if record:
        V = Ms.shape[0]
        for k in ("params", "grad", "m", "v"):
            out[k] = np.zeros((n_iters, 9), np.float32)
        out["arg"] = np.zeros((n_iters, V, 4), np.int64)
        out["pred"] = np.zeros((n_iters, V, 4), np.float32)
        out["etas"] = np.zeros((n_iters, 1000), np.float32)
        out["omegas"] = np.zeros((n_iters, 1000), np.float32)
        out["ae"] = np.zeros((n_iters, 5), np.float32)   # the (a, e) handed to the sampler
    guard = torch.autograd.set_detect_anomaly if anomaly else _NoAnomaly
    try:
        for it in range(n_iters):
            guard(True)
            opt.zero_grad()
            pts, etas, omegas = surface_points(t, ang, scales, shapes, sampler)
            loss, pred, arg = box_loss(pts, Ms, target, maskt, want_arg=record)
            if A is not None:
                # two separate difference nodes, as at sq_libs.py:465 (keeps autograd's
                # accumulation order into scales.grad identical)
                loss3 = (scales_init - scales)[None, :] @ A @ (scales_init - scales)[None, :].T
                loss = loss + loss3[0, 0] * 20
            loss.backward()
            out["loss"][it] = float(loss.detach())
            if record:
                g = [x.grad if x.grad is not None else torch.zeros_like(x) for x in leaves]
                out["grad"][it] = flat(g)
                out["arg"][it] = arg.numpy()
                out["pred"][it] = pred.numpy()
                out["etas"][it] = etas.numpy().ravel()
                out["omegas"][it] = omegas.numpy().ravel()
                out["ae"][it] = np.concatenate(surface_points.last_ae)
            opt.step()
            if record:
                out["params"][it] = flat(leaves)
                for key, name in (("m", "exp_avg"), ("v", "exp_avg_sq")):
                    out[key][it] = flat([opt.state[x][name] if x in opt.state else torch.zeros_like(x)
                                         for x in leaves])
    finally:
        if anomaly:
            torch.autograd.set_detect_anomaly(False)
    out["final"] = flat(leaves)
    return out


def points(p9, sampler=None):
    """compute_ellipsoid_points(use_numpy=True) for a packed parameter vector (t3, angle, s3, h2)."""
    sampler = sampler or default_sampler()
    p = torch.tensor(np.asarray(p9, np.float32))
    with torch.no_grad():
        pts, _, _ = surface_points(p[0:3], p[3], p[4:7], p[7:9], sampler)
    return pts.numpy()
