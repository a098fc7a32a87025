"""Helpers shared by the oracle and GPU parity tests: golden reference trajectories -> packed problems."""
import numpy as np

from odam_b200.api import PackedTracks

TOL_PARAM = 1e-4   # BASELINE.json: relative error on parameters (denominator floored at 1e-3)
TOL_LOSS = 1e-5    # BASELINE.json: relative error on the per-iteration loss
FLOOR = 1e-3


def rel_param(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), FLOOR)


def rel_loss(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-30)


class Case:
    def __init__(self, G, k):
        self.k = k
        self.obj = int(G["case_obj"][k])
        self.repr = str(G["case_repr"][k])
        self.use_prior = bool(G["case_prior"][k])
        self.iters = int(G["case_iters"][k])
        self.V = int(G["case_views"][k])
        i, V = self.obj, self.V
        self.Ms = np.ascontiguousarray(G["P_cws"][i][:V].reshape(V, 12), np.float32)
        self.box = np.ascontiguousarray(G["box"][i][:V], np.float32)
        self.mask = np.ascontiguousarray(G["mask"][i][:V], np.uint8)
        self.cls = int(G["cls"][i])
        self.prior33 = G["prior_by_class"][self.cls] if self.use_prior else None
        self.prior_table = np.stack([G["prior_by_class"][c].astype(np.float32).reshape(9) for c in range(8)]) \
            if self.use_prior else None
        self.init = G[f"c{k}_init"]
        for x in ("params", "grad", "m", "v", "loss", "final_points", "arg", "eta_idx", "resid_sign"):
            setattr(self, x, G[f"c{k}_{x}"])

    def states_before(self):
        """(params, m, v) BEFORE each step s = 0..iters-1 (teacher-forcing inputs)."""
        z = np.zeros((1, 9), np.float32)
        return (np.vstack([self.init[None], self.params[:-1]]), np.vstack([z, self.m[:-1]]),
                np.vstack([z, self.v[:-1]]))

    def tracks(self, inits):
        """n = len(inits) copies of this object's views, one per initial state."""
        n, V = len(inits), self.V
        return PackedTracks(init=np.ascontiguousarray(inits, np.float32), cls=np.full(n, self.cls, np.int32),
                            view_off=(np.arange(n + 1) * V).astype(np.int32), Ms=np.tile(self.Ms, (n, 1)),
                            box=np.tile(self.box, (n, 1)), mask=np.tile(self.mask, (n, 1)))


def all_cases(G):
    return [Case(G, k) for k in range(len(G["case_obj"]))]


class WideCase(Case):
    """A case of tests/golden/ref_runs_wide.npz (round 2: the BASELINE view counts and the loop's edge cases); every
    case carries its own views, so nothing is indexed through a shared scene."""

    def __init__(self, W, k):
        self.k = k
        self.name = str(W["names"][k])
        self.repr = str(W["case_repr"][k])
        self.use_prior = bool(W["case_prior"][k])
        self.iters = int(W["case_iters"][k])
        self.V = V = int(W["case_views"][k])
        self.Ms64, self.box64 = W[f"w{k}_Ms"], W[f"w{k}_box"]
        self.Ms = np.ascontiguousarray(self.Ms64.reshape(V, 12), np.float32)
        self.box = np.ascontiguousarray(self.box64, np.float32)
        self.mask = np.ascontiguousarray(W[f"w{k}_mask"], np.uint8)
        self.cls = int(W[f"w{k}_cls"])
        self.translate, self.angle, self.dims = W[f"w{k}_translate"], float(W[f"w{k}_angle"]), W[f"w{k}_dims"]
        self.prior33 = W["prior_by_class"][self.cls] if self.use_prior else None
        self.prior_table = np.stack([W["prior_by_class"][c].astype(np.float32).reshape(9) for c in range(8)]) \
            if self.use_prior else None
        self.init = W[f"w{k}_init"]
        for x in ("params", "grad", "m", "v", "loss", "final_points", "arg", "eta_idx", "resid_sign"):
            setattr(self, x, W[f"w{k}_{x}"])


def wide_cases(W):
    return [WideCase(W, k) for k in range(len(W["names"]))]
