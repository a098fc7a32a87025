"""ctypes binding of oracle/sq_oracle.c and of the reference's own C++ sampler.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
N_SAMPLES = 1000
GRID = 201

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def build(quiet=True):
    """(Re)build the checker libraries; a no-op when they are up to date."""
    subprocess.run(["make", "-C", _HERE] + (["-s"] if quiet else []), check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _load(name):
    path = os.path.join(_HERE, "_build", name)
    if not os.path.exists(path):
        build()
    return C.CDLL(path)


_libs = {}


def lib(f64=False):
    key = "f64" if f64 else "f32"
    if key not in _libs:
        L = _load("libsq_oracle_f64.so" if f64 else "libsq_oracle.so")
        L.sq_uniform_stream.argtypes = [C.c_uint32, C.c_int, C.c_void_p]
        L.sq_oracle_sample.argtypes = [C.c_void_p] * 9
        L.sq_oracle_sample.restype = C.c_int
        L.sq_oracle_run.restype = C.c_int
        L.sq_oracle_points.argtypes = [C.c_void_p, C.c_void_p]
        L.sq_oracle_points.restype = C.c_int
        L.sq_dc_grid.argtypes = [C.c_float] * 5 + [C.c_void_p]
        L.sq_dc_grid.restype = C.c_int
        _libs[key] = L
    return _libs[key]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def uniform_stream(n=2 * N_SAMPLES, seed=0):
    out = np.empty(n, np.float32)
    lib().sq_uniform_stream(seed, n, _ptr(out))
    return out


def sample(a, e):
    """a[3], e[2] fp32 -> dict(eta_grid, omega_grid, eta_idx, omega_idx, etas, omegas, cdf, depth)."""
    a = np.ascontiguousarray(a, np.float32).reshape(3)
    e = np.ascontiguousarray(e, np.float32).reshape(2)
    o = dict(eta_grid=np.empty(GRID, np.float32), omega_grid=np.empty(GRID, np.float32),
             eta_idx=np.empty(N_SAMPLES, np.int32), omega_idx=np.empty(N_SAMPLES, np.int32),
             etas=np.empty(N_SAMPLES, np.float32), omegas=np.empty(N_SAMPLES, np.float32),
             cdf=np.empty(GRID, np.float32))
    o["depth"] = lib().sq_oracle_sample(_ptr(a), _ptr(e), _ptr(o["eta_grid"]), _ptr(o["omega_grid"]),
                                        _ptr(o["eta_idx"]), _ptr(o["omega_idx"]),
                                        _ptr(o["etas"]), _ptr(o["omegas"]), _ptr(o["cdf"]))
    return o


def sample_on_batch(shapes, epsilons, n=N_SAMPLES):
    """Same call shape as the reference's fast_sample_on_batch (B=M=1), served by the C restatement."""
    assert n == N_SAMPLES
    o = sample(np.asarray(shapes).reshape(3), np.asarray(epsilons).reshape(2))
    return o["etas"].reshape(1, 1, n).copy(), o["omegas"].reshape(1, 1, n).copy()


def ref_sampler_path():
    return os.path.join(_HERE, "_ref", "libref_sampler.so")


_ref = None


def ref_sample_on_batch(shapes, epsilons, n=N_SAMPLES):
    """The reference's own compiled sampling.cpp (oracle/_ref), called as _sampler.pyx:413-441 does."""
    global _ref
    if _ref is None:
        _ref = C.CDLL(ref_sampler_path())
        _ref.sample_on_batch.argtypes = [C.c_void_p] * 4 + [C.c_int] * 5
        _ref.sample_on_batch.restype = None
    shapes = np.ascontiguousarray(shapes, np.float32)
    epsilons = np.ascontiguousarray(epsilons, np.float32)
    B, M = shapes.shape[0], shapes.shape[1]
    etas = np.zeros((B, M, n), np.float32)
    omegas = np.zeros((B, M, n), np.float32)
    _ref.sample_on_batch(_ptr(shapes), _ptr(epsilons), _ptr(etas), _ptr(omegas), B, M, n, GRID, 0)
    return etas, omegas


def have_ref_sampler():
    return os.path.exists(ref_sampler_path())


def run(init9, Ms, box, mask, prior9=None, n_iters=200, optimize_shapes=True, lr=0.01, lr_shape=0.1,
        m0=None, v0=None, step0=0, s0=None, f64=False, record_indices=False):
    """One object's optimisation trajectory on the CPU restatement.

    Returns dict(params[n_iters,9] (after each step), loss[n_iters] (before each step, incl. prior),
    grad[n_iters,9], arg[n_iters,V,4], pred[n_iters,V,4], m[9], v[9], rc[, eta_idx, omega_idx]).
    """
    init9 = np.ascontiguousarray(init9, np.float32).reshape(9)
    Ms = np.ascontiguousarray(Ms, np.float32).reshape(-1, 12)
    V = Ms.shape[0]
    box = np.ascontiguousarray(box, np.float32).reshape(V, 4)
    mask = np.ascontiguousarray(mask, np.uint8).reshape(V, 4)
    prior9 = None if prior9 is None else np.ascontiguousarray(prior9, np.float32).reshape(9)
    f = lambda x, n: None if x is None else np.ascontiguousarray(x, np.float32).reshape(n)
    m0, v0, s0 = f(m0, 9), f(v0, 9), f(s0, 3)
    o = dict(params=np.zeros((n_iters, 9), np.float32), loss=np.zeros(n_iters, np.float32),
             grad=np.zeros((n_iters, 9), np.float32), arg=np.zeros((n_iters, V, 4), np.int32),
             pred=np.zeros((n_iters, V, 4), np.float32), m=np.zeros(9, np.float32), v=np.zeros(9, np.float32))
    if record_indices:
        o["eta_idx"] = np.zeros((n_iters, N_SAMPLES), np.int32)
        o["omega_idx"] = np.zeros((n_iters, N_SAMPLES), np.int32)
    L = lib(f64)
    L.sq_oracle_run.argtypes = ([C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_double, C.c_double]
                                + [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p] + [C.c_void_p] * 9)
    o["rc"] = L.sq_oracle_run(_ptr(init9), V, _ptr(Ms), _ptr(box), _ptr(mask), _ptr(prior9), n_iters,
                              int(optimize_shapes), lr, lr_shape, _ptr(m0), _ptr(v0), step0, _ptr(s0),
                              _ptr(o["params"]), _ptr(o["loss"]), _ptr(o["grad"]), _ptr(o["arg"]),
                              _ptr(o["pred"]), _ptr(o.get("eta_idx")), _ptr(o.get("omega_idx")),
                              _ptr(o["m"]), _ptr(o["v"]))
    return o


def points(p9):
    p9 = np.ascontiguousarray(p9, np.float32).reshape(9)
    out = np.empty((N_SAMPLES, 3), np.float32)
    rc = lib().sq_oracle_points(_ptr(p9), _ptr(out))
    if rc != 0:
        raise FloatingPointError("sampler hit NaN geometry")
    return out
