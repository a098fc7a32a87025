"""Shard independent object tracks across the GPUs of one box.

Objects (and scenes) are independent (reference src/scripts/run_multi_view.py:44-69 has no cross-object
state), so the data path has NO collective: every rank optimises a contiguous block of objects, balanced by
the amount of work (sum of views, since cost ~ V per object).  The only communication is one all-gather of
the final [n, 9] parameters (+ final loss) at the end -- NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
import numpy as np


def partition_by_views(view_off, world_size):
    """Contiguous object ranges [lo, hi) per rank, balanced by sum of views.  Returns list of (lo, hi)."""
    view_off = np.asarray(view_off, np.int64)
    n = len(view_off) - 1
    total = int(view_off[-1] - view_off[0])
    bounds = [0]
    for r in range(1, world_size):
        # first object whose prefix work reaches r/world of the total (keeps blocks contiguous and ordered)
        target = view_off[0] + total * r / world_size
        k = int(np.searchsorted(view_off, target, side="left"))
        bounds.append(min(max(k, bounds[-1]), n))
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def all_gather_rows(local, counts, dist, device=None):
    """Gather per-rank row blocks [counts[r], C] into the full [sum(counts), C] array on every rank.

    Uses one all_gather_into_tensor over blocks padded to the largest shard (a few KB..MB: latency-bound,
    NVSwitch makes it uniform), issued on the current stream of `local`'s device.
    """
    import torch
    world = dist.get_world_size()
    pad = int(max(counts))
    C = local.shape[1]
    buf = torch.zeros((pad, C), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = torch.empty((world * pad, C), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf)
    out = out.view(world, pad, C)
    return torch.cat([out[r, : counts[r]] for r in range(world)], 0)


def optimize_sharded(tracks, optimize_fn, dist, rank, world_size):
    """tracks: the FULL PackedTracks on every rank.  optimize_fn(shard) -> torch [n_shard, C] rows (params,
    optionally with extra columns) on the rank's device.  Returns the gathered [n, C] tensor."""
    parts = partition_by_views(tracks.view_off, world_size)
    lo, hi = parts[rank]
    local = optimize_fn(tracks.slice(lo, hi))
    return all_gather_rows(local, [h - l for l, h in parts], dist)


def optimize_on_devices(tracks, prior, n_iters, representation, devices, **kw):
    """ONE process driving several GPUs (the multi-device form of the call site): the packed tracks are split into
    contiguous object blocks balanced by views, one host thread per device runs its block through the host-buffer
    C ABI (ctypes releases the GIL; the library serialises per device, not globally), and the per-block results are
    concatenated in object order.  Objects are batch-invariant, so the result is bit-identical to a single-device call."""
    import threading
    from . import api
    parts = partition_by_views(tracks.view_off, len(devices))
    outs, errs = [None] * len(devices), [None] * len(devices)
    # the launch configuration (CTA size, slices, cluster, code layout) is chosen from the WHOLE batch and handed to every
    # block, so that the result does not depend on how many devices share the work (cluster sizes sum in different orders)
    cfg = api.query_launch(tracks.view_off)
    for k in ("threads", "max_slices", "cluster", "code_layout"):
        kw.setdefault(k, cfg[k])

    def work(r):
        lo, hi = parts[r]
        try:
            if hi > lo:
                outs[r] = api.optimize_host(tracks.slice(lo, hi), prior=prior, n_iters=n_iters,
                                            representation=representation, device=devices[r], **kw)
        except Exception as e:   # re-raised in the caller's thread
            errs[r] = e
    threads = [threading.Thread(target=work, args=(r,)) for r in range(len(devices))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errs:
        if e is not None:
            raise e
    got = [o for o in outs if o is not None]
    return {k: np.concatenate([o[k] for o in got], 0) for k in got[0]}


def optimize_sharded_device(tracks, prior, n_iters, dist, rank, world_size, device, representation="super_quadric",
                            out=None):
    """One process per GPU (torchrun): every rank holds the FULL packed batch, optimises its own contiguous block on
    `device` (one persistent launch on the current stream) and all ranks exchange the final [n, 9] parameters with one
    NCCL all-gather issued on the same stream -- the only communication of the path.  Returns (gathered [n, 9] CUDA
    tensor, this rank's DeviceTracks, its output dict) so that a caller can re-run the step without re-uploading."""
    import torch
    from . import api
    parts = partition_by_views(tracks.view_off, world_size)
    lo, hi = parts[rank]
    dt = api.DeviceTracks(tracks.slice(lo, hi), device, prior)
    res = api.optimize_device(dt, n_iters=n_iters, representation=representation, out=out)
    full = all_gather_rows(res["params"], [h - l for l, h in parts], dist)
    return full, dt, res
