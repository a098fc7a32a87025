"""Where does the multi-GPU step go?  Per-rank, per-step kernel time and all-gather time (torchrun, one rank per GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import bench
from odam_b200 import api

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
cfg, scene, tracks, prior = bench.workload(2, rank, None, device=dev)
dt = api.DeviceTracks(tracks, dev, prior)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
gath = torch.empty((world * tracks.n, 9), dtype=torch.float32, device=dev)
out = api.optimize_device(dt, n_iters=200)
for _ in range(3):
    api.optimize_device(dt, n_iters=200, out=out); dist.all_gather_into_tensor(gath, out["params"])
torch.cuda.synchronize(); dist.barrier()
for use_flush in (True, False):
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(12)]
    for k in range(12):
        if use_flush:
            flush.zero_()
        ev[k][0].record(); api.optimize_device(dt, n_iters=200, out=out); ev[k][1].record()
        dist.all_gather_into_tensor(gath, out["params"]); ev[k][2].record()
    torch.cuda.synchronize(); dist.barrier()
    kern = np.array([e[0].elapsed_time(e[1]) for e in ev]); gat = np.array([e[1].elapsed_time(e[2]) for e in ev])
    gap = np.array([ev[k][2].elapsed_time(ev[k + 1][0]) for k in range(11)])
    for r in range(world):
        if r == rank:
            print(f"rank {rank} flush={use_flush}: kernel ms {np.round(kern[2:8], 3)} gather ms {np.round(gat[2:8], 3)} "
                  f"between steps ms {np.round(gap[2:8], 3)}", flush=True)
        dist.barrier()
dist.destroy_process_group()
