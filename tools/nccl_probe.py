"""Latency of the [n,9] parameter all-gather on this box (torchrun, one rank per GPU): NCCL vs the time budget of a step."""
import os
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
n = 50
src = torch.randn(n, 9, device="cuda")
dst = torch.empty(world * n, 9, device="cuda")
for _ in range(20):
    dist.all_gather_into_tensor(dst, src)
torch.cuda.synchronize()
dist.barrier()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(101)]
ev[0].record()
for k in range(100):
    dist.all_gather_into_tensor(dst, src)
    ev[k + 1].record()
torch.cuda.synchronize()
t = sorted(ev[k].elapsed_time(ev[k + 1]) * 1e3 for k in range(100))
# with a 2.4 ms kernel in front of it on every rank (the bench's pattern)
spin = torch.empty(1 << 26, device="cuda")
ev2 = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(30)]
for k in range(30):
    spin.zero_(); spin.add_(1.0); spin.mul_(0.5)
    ev2[k][0].record()
    dist.all_gather_into_tensor(dst, src)
    ev2[k][1].record()
torch.cuda.synchronize()
t2 = sorted(a.elapsed_time(b) * 1e3 for a, b in ev2)
if rank == 0:
    print(f"all_gather_into_tensor [{n},9] x{world}: back-to-back median {t[50]:.1f} us (min {t[0]:.1f}); "
          f"after compute median {t2[15]:.1f} us (min {t2[0]:.1f})")
dist.destroy_process_group()
