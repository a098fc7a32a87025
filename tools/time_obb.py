"""Device time of the oriented-box kernel (odam_sq_oriented_boxes, device pointers) for n objects."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from odam_b200 import _lib, api, synthetic  # noqa: E402

L = _lib.load()
L.odam_sq_init(0)
for n in (1, 50, 148, 592, 2000):
    scene = synthetic.make_scene(n, 4, seed=3)
    tracks = api.pack_scene(scene)
    rng = np.random.default_rng(0)
    P = tracks.init.copy()
    P[:, 7:9] = rng.uniform(-2, 2, (n, 2))
    p = torch.from_numpy(P).cuda()
    corners = torch.zeros((n, 8, 3), dtype=torch.float64, device="cuda")
    flags = torch.zeros(n, dtype=torch.int32, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    ev[0].record()
    for k in range(5):
        _lib.check(L.odam_sq_oriented_boxes(p.data_ptr(), n, corners.data_ptr(), flags.data_ptr(), None, st))
        ev[k + 1].record()
    torch.cuda.synchronize()
    ts = [ev[k].elapsed_time(ev[k + 1]) for k in range(5)]
    print(f"n={n:5d}: oriented-box kernel {min(ts[1:]) * 1e3:8.1f} us  (first {ts[0] * 1e3:.1f} us)  flagged {int((flags != 0).sum())}")
