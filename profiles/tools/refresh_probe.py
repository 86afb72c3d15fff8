#!/usr/bin/env python
"""Where does the first occupancy refresh after graph capture spend its time?  Prints per-step device times of 40
graphed steps and a host/device breakdown of every refresh (debug tool, not part of the bench)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for _p in (ROOT, os.path.join(ROOT, "single-stable-dreamfusion_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import torch  # noqa: E402

import bench as B  # noqa: E402
from ngp_b200 import provider  # noqa: E402
from ngp_b200.trainer import TrainStep  # noqa: E402

dev = torch.device("cuda:0")
model = B.build_model(dev)
views = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ro, rd = provider.make_training_views(views, 64, 64, seed=0, pin=False)
G = torch.randn(views, 3, 64, 64) * 1e-2
step = TrainStep(model, 64, 64, lr=1e-5, graph=True)
packed = step.pack_inputs(ro, rd, G).to(dev)
orig = model.update_extra_state


def timed_update(*a, **k):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = orig(*a, **k)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("  refresh: host %.2f ms, +device drain %.2f ms, allocator reserved %.0f MB" % (
        (t1 - t0) * 1e3, (t2 - t1) * 1e3, torch.cuda.memory_reserved() / 2**20))
    return out


model.update_extra_state = timed_update
evs = []
for i in range(40):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    evs.append(e)
    step(packed)
e = torch.cuda.Event(enable_timing=True)
e.record()
evs.append(e)
torch.cuda.synchronize()
print(" ".join("%.2f" % a.elapsed_time(b) for a, b in zip(evs, evs[1:])))
