"""How many marched samples of a bench-config step carry exactly zero gradient (they lie behind the early-termination
point of their ray, raymarching.cu:557/:672), per sample, per aligned warp of 32 and per 128-sample tile."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "single-stable-dreamfusion_b200"))
import torch
import bench
from ngp_b200 import provider
from ngp_b200.trainer import TrainStep
dev = torch.device("cuda", 0)
model = bench.build_model(dev)
views = 8
step = TrainStep(model, 64, 64, lr=1e-5, max_steps=1024, graph=False, manual=True, n_chunks=1)
ro, rd = provider.make_training_views(views, 64, 64, seed=0, pin=False)
G = torch.randn(views, 3, 64, 64, generator=torch.Generator().manual_seed(2)) * 1e-2
step(ro.to(dev), rd.to(dev), G.to(dev))
torch.cuda.synchronize()
ws = step._mws["chunks"][0][2]
n = int(ws.counter[0].item())
dead = (ws.d_sigma[:n] == 0) & (ws.d_rgb[:n] == 0).all(-1)
pad = (-n) % 128
d = torch.cat([dead, torch.ones(pad, dtype=torch.bool, device=dev)])
out = {"samples": n, "dead_fraction": dead.float().mean().item(),
       "dead_warps32": d.view(-1, 32).all(-1).float().mean().item(),
       "dead_tiles128": d.view(-1, 128).all(-1).float().mean().item(),
       "denc_zero_rows": (ws.d_enc[:n] == 0).all(-1).float().mean().item()}
print(json.dumps(out))
