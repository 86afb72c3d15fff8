"""Debug helper (not a test): fused vs unfused field gradients against an fp64 autograd reference."""
import argparse, sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "single-stable-dreamfusion_b200"))
from ngp_b200.network_grid import NeRFNetwork
DEV = "cuda:0"
opt = argparse.Namespace(bound=1, cuda_ray=True, min_near=0.1, density_thresh=10, bg_radius=1.4)
torch.manual_seed(0)
m = NeRFNetwork(opt).to(DEV)
with torch.no_grad():
    m.encoder.embeddings.uniform_(-0.5, 0.5)
M = 70000
g = torch.Generator(device=DEV).manual_seed(3)
x = (torch.rand(M, 3, device=DEV, generator=g) * 2 - 1) * 0.9
gs = torch.randn(M, device=DEV, generator=g) * 0.1
ga = torch.randn(M, 3, device=DEV, generator=g)
grads = {}
for fused in (True, False):
    m.fused = fused
    m.zero_grad(set_to_none=True)
    with torch.autocast("cuda", torch.float16):
        s, a = m.common_forward(x)
        (s * gs).sum().add((a.float() * ga).sum()).backward()
    grads[fused] = {n: p.grad.detach().clone().double() for n, p in m.named_parameters() if p.grad is not None}
# fp64 reference with fp16-quantised weights/table, enc from the (bit-exact) encoder in fp32 path
m.fused = False
m.zero_grad(set_to_none=True)
emb = m.encoder.embeddings
with torch.autocast("cuda", torch.float16):
    enc = m.encoder(x, bound=1)          # half, differentiable wrt embeddings (fp32 accumulate backward)
W = [l.weight.half().double().requires_grad_(True) for l in m.sigma_net.net]
B = [l.bias.half().double().requires_grad_(True) for l in m.sigma_net.net]
e64 = enc.double()
h = torch.relu(e64 @ W[0].T + B[0]); h = torch.relu(h @ W[1].T + B[1]); o = h @ W[2].T + B[2]
blob = 5 * torch.exp(-(x.double() ** 2).sum(-1) / 0.08)
s64 = torch.exp(o[:, 0] + blob); a64 = torch.sigmoid(o[:, 1:])
loss = (s64 * gs.double()).sum() + (a64 * ga.double()).sum()
loss.backward()
ref = {"encoder.embeddings": emb.grad.double()}
for i in range(3):
    ref["sigma_net.net.%d.weight" % i] = W[i].grad; ref["sigma_net.net.%d.bias" % i] = B[i].grad
for n in ref:
    r = ref[n]
    f = (grads[True][n] - r).norm() / r.norm(); u = (grads[False][n] - r).norm() / r.norm()
    fu = (grads[True][n] - grads[False][n]).norm() / grads[False][n].norm()
    print("%-28s fused-vs-fp64 %.3e  unfused-vs-fp64 %.3e  fused-vs-unfused %.3e" % (n, f.item(), u.item(), fu.item()))
