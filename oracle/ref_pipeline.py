"""TEST INFRASTRUCTURE / MEASUREMENT: the reference's `-O` train step driven through the reference's OWN
CUDA extensions (oracle/_ref, built unmodified from /root/reference), with the reference's host-side
call pattern restated around them.  This is the ">= 50x the reference CUDA extensions" denominator and an
end-to-end parity checker; bench.py reports it as ``ref_cuda_ext``.  The product never imports it.

What is restated (not imported - /root/reference does not exist on the GPU box):
  * gridencoder/grid.py:19-84   : per-call full-table fp32->fp16 cast, [L,B,C] output + permute copy,
                                  backward permute copy + zeros_like(half table) + fp16 atomics;
  * raymarching/raymarching.py:161-288 : zero-filled N*max_steps buffers, `.item()` sync,
                                  `torch.cuda.empty_cache()` every step, composite fwd/bwd wrappers;
  * nerf/network_grid.py:76-87,158-167 and activation.py : nn.Linear MLPs under autocast, trunc_exp, sigmoid;
  * nerf/renderer.py:446-559,562-613 : run_cuda (training branch) and update_extra_state;
  * nerf/utils.py:337-403,696-713 : the Trainer's step (two backward passes, GradScaler, Adam).
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function

from . import ref_ext


class _RefGridEncode(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, inputs, embeddings, offsets, S, Hres, gridtype, ns):
        inputs = inputs.contiguous()
        B, D = inputs.shape
        L = offsets.shape[0] - 1
        C = embeddings.shape[1]
        if torch.is_autocast_enabled("cuda") and C % 2 == 0:
            embeddings = embeddings.to(torch.half)          # whole-table cast, every forward (grid.py:38-39)
        outputs = torch.empty(L, B, C, device=inputs.device, dtype=embeddings.dtype)
        ns.grid.grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, Hres, None, gridtype, False)
        outputs = outputs.permute(1, 0, 2).reshape(B, L * C)  # a real copy (grid.py:52)
        ctx.save_for_backward(inputs, embeddings, offsets)
        ctx.meta = (B, D, C, L, S, Hres, gridtype, ns)
        return outputs

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        inputs, embeddings, offsets = ctx.saved_tensors
        B, D, C, L, S, Hres, gridtype, ns = ctx.meta
        grad = grad.view(B, L, C).permute(1, 0, 2).contiguous()
        grad_embeddings = torch.zeros_like(embeddings)
        ns.grid.grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, Hres, None, None,
                                     gridtype, False)
        return None, grad_embeddings, None, None, None, None, None


class _RefCompositeTrain(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh, ns):
        sigmas = sigmas.contiguous()
        rgbs = rgbs.contiguous()
        M, N = sigmas.shape[0], rays.shape[0]
        ws = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        depth = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        image = torch.empty(N, 3, dtype=sigmas.dtype, device=sigmas.device)
        ns.march.composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, T_thresh, ws, depth, image)
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, ws, image)
        ctx.meta = (M, N, T_thresh, ns)
        return ws, depth, image

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g_ws, g_depth, g_image):
        sigmas, rgbs, deltas, rays, ws, image = ctx.saved_tensors
        M, N, T_thresh, ns = ctx.meta
        gs = torch.zeros_like(sigmas)
        gc = torch.zeros_like(rgbs)
        ns.march.composite_rays_train_backward(g_ws.contiguous(), g_image.contiguous(), sigmas, rgbs, deltas, rays, ws, image,
                                               M, N, T_thresh, gs, gc)
        return gs, gc, None, None, None, None


class _TruncExp(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float)
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g):
        return g * torch.exp(ctx.saved_tensors[0].clamp(-15, 15))


class _RefFreq(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, inputs, degree, out_dim, ns):
        inputs = inputs.contiguous()
        B, D = inputs.shape
        out = torch.empty(B, out_dim, dtype=inputs.dtype, device=inputs.device)
        ns.freq.freq_encode_forward(inputs, B, D, degree, out_dim, out)
        ctx.save_for_backward(inputs, out)
        ctx.meta = (B, D, degree, out_dim, ns)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        inputs, out = ctx.saved_tensors
        B, D, degree, out_dim, ns = ctx.meta
        gi = torch.zeros_like(inputs)
        ns.freq.freq_encode_backward(grad.contiguous(), out, B, D, degree, out_dim, gi)
        return gi, None, None, None


def _mlp(dims):
    return nn.ModuleList([nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])


def _run_mlp(layers, x):
    for i, l in enumerate(layers):
        x = l(x)
        if i != len(layers) - 1:
            x = F.relu(x, inplace=True)
    return x


class RefGridNeRF(nn.Module):
    """network_grid.NeRFNetwork + NeRFRenderer (cuda_ray, bound 1) on the reference's kernels."""

    def __init__(self, ns, bound=1.0, density_thresh=10.0):
        super().__init__()
        self.ns, self.bound, self.density_thresh = ns, bound, density_thresh
        self.cascade = 1 + math.ceil(math.log2(bound))
        self.grid_size = 128
        L, C, base, log2, desired = 16, 2, 16, 16, 2048 * bound
        self.pls = np.exp2(np.log2(desired / base) / (L - 1))
        offs, off = [], 0
        for i in range(L):
            res = int(np.ceil(base * self.pls ** i))
            n = int(np.ceil(min(2 ** log2, (res + 1) ** 3) / 8) * 8)
            offs.append(off)
            off += n
        offs.append(off)
        self.register_buffer("offsets", torch.from_numpy(np.array(offs, np.int32)))
        self.embeddings = nn.Parameter(torch.empty(off, C).uniform_(-1e-4, 1e-4))
        self.sigma_net = _mlp([32, 64, 64, 4])
        self.bg_net = _mlp([39, 64, 3])
        self.register_buffer("aabb", torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32))
        self.register_buffer("density_grid", torch.zeros(self.cascade, 128 ** 3))
        self.register_buffer("density_bitfield", torch.zeros(self.cascade * 128 ** 3 // 8, dtype=torch.uint8))
        self.register_buffer("step_counter", torch.zeros(16, 2, dtype=torch.int32))
        self.mean_density, self.mean_count, self.local_step = 0, 0, 0

    def common_forward(self, x):
        x01 = (x + self.bound) / (2 * self.bound)
        h = _RefGridEncode.apply(x01.view(-1, 3), self.embeddings, self.offsets, float(np.log2(self.pls)), 16, 1, self.ns)
        h = _run_mlp(self.sigma_net, h)
        blob = 5 * torch.exp(-(x ** 2).sum(-1) / (2 * 0.2 ** 2))
        return _TruncExp.apply(h[..., 0] + blob), torch.sigmoid(h[..., 1:])

    def background(self, d):
        return torch.sigmoid(_run_mlp(self.bg_net, _RefFreq.apply(d.reshape(-1, 3), 6, 39, self.ns)))

    # -- shading (nerf/network_grid.py:90-144) ---------------------------------------------------------------------
    def finite_difference_normal(self, x, epsilon=1e-2):
        vals = []
        for axis in range(3):
            for sign in (1.0, -1.0):
                off = torch.zeros(1, 3, device=x.device)
                off[0, axis] = sign * epsilon
                vals.append(self.common_forward((x + off).clamp(-self.bound, self.bound))[0])
        dx_pos, dx_neg, dy_pos, dy_neg, dz_pos, dz_neg = vals
        normal = torch.stack([0.5 * (dx_pos - dx_neg) / epsilon, 0.5 * (dy_pos - dy_neg) / epsilon,
                              0.5 * (dz_pos - dz_neg) / epsilon], dim=-1)
        return -normal

    def normal(self, x):
        normal = self.finite_difference_normal(x)
        normal = normal / torch.sqrt(torch.clamp(torch.sum(normal * normal, -1, keepdim=True), min=1e-20))  # safe_normalize
        normal[torch.isnan(normal)] = 0
        return normal

    def forward(self, x, d, l=None, ratio=1, shading="albedo"):
        if shading == "albedo":
            sigma, color = self.common_forward(x)
            return sigma, color, None
        sigma, albedo = self.common_forward(x)
        normal = self.normal(x)
        lambertian = ratio + (1 - ratio) * (normal @ l).clamp(min=0)
        if shading == "textureless":
            color = lambertian.unsqueeze(-1).repeat(1, 3)
        elif shading == "normal":
            color = (normal + 1) / 2
        else:
            color = albedo * lambertian.unsqueeze(-1)
        return sigma, color, normal

    def march_train(self, rays_o, rays_d, nears, fars, counter, max_steps, perturb=True):
        ns = self.ns
        N = rays_o.shape[0]
        M = N * max_steps
        dev = rays_o.device
        xyzs = torch.zeros(M, 3, device=dev)            # 134 MB of memsets per 4096 rays (raymarching.py:205-207)
        dirs = torch.zeros(M, 3, device=dev)
        deltas = torch.zeros(M, 2, device=dev)
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        noises = torch.rand(N, device=dev) if perturb else torch.zeros(N, device=dev)
        ns.march.march_rays_train(rays_o, rays_d, self.density_bitfield, self.bound, 0.0, max_steps, N, self.cascade, 128, M,
                                  nears, fars, xyzs, dirs, deltas, rays, counter, noises)
        m = counter[0].item()                            # host sync (raymarching.py:224)
        m += 128 - m % 128
        torch.cuda.empty_cache()                         # allocator flush every step (raymarching.py:231)
        return xyzs[:m], dirs[:m], deltas[:m], rays

    def render_train(self, rays_o, rays_d, max_steps=1024, T_thresh=1e-4, shading="albedo", ambient_ratio=1.0):
        """run_cuda, training branch (nerf/renderer.py:446-494,535-557)."""
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        nears, fars = ref_ext.near_far_from_aabb(self.ns, rays_o, rays_d, self.aabb, 0.2)
        light_d = rays_o[0] + torch.randn(3, device=rays_o.device)   # drawn even for albedo shading (:461-464)
        light_d = light_d / torch.sqrt(torch.clamp(torch.sum(light_d * light_d, -1, keepdim=True), min=1e-20))
        counter = self.step_counter[self.local_step % 16]
        counter.zero_()
        self.local_step += 1
        xyzs, dirs, deltas, rays = self.march_train(rays_o, rays_d, nears, fars, counter, max_steps)
        sigmas, rgbs, normals = self(xyzs, dirs, light_d, ratio=ambient_ratio, shading=shading)
        ws, depth, image = _RefCompositeTrain.apply(sigmas, rgbs, deltas, rays, T_thresh, self.ns)
        results = {}
        if normals is not None:   # orientation + smoothness regularisers (:485-494)
            weights = 1 - torch.exp(-sigmas)
            results["loss_orient"] = (weights.detach() * (normals * dirs).sum(-1).clamp(min=0) ** 2).mean()
            normals_perturb = self.normal(xyzs + torch.randn_like(xyzs) * 1e-2)
            results["loss_smooth"] = (normals - normals_perturb).abs().mean()
        bg = self.background(rays_d)
        image = image + (1 - ws).unsqueeze(-1) * bg
        depth = torch.clamp(depth - nears, min=0) / (fars - nears)
        results.update({"image": image.view(*prefix, 3), "depth": depth.view(*prefix), "weights_sum": ws.reshape(*prefix)})
        return results

    @torch.no_grad()
    def render_eval(self, rays_o, rays_d, max_steps=1024, T_thresh=1e-4, perturb=False, bg_color=None, use_bg_net=True):
        """run_cuda, inference branch (nerf/renderer.py:496-557) on the reference's march_rays / composite_rays, with the
        reference's boolean-mask compaction and one host sync per iteration."""
        ns = self.ns
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3).float()
        rays_d = rays_d.contiguous().view(-1, 3).float()
        N, dev = rays_o.shape[0], rays_o.device
        nears, fars = ref_ext.near_far_from_aabb(ns, rays_o, rays_d, self.aabb, 0.2)
        _ = torch.randn(3, device=dev)
        weights_sum = torch.zeros(N, device=dev)
        depth = torch.zeros(N, device=dev)
        image = torch.zeros(N, 3, device=dev)
        rays_alive = torch.arange(N, dtype=torch.int32, device=dev)
        rays_t = nears.clone()
        step = 0
        iters = 0
        while step < max_steps:
            n_alive = rays_alive.shape[0]
            if n_alive <= 0:
                break
            n_step = max(min(N // n_alive, 8), 1)
            M = n_alive * n_step
            M += 128 - (M % 128)
            xyzs = torch.zeros(M, 3, device=dev)
            dirs = torch.zeros(M, 3, device=dev)
            deltas = torch.zeros(M, 2, device=dev)
            noises = torch.rand(n_alive, device=dev) if (perturb and step == 0) else torch.zeros(n_alive, device=dev)
            ns.march.march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, self.bound, 0.0, max_steps, self.cascade, 128,
                                self.density_bitfield, nears, fars, xyzs, dirs, deltas, noises)
            sigmas, rgbs = self.common_forward(xyzs)
            ns.march.composite_rays(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas.float().contiguous(),
                                    rgbs.float().contiguous(), deltas, weights_sum, depth, image)
            rays_alive = rays_alive[rays_alive >= 0]
            step += n_step
            iters += 1
        bg = self.background(rays_d) if use_bg_net else (1 if bg_color is None else bg_color)
        image = image + (1 - weights_sum).unsqueeze(-1) * bg
        depth = torch.clamp(depth - nears, min=0) / (fars - nears)
        return {"image": image.view(*prefix, 3), "depth": depth.view(*prefix), "weights_sum": weights_sum.reshape(*prefix),
                "mask": (nears < fars).reshape(*prefix), "iterations": iters}

    @torch.no_grad()
    def update_extra_state(self, decay=0.95, noise=None):
        ns = self.ns
        Hh = self.grid_size
        dev = self.density_bitfield.device
        tmp_grid = -torch.ones_like(self.density_grid)
        ar = torch.arange(Hh, dtype=torch.int32, device=dev)
        xx, yy, zz = torch.meshgrid(ar, ar, ar, indexing="ij")
        coords = torch.cat([xx.reshape(-1, 1), yy.reshape(-1, 1), zz.reshape(-1, 1)], dim=-1)
        indices = torch.empty(coords.shape[0], dtype=torch.int32, device=dev)
        ns.march.morton3D(coords.int().contiguous(), coords.shape[0], indices)
        indices = indices.long()
        xyzs = 2 * coords.float() / (Hh - 1) - 1
        for cas in range(self.cascade):
            bound = min(2 ** cas, self.bound)
            hgs = bound / Hh
            cas_xyzs = xyzs * (bound - hgs)
            u = torch.rand_like(cas_xyzs) if noise is None else noise[cas]
            cas_xyzs += (u * 2 - 1) * hgs
            sig, _ = self.common_forward(cas_xyzs)
            tmp_grid[cas, indices] = sig.reshape(-1).detach().float()
        valid = self.density_grid >= 0
        self.density_grid[valid] = torch.maximum(self.density_grid[valid] * decay, tmp_grid[valid])
        self.mean_density = torch.mean(self.density_grid[valid]).item()
        thresh = min(self.mean_density, self.density_thresh)
        ns.march.packbits(self.density_grid, self.density_grid.numel() // 8, thresh, self.density_bitfield)
        total = min(16, self.local_step)
        if total > 0:
            self.mean_count = int(self.step_counter[:total, 0].sum().item() / total)
        self.local_step = 0

    def param_groups(self, lr):
        return [{"params": [self.embeddings], "lr": lr * 10}, {"params": self.sigma_net.parameters(), "lr": lr},
                {"params": self.bg_net.parameters(), "lr": lr}]


def time_reference_train_step(device, views=1, steps=200, warmup=50, Hh=64, Ww=64, max_steps=1024, lr=1e-5):
    """Times the reference's -O train step (reference kernels + reference call pattern) on `device`.
    The reference renders one view per step (provider.py:240 batch_size=1).  lr: as in bench.py a SMALL step, so that the
    synthetic random-gradient "training" leaves the scene at its random-init occupancy and every timed step marches the
    same ~0.4 M samples per view on both sides of the comparison."""
    import sys, os
    ns = ref_ext.load()
    if ns is None:
        return {"unavailable": "oracle/_ref not built"}
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "single-stable-dreamfusion_b200"))
    from ngp_b200 import provider
    torch.manual_seed(0)
    model = RefGridNeRF(ns).to(device)
    model.train()
    opt = torch.optim.Adam(model.param_groups(lr), betas=(0.9, 0.99), eps=1e-15)
    scaler = torch.amp.GradScaler("cuda")
    ro_all, rd_all = provider.make_training_views(32 * views, Hh, Ww, seed=0)
    ro_all = ro_all.view(32, views, Hh * Ww, 3).to(device)
    rd_all = rd_all.view(32, views, Hh * Ww, 3).to(device)
    G = (torch.randn(views, 3, Hh, Ww, generator=torch.Generator().manual_seed(2)) * 1e-2).to(device)
    samples = 0
    gstep = 0

    def one(i):
        nonlocal samples, gstep
        if gstep % 16 == 0:
            with torch.autocast("cuda", torch.float16):
                model.update_extra_state()
        gstep += 1
        opt.zero_grad()
        with torch.autocast("cuda", torch.float16):
            out = model.render_train(ro_all[i % 32], rd_all[i % 32], max_steps)
            pred = out["image"].reshape(views, Hh, Ww, 3).permute(0, 3, 1, 2).contiguous()
            pred.backward(gradient=G, retain_graph=True)
            a = out["weights_sum"].reshape(views, 1, Hh, Ww).clamp(1e-5, 1 - 1e-5)
            loss = 1e-4 * (-a * torch.log2(a) - (1 - a) * torch.log2(1 - a)).mean()
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        samples += int(model.step_counter[(model.local_step - 1) % 16, 0].item())
        return loss.item()                               # nerf/utils.py:715

    for i in range(warmup):
        one(i)
    torch.cuda.synchronize()
    samples = 0
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record()
    for i in range(steps):
        one(100 + i)
        evs[i + 1].record()
    torch.cuda.synchronize()
    ms = evs[0].elapsed_time(evs[-1])
    per = sorted(a.elapsed_time(b) for a, b in zip(evs, evs[1:]))
    med = per[len(per) // 2]
    return {"value": samples / (ms * 1e-3), "value_median": (samples / steps) / (med * 1e-3), "unit": "samples/s",
            "ms_per_step": ms / steps, "ms_per_step_median": med, "ms_per_step_p10": per[len(per) // 10],
            "ms_per_step_p90": per[(len(per) * 9) // 10], "views_per_step": views,
            "samples_per_step": samples / steps, "steps": steps, "warmup": warmup, "lr": lr,
            "what": "reference CUDA extensions (gridencoder/raymarching/freqencoder rebuilt unmodified for sm_100) "
                    "+ the reference's host call pattern, same -O train step, 1 GPU"}


def time_reference_inference(device, frames=20, res=800, max_steps=1024):
    """Times the reference's inference loop (reference kernels + its host loop with boolean-mask compaction and a
    per-iteration sync, nerf/renderer.py:496-557) on the test orbit's cameras, plus the per-frame device->host read and uint8
    conversion of Trainer.test (nerf/utils.py:526-537)."""
    import sys, os
    ns = ref_ext.load()
    if ns is None:
        return {"unavailable": "oracle/_ref not built"}
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "single-stable-dreamfusion_b200"))
    from ngp_b200 import provider
    torch.manual_seed(0)
    model = RefGridNeRF(ns).to(device)
    with torch.autocast("cuda", torch.float16):
        model.update_extra_state()
    model.eval()
    views = provider.make_orbit_views(frames, res, res)
    rays = [(torch.from_numpy(o).to(device)[None], torch.from_numpy(d).to(device)[None]) for o, d in views]

    def one(i):
        with torch.no_grad(), torch.autocast("cuda", torch.float16):
            out = model.render_eval(rays[i][0], rays[i][1], max_steps=max_steps, T_thresh=1e-4, perturb=False)
        pred = (out["image"].reshape(res, res, 3).detach().cpu().numpy() * 255).astype(np.uint8)
        with np.errstate(invalid="ignore"):
            (out["depth"].reshape(res, res).detach().cpu().numpy() * 255).astype(np.uint8)
        return out["iterations"], pred

    one(0); one(1)
    torch.cuda.synchronize()
    import time
    per, its = [], []
    for i in range(frames):
        t0 = time.perf_counter()
        it, _ = one(i)
        torch.cuda.synchronize()
        per.append((time.perf_counter() - t0) * 1e3)
        its.append(it)
    per_s = sorted(per)
    return {"ms_per_frame": sum(per) / len(per), "ms_per_frame_median": per_s[len(per_s) // 2], "frames": frames,
            "frames_per_s_median": 1e3 / per_s[len(per_s) // 2], "loop_iterations_per_frame": its[:: max(1, frames // 5)],
            "what": "reference CUDA extensions + the reference's host inference loop and per-frame host conversion, %dx%d" % (res, res)}


def time_reference_encoder(device, points=1 << 22, iters=8):
    """The reference's GridEncoder path (its extension + its wrapper's casts / permutes / zero-fills, grid.py:19-84) on BASELINE
    configs[1]: hash 2^19, 16 x 2, base 16 -> 2048, fp16 autocast, forward and forward + backward."""
    ns = ref_ext.load()
    if ns is None:
        return {"unavailable": "oracle/_ref not built"}
    L, C, base, log2, desired = 16, 2, 16, 19, 2048
    pls = np.exp2(np.log2(desired / base) / (L - 1))
    offs, off = [], 0
    for i in range(L):
        res = int(np.ceil(base * pls ** i))
        offs.append(off)
        off += int(np.ceil(min(2 ** log2, (res + 1) ** 3) / 8) * 8)
    offs.append(off)
    offsets = torch.from_numpy(np.array(offs, np.int32)).to(device)
    torch.manual_seed(0)
    emb = nn.Parameter(torch.empty(off, C, device=device).uniform_(-1, 1))
    xs = [torch.rand(points, 3, device=device, generator=torch.Generator(device=device).manual_seed(1 + k)) for k in range(4)]
    S = float(np.log2(pls))
    with torch.autocast("cuda", torch.float16):
        out = _RefGridEncode.apply(xs[0], emb, offsets, S, base, 0, ns)
    gs = [torch.randn(out.shape, device=device, dtype=out.dtype, generator=torch.Generator(device=device).manual_seed(20 + k)) for k in range(4)]

    def fwd(i):
        with torch.no_grad(), torch.autocast("cuda", torch.float16):
            _RefGridEncode.apply(xs[i % 4], emb, offsets, S, base, 0, ns)

    def fwd_bwd(i):
        emb.grad = None
        with torch.autocast("cuda", torch.float16):
            o = _RefGridEncode.apply(xs[i % 4], emb, offsets, S, base, 0, ns)
        o.backward(gs[i % 4])

    def timed(fn):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(3 + i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    t_f, t_fb = timed(fwd), timed(fwd_bwd)
    return {"fwd_ms": t_f, "fwd_bwd_ms": t_fb, "points": points, "fwd_bwd_points_per_s": points / (t_fb * 1e-3),
            "what": "reference gridencoder extension + its Python wrapper (table cast, [L,B,C] permute copies, fp16 atomics)"}


if __name__ == "__main__":
    # run in a fresh process (own CUDA context / caching allocator): the reference empties the allocator cache every
    # step (raymarching.py:231), which must not be charged for another workload's cached blocks
    import json
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == "encoder":
        print("REF_PIPELINE_JSON " + json.dumps(time_reference_encoder(torch.device("cuda", 0))), flush=True)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "infer":
        res = time_reference_inference(torch.device("cuda", 0), frames=int(sys.argv[2]) if len(sys.argv) > 2 else 20,
                                       res=int(sys.argv[3]) if len(sys.argv) > 3 else 800)
        print("REF_PIPELINE_JSON " + json.dumps(res), flush=True)
        sys.exit(0)
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    warm = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    lr = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-5
    res = time_reference_train_step(torch.device("cuda", 0), views=1, steps=steps, warmup=warm, lr=lr)
    print("REF_PIPELINE_JSON " + json.dumps(res), flush=True)
