"""TEST INFRASTRUCTURE / CPU BASELINE: a pure-PyTorch port of the reference's non-``--cuda_ray``
renderer and its "vanilla" field network, runnable on host cores.

It restates, with the same arithmetic and the same RNG consumption order,
  * ``NeRFRenderer.run``       (nerf/renderer.py:301-443) incl. ``sample_pdf`` (:15-49),
  * ``nerf/network.py``        ``ResBlock`` (:13-41), ``MLP`` (:44-67), ``NeRFNetwork`` (:70-221, albedo path),
  * the two CUDA-only ops that path touches, as plain torch: ``near_far_from_aabb``
    (raymarching.cu:92-145) and ``FreqEncoder`` (freqencoder.cu:30-58),
and is pinned against the real reference imported from /root/reference by
``oracle/make_golden_cpu.py`` (fixture: tests/golden/cpu_renderer_golden.npz).

BASELINE.json configs[0]: 64x64 rays x (64+32) samples, random init, fp32 fwd+bwd, no SD guidance.
bench.py times this on the GPU box's host cores as ``cpu_baseline`` (kind "port") and as
``--impl reference``.  The product never imports it.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

FLT_MAX = 3.4028234663852886e38


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    """Slab test; both outputs FLT_MAX on a miss; near clamped to min_near (raymarching.cu:100-145)."""
    rd = 1.0 / rays_d
    lo = (aabb[:3] - rays_o) * rd
    hi = (aabb[3:] - rays_o) * rd
    tmin = torch.minimum(lo, hi)
    tmax = torch.maximum(lo, hi)
    # sequential axis merge with the reference's miss tests
    near, far = tmin[:, 0], tmax[:, 0]
    miss = (near > tmax[:, 1]) | (tmin[:, 1] > far)
    near = torch.maximum(near, tmin[:, 1])
    far = torch.minimum(far, tmax[:, 1])
    miss = miss | (near > tmax[:, 2]) | (tmin[:, 2] > far)
    near = torch.maximum(near, tmin[:, 2])
    far = torch.minimum(far, tmax[:, 2])
    near = torch.clamp(near, min=min_near)
    big = torch.full_like(near, FLT_MAX)
    return torch.where(miss, big, near), torch.where(miss, big, far)


class FreqEncoder(nn.Module):
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(deg-1) x), cos(2^(deg-1) x)] (freqencoder.cu:46-56)."""

    def __init__(self, input_dim=3, degree=6):
        super().__init__()
        self.input_dim, self.degree = input_dim, degree
        self.output_dim = input_dim + input_dim * 2 * degree

    def forward(self, x, **kwargs):
        parts = [x]
        for f in range(self.degree):
            parts.append(torch.sin(x * 2.0 ** f))
            parts.append(torch.cos(x * 2.0 ** f))
        return torch.cat(parts, dim=-1)


class ResBlock(nn.Module):
    def __init__(self, dim_in, dim_out, bias=True):
        super().__init__()
        self.dense = nn.Linear(dim_in, dim_out, bias=bias)
        self.norm = nn.LayerNorm(dim_out)
        self.activation = nn.SiLU()
        self.skip = nn.Linear(dim_in, dim_out, bias=False) if dim_in != dim_out else None

    def forward(self, x):
        out = self.norm(self.dense(x))
        out = out + (self.skip(x) if self.skip is not None else x)
        return self.activation(out)


class VanillaMLP(nn.Module):
    def __init__(self, dim_in, dim_out, dim_hidden, num_layers, bias=True):
        super().__init__()
        layers = []
        for l in range(num_layers):
            if l != num_layers - 1:
                layers.append(ResBlock(dim_in if l == 0 else dim_hidden, dim_hidden, bias=bias))
            else:
                layers.append(nn.Linear(dim_hidden, dim_out, bias=bias))
        self.net = nn.ModuleList(layers)

    def forward(self, x):
        for layer in self.net:
            x = layer(x)
        return x


def sample_pdf(bins, weights, n_samples, det=False):
    """Inverse-CDF resampling of z values (nerf/renderer.py:15-49)."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    if det:
        u = torch.linspace(0. + 0.5 / n_samples, 1. - 0.5 / n_samples, steps=n_samples).to(weights.device)
        u = u.expand(list(cdf.shape[:-1]) + [n_samples])
    else:
        u = torch.rand(list(cdf.shape[:-1]) + [n_samples]).to(weights.device)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_lo, cdf_hi = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bin_lo, bin_hi = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = cdf_hi - cdf_lo
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_lo) / denom
    return bin_lo + t * (bin_hi - bin_lo)


class VanillaNeRF(nn.Module):
    """nerf/network.py:NeRFNetwork + NeRFRenderer.run, albedo shading, training mode."""

    def __init__(self, bound=1.0, min_near=0.1, bg_radius=1.4, num_layers=5, hidden_dim=128, num_layers_bg=2,
                 hidden_dim_bg=64):
        super().__init__()
        self.bound, self.min_near, self.bg_radius = bound, min_near, bg_radius
        self.register_buffer("aabb_train", torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32))
        self.encoder = FreqEncoder(3, 6)
        self.sigma_net = VanillaMLP(self.encoder.output_dim, 4, hidden_dim, num_layers)
        self.encoder_bg = FreqEncoder(3, 6)
        self.bg_net = VanillaMLP(self.encoder_bg.output_dim, 3, hidden_dim_bg, num_layers_bg)

    def common_forward(self, x):
        h = self.sigma_net(self.encoder(x))
        blob = 5 * torch.exp(-(x ** 2).sum(-1) / (2 * 0.2 ** 2))
        sigma = torch.exp(h[..., 0] + blob)  # trunc_exp forward; its clamped backward matters only past e^15
        albedo = torch.sigmoid(h[..., 1:])
        return sigma, albedo

    def background(self, d):
        return torch.sigmoid(self.bg_net(self.encoder_bg(d)))

    def run(self, rays_o, rays_d, num_steps=64, upsample_steps=32, perturb=True):
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        N = rays_o.shape[0]
        aabb = self.aabb_train
        nears, fars = near_far_from_aabb(rays_o, rays_d, aabb, self.min_near)
        nears = nears.unsqueeze(-1)
        fars = fars.unsqueeze(-1)
        # light direction is sampled even on the albedo path (renderer.py:324-327): keeps the RNG stream aligned
        _ = rays_o[0] + torch.randn(3, dtype=torch.float)

        z_vals = torch.linspace(0.0, 1.0, num_steps).unsqueeze(0).expand((N, num_steps))
        z_vals = nears + (fars - nears) * z_vals
        sample_dist = (fars - nears) / num_steps
        if perturb:
            z_vals = z_vals + (torch.rand(z_vals.shape) - 0.5) * sample_dist
        xyzs = rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * z_vals.unsqueeze(-1)
        xyzs = torch.min(torch.max(xyzs, aabb[:3]), aabb[3:])

        sigma_c, albedo_c = self.common_forward(xyzs.reshape(-1, 3))
        dens = {"sigma": sigma_c.view(N, num_steps, -1), "albedo": albedo_c.view(N, num_steps, -1)}

        if upsample_steps > 0:
            with torch.no_grad():
                deltas = z_vals[..., 1:] - z_vals[..., :-1]
                deltas = torch.cat([deltas, sample_dist * torch.ones_like(deltas[..., :1])], dim=-1)
                alphas = 1 - torch.exp(-deltas * dens["sigma"].squeeze(-1))
                alphas_shifted = torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas + 1e-15], dim=-1)
                weights = alphas * torch.cumprod(alphas_shifted, dim=-1)[..., :-1]
                z_mid = z_vals[..., :-1] + 0.5 * deltas[..., :-1]
                new_z = sample_pdf(z_mid, weights[:, 1:-1], upsample_steps, det=not self.training).detach()
                new_xyzs = rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * new_z.unsqueeze(-1)
                new_xyzs = torch.min(torch.max(new_xyzs, aabb[:3]), aabb[3:])
            sigma_f, albedo_f = self.common_forward(new_xyzs.reshape(-1, 3))
            new_dens = {"sigma": sigma_f.view(N, upsample_steps, -1), "albedo": albedo_f.view(N, upsample_steps, -1)}
            z_vals = torch.cat([z_vals, new_z], dim=1)
            z_vals, z_index = torch.sort(z_vals, dim=1)
            xyzs = torch.cat([xyzs, new_xyzs], dim=1)
            xyzs = torch.gather(xyzs, dim=1, index=z_index.unsqueeze(-1).expand_as(xyzs))
            for k in dens:
                tmp = torch.cat([dens[k], new_dens[k]], dim=1)
                dens[k] = torch.gather(tmp, dim=1, index=z_index.unsqueeze(-1).expand_as(tmp))

        deltas = z_vals[..., 1:] - z_vals[..., :-1]
        deltas = torch.cat([deltas, sample_dist * torch.ones_like(deltas[..., :1])], dim=-1)
        alphas = 1 - torch.exp(-deltas * dens["sigma"].squeeze(-1))
        alphas_shifted = torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas + 1e-15], dim=-1)
        weights = alphas * torch.cumprod(alphas_shifted, dim=-1)[..., :-1]

        # the reference evaluates the full field a second time on the merged samples (renderer.py:399)
        _, rgbs = self.common_forward(xyzs.reshape(-1, 3))
        rgbs = rgbs.view(N, -1, 3)

        weights_sum = weights.sum(dim=-1)
        ori_z = ((z_vals - nears) / (fars - nears)).clamp(0, 1)
        depth = torch.sum(weights * ori_z, dim=-1)
        image = torch.sum(weights.unsqueeze(-1) * rgbs, dim=-2)
        bg = self.background(rays_d.reshape(-1, 3))
        image = image + (1 - weights_sum).unsqueeze(-1) * bg
        return {"image": image.view(*prefix, 3), "depth": depth.view(*prefix), "weights_sum": weights_sum,
                "mask": (nears < fars).reshape(*prefix)}


def make_view(H=64, W=64, seed=0):
    """One random training view in the style of provider.rand_poses + get_rays (deterministic, numpy-free)."""
    g = torch.Generator().manual_seed(seed)
    radius = 1.0 + 0.5 * torch.rand(1, generator=g).item()
    theta = math.radians(100.0 * torch.rand(1, generator=g).item())
    phi = math.radians(360.0 * torch.rand(1, generator=g).item())
    fov = 40.0 + 30.0 * torch.rand(1, generator=g).item()
    centre = torch.tensor([radius * math.sin(theta) * math.sin(phi), radius * math.cos(theta),
                           radius * math.sin(theta) * math.cos(phi)])
    fwd = F.normalize(-centre, dim=0)
    up0 = torch.tensor([0.0, -1.0, 0.0])
    right = F.normalize(torch.linalg.cross(fwd, up0), dim=0)
    up = F.normalize(torch.linalg.cross(right, fwd), dim=0)
    R = torch.stack([right, up, fwd], dim=-1)
    focal = H / (2 * math.tan(math.radians(fov) / 2))
    j, i = torch.meshgrid(torch.arange(H, dtype=torch.float32) + 0.5, torch.arange(W, dtype=torch.float32) + 0.5,
                          indexing="ij")
    d = torch.stack([(i - W / 2) / focal, (j - H / 2) / focal, torch.ones_like(i)], -1).reshape(-1, 3)
    d = F.normalize(d, dim=-1)
    rays_d = d @ R.T
    rays_o = centre.expand_as(rays_d)
    return rays_o[None].contiguous(), rays_d[None].contiguous()


def train_step(model, rays_o, rays_d, grad_image, num_steps=64, upsample_steps=32):
    """fwd + bwd of one view; returns (#composited samples, results)."""
    for p in model.parameters():
        p.grad = None
    out = model.run(rays_o, rays_d, num_steps=num_steps, upsample_steps=upsample_steps, perturb=True)
    (out["image"] * grad_image).sum().backward()
    n_rays = rays_o.shape[0] * rays_o.shape[1]
    return n_rays * (num_steps + upsample_steps), out
