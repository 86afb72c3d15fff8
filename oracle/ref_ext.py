"""TEST INFRASTRUCTURE: loads the reference's own CUDA extensions built by ``build_ref.py``.

The modules are the reference's unmodified pybind targets (``_gridencoder``, ``_raymarching``,
``_freqencoder``); they only run on a GPU box.  The helpers below restate the allocation /
zero-fill contract of the reference's Python wrappers (gridencoder/grid.py:22-84,
raymarching/raymarching.py:19-373, freqencoder/freq.py:15-52) around the raw native calls so
tests and the bench can drive "the reference" without importing /root/reference at run time.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_mods = None


def load():
    """Returns a namespace with .grid, .march, .freq modules, or None if unavailable."""
    global _mods
    if _mods is not None:
        return _mods or None
    names = {"grid": "_gridencoder", "march": "_raymarching", "freq": "_freqencoder"}
    if not all(os.path.exists(os.path.join(REF_DIR, n + ".so")) for n in names.values()):
        _mods = False
        return None
    import torch  # noqa: F401  (the extensions link against libtorch)
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    ns = types.SimpleNamespace()
    try:
        for k, n in names.items():
            setattr(ns, k, importlib.import_module(n))
    except Exception as e:  # pragma: no cover
        print("[ref_ext] could not import reference extensions:", e)
        _mods = False
        return None
    _mods = ns
    return ns


# ---- the reference wrappers' calling conventions, restated ------------------------------------------
def grid_encode_forward(ns, inputs, embeddings, offsets, S, H, calc_grad_inputs=False, gridtype=0, align_corners=False):
    """gridencoder/grid.py:22-58 without autograd.  Returns (outputs [B, L*C], dy_dx|None, outputs_LBC)."""
    import torch
    inputs = inputs.contiguous()
    B, D = inputs.shape
    L = offsets.shape[0] - 1
    C = embeddings.shape[1]
    outputs = torch.empty(L, B, C, device=inputs.device, dtype=embeddings.dtype)
    dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=embeddings.dtype) if calc_grad_inputs else None
    ns.grid.grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype, align_corners)
    return outputs.permute(1, 0, 2).reshape(B, L * C), dy_dx, outputs


def grid_encode_backward(ns, grad, inputs, embeddings, offsets, S, H, dy_dx=None, gridtype=0, align_corners=False):
    """gridencoder/grid.py:60-84.  grad [B, L*C]; returns (grad_embeddings in embeddings.dtype, grad_inputs|None)."""
    import torch
    B, D = inputs.shape
    L = offsets.shape[0] - 1
    C = embeddings.shape[1]
    grad = grad.view(B, L, C).permute(1, 0, 2).contiguous()
    grad_embeddings = torch.zeros_like(embeddings)
    grad_inputs = torch.zeros_like(inputs, dtype=embeddings.dtype) if dy_dx is not None else None
    ns.grid.grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, dy_dx, grad_inputs,
                                 gridtype, align_corners)
    return grad_embeddings, grad_inputs


def near_far_from_aabb(ns, rays_o, rays_d, aabb, min_near=0.2):
    import torch
    N = rays_o.shape[0]
    nears = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
    fars = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
    ns.march.near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars)
    return nears, fars


def march_rays_train(ns, rays_o, rays_d, bound, bitfield, C, H, nears, fars, noises, dt_gamma=0.0, max_steps=1024, M=None):
    """raymarching.py:164-235 with injected noises; returns full-capacity buffers + counter (no slicing)."""
    import torch
    N = rays_o.shape[0]
    if M is None:
        M = N * max_steps
    dev = rays_o.device
    xyzs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
    dirs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
    deltas = torch.zeros(M, 2, dtype=torch.float32, device=dev)
    rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    ns.march.march_rays_train(rays_o, rays_d, bitfield, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs,
                              deltas, rays, counter, noises)
    return xyzs, dirs, deltas, rays, counter


def composite_rays_train_forward(ns, sigmas, rgbs, deltas, rays, T_thresh=1e-4):
    import torch
    M, N = sigmas.shape[0], rays.shape[0]
    ws = torch.empty(N, dtype=torch.float32, device=sigmas.device)
    depth = torch.empty(N, dtype=torch.float32, device=sigmas.device)
    image = torch.empty(N, 3, dtype=torch.float32, device=sigmas.device)
    ns.march.composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, T_thresh, ws, depth, image)
    return ws, depth, image


def composite_rays_train_backward(ns, grad_ws, grad_image, sigmas, rgbs, deltas, rays, ws, image, T_thresh=1e-4):
    import torch
    M, N = sigmas.shape[0], rays.shape[0]
    gs = torch.zeros_like(sigmas)
    gc = torch.zeros_like(rgbs)
    ns.march.composite_rays_train_backward(grad_ws, grad_image, sigmas, rgbs, deltas, rays, ws, image, M, N, T_thresh, gs, gc)
    return gs, gc


def march_rays(ns, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, C, H, nears, fars, noises,
               dt_gamma=0.0, max_steps=1024, align=-1):
    import torch
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)
    dev = rays_o.device
    xyzs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
    dirs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
    deltas = torch.zeros(M, 2, dtype=torch.float32, device=dev)
    ns.march.march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, bitfield,
                        nears, fars, xyzs, dirs, deltas, noises)
    return xyzs, dirs, deltas


def canonical_rays(rays, *per_sample):
    """Sort the reference's nondeterministically ordered `rays` rows by ray id and gather each ray's samples into
    ray order.  Returns (counts[N], [gathered per-sample tensors...])."""
    import torch
    order = torch.argsort(rays[:, 0].long())
    r = rays[order]
    counts = r[:, 2].long()
    offs = r[:, 1].long()
    total = int(counts.sum().item())
    idx = torch.repeat_interleave(offs, counts) + (torch.arange(total, device=rays.device) -
                                                   torch.repeat_interleave(torch.cumsum(counts, 0) - counts, counts))
    return counts, [t[idx] for t in per_sample]
