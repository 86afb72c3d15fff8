"""Drop-in replacement for the reference's ``raymarching/raymarching.py`` on top of libngp_b200.so.

Function names, argument order, defaults, returned shapes / dtypes and autocast behaviour follow
raymarching.py:19-373.  Differences that a caller can observe are limited to what the reference
itself leaves unspecified: ``march_rays_train`` returns its ``rays`` rows in ray order with
prefix-sum offsets (the reference's order depends on atomic timing, raymarching.cu:405-406).
"""
import torch
from torch.autograd import Function

from ngp_b200 import _cabi

_fwd32 = torch.amp.custom_fwd(device_type='cuda', cast_inputs=torch.float32)
_bwd = torch.amp.custom_bwd(device_type='cuda')


def _workspace(nbytes, device):
    ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
    ws[:256].zero_()   # ngp_march_rays_train: the head of a new workspace (block-election word) must be zero
    return ws


# ----------------------------------------
# utils
# ----------------------------------------

class _near_far_from_aabb(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, aabb, min_near=0.2):
        ''' rays_o/rays_d: float [N, 3]; aabb: float [6] (xmin, ymin, zmin, xmax, ymax, zmax)
        Returns nears, fars: float [N] (both FLT_MAX for a ray that misses the box). '''
        if not rays_o.is_cuda: rays_o = rays_o.cuda()
        if not rays_d.is_cuda: rays_d = rays_d.cuda()
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        aabb = aabb.to(rays_o.device, torch.float32).contiguous()
        N = rays_o.shape[0]
        nears = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        fars = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        _cabi.call("ngp_near_far_from_aabb", rays_o.device, _cabi.ptr(rays_o), _cabi.ptr(rays_d), _cabi.ptr(aabb), N,
                   float(min_near), _cabi.ptr(nears), _cabi.ptr(fars))
        return nears, fars

near_far_from_aabb = _near_far_from_aabb.apply


class _sph_from_ray(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, radius):
        ''' Spherical coordinates (theta, phi in [-1, 1]) where each ray leaves Sphere(radius). -> [N, 2] '''
        if not rays_o.is_cuda: rays_o = rays_o.cuda()
        if not rays_d.is_cuda: rays_d = rays_d.cuda()
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        N = rays_o.shape[0]
        coords = torch.empty(N, 2, dtype=rays_o.dtype, device=rays_o.device)
        _cabi.call("ngp_sph_from_ray", rays_o.device, _cabi.ptr(rays_o), _cabi.ptr(rays_d), float(radius), N,
                   _cabi.ptr(coords))
        return coords

sph_from_ray = _sph_from_ray.apply


class _morton3D(Function):
    @staticmethod
    def forward(ctx, coords):
        ''' coords: int32 [N, 3] in [0, 1024) -> Morton indices int32 [N] '''
        if not coords.is_cuda: coords = coords.cuda()
        N = coords.shape[0]
        indices = torch.empty(N, dtype=torch.int32, device=coords.device)
        coords = coords.int().contiguous()
        _cabi.call("ngp_morton3D", coords.device, _cabi.ptr(coords), N, _cabi.ptr(indices))
        return indices

morton3D = _morton3D.apply


class _morton3D_invert(Function):
    @staticmethod
    def forward(ctx, indices):
        ''' indices: int32 [N] -> coords int32 [N, 3] '''
        if not indices.is_cuda: indices = indices.cuda()
        N = indices.shape[0]
        coords = torch.empty(N, 3, dtype=torch.int32, device=indices.device)
        indices = indices.int().contiguous()
        _cabi.call("ngp_morton3D_invert", indices.device, _cabi.ptr(indices), N, _cabi.ptr(coords))
        return coords

morton3D_invert = _morton3D_invert.apply


class _packbits(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, grid, thresh, bitfield=None):
        ''' grid: float [C, H*H*H]; bit i of byte n = grid.flat[8n + i] > thresh -> uint8 [C*H*H*H/8] '''
        if not grid.is_cuda: grid = grid.cuda()
        grid = grid.contiguous()
        C = grid.shape[0]
        H3 = grid.shape[1]
        N = C * H3 // 8
        if bitfield is None:
            bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
        _cabi.call("ngp_packbits", grid.device, _cabi.ptr(grid), N, float(thresh), _cabi.ptr(bitfield))
        return bitfield

packbits = _packbits.apply

# ----------------------------------------
# train functions
# ----------------------------------------

class _march_rays_train(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1,
                perturb=False, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024):
        ''' March rays through the occupancy bitfield (forward only).  Arguments as raymarching.py:164-183.
        Returns xyzs [M, 3], dirs [M, 3], deltas [M, 2] (dt, t - last_t) and rays int32 [N, 3]
        (ray id, first row, row count). '''
        if not rays_o.is_cuda: rays_o = rays_o.cuda()
        if not rays_d.is_cuda: rays_d = rays_d.cuda()
        if not density_bitfield.is_cuda: density_bitfield = density_bitfield.cuda()

        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        density_bitfield = density_bitfield.contiguous()
        nears = nears.contiguous()
        fars = fars.contiguous()
        device = rays_o.device

        N = rays_o.shape[0]
        M = N * max_steps  # capacity when every ray is kept

        sliced = force_all_rays or mean_count <= 0
        if not sliced:
            # running-average capacity (raymarching.py:200-203): rays that do not fit are dropped
            if align > 0:
                mean_count += align - mean_count % align
            M = mean_count
            xyzs = torch.zeros(M, 3, dtype=rays_o.dtype, device=device)
            dirs = torch.zeros(M, 3, dtype=rays_o.dtype, device=device)
            deltas = torch.zeros(M, 2, dtype=rays_o.dtype, device=device)
        else:
            # rows [0, total) are fully written by the kernel and the pad rows are zeroed below, so
            # the reference's 134 MB memset of the full capacity (raymarching.py:205-207) is not needed
            xyzs = torch.empty(M, 3, dtype=rays_o.dtype, device=device)
            dirs = torch.empty(M, 3, dtype=rays_o.dtype, device=device)
            deltas = torch.empty(M, 2, dtype=rays_o.dtype, device=device)
        rays = torch.empty(N, 3, dtype=torch.int32, device=device)  # id, offset, num_steps

        if step_counter is None:
            step_counter = torch.zeros(2, dtype=torch.int32, device=device)  # point counter, ray counter

        if perturb:
            noises = torch.rand(N, dtype=rays_o.dtype, device=device)
        else:
            noises = torch.zeros(N, dtype=rays_o.dtype, device=device)

        lib = _cabi.load()
        ws = _workspace(lib.ngp_march_rays_train_workspace(N, int(max_steps)), device)
        _cabi.call("ngp_march_rays_train", device, _cabi.ptr(rays_o), _cabi.ptr(rays_d), _cabi.ptr(density_bitfield),
                   float(bound), float(dt_gamma), int(max_steps), N, int(C), int(H), M, _cabi.ptr(nears), _cabi.ptr(fars),
                   _cabi.ptr(xyzs), _cabi.ptr(dirs), _cabi.ptr(deltas), _cabi.ptr(rays), _cabi.ptr(step_counter),
                   _cabi.ptr(noises), _cabi.ptr(ws), ws.numel())

        if sliced:
            total = step_counter[0].item()  # D2H copy, as raymarching.py:224 (the API returns data-dependent shapes)
            m = total
            if align > 0:
                m += align - m % align
            m = min(m, M)
            xyzs = xyzs[:m]
            dirs = dirs[:m]
            deltas = deltas[:m]
            if m > total:
                xyzs[total:].zero_()
                dirs[total:].zero_()
                deltas[total:].zero_()

        return xyzs, dirs, deltas, rays

march_rays_train = _march_rays_train.apply


class _composite_rays_train(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh=1e-4):
        ''' sigmas [M], rgbs [M, 3], deltas [M, 2], rays int32 [N, 3]
        Returns weights_sum [N], depth [N], image [N, 3] (colour pre-multiplied by alpha). '''
        sigmas = sigmas.contiguous()
        rgbs = rgbs.contiguous()
        deltas = deltas.contiguous()
        rays = rays.contiguous()
        M = sigmas.shape[0]
        N = rays.shape[0]
        weights_sum = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        depth = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        image = torch.empty(N, 3, dtype=sigmas.dtype, device=sigmas.device)
        _cabi.call("ngp_composite_rays_train_forward", sigmas.device, _cabi.ptr(sigmas), _cabi.ptr(rgbs), _cabi.ptr(deltas),
                   _cabi.ptr(rays), M, N, float(T_thresh), _cabi.ptr(weights_sum), _cabi.ptr(depth), _cabi.ptr(image))
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, weights_sum, depth, image)
        ctx.dims = [M, N, T_thresh]
        return weights_sum, depth, image

    @staticmethod
    @_bwd
    def backward(ctx, grad_weights_sum, grad_depth, grad_image):
        # grad_depth is not propagated (raymarching.py:275)
        grad_weights_sum = grad_weights_sum.contiguous()
        grad_image = grad_image.contiguous()
        sigmas, rgbs, deltas, rays, weights_sum, depth, image = ctx.saved_tensors
        M, N, T_thresh = ctx.dims
        grad_sigmas = torch.zeros_like(sigmas)
        grad_rgbs = torch.zeros_like(rgbs)
        _cabi.call("ngp_composite_rays_train_backward", sigmas.device, _cabi.ptr(grad_weights_sum), _cabi.ptr(grad_image),
                   _cabi.ptr(sigmas), _cabi.ptr(rgbs), _cabi.ptr(deltas), _cabi.ptr(rays), _cabi.ptr(weights_sum),
                   _cabi.ptr(image), M, N, float(T_thresh), _cabi.ptr(grad_sigmas), _cabi.ptr(grad_rgbs))
        return grad_sigmas, grad_rgbs, None, None, None

composite_rays_train = _composite_rays_train.apply

# ----------------------------------------
# infer functions
# ----------------------------------------

class _march_rays(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far,
                align=-1, perturb=False, dt_gamma=0, max_steps=1024):
        ''' March every alive ray for up to n_step samples (inference).  Arguments as raymarching.py:300-320.
        Returns xyzs, dirs [n_alive * n_step (+pad), 3] and deltas [.., 2]; unused slots stay zero. '''
        if not rays_o.is_cuda: rays_o = rays_o.cuda()
        if not rays_d.is_cuda: rays_d = rays_d.cuda()
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        device = rays_o.device

        M = n_alive * n_step
        if align > 0:
            M += align - (M % align)

        xyzs = torch.zeros(M, 3, dtype=rays_o.dtype, device=device)
        dirs = torch.zeros(M, 3, dtype=rays_o.dtype, device=device)
        deltas = torch.zeros(M, 2, dtype=rays_o.dtype, device=device)  # (dt for rgb, t - last_t for depth)

        if perturb:
            noises = torch.rand(n_alive, dtype=rays_o.dtype, device=device)
        else:
            noises = torch.zeros(n_alive, dtype=rays_o.dtype, device=device)

        _cabi.call("ngp_march_rays", device, int(n_alive), int(n_step), _cabi.ptr(rays_alive), _cabi.ptr(rays_t),
                   _cabi.ptr(rays_o), _cabi.ptr(rays_d), float(bound), float(dt_gamma), int(max_steps), int(C), int(H),
                   _cabi.ptr(density_bitfield), _cabi.ptr(near), _cabi.ptr(far), _cabi.ptr(xyzs), _cabi.ptr(dirs),
                   _cabi.ptr(deltas), _cabi.ptr(noises))
        return xyzs, dirs, deltas

march_rays = _march_rays.apply


class _composite_rays(Function):
    @staticmethod
    @_fwd32  # sigmas & rgbs arrive as half under autocast
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
        ''' In-place accumulation of up to n_step samples per alive ray (inference); raymarching.py:354-370. '''
        sigmas = sigmas.contiguous()
        rgbs = rgbs.contiguous()
        _cabi.call("ngp_composite_rays", sigmas.device, int(n_alive), int(n_step), float(T_thresh), _cabi.ptr(rays_alive),
                   _cabi.ptr(rays_t), _cabi.ptr(sigmas), _cabi.ptr(rgbs), _cabi.ptr(deltas), _cabi.ptr(weights_sum),
                   _cabi.ptr(depth), _cabi.ptr(image))
        return tuple()

composite_rays = _composite_rays.apply


def compact_alive(rays_alive, n_alive=None):
    ''' Device-side `rays_alive[rays_alive >= 0]` (nerf/renderer.py:529).  Returns (buffer, n_out) where
    n_out is a 1-element int32 CUDA tensor and buffer[:n_out] holds the survivors in order. '''
    rays_alive = rays_alive.contiguous()
    n = rays_alive.shape[0] if n_alive is None else int(n_alive)
    out = torch.empty_like(rays_alive)
    n_out = torch.empty(1, dtype=torch.int32, device=rays_alive.device)
    lib = _cabi.load()
    ws = _workspace(lib.ngp_compact_alive_workspace(n), rays_alive.device)
    _cabi.call("ngp_compact_alive", rays_alive.device, _cabi.ptr(rays_alive), n, _cabi.ptr(out), _cabi.ptr(n_out),
               _cabi.ptr(ws), ws.numel())
    return out, n_out
