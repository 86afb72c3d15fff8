"""``NeRFRenderer`` - host-side mirror of the reference's cuda-ray renderer (nerf/renderer.py:64-653).

The contract of ``run_cuda`` / ``update_extra_state`` / ``render`` is the reference's: same
arguments (including the swallowed ``**kwargs`` the Trainer splats in, nerf/utils.py:363), same
``results`` dict, same registered buffers (``aabb_train``, ``aabb_infer``, ``density_grid``,
``density_bitfield``, ``step_counter``) so reference checkpoints load, same side effects on
``local_step`` / ``step_counter``.  The plumbing underneath is new:

* the occupancy update builds its query points directly in Morton order and does EMA-max, mean and
  bit packing on the device without a host round trip (``mean_density`` / ``mean_count`` are read
  back lazily, only if somebody asks for the python value);
* the inference loop compacts alive rays with a device kernel instead of boolean-mask indexing.

``export_mesh`` (offline marching cubes + UV unwrap) is outside the hot path and not provided.
"""
import math

import torch
import torch.nn as nn

import raymarching
from . import _cabi


def safe_normalize(x, eps=1e-20):
    return x / torch.sqrt(torch.clamp(torch.sum(x * x, -1, keepdim=True), min=eps))


class _LazyScalar:
    """A device scalar whose python value is fetched (one sync) only on first use."""

    def __init__(self, tensor=None, value=None, fn=None):
        self.tensor, self.value, self.fn = tensor, value, fn

    def get(self):
        if self.value is None:
            v = self.tensor.item() if self.tensor is not None else 0
            self.value = self.fn(v) if self.fn is not None else v
        return self.value


class _LazyCount:
    """mean_count = int(sum(step_counter[:n, 0]) / n), evaluated (one reduction + sync) on first use."""

    def __init__(self, snapshot, n):
        self.snapshot, self.n, self.value = snapshot, n, None

    def get(self):
        if self.value is None:
            self.value = int(int(self.snapshot[:self.n].sum().item()) / self.n)
        return self.value


class NeRFRenderer(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.bound = opt.bound
        self.cascade = 1 + math.ceil(math.log2(opt.bound))
        self.grid_size = 128
        self.cuda_ray = opt.cuda_ray
        self.min_near = opt.min_near
        self.density_thresh = opt.density_thresh
        self.bg_radius = opt.bg_radius

        # (xmin, ymin, zmin, xmax, ymax, zmax); only used to clip rays - hashing uses the cubic bound
        aabb_train = torch.FloatTensor([-opt.bound, -opt.bound, -opt.bound, opt.bound, opt.bound, opt.bound])
        self.register_buffer('aabb_train', aabb_train)
        self.register_buffer('aabb_infer', aabb_train.clone())

        if self.cuda_ray:
            self.register_buffer('density_grid', torch.zeros([self.cascade, self.grid_size ** 3]))
            self.register_buffer('density_bitfield',
                                 torch.zeros(self.cascade * self.grid_size ** 3 // 8, dtype=torch.uint8))
            self._mean_density = _LazyScalar(value=0)
            self.iter_density = 0
            self.register_buffer('step_counter', torch.zeros(16, 2, dtype=torch.int32))  # 16 steps averaged
            self._mean_count = _LazyScalar(value=0)
            self.local_step = 0

    # python-visible scalars of the reference (renderer.py:91,96), fetched lazily from the device
    @property
    def mean_density(self):
        return self._mean_density.get()

    @mean_density.setter
    def mean_density(self, v):
        self._mean_density = _LazyScalar(value=v)

    @property
    def mean_count(self):
        return self._mean_count.get()

    @mean_count.setter
    def mean_count(self, v):
        self._mean_count = _LazyScalar(value=v)

    def forward(self, x, d):
        raise NotImplementedError()

    def density(self, x):
        raise NotImplementedError()

    def color(self, x, d, mask=None, **kwargs):
        raise NotImplementedError()

    def reset_extra_state(self):
        if not self.cuda_ray:
            return
        self.density_grid.zero_()
        self.mean_density = 0
        self.iter_density = 0
        self.step_counter.zero_()
        self.mean_count = 0
        self.local_step = 0

    def export_mesh(self, *args, **kwargs):
        raise NotImplementedError("export_mesh (offline mesh extraction) is outside the B200 hot path")

    fused_train = True  # take the sync-free fused training render when the field has the reference's shape

    def _fused_training_path(self, rays_o):
        if not self.fused_train or not getattr(self, "fused", False):
            return False
        from . import field as _field
        enc, net = getattr(self, "encoder", None), getattr(self, "sigma_net", None)
        if enc is None or net is None:
            return False
        probe = rays_o.detach().view(-1, 3)
        return _field.can_fuse(probe, enc, net)

    def run(self, *args, **kwargs):
        raise NotImplementedError(
            "the pure-PyTorch sampler (nerf/renderer.py:301) is not part of the B200 hot path; construct the model "
            "with opt.cuda_ray=True")

    # --------------------------------------------------------------------------------------------
    def run_cuda(self, rays_o, rays_d, dt_gamma=0, light_d=None, ambient_ratio=1.0, shading='albedo', bg_color=None,
                 perturb=False, force_all_rays=False, max_steps=1024, T_thresh=1e-4, **kwargs):
        # rays_o, rays_d: [B, N, 3] -> image [B, N, 3], depth [B, N], weights_sum [B, N], mask [B, N]
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        N = rays_o.shape[0]
        device = rays_o.device

        # NOTE the reference passes no min_near here, so the wrapper default 0.2 applies (renderer.py:458)
        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, self.aabb_train if self.training else self.aabb_infer)

        if light_d is None:
            # light roughly from the camera side so the visible face is lit (renderer.py:461-464)
            light_d = safe_normalize(rays_o[0] + torch.randn(3, device=device, dtype=torch.float))

        results = {}

        if self.training:
            counter = self.step_counter[self.local_step % 16]
            self.local_step += 1
            normals = None

            if shading == 'albedo' and self._fused_training_path(rays_o):
                # sync-free path: fixed-capacity buffers + device-side sample count (render_train.py); draws the
                # same torch.rand(N) the reference's wrapper draws (raymarching.py:213-216)
                from .render_train import render_train
                if perturb:
                    noises = torch.rand(N, dtype=torch.float32, device=device)
                else:
                    noises = torch.zeros(N, dtype=torch.float32, device=device)
                weights_sum, depth, image = render_train(self, rays_o.float(), rays_d.float(), nears, fars, noises,
                                                         dt_gamma, max_steps, T_thresh)
                counter.copy_(self._train_ws.counter)
            else:
                counter.zero_()
                xyzs, dirs, deltas, rays = raymarching.march_rays_train(
                    rays_o, rays_d, self.bound, self.density_bitfield, self.cascade, self.grid_size, nears, fars, counter,
                    self.mean_count if not force_all_rays else -1, perturb, 128, force_all_rays, dt_gamma, max_steps)

                sigmas, rgbs, normals = self(xyzs, dirs, light_d, ratio=ambient_ratio, shading=shading)

                weights_sum, depth, image = raymarching.composite_rays_train(sigmas, rgbs, deltas, rays, T_thresh)

            if normals is not None:
                # orientation + smoothness regularisers (renderer.py:485-494)
                weights = 1 - torch.exp(-sigmas)
                loss_orient = weights.detach() * (normals * dirs).sum(-1).clamp(min=0) ** 2
                results['loss_orient'] = loss_orient.mean()
                normals_perturb = self.normal(xyzs + torch.randn_like(xyzs) * 1e-2)
                results['loss_smooth'] = (normals - normals_perturb).abs().mean()
        else:
            dtype = torch.float32
            weights_sum = torch.zeros(N, dtype=dtype, device=device)
            depth = torch.zeros(N, dtype=dtype, device=device)
            image = torch.zeros(N, 3, dtype=dtype, device=device)

            done = False
            if shading == 'albedo' and self.infer_loop == 'graph' and self._fused_training_path(rays_o):
                # the whole loop as one CUDA-graph launch with a conditional WHILE node: no host sync per iteration
                from . import _nvtx
                with _nvtx.range("ngp.infer.loop"):
                    res = self._run_infer_loop(rays_o, rays_d, nears, fars, perturb, dt_gamma, max_steps, T_thresh)
                if res is not None:
                    weights_sum, depth, image = res
                    done = True

            n_alive = 0 if done else N
            rays_alive = torch.arange(n_alive, dtype=torch.int32, device=device)
            rays_t = nears.clone()

            step = 0
            while step < max_steps and n_alive > 0:
                # more steps per launch as rays die off (renderer.py:521)
                n_step = max(min(N // n_alive, 8), 1)
                xyzs, dirs, deltas = raymarching.march_rays(
                    n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, self.bound, self.density_bitfield, self.cascade,
                    self.grid_size, nears, fars, 128, perturb if step == 0 else False, dt_gamma, max_steps)
                sigmas, rgbs, normals = self(xyzs, dirs, light_d, ratio=ambient_ratio, shading=shading)
                raymarching.composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth,
                                           image, T_thresh)
                rays_alive, n_out = raymarching.compact_alive(rays_alive, n_alive)
                n_alive = int(n_out.item())  # the loop shape is data dependent (one 4-byte D2H per iteration)
                step += n_step

        if self.bg_radius > 0:
            bg_color = self.background(rays_d)  # [N, 3]
        elif bg_color is None:
            bg_color = 1

        # blend + depth normalisation + hit mask (renderer.py:541-551) in one launch; depth is NaN for rays that miss
        # the box (0/0), as in the reference
        from .step_ops import blend_background
        image, depth, mask = blend_background(image.float(), weights_sum, depth, bg_color, nears, fars)
        image = image.view(*prefix, 3)
        depth = depth.view(*prefix)
        weights_sum = weights_sum.reshape(*prefix)
        mask = mask.reshape(*prefix)

        results['image'] = image
        results['depth'] = depth
        results['weights_sum'] = weights_sum
        results['mask'] = mask
        return results

    # 'graph': the inference loop runs on the device (csrc/raymarch.cu ngp_render_infer_loop); 'host': the reference's
    # host loop with one 4-byte read per iteration (also the fallback for shaded renders and non-fusable fields)
    infer_loop = 'graph'

    @torch.no_grad()
    def _run_infer_loop(self, rays_o, rays_d, nears, fars, perturb, dt_gamma, max_steps, T_thresh):
        import numpy as np
        from .field import cached_half
        N, device = rays_o.shape[0], rays_o.device
        lib = _cabi.load()
        ws = getattr(self, "_infer_ws", None)
        if ws is None or ws["N"] != N or ws["work"].device != device:
            e = lambda *shape: torch.empty(*shape, device=device, dtype=torch.float32)  # noqa: E731
            ws = dict(N=N, rays_o=e(N, 3), rays_d=e(N, 3), nears=e(N), fars=e(N), noises=e(N), weights_sum=e(N), depth=e(N),
                      image=e(N, 3), work=torch.empty(int(lib.ngp_render_infer_workspace(N)), device=device, dtype=torch.uint8))
            self._infer_ws = ws
        # the graph bakes its pointers in: inputs are staged in the persistent workspace (one graph per resolution)
        ws["rays_o"].copy_(rays_o); ws["rays_d"].copy_(rays_d); ws["nears"].copy_(nears); ws["fars"].copy_(fars)
        if perturb:
            ws["noises"].uniform_()          # the torch.rand(n_alive) of the first march_rays call (raymarching.py:337-339)
        enc = self.encoder
        l0, l1, l2 = self.sigma_net.net
        hw = [cached_half(t) for t in (l0.weight, l0.bias, l1.weight, l1.bias, l2.weight, l2.bias)]
        table = cached_half(enc.embeddings)
        P = _cabi.ptr
        L, S = enc.offsets.shape[0] - 1, float(np.log2(enc.per_level_scale))
        # quad table of the current fp16 embeddings (csrc/field_mlp.cu: one or two 16-byte gathers per level instead of four
        # to eight 4-byte ones), rebuilt per frame: 13 us next to a ~7 ms frame, and never stale
        quads = ws.get("quads")
        if quads is None or quads.shape[0] != table.shape[0]:
            quads = ws["quads"] = torch.zeros(table.shape[0], 4, dtype=torch.int32, device=device)
        _cabi.call("ngp_grid_quad_table", device, P(table), P(enc.offsets), L, table.shape[0], S, int(enc.base_resolution),
                   int(enc.gridtype_id), int(bool(enc.align_corners)), P(quads))
        rc = _cabi.call_rc("ngp_render_infer_loop_quads", device, P(ws["rays_o"]), P(ws["rays_d"]), P(ws["nears"]), P(ws["fars"]), N,
                           float(self.bound), float(dt_gamma), int(max_steps), int(self.cascade), int(self.grid_size),
                           P(self.density_bitfield), float(T_thresh), P(ws["noises"]) if perturb else None, P(table), P(quads),
                           P(enc.offsets), L, 2, S,
                           int(enc.base_resolution), int(enc.gridtype_id), int(bool(enc.align_corners)), *[P(t) for t in hw],
                           64, 4, P(ws["weights_sum"]), P(ws["depth"]), P(ws["image"]), P(ws["work"]), ws["work"].numel(),
                           launches=1)
        if rc != 0:
            if rc == -2:   # NGP_ERR_UNSUPPORTED: this driver cannot build the conditional graph - host loop from now on
                type(self).infer_loop = 'host'
                return None
            _cabi.check(rc, "ngp_render_infer_loop")
        return ws["weights_sum"].clone(), ws["depth"], ws["image"]

    def infer_loop_iterations(self):
        """Iterations the last device-driven inference loop ran (one device->host read; diagnostics / bench only)."""
        import ctypes
        ws = getattr(self, "_infer_ws", None)
        if ws is None:
            return None
        st = (ctypes.c_int * 8)()
        _cabi.check(_cabi.load().ngp_render_infer_state(_cabi.ptr(ws["work"]), st, torch.cuda.current_stream().cuda_stream),
                    "ngp_render_infer_state")
        return int(st[4])

    # --------------------------------------------------------------------------------------------
    @torch.no_grad()
    def update_extra_state(self, decay=0.95, S=128, noise=None):
        """Occupancy-grid refresh (renderer.py:562-613).  `noise` ([cascade, H^3, 3] uniform in [0,1), indexed by
        the linear cell id) can be injected for reproducible comparisons; by default it is drawn with
        torch.rand exactly where the reference draws its jitter (one rand per cascade, same shape)."""
        if not self.cuda_ray:
            return
        H = self.grid_size
        device = self.density_bitfield.device
        n_cells = H ** 3
        shard = getattr(self, "dp_shard", None)  # (rank, world[, group]): data-parallel refresh, see below
        sharded = shard is not None and noise is None and shard[1] > 1 and n_cells % shard[1] == 0
        n_query = n_cells // shard[1] if sharded else n_cells
        # persistent scratch: the refresh allocates nothing after its first call (an allocation after a CUDA graph has
        # been captured costs tens of milliseconds of cudaMalloc on the step it lands on)
        ws = getattr(self, "_occ_ws", None)
        if ws is None or ws["n"] != n_query or ws["u"].device != device:
            e = lambda *shape, dtype=torch.float32: torch.empty(*shape, device=device, dtype=dtype)  # noqa: E731
            ws = dict(n=n_query, u=e(n_query, 3), xyzs=e(n_query, 3), sigma=e(n_query), rgb=e(n_query, 3),
                      tmp=torch.empty_like(self.density_grid), mean=e(1), acc=e(16, dtype=torch.uint8))
            self._occ_ws = ws
        tmp_grid = ws["tmp"]

        for cas in range(self.cascade):
            bound = min(2 ** cas, self.bound)
            half_cell = bound / H
            if noise is None:
                u = ws["u"].uniform_()  # the torch.rand of renderer.py:586, drawn in place
            else:
                u = noise[cas].to(device, torch.float32).contiguous()
            xyzs = ws["xyzs"]
            if sharded:
                # every rank evaluates the density of 1/world of the cells (a contiguous Morton range, with its own
                # jitter) and the ranges are all-gathered: identical grids on all ranks at 1/world of the field work
                import torch.distributed as dist
                _cabi.call("ngp_occupancy_cell_points_range", device, H, float(bound - half_cell), float(half_cell),
                           _cabi.ptr(u), shard[0] * n_query, n_query, _cabi.ptr(xyzs))
                dist.all_gather_into_tensor(tmp_grid[cas], self._density_into(xyzs, ws),
                                            group=shard[2] if len(shard) > 2 else None)
            else:
                _cabi.call("ngp_occupancy_cell_points", device, H, float(bound - half_cell), float(half_cell), _cabi.ptr(u),
                           _cabi.ptr(xyzs))
                # Morton-ordered queries: the densities land in density_grid order, no index scatter needed
                sig = self._density_into(xyzs, ws)
                if self.cascade == 1:
                    tmp_grid = sig.view(1, -1)  # the scratch itself: no copy
                else:
                    tmp_grid[cas].copy_(sig)

        mean = ws["mean"]
        _cabi.call("ngp_update_density_grid", device, _cabi.ptr(self.density_grid), _cabi.ptr(tmp_grid),
                   self.cascade * n_cells, float(decay), float(self.density_thresh), _cabi.ptr(mean),
                   _cabi.ptr(self.density_bitfield), _cabi.ptr(ws["acc"]), ws["acc"].numel())
        self._mean_density = _LazyScalar(tensor=mean)
        self.iter_density += 1

        # mean sample count of the last (up to 16) steps (renderer.py:610-613): a snapshot of the counters now, the
        # sum and the device->host read only if somebody asks for the value (nobody does with force_all_rays)
        total_step = min(16, self.local_step)
        if total_step > 0:
            snap = ws.setdefault("count_snap", torch.zeros(16, dtype=torch.int32, device=device))
            snap.copy_(self.step_counter[:, 0])
            self._mean_count = _LazyCount(snap, total_step)
        self.local_step = 0

    def _density_into(self, xyzs, ws):
        """sigma (fp32 [n]) of the query points, written into the refresh scratch when the fused field applies."""
        from . import field as _field
        enc, net = getattr(self, "encoder", None), getattr(self, "sigma_net", None)
        if getattr(self, "fused", False) and enc is not None and net is not None and _field.can_fuse(xyzs, enc, net):
            _field.fused_field_into(xyzs, enc, net, self.bound, ws["sigma"], ws["rgb"])
            return ws["sigma"]
        return self.density(xyzs)['sigma'].reshape(-1).detach().float().contiguous()

    def render(self, rays_o, rays_d, staged=False, max_ray_batch=4096, **kwargs):
        # rays_o, rays_d: [B, N, 3]; never staged in cuda_ray mode (renderer.py:630-652)
        if not self.cuda_ray:
            return self.run(rays_o, rays_d, **kwargs)
        return self.run_cuda(rays_o, rays_d, **kwargs)
