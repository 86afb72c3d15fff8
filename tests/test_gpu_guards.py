"""Out-of-bounds evidence without compute-sanitizer (closed on this GPU pool): every CUDA tensor the Python layer allocates
while the hot path runs is placed between two 512-byte guard bands filled with a sentinel; after the run every band must
be intact.  Covers the drop-in ops (grid encode, march, composite, inference loop, occupancy refresh), the fused tcgen05
field kernels, the shading stencil and the hand-scheduled train step, i.e. every kernel that writes through a pointer the
host handed it.  (Reads past a buffer are not caught this way; sizes at tile / warp boundaries are chosen to provoke
tail-handling bugs: ray counts and sample counts that are not multiples of 32 / 128.)"""
import argparse
import contextlib

import numpy as np
import pytest
import torch

import ngp_testutil as util

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
PAD = 512
SENTINEL = 0xA5


class Guards:
    def __init__(self):
        self.records = []
        self.orig = {}

    def _wrap(self, shape, dtype, device):
        dtype = dtype or torch.get_default_dtype()
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        n_al = (n + 15) // 16 * 16
        raw = self.orig["empty"](n_al + 2 * PAD, dtype=torch.uint8, device=device)
        raw.fill_(SENTINEL)
        self.records.append((raw, n))
        return raw[PAD:PAD + n].view(dtype).view(tuple(shape))

    @staticmethod
    def _shape(args):
        if len(args) == 1 and isinstance(args[0], (tuple, list, torch.Size)):
            return tuple(int(v) for v in args[0])
        return tuple(int(v) for v in args)

    @contextlib.contextmanager
    def active(self):
        g = self
        self.orig = dict(empty=torch.empty, zeros=torch.zeros, empty_like=torch.empty_like, zeros_like=torch.zeros_like)

        def is_cuda(device):
            return device is not None and torch.device(device).type == "cuda"

        def empty(*args, dtype=None, device=None, **kw):
            if not is_cuda(device) or kw.get("pin_memory") or "out" in kw:
                return g.orig["empty"](*args, dtype=dtype, device=device, **kw)
            return g._wrap(g._shape(args), dtype, device)

        def zeros(*args, dtype=None, device=None, **kw):
            if not is_cuda(device) or "out" in kw:
                return g.orig["zeros"](*args, dtype=dtype, device=device, **kw)
            t = g._wrap(g._shape(args), dtype, device)
            t.zero_()
            return t

        def empty_like(t, dtype=None, device=None, **kw):
            dev = device if device is not None else t.device
            if not is_cuda(dev):
                return g.orig["empty_like"](t, dtype=dtype, device=device, **kw)
            return g._wrap(tuple(t.shape), dtype or t.dtype, dev)

        def zeros_like(t, dtype=None, device=None, **kw):
            out = empty_like(t, dtype=dtype, device=device)
            out.zero_()
            return out

        torch.empty, torch.zeros, torch.empty_like, torch.zeros_like = empty, zeros, empty_like, zeros_like
        try:
            yield self
        finally:
            torch.empty, torch.zeros = self.orig["empty"], self.orig["zeros"]
            torch.empty_like, torch.zeros_like = self.orig["empty_like"], self.orig["zeros_like"]

    def check(self):
        torch.cuda.synchronize()
        bad = 0
        for raw, n in self.records:
            head, tail = raw[:PAD], raw[PAD + n:]
            if not (bool((head == SENTINEL).all()) and bool((tail == SENTINEL).all())):
                bad += 1
        return len(self.records), bad


def _model():
    from ngp_b200.network_grid import NeRFNetwork
    opt = argparse.Namespace(bound=1, cuda_ray=True, min_near=0.1, density_thresh=10, bg_radius=1.4)
    torch.manual_seed(0)
    m = NeRFNetwork(opt).to(DEV).train()
    with torch.no_grad():
        m.encoder.embeddings.uniform_(-0.3, 0.3)
    return m


def test_no_kernel_writes_outside_its_buffers():
    from ngp_b200.trainer import TrainStep
    g = Guards()
    side = 27                                   # 729 rays: not a multiple of 32 or 128
    rays_o, rays_d = util.look_at_rays(side, radius=1.4, phi_deg=75)
    with g.active():
        model = _model()
        ro, rd = torch.from_numpy(rays_o).to(DEV)[None], torch.from_numpy(rays_d).to(DEV)[None]
        with torch.autocast("cuda", torch.float16):
            model.update_extra_state()
            # modular drop-in path (grid encode / march / composite wrappers) and the fused render, forward + backward
            for fused in (False, True):
                model.fused = fused
                out = model.render(ro, rd, staged=False, perturb=True, force_all_rays=True, max_steps=200, shading="albedo")
                (out["image"].sum() + out["weights_sum"].sum()).backward()
            # shading stencil (7-point + 6-point batches, shade kernels)
            out = model.render(ro, rd, staged=False, perturb=True, force_all_rays=False, max_steps=96, shading="lambertian",
                               ambient_ratio=0.1)
            (out["image"].sum() + out["loss_orient"]).backward()
        # inference: device-driven loop and host loop
        model.eval()
        for mode in ("graph", "host"):
            model.infer_loop = mode
            with torch.no_grad(), torch.autocast("cuda", torch.float16):
                model.render(ro, rd, staged=True, perturb=True, max_steps=200)
        model.train()
        # hand-scheduled train step: eager, then as graph replays, rays from poses (ngp_train_prologue_rays)
        from ngp_b200 import provider
        poses, intr = provider.make_training_poses(1, side, side, seed=1)
        G = torch.randn(1, 3, side, side, device=DEV) * 1e-2
        for graph in (False, True):
            step = TrainStep(model, side, side, lr=1e-4, max_steps=200, graph=graph, manual=True, n_chunks=2,
                             device_rays=(side, 0, 1))
            for _ in range(3):
                step(poses.to(DEV), intr.to(DEV), G)
            step.flush()
        n, bad = g.check()
    assert n > 100, n
    assert bad == 0, "%d of %d guarded buffers were written outside their bounds" % (bad, n)
