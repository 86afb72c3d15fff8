"""The `-O` train step around ``NeRFRenderer.render`` - eager, or captured once into a CUDA graph.

``TrainStep`` restates the reference Trainer's per-step call pattern (nerf/utils.py:337-403 train_step,
:696-713 train_one_epoch) for a synthetic guidance gradient: occupancy refresh every 16 steps, render under
fp16 autocast, the guidance's manual ``pred_rgb.backward(gradient=G, retain_graph=True)`` (nerf/sd.py:115),
the entropy regulariser through ``GradScaler``, Adam (betas 0.9/0.99, eps 1e-15, 10x lr for the encoder:
main.py:128, network_grid.py:170-181).  Data parallel: views are sharded across ranks and ONE all-reduce of
the flat gradient bucket precedes the optimizer step (parallel.py).

With ``graph=True`` everything between "inputs are in the static buffers" and "parameters are updated" is
one ``cudaGraphLaunch``: the training render has static shapes and no host sync (render_train.py), the
optimizer is torch's fused capturable Adam driven by the scaler's device-side found_inf, and NCCL
all-reduce is graph-capturable.  The occupancy refresh stays outside the graph (it runs every 16th step).
"""
import torch
import torch.distributed as dist

from . import _cabi
from .optim import FusedAdamScaler
from .parallel import FlatGradBucket
from .step_ops import entropy_loss as fused_entropy_loss


def entropy_loss(weights_sum, lam=1e-4):
    """lambda_entropy * binary entropy of the per-ray opacity (nerf/utils.py:389-394), as the reference's chain of
    torch ops (the one-launch version is step_ops.entropy_loss)."""
    alphas = weights_sum.clamp(1e-5, 1 - 1e-5)
    return lam * (-alphas * torch.log2(alphas) - (1 - alphas) * torch.log2(1 - alphas)).mean()


class TrainStep:
    def __init__(self, model, H, W, lr=1e-3, max_steps=1024, lambda_entropy=1e-4, update_interval=16, graph=False,
                 world_size=1, fused_optimizer=True, lr_decay=None):
        self.model, self.H, self.W = model, H, W
        self.max_steps, self.lam, self.update_interval = max_steps, lambda_entropy, update_interval
        self.world = world_size
        self.use_graph = graph
        self.single_backward = True  # one engine pass for both roots (see _body); False = the reference's two passes
        device = next(model.parameters()).device
        self.device = device
        self.fused_optimizer = fused_optimizer
        if fused_optimizer:
            # one flat buffer each for params / grads / moments / fp16 shadow; unscale + Adam + scaler.update + shadow
            # refresh + zero_grad in two launches (optim.py); the all-reduce mean is folded into the unscale
            self.opt = FusedAdamScaler(model.get_params(lr), betas=(0.9, 0.99), eps=1e-15, grad_div=float(world_size),
                                       lr_decay=lr_decay)
            self.scaler = self.opt
            self.bucket = None
            self.flat_grads = self.opt.flat_grads
        else:
            self.opt = torch.optim.Adam(model.get_params(lr), betas=(0.9, 0.99), eps=1e-15, fused=True, capturable=graph)
            self.scaler = torch.amp.GradScaler("cuda")
            self.bucket = FlatGradBucket(list(model.parameters()), device)
            self.flat_grads = self.bucket.flat
        self.global_step = 0
        self.n_updates = 0
        self.samples = torch.zeros(1, dtype=torch.int64, device=device)  # running count of marched samples
        self._graph = None
        self._graph_launches = 0
        self._static = None
        self.loss = None

    # -- one step's device work, shape-static when the fused training render is active ---------------------------
    def _body(self, rays_o, rays_d, G):
        model = self.model
        B = rays_o.shape[0]
        if not self.fused_optimizer:
            self.bucket.zero()  # (the fused optimizer kernel leaves the gradient buffer zeroed)
        with torch.autocast("cuda", torch.float16):
            out = model.render(rays_o, rays_d, staged=False, perturb=True, bg_color=None, ambient_ratio=1.0,
                               shading="albedo", force_all_rays=True, max_steps=self.max_steps, dt_gamma=0)
            pred_rgb = out["image"].reshape(B, self.H, self.W, 3).permute(0, 3, 1, 2).contiguous()
            ws = out["weights_sum"].reshape(B, 1, self.H, self.W)
            loss = fused_entropy_loss(ws, self.lam) if self.fused_optimizer else entropy_loss(ws, self.lam)
        # The reference runs TWO backward passes over the render graph per step: the guidance's manual
        # `pred_rgb.backward(gradient=G, retain_graph=True)` (nerf/sd.py:115) and `scaler.scale(loss).backward()`
        # (nerf/utils.py:708).  Back-propagation is linear in the upstream gradient, so both roots are handed to the
        # engine at once: it sums d(image), d(weights_sum) at the render node and walks the heavy part (composite,
        # MLP, grid scatter) once.  Same accumulated .grad up to fp32 summation order (tests/test_gpu_pipeline.py).
        if self.single_backward:
            torch.autograd.backward([pred_rgb, self.scaler.scale(loss)], [G, None])
        else:
            pred_rgb.backward(gradient=G, retain_graph=True)
            self.scaler.scale(loss).backward()
        if self.world > 1:
            dist.all_reduce(self.flat_grads, op=dist.ReduceOp.SUM)
        if self.fused_optimizer:
            self.opt.step(zero_grads=True)
        else:
            if self.world > 1:
                self.bucket.flat.div_(self.world)
            self.scaler.step(self.opt)
            self.scaler.update()
        return loss

    def _bookkeeping_after(self, local_step_before):
        """Python-side effects of run_cuda that a graph replay does not re-execute."""
        model = self.model
        row = local_step_before % 16
        ws = getattr(model, "_train_ws", None)
        if ws is not None:
            model.step_counter[row].copy_(ws.counter)
        model.local_step = local_step_before + 1

    def __call__(self, rays_o, rays_d, G):
        model = self.model
        if self.global_step % self.update_interval == 0:
            if self.use_graph and not self.fused_optimizer:
                from . import field
                field.invalidate_half_cache()  # graph replays update the parameters without bumping ._version
            with torch.autocast("cuda", torch.float16):
                model.update_extra_state()
            self.n_updates += 1
        self.global_step += 1
        (self.opt.attach_grads if self.fused_optimizer else self.bucket.attach)()

        if not self.use_graph:
            loss = self._body(rays_o, rays_d, G)
            self.samples.add_(model.step_counter[(model.local_step - 1) % 16, 0].long())
            self.loss = loss
            return loss

        if self._graph is None:
            self._capture(rays_o, rays_d, G)
        ro_s, rd_s, g_s = self._static
        ro_s.copy_(rays_o, non_blocking=True)
        rd_s.copy_(rays_d, non_blocking=True)
        g_s.copy_(G, non_blocking=True)
        before = model.local_step
        self._graph.replay()
        _cabi.LAUNCHES += self._graph_launches  # our kernels inside the replayed graph
        self._bookkeeping_after(before)
        self.samples.add_(model._train_ws.counter[0].long())
        return self.loss

    def _capture(self, rays_o, rays_d, G):
        model = self.model
        self._static = (torch.empty_like(rays_o, device=self.device).copy_(rays_o),
                        torch.empty_like(rays_d, device=self.device).copy_(rays_d),
                        torch.empty_like(G, device=self.device).copy_(G))
        ro_s, rd_s, g_s = self._static
        # warm up on a side stream (allocator, lazy initialisation, cudaFuncSetAttribute) before capturing
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        saved_step = model.local_step
        with torch.cuda.stream(side):
            for _ in range(3):
                self._body(ro_s, rd_s, g_s)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        model.local_step = saved_step
        self._graph = torch.cuda.CUDAGraph()
        launches0 = _cabi.LAUNCHES
        with torch.cuda.graph(self._graph, capture_error_mode="thread_local"):
            self.loss = self._body(ro_s, rd_s, g_s)
        self._graph_launches = _cabi.LAUNCHES - launches0
        _cabi.LAUNCHES = launches0  # capture launches nothing
        model.local_step = saved_step
