#!/usr/bin/env python
"""bench.py - NeRF train-step samples/s (march + hash-grid + MLP + composite, fwd + bwd) on 1..8 B200.

    python bench.py --gpus N --steps K --warmup W           # the B200-native arm (this repo)
    python bench.py --impl reference --steps K --warmup W   # the reference's CPU renderer on host cores

Workload (BASELINE.json configs[2]/[3]): the `-O` train step of
nerf/network_grid.py (tiled 16x2 grid, 2^16 rows/level, 32->64->64->4 MLP, 128^3 occupancy grid,
max_steps 1024), 8 camera views of 64x64 rays per step sharded over the N ranks, fp16 autocast +
GradScaler, synthetic SDS gradient in place of the U-Net (`pred_rgb.backward(gradient=G,
retain_graph=True)`), entropy regulariser backward, flat-bucket gradient all-reduce, Adam step,
occupancy update every 16 steps (inside the timed region).  A "sample" is one marched point
actually produced (sum of step_counter[:, 0]), padding excluded.

One JSON line is printed by rank 0; see the task contract for the keys.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "single-stable-dreamfusion_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "NeRF train-step samples/s (march+hashgrid+MLP+composite fwd/bwd)"
UNIT = "samples/s"
VIEWS_PER_STEP = 8
H = W = 64
MAX_STEPS = 1024


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--views", type=int, default=VIEWS_PER_STEP, help="camera views per step, whole job")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip timing the reference's CUDA extensions")
    ap.add_argument("--cpu-sample-steps", type=int, default=3)
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's pure-PyTorch non-cuda-ray renderer on host cores
# ---------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup):
    """BASELINE.json configs[0]: 64x64 rays x (64+32) samples, vanilla NeRF, fp32 fwd+bwd on CPU.
    Returns (samples/s, ms/step, cores, description)."""
    import torch
    from oracle import torch_renderer as TR
    torch.manual_seed(0)
    model = TR.VanillaNeRF(bound=1.0, min_near=0.1, bg_radius=1.4)
    model.train()
    views = [TR.make_view(H, W, seed=s) for s in range(4)]
    G = torch.randn(1, H * W, 3, generator=torch.Generator().manual_seed(1))
    for i in range(warmup):
        TR.train_step(model, *views[i % 4], G)
    t0 = time.perf_counter()
    n = 0
    for i in range(steps):
        k, _ = TR.train_step(model, *views[i % 4], G)
        n += k
    dt = time.perf_counter() - t0
    cores = torch.get_num_threads()
    desc = ("%d fwd+bwd steps of one 64x64-ray view x (64+32) samples through a port of nerf/renderer.py:run + "
            "nerf/network.py (fp32, %d torch threads of %d host cpus)" % (steps, cores, os.cpu_count() or 0))
    return n / dt, dt / max(steps, 1) * 1e3, cores, desc


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 12))
    warm = max(1, min(args.warmup, 2))
    v, ms, cores, desc = cpu_reference_run(steps, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "reference CPU path: nerf/network.py vanilla NeRF + NeRFRenderer.run, 64x64 rays x "
                               "(64+32) samples, fwd+bwd (BASELINE configs[0]); sample = composited sample"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
            self.f.close()
            rows = [r.split(",") for r in open(self.path).read().strip().splitlines() if r.strip()]
            sm = sorted(float(r[1]) for r in rows if len(r) >= 9)
            if sm:
                out["sm_mhz"] = sm[len(sm) // 2]
                out["sm_max_mhz"] = max(float(r[2]) for r in rows if len(r) >= 9)
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for k, nm in enumerate(names):
                    if any("Active" == r[5 + k].strip() for r in rows if len(r) >= 9):
                        out["reasons"].append(nm)
            os.remove(self.path)
        except Exception:
            pass
        return out


# ---------------------------------------------------------------------------------------------------
# the B200 arm
# ---------------------------------------------------------------------------------------------------
def build_model(device):
    import torch
    from ngp_b200.network_grid import NeRFNetwork
    opt = argparse.Namespace(bound=1, cuda_ray=True, min_near=0.1, density_thresh=10, bg_radius=1.4)
    torch.manual_seed(0)
    model = NeRFNetwork(opt).to(device)
    model.train()
    return model


def entropy_loss(ws, lam=1e-4):
    import torch
    alphas = ws.clamp(1e-5, 1 - 1e-5)
    return lam * (-alphas * torch.log2(alphas) - (1 - alphas) * torch.log2(1 - alphas)).mean()


class TrainStep:
    """The reference Trainer's call pattern (nerf/utils.py:337-403, 696-713) around the B200 renderer."""

    def __init__(self, model, device, world_size):
        import torch
        from ngp_b200.parallel import FlatGradBucket
        self.model, self.device, self.world = model, device, world_size
        self.opt = torch.optim.Adam(model.get_params(1e-3), betas=(0.9, 0.99), eps=1e-15)
        self.scaler = torch.amp.GradScaler("cuda")
        self.bucket = FlatGradBucket(list(model.parameters()), device)
        self.global_step = 0
        self.n_updates = 0

    def __call__(self, rays_o, rays_d, G):
        import torch
        model = self.model
        if self.global_step % 16 == 0:
            with torch.autocast("cuda", torch.float16):
                model.update_extra_state()
            self.n_updates += 1
        self.global_step += 1
        self.bucket.zero()
        self.bucket.attach()
        B = rays_o.shape[0]
        with torch.autocast("cuda", torch.float16):
            out = model.render(rays_o, rays_d, staged=False, perturb=True, bg_color=None, ambient_ratio=1.0,
                               shading="albedo", force_all_rays=True, max_steps=MAX_STEPS, dt_gamma=0)
            pred_rgb = out["image"].reshape(B, H, W, 3).permute(0, 3, 1, 2).contiguous()
            # synthetic SDS: the guidance back-propagates a given gradient through the NeRF graph (nerf/sd.py:115)
            pred_rgb.backward(gradient=G, retain_graph=True)
            loss = entropy_loss(out["weights_sum"].reshape(B, 1, H, W))
        self.scaler.scale(loss).backward()
        self.bucket.all_reduce(average=True)
        self.scaler.step(self.opt)
        self.scaler.update()
        return loss


def measure_l2_peaks(device):
    """Chip ceilings for the encoder's access shapes: random 4-byte gathers / 8-byte red.add over a 24 MB
    (L2-resident) table.  Returns (gathers/s, reds/s)."""
    import torch
    from ngp_b200 import _cabi
    words = 1 << 23  # 32 MB of u32; power of two; L2 is 126 MB
    table = torch.zeros(words, dtype=torch.int32, device=device)
    n_threads, iters = 148 * 2048 * 4, 16
    sink = torch.empty(n_threads, dtype=torch.int32, device=device)
    ftable = torch.zeros(words, dtype=torch.float32, device=device)
    res = []
    for name, args in (("ngp_bench_gather4", (_cabi.ptr(table), words, _cabi.ptr(sink), n_threads, iters, 1)),
                       ("ngp_bench_red8", (_cabi.ptr(ftable), words, n_threads, iters, 2))):
        best = 0.0
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _cabi.call(name, device, *args)
            e1.record()
            torch.cuda.synchronize()
            ops = n_threads * iters * 8
            best = max(best, ops / (e0.elapsed_time(e1) * 1e-3))
        res.append(best)
    return res[0], res[1]


def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    from ngp_b200 import _cabi, provider
    from ngp_b200.parallel import shard_views

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "the B200 arm needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    _cabi.load()

    first, n_local = shard_views(args.views, rank, world)
    n_pool = 64
    ro_all, rd_all = provider.make_training_views(n_pool * args.views, H, W, seed=0)
    ro_all = ro_all.view(n_pool, args.views, H * W, 3)
    rd_all = rd_all.view(n_pool, args.views, H * W, 3)
    g_host = (torch.randn(n_pool, args.views, 3, H, W, generator=torch.Generator().manual_seed(2)) * 1e-2).pin_memory()

    model = build_model(device)
    step_fn = TrainStep(model, device, world)

    def host_batch(i):
        k = i % n_pool
        return (ro_all[k, first:first + n_local], rd_all[k, first:first + n_local], g_host[k, first:first + n_local])

    # inputs resident in HBM for the `value` measurement
    dev_pool = [tuple(t.to(device, non_blocking=True) for t in host_batch(i)) for i in range(n_pool)]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sample_acc = torch.zeros(1, dtype=torch.int64, device=device)

    def run_steps(n, e2e, start_index):
        for s in range(n):
            i = start_index + s
            if e2e:
                ro, rd, G = (t.to(device, non_blocking=True) for t in host_batch(i))
            else:
                ro, rd, G = dev_pool[i % n_pool]
            loss = step_fn(ro, rd, G)
            sample_acc.add_(model.step_counter[(model.local_step - 1) % 16, 0].long())
            if e2e:
                loss.item()  # device -> host read of the step's result

    # ---- warm-up ------------------------------------------------------------------------------------
    run_steps(max(args.warmup, 3), False, 0)
    barrier()

    # ---- timed: inputs resident ----------------------------------------------------------------------
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    _cabi.PROFILE = {"ngp_grid_encode_forward": [], "ngp_grid_encode_backward": []}
    sample_acc.zero_()
    updates0 = step_fn.n_updates
    launches0 = _cabi.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    run_steps(args.steps, False, 1000)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _cabi.LAUNCHES - launches0
    n_updates = step_fn.n_updates - updates0
    prof = _cabi.PROFILE
    _cabi.PROFILE = None
    samples = int(sample_acc.item())
    kern = {}
    for name, evs in prof.items():
        tot = sum(a.elapsed_time(b) for a, b in evs)
        kern[name] = (tot, len(evs))

    # ---- timed: end to end (pinned host inputs, H2D inside, loss read back) -----------------------------
    sample_acc.zero_()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    f0.record()
    run_steps(args.steps, True, 2000)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    samples_e2e = int(sample_acc.item())
    clk = clocks.stop() if rank == 0 else None

    # ---- reduce over ranks: time = max, samples = sum ------------------------------------------------------
    stats = torch.tensor([ms, ms_e2e, float(samples), float(samples_e2e)], dtype=torch.float64, device=device)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, ms_e2e, samples, samples_e2e = mx[0].item(), mx[1].item(), sm[2].item(), sm[3].item()

    if rank == 0:
        value = samples / (ms * 1e-3)
        e2e_value = samples_e2e / (ms_e2e * 1e-3)
        h2d = sum(t.numel() * t.element_size() for t in host_batch(0)) * world
        # roofline of the dominant kernel (the grid-encode scatter or gather), against the measured L2 ceilings
        gathers_s, reds_s = measure_l2_peaks(device)
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
            os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        dom = max(kern, key=lambda k: kern[k][0])
        tot_ms, n_calls = kern[dom]
        # algorithmic L2 bytes per encoded point: 16 levels x 8 corners x 4 B (SURVEY 8d / BASELINE.md 4)
        # (rank 0's launches: its share of the marched samples, padded to 128, plus the 128^3 occupancy queries)
        local_points = samples / world + (2097152 * n_updates if "forward" in dom else 0)
        achieved = 512.0 * local_points / (tot_ms * 1e-3) / 1e9 if tot_ms > 0 else 0.0
        peak = (gathers_s if "forward" in dom else reds_s) * 4 / 1e9
        roofline = {
            "bound": "l2", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak if peak else None, "traffic": None,
            "avg_launch_ms": tot_ms / max(n_calls, 1), "launches": n_calls,
            "share_of_step": tot_ms / ms,
            "peak_source": "measured in this run: random 4-B gathers / 8-B red.add over a 32 MB L2-resident table "
                           "(ngp_bench_gather4 / ngp_bench_red8), counted at 4 algorithmic bytes per access",
            "hbm_peak_gbs": peaks.get("hbm_gbs"),
            "other_kernels_ms": {k: v[0] for k, v in kern.items()},
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {
                "workload": "-O train step (BASELINE configs[2]/[3]): network_grid tiledgrid 16x2 @2^16 + 64-wide MLP, "
                            "128^3 occupancy grid, %d views x 64x64 rays per step over %d GPU(s), max_steps 1024, "
                            "synthetic SDS grad + entropy backward, grad all-reduce, Adam + GradScaler, occupancy "
                            "update every 16 steps" % (args.views, world),
                "views_per_step": args.views, "rays_per_step": args.views * H * W,
                "samples_per_step": samples / args.steps, "timing": "inputs (3.5 MB/step) and the 7 MB table are "
                "smaller than L2 by nature of the workload; each step runs on a different view batch (64-batch pool), "
                "the 134+ MB/step of sample buffers exceed L2",
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * world,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches * world,
            "clocks": clk,
            "roofline": roofline,
        }
        if not args.no_cpu_baseline:
            try:
                v, cms, cores, desc = cpu_reference_run(args.cpu_sample_steps, 1)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
                                        "ms_per_step": cms}
            except Exception as e:  # pragma: no cover
                line["cpu_baseline"] = {"error": repr(e)}
        if not args.no_ref_cuda:
            try:
                from oracle import ref_pipeline
                line["ref_cuda_ext"] = ref_pipeline.time_reference_train_step(device, views=1, steps=20, warmup=5)
            except Exception as e:
                line["ref_cuda_ext"] = {"unavailable": repr(e)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
