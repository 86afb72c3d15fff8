#!/usr/bin/env python
"""Per-kernel timing of the train-step hot kernels on the real marched samples of the bench workload.

    python profiles/kbench.py [--views 8] [--iters 20] [--sweep]

Builds the bench model (random-init scene, one occupancy refresh), renders one 8-view batch through the fused
training path so that the capacity buffers hold real samples, then times each entry point alone with CUDA events
(20 launches, L2 flushed by a 256 MB memset between launches unless --no-flush).  --sweep also tries the forward
kernel's occupancy / carveout settings (ngp_field_set_option).  Prints one JSON line per measurement.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "single-stable-dreamfusion_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--red", action="store_true", help="red.global.add width / active-lane micro-benchmark")
    ap.add_argument("--no-flush", action="store_true")
    args = ap.parse_args()

    from ngp_b200 import _cabi, provider
    from ngp_b200.field import cached_half
    import bench as B
    dev = torch.device("cuda:0")
    lib = _cabi.load()
    model = B.build_model(dev)
    with torch.autocast("cuda", torch.float16):
        model.update_extra_state()
    ro, rd = provider.make_training_views(args.views, 64, 64, seed=0, pin=False)
    ro, rd = ro.view(args.views, 4096, 3).to(dev), rd.view(args.views, 4096, 3).to(dev)
    with torch.autocast("cuda", torch.float16):
        out = model.render(ro, rd, staged=False, perturb=True, force_all_rays=True, max_steps=1024, dt_gamma=0,
                           ambient_ratio=1.0, shading="albedo")
    out["image"].backward(torch.randn_like(out["image"]) * 1e-2)
    torch.cuda.synchronize()
    ws = model._train_ws
    M = int(ws.counter[0].item())
    N = ws.n_rays
    enc = model.encoder
    L, S = 16, float(np.log2(enc.per_level_scale))
    table = cached_half(enc.embeddings)
    l0, l1, l2 = model.sigma_net.net
    hw = [cached_half(t) for t in (l0.weight, l0.bias, l1.weight, l1.bias, l2.weight, l2.bias)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    p = _cabi.ptr
    grad_table = torch.zeros_like(enc.embeddings)
    grad_odd = torch.zeros(grad_table.numel() + 4, device=dev)[2:2 + grad_table.numel()]
    gflat = torch.zeros(64 * 32 + 64 + 64 * 64 + 64 + 256 + 4, device=dev)
    sizes = [64 * 32, 64, 64 * 64, 64, 4 * 64, 4]
    gw = torch.split(gflat, sizes)
    wsum = torch.empty(N, device=dev)
    depth = torch.empty(N, device=dev)
    image = torch.empty(N, 3, device=dev)
    g_ws = torch.randn(N, device=dev) * 1e-3
    g_img = torch.randn(N, 3, device=dev) * 1e-2
    ro_f, rd_f = ro.reshape(-1, 3).contiguous(), rd.reshape(-1, 3).contiguous()
    import raymarching
    nears, fars = raymarching.near_far_from_aabb(ro_f, rd_f, model.aabb_train)
    noises = torch.rand(N, device=dev)

    quads = torch.zeros(table.shape[0], 4, dtype=torch.int32, device=dev)
    _cabi.call("ngp_grid_quad_table", dev, p(table), p(enc.offsets), L, table.shape[0], S, int(enc.base_resolution),
               int(enc.gridtype_id), int(bool(enc.align_corners)), p(quads))
    calls = {
        "ngp_field_forward_quads": lambda: _cabi.call(
            "ngp_field_forward_quads", dev, p(ws.xyzs), ws.cap, p(ws.counter), p(table), p(quads), p(enc.offsets), L, 2, S,
            int(enc.base_resolution), int(enc.gridtype_id), int(bool(enc.align_corners)), float(model.bound),
            *[p(t) for t in hw], 64, 4, p(ws.sigma), p(ws.rgb), p(ws.enc), p(ws.h1), p(ws.h2)),
        "ngp_grid_quad_table": lambda: _cabi.call(
            "ngp_grid_quad_table", dev, p(table), p(enc.offsets), L, table.shape[0], S, int(enc.base_resolution),
            int(enc.gridtype_id), int(bool(enc.align_corners)), p(quads)),
        "ngp_field_forward": lambda: _cabi.call(
            "ngp_field_forward", dev, p(ws.xyzs), ws.cap, p(ws.counter), p(table), p(enc.offsets), L, 2, S,
            int(enc.base_resolution), int(enc.gridtype_id), int(bool(enc.align_corners)), float(model.bound),
            *[p(t) for t in hw], 64, 4, p(ws.sigma), p(ws.rgb), p(ws.enc), p(ws.h1), p(ws.h2)),
        "ngp_field_backward": lambda: _cabi.call(
            "ngp_field_backward", dev, ws.cap, p(ws.counter), p(hw[0]), p(hw[2]), p(hw[4]), 64, 4, p(ws.d_sigma),
            p(ws.d_rgb), p(ws.sigma), p(ws.rgb), p(ws.enc), p(ws.h1), p(ws.h2), p(ws.d_enc), *[p(t) for t in gw]),
        "ngp_grid_scatter_samples": lambda: _cabi.call(
            "ngp_grid_scatter_samples", dev, p(ws.d_enc), p(ws.xyzs), float(model.bound), p(ws.counter), ws.cap,
            p(enc.offsets), L, 2, S, int(enc.base_resolution), int(enc.gridtype_id), int(bool(enc.align_corners)),
            p(grad_table)),
        "ngp_grid_scatter_samples_split": lambda: _cabi.call(
            "ngp_grid_scatter_samples_split", dev, p(ws.d_enc), p(ws.xyzs), float(model.bound), p(ws.counter), ws.cap,
            p(enc.offsets), L, 2, S, int(enc.base_resolution), int(enc.gridtype_id), int(bool(enc.align_corners)),
            p(grad_table), p(grad_odd)),
        "ngp_grid_fold_odd": lambda: _cabi.call("ngp_grid_fold_odd", dev, p(grad_table), p(grad_odd), grad_table.numel()),
        "ngp_train_ray_loss": lambda: _cabi.call(
            "ngp_train_ray_loss", dev, p(ws.sigma), p(ws.rgb), p(ws.deltas), p(ws.rays), ws.cap, N, 1e-4, None, 1.0, p(g_img), 0,
            0, N, 1e-4, None, p(wsum), p(depth), p(image), None, p(ws.d_sigma), p(ws.d_rgb), None, None, None, None, None),
        "ngp_composite_rays_train_forward": lambda: _cabi.call(
            "ngp_composite_rays_train_forward", dev, p(ws.sigma), p(ws.rgb), p(ws.deltas), p(ws.rays), ws.cap, N, 1e-4,
            p(wsum), p(depth), p(image)),
        "ngp_composite_rays_train_backward": lambda: _cabi.call(
            "ngp_composite_rays_train_backward", dev, p(g_ws), p(g_img), p(ws.sigma), p(ws.rgb), p(ws.deltas), p(ws.rays),
            p(wsum), p(image), ws.cap, N, 1e-4, p(ws.d_sigma), p(ws.d_rgb)),
    }

    def march():
        ws.counter.zero_()
        _cabi.call("ngp_march_rays_train", dev, p(ro_f), p(rd_f), p(model.density_bitfield), float(model.bound), 0.0, 1024, N,
                   int(model.cascade), int(model.grid_size), ws.cap, p(nears), p(fars), p(ws.xyzs), None, p(ws.deltas),
                   p(ws.rays), p(ws.counter), p(noises), p(ws.march_ws), ws.march_ws.numel())

    def march_packed():
        ws.counter.zero_()
        _cabi.call("ngp_march_rays_train_packed", dev, p(ro_f), p(rd_f), p(model.density_bitfield), float(model.bound), 0.0, 1024,
                   N, int(model.cascade), int(model.grid_size), ws.cap, p(nears), p(fars), p(ws.xyzs), None, p(ws.deltas),
                   p(ws.rays), p(ws.counter), p(noises))
    calls["ngp_march_rays_train_packed"] = march_packed

    def time_call(fn, iters):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            if not args.no_flush:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2], ts[0]

    def report(name, extra=None):
        med, best = time_call(calls[name] if name in calls else march, args.iters)
        row = {"kernel": name, "samples": M, "rays": N, "median_ms": round(med, 4), "min_ms": round(best, 4),
               "Gsamples_per_s": round(M / med / 1e6, 3)}
        if extra:
            row.update(extra)
        print(json.dumps(row), flush=True)
        return med

    for name in calls:
        report(name)
    for cap in (4, 8, 16, 32):
        assert lib.ngp_grid_set_option(2, cap) == 0
        report("ngp_grid_scatter_samples_split", {"max_run": cap})
    assert lib.ngp_grid_set_option(2, 8) == 0
    for min_rays in (0, 1):
        assert lib.ngp_march_set_option(2, min_rays) == 0
        report("ngp_march_rays_train", {"walk": "thread per ray, closed-form jumps" if min_rays else "warp per ray"})
    assert lib.ngp_march_set_option(2, 0) == 0
    if args.red:
        words = 1 << 23
        ftable = torch.zeros(words, dtype=torch.float32, device=dev)
        n_threads, iters = 148 * 2048 * 4, 16
        for width in (1, 2, 4):
            for lane_stride in (1, 2, 4):
                best = 1e9
                for _ in range(4):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    _cabi.call("ngp_bench_red_width", dev, p(ftable), words, n_threads, iters, 3, width, lane_stride)
                    e1.record()
                    torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1))
                ops = n_threads // lane_stride * iters * 8
                print(json.dumps({"microbench": "red", "floats_per_op": width, "active_lanes": 32 // lane_stride,
                                  "Gops_per_s": round(ops / best / 1e6, 1), "GB_per_s": round(ops * width * 4 / best / 1e6, 1)}), flush=True)
    if args.sweep:
        for groups, ctas in ((1, 7), (1, 5), (2, 4), (2, 3), (4, 2)):
            for carve in (-1, 100):
                assert lib.ngp_field_set_option(2, groups) == 0 and lib.ngp_field_set_option(0, ctas) == 0
                assert lib.ngp_field_set_option(1, carve) == 0
                report("ngp_field_forward", {"groups": groups, "ctas_per_sm": ctas, "carveout": carve})
                report("ngp_field_forward_quads", {"groups": groups, "ctas_per_sm": ctas, "carveout": carve})
        assert lib.ngp_field_set_option(2, 2) == 0 and lib.ngp_field_set_option(0, 0) == 0 and lib.ngp_field_set_option(1, -1) == 0


if __name__ == "__main__":
    main()
