#!/usr/bin/env python
"""BASELINE configs[1] (GridEncoder standalone, 2^22 points, fp16 fwd+bwd) and configs[4] (800x800 inference frames) in
one short run; prints one JSON line per measurement.  Every part is guarded: a failure prints an error line and the
script moves on.   usage: python profiles/tools/cfg_bench.py [--frames 3]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for _p in (ROOT, os.path.join(ROOT, "single-stable-dreamfusion_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def cfg2():
    from gridencoder import GridEncoder
    torch.manual_seed(0)
    enc = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19,
                      desired_resolution=2048, gridtype="hash").to(dev)
    with torch.no_grad():
        enc.embeddings.uniform_(-1, 1)
    B = 1 << 22
    x = torch.rand(B, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(1)) * 2 - 1
    with torch.autocast("cuda", torch.float16):
        out = enc(x, bound=1)
    g = torch.randn(out.shape, device=dev, dtype=out.dtype, generator=torch.Generator(device=dev).manual_seed(2))

    def fwd():
        with torch.no_grad(), torch.autocast("cuda", torch.float16):
            enc(x, bound=1)

    def fwd_bwd():
        enc.embeddings.grad = None
        with torch.autocast("cuda", torch.float16):
            o = enc(x, bound=1)
        o.backward(g)

    t_f = timed(fwd, 5)
    t_fb = timed(fwd_bwd, 5)
    res = {"config": "cfg2 GridEncoder 16x2 hash 2^19 base16 res2048, 2^22 points, fp16 autocast", "points": B,
           "fwd_ms": t_f, "fwd_bwd_ms": t_fb, "fwd_points_per_s": B / (t_f * 1e-3), "fwd_bwd_points_per_s": B / (t_fb * 1e-3),
           "fwd_l2_gather_gbs_algorithmic": 512.0 * B / (t_f * 1e-3) / 1e9}
    try:  # the reference's own extension on the same tensors
        from oracle import ref_ext as R
        ns = R.load()
        if ns is not None:
            emb_h = enc.embeddings.detach().half()
            x01 = ((x + 1) / 2).contiguous()
            S = float(np.log2(enc.per_level_scale))
            t_rf = timed(lambda: R.grid_encode_forward(ns, x01, emb_h, enc.offsets, S, 16, False, 0, False), 3, 1)
            gb = g.half().contiguous()
            t_rb = timed(lambda: R.grid_encode_backward(ns, gb, x01, emb_h, enc.offsets, S, 16, None, 0, False), 3, 1)
            res.update(ref_ext_fwd_ms=t_rf, ref_ext_bwd_ms=t_rb, speedup_fwd=t_rf / t_f, speedup_fwd_bwd=(t_rf + t_rb) / t_fb)
    except Exception as e:  # noqa: BLE001
        res["ref_ext_error"] = repr(e)[:200]
    return res


def cfg5(frames):
    import bench as Bm
    from ngp_b200 import provider
    model = Bm.build_model(dev)
    with torch.autocast("cuda", torch.float16):
        for _ in range(4):
            model.update_extra_state()
    model.eval()
    views = provider.make_orbit_views(frames + 1, 800, 800)
    times, samples = [], 0
    for i, (ro, rd) in enumerate(views):
        ro_t = torch.from_numpy(ro).to(dev)[None]
        rd_t = torch.from_numpy(rd).to(dev)[None]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad(), torch.autocast("cuda", torch.float16):
            out = model.render(ro_t, rd_t, staged=True, perturb=False, bg_color=torch.ones(3, device=dev), max_steps=1024,
                               dt_gamma=0, ambient_ratio=1.0, shading="albedo")
        torch.cuda.synchronize()
        if i > 0:
            times.append(time.perf_counter() - t0)
        assert out["image"].shape == (1, 640000, 3) and torch.isfinite(out["image"]).all()
    ms = 1e3 * sum(times) / len(times)
    return {"config": "cfg5 inference 800x800 orbit frames (random-init scene, occupancy grid refreshed 4x), wall clock incl. the "
                      "host loop of march_rays / composite_rays", "frames": len(times), "ms_per_frame": ms,
            "frames_per_s": 1e3 / ms, "rays_per_s": 640000 / (ms * 1e-3)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    for name, fn in (("cfg2", cfg2), ("cfg5", lambda: cfg5(a.frames))):
        if a.only and a.only != name:
            continue
        try:
            print(json.dumps(fn()), flush=True)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"config": name, "error": repr(e)[:400]}), flush=True)
