#!/usr/bin/env python
"""TEST INFRASTRUCTURE: generates tests/golden/ref_golden.npz by running the REFERENCE's own CUDA
extensions (oracle/_ref, built from /root/reference by build_ref.py) on the seeded cases of
tests/ngp_testutil.py.  Must run on a GPU box:

    gpurun -- 'python oracle/make_golden.py gpurun_out/ref_golden.npz'

and the result is then committed as tests/golden/ref_golden.npz.  Only OUTPUTS of the reference are
stored; inputs are regenerated from their seeds by the tests.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ngp_testutil as util  # noqa: E402
from oracle import ref_ext  # noqa: E402


def main(out_path):
    import torch
    ns = ref_ext.load()
    assert ns is not None, "oracle/_ref is not built"
    dev = torch.device("cuda:0")
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    out = {}

    # ---- per-level scales exp2f(l*S)*H-1 as the DEVICE evaluates them (gridencoder.cu:125; MUFU.EX2 != libm) ----
    c0 = util.golden_grid_case(0, "f32")
    sys.path.insert(0, os.path.join(ROOT, "single-stable-dreamfusion_b200"))
    from ngp_b200 import _cabi
    sc = torch.empty(16, device=dev)
    rs = torch.empty(16, dtype=torch.int32, device=dev)
    _cabi.call("ngp_grid_level_params", dev, 16, float(c0["S"]), 16, _cabi.ptr(sc), _cabi.ptr(rs))
    out["grid_scales"] = sc.cpu().numpy()

    # ---- grid encoder: forward (+dy_dx) and backward, fp32 and fp16, hash and tiled -----------------
    for gridtype in (0, 1):
        for dt in ("f32", "f16"):
            c = util.golden_grid_case(gridtype, dt)
            key = "grid_g%d_%s_" % (gridtype, dt)
            x, emb, offs = T(c["x"]), T(c["emb"]), T(c["offs"])
            o, dydx, _ = ref_ext.grid_encode_forward(ns, x, emb, offs, float(c["S"]), c["H"], True, gridtype, False)
            out[key + "out"] = o.cpu().numpy()
            out[key + "dydx"] = dydx.cpu().numpy()
            ge, gi = ref_ext.grid_encode_backward(ns, T(c["grad"]), x, emb, offs, float(c["S"]), c["H"], dydx, gridtype, False)
            ge = ge.float().cpu().numpy()
            rows = np.random.default_rng(5).integers(0, ge.shape[0], 4096)
            out[key + "gemb_rows"] = ge[rows]
            out[key + "gemb_l1"] = np.array([np.abs(ge.astype(np.float64)).sum()])
            out[key + "ginp"] = gi.cpu().numpy()

    # ---- raymarching ------------------------------------------------------------------------------------
    for name, kw in (("m1", {}), ("m2", dict(cascade=2, bound=2.0, dt_gamma=1.0 / 128, max_steps=128, seed=21))):
        c = util.golden_march_case(**kw)
        ro, rd = T(c["rays_o"]), T(c["rays_d"])
        grid = T(c["grid"])
        bits = torch.empty(grid.numel() // 8, dtype=torch.uint8, device=dev)
        ns.march.packbits(grid, grid.numel() // 8, c["thresh"], bits)
        nears, fars = ref_ext.near_far_from_aabb(ns, ro, rd, T(c["aabb"]), 0.2)
        xyzs, dirs, deltas, rays, counter = ref_ext.march_rays_train(
            ns, ro, rd, c["bound"], bits, c["cascade"], 128, nears, fars, T(c["noises"]), c["dt_gamma"], c["max_steps"])
        counts, (cx, cd, cl) = ref_ext.canonical_rays(rays, xyzs, dirs, deltas)
        out[name + "_bits_sum"] = np.array([int(bits.long().sum().item())])
        out[name + "_bits_head"] = bits[:4096].cpu().numpy()
        out[name + "_nears"] = nears.cpu().numpy()
        out[name + "_fars"] = fars.cpu().numpy()
        out[name + "_counts"] = counts.int().cpu().numpy()
        out[name + "_counter"] = counter.cpu().numpy()
        out[name + "_xyzs"] = cx.cpu().numpy()
        out[name + "_deltas"] = cl.cpu().numpy()
        # composite on a deterministic pseudo-field, in canonical (ray-ordered) layout
        sig, rgb = util.pseudo_field(cx.cpu().numpy())
        N = ro.shape[0]
        cr = torch.stack([torch.arange(N, device=dev), torch.cumsum(counts, 0) - counts, counts], 1).int().contiguous()
        ws, depth, image = ref_ext.composite_rays_train_forward(ns, T(sig), T(rgb), cl.contiguous(), cr, 1e-4)
        out[name + "_ws"] = ws.cpu().numpy()
        out[name + "_depth"] = depth.cpu().numpy()
        out[name + "_image"] = image.cpu().numpy()
        rng = np.random.default_rng(33)
        gws = rng.standard_normal(N).astype(np.float32)
        gim = rng.standard_normal((N, 3)).astype(np.float32)
        gs, gc = ref_ext.composite_rays_train_backward(ns, T(gws), T(gim), T(sig), T(rgb), cl.contiguous(), cr, ws, image, 1e-4)
        out[name + "_gsig"] = gs.cpu().numpy()
        out[name + "_grgb"] = gc.cpu().numpy()
        # one inference marching call from the near plane, 4 steps per ray
        alive = torch.arange(N, dtype=torch.int32, device=dev)
        ix, idr, idl = ref_ext.march_rays(ns, N, 4, alive, nears.clone(), ro, rd, c["bound"], bits, c["cascade"], 128, nears,
                                          fars, torch.zeros(N, device=dev), c["dt_gamma"], c["max_steps"], 128)
        out[name + "_inf_xyzs"] = ix.cpu().numpy()
        out[name + "_inf_deltas"] = idl.cpu().numpy()

    # ---- morton / freq ---------------------------------------------------------------------------------------
    coords = np.random.default_rng(41).integers(0, 128, (512, 3)).astype(np.int32)
    idx = torch.empty(512, dtype=torch.int32, device=dev)
    ns.march.morton3D(T(coords), 512, idx)
    out["morton"] = idx.cpu().numpy()
    fx = np.random.default_rng(42).uniform(-1, 1, (128, 3)).astype(np.float32)
    fo = torch.empty(128, 39, device=dev)
    ns.freq.freq_encode_forward(T(fx), 128, 3, 6, 39, fo)
    out["freq_out"] = fo.cpu().numpy()
    fg = np.random.default_rng(43).standard_normal((128, 39)).astype(np.float32)
    fgi = torch.zeros(128, 3, device=dev)
    ns.freq.freq_encode_backward(T(fg), fo, 128, 3, 6, 39, fgi)
    out["freq_gin"] = fgi.cpu().numpy()

    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, "with", len(out), "arrays,", os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_golden.npz"))
