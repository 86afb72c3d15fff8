"""GPU parity tests (run with -m gpu on a B200).  Every op goes through the C ABI (ctypes) and is compared

  (a) with the CPU oracle (oracle/ngp_oracle.c) on the same seeded inputs, and
  (b) with the reference's own CUDA extensions (oracle/_ref), which also pins the oracle.

Tolerances (BASELINE.json north_star): integer / byte / index outputs and sample positions bit-exact;
fp32 encodings and composited outputs <= 1e-5 relative (we assert bit-equality where the evaluation
order is mirrored); fp16 encodings <= 1e-3 relative (asserted bit-equal); gradients <= 1e-3 relative.
"""
import numpy as np
import pytest
import torch

import ngp_testutil as util
from oracle import oracle as O
from oracle import ref_ext as R

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def N_(t):
    return t.detach().cpu().numpy()


def cabi():
    from ngp_b200 import _cabi
    return _cabi


def device_scales(L, S, H):
    c = cabi()
    sc = torch.empty(L, device=DEV)
    rs = torch.empty(L, dtype=torch.int32, device=DEV)
    c.call("ngp_grid_level_params", sc.device, L, float(S), H, c.ptr(sc), c.ptr(rs))
    return N_(sc), N_(rs).astype(np.uint32)


def my_grid_forward(x, emb, offs, S, H, gridtype, dydx=False, layout=1, align=False):
    c = cabi()
    B, D = x.shape
    L = offs.shape[0] - 1
    C = emb.shape[1]
    out = torch.empty((B, L * C) if layout == 1 else (L, B, C), dtype=emb.dtype, device=DEV)
    j = torch.empty(B, L * D * C, dtype=emb.dtype, device=DEV) if dydx else None
    c.call("ngp_grid_encode_forward", x.device, c.ptr(x), c.ptr(emb), c.ptr(offs), c.ptr(out), B, D, C, L, float(S), H,
           c.ptr(j), gridtype, int(align), c.dtype_code(emb.dtype), layout)
    return out, j


def my_grid_backward(grad, x, offs, n_rows, C, S, H, gridtype, gt_dtype=torch.float32, dydx=None, layout=1):
    c = cabi()
    B, D = x.shape
    L = offs.shape[0] - 1
    ge = torch.zeros(n_rows, C, dtype=gt_dtype, device=DEV)
    gi = torch.zeros(B, D, dtype=grad.dtype, device=DEV) if dydx is not None else None
    c.call("ngp_grid_encode_backward", x.device, c.ptr(grad), c.ptr(x), None, c.ptr(offs), c.ptr(ge), B, D, C, L, float(S),
           H, c.ptr(dydx), c.ptr(gi), gridtype, 0, c.dtype_code(grad.dtype), layout, c.dtype_code(gt_dtype))
    return ge, gi


# ---------------------------------------------------------------------------------------------------------
# grid encoder
# ---------------------------------------------------------------------------------------------------------
def test_device_level_scales_vs_libm():
    offs, S = util.make_offsets()
    dev_sc, dev_rs = device_scales(16, np.float32(S), 16)
    cpu_sc, cpu_rs = O.grid_level_params(16, np.float32(S), 16)
    assert np.array_equal(dev_rs, cpu_rs)                    # resolutions decide every index
    ulp = np.abs(dev_sc.view(np.int32).astype(np.int64) - cpu_sc.view(np.int32).astype(np.int64))
    assert ulp.max() <= 2, ulp                               # MUFU.EX2 vs libm exp2f
    assert dev_sc[0] == 15.0 and dev_sc[15] == 2047.0


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
@pytest.mark.parametrize("gridtype,log2", [(0, 19), (1, 16)])
def test_grid_forward_vs_oracle(dtype, gridtype, log2):
    rng = np.random.default_rng(100 + gridtype)
    offs, S = util.make_offsets(log2_hashmap_size=log2)
    emb = rng.uniform(-1, 1, (offs[-1], 2)).astype(dtype)
    x = rng.uniform(0, 1, (20000, 3)).astype(np.float32)
    x[:3] = [[0, 0, 0], [1, 1, 1], [0.5, 1.0000001, 0.5]]
    dev_sc, _ = device_scales(16, np.float32(S), 16)
    out, j = my_grid_forward(T(x), T(emb), T(offs), np.float32(S), 16, gridtype, dydx=True)
    want, wj = O.grid_encode_forward(x, emb, offs, np.float32(S), 16, calc_dydx=True, gridtype=gridtype,
                                     scale_override=dev_sc)
    got = N_(out)
    assert np.array_equal(got.view(np.uint16 if dtype == np.float16 else np.uint32),
                          want.view(np.uint16 if dtype == np.float16 else np.uint32))
    # dy_dx: same evaluation order -> bit equal too
    assert np.array_equal(N_(j), wj)
    # [L,B,C] layout is the same numbers
    out_lbc, _ = my_grid_forward(T(x), T(emb), T(offs), np.float32(S), 16, gridtype, layout=0)
    assert torch.equal(out_lbc.permute(1, 0, 2).reshape(20000, -1), out)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("gridtype,log2", [(0, 19), (1, 16)])
def test_grid_forward_backward_vs_reference_ext(ref_ext, dtype, gridtype, log2):
    g = torch.Generator(device="cpu").manual_seed(7 + gridtype)
    offs_np, S = util.make_offsets(log2_hashmap_size=log2)
    offs = T(offs_np)
    emb = (torch.rand(int(offs_np[-1]), 2, generator=g) * 2 - 1).to(DEV).to(dtype)
    B = 1 << 18
    x = torch.rand(B, 3, generator=g).to(DEV)
    ref_out, ref_j, _ = R.grid_encode_forward(ref_ext, x, emb, offs, float(np.float32(S)), 16, True, gridtype, False)
    out, j = my_grid_forward(x, emb, offs, np.float32(S), 16, gridtype, dydx=True)
    assert torch.equal(out, ref_out), (out.float() - ref_out.float()).abs().max().item()
    assert torch.equal(j, ref_j)

    grad = torch.randn(B, 32, generator=g).to(DEV).to(dtype)
    ref_ge, ref_gi = R.grid_encode_backward(ref_ext, grad, x, emb, offs, float(np.float32(S)), 16, ref_j, gridtype, False)
    ge, gi = my_grid_backward(grad, x, offs, emb.shape[0], 2, np.float32(S), 16, gridtype, torch.float32, dydx=j)
    assert torch.equal(gi, ref_gi.to(gi.dtype))
    rel = util.rel_l2(N_(ge), N_(ref_ge.float()))
    # fp32: only the atomic order differs.  fp16: the reference ACCUMULATES in half (gridencoder.cu:298-304)
    assert rel < (1e-5 if dtype == torch.float32 else 2e-3), rel
    # same-dtype table through the C ABI (the strict drop-in mode) behaves like the reference's
    ge_same, _ = my_grid_backward(grad, x, offs, emb.shape[0], 2, np.float32(S), 16, gridtype, dtype)
    rel2 = util.rel_l2(N_(ge_same.float()), N_(ref_ge.float()))
    assert rel2 < (1e-5 if dtype == torch.float32 else 2e-3), rel2


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
def test_grid_backward_vs_exact_sum(dtype, ref_ext):
    """fp32-accumulated gradients vs the oracle's exact (float64) sum; also shows we are no worse than the reference."""
    rng = np.random.default_rng(5)
    offs, S = util.make_offsets(log2_hashmap_size=16)
    B = 30000
    x = rng.uniform(0, 1, (B, 3)).astype(np.float32)
    g = rng.standard_normal((B, 32)).astype(dtype)
    dev_sc, _ = device_scales(16, np.float32(S), 16)
    truth = O.grid_encode_backward(g, x, offs, offs[-1], 2, np.float32(S), 16, gridtype=1, scale_override=dev_sc)
    ge, _ = my_grid_backward(T(g), T(x), T(offs), int(offs[-1]), 2, np.float32(S), 16, 1, torch.float32)
    mine = util.rel_l2(N_(ge), truth)
    assert mine < 1e-6, mine
    emb = torch.zeros(int(offs[-1]), 2, device=DEV, dtype=torch.float16 if dtype == np.float16 else torch.float32)
    ref_ge, _ = R.grid_encode_backward(ref_ext, T(g), T(x), emb, T(offs), float(np.float32(S)), 16, None, 1, False)
    theirs = util.rel_l2(N_(ref_ge.float()), truth)
    # fp32 gradients: both sit at fp32 rounding noise (~6e-8), where the order of the last few additions decides the
    # digit; fp16 gradients: the reference accumulates in half and is orders of magnitude worse
    assert mine <= theirs * (1.25 if dtype == np.float32 else 1.01) + 1e-9, (mine, theirs)


def test_grid_other_dims_and_channels_vs_oracle():
    """The reference instantiates D in 1..5 x C in {1,2,4,8}; spot-check the non-hot combinations."""
    rng = np.random.default_rng(6)
    for D, C, dtype in [(2, 2, np.float16), (2, 4, np.float32), (3, 1, np.float32), (3, 8, np.float16), (3, 4, np.float16),
                        (4, 2, np.float32), (1, 2, np.float32), (5, 1, np.float32)]:
        offs, S = util.make_offsets(num_levels=6, desired_resolution=96, log2_hashmap_size=12, input_dim=D)
        emb = rng.uniform(-1, 1, (offs[-1], C)).astype(dtype)
        x = rng.uniform(0, 1, (3000, D)).astype(np.float32)
        dev_sc, _ = device_scales(6, np.float32(S), 16)
        for gridtype in (0, 1):
            out, _ = my_grid_forward(T(x), T(emb), T(offs), np.float32(S), 16, gridtype)
            want, _ = O.grid_encode_forward(x, emb, offs, np.float32(S), 16, gridtype=gridtype, scale_override=dev_sc)
            assert np.array_equal(N_(out), want), (D, C, dtype, gridtype)
            if D == 3 and C in (2, 4):
                # align_corners (grid.py:91): the +1 corner of x = 1 lands on `resolution` and must wrap like the reference
                xa = x.copy(); xa[:5] = 1.0; xa[5:9] = 0.0
                offs_a, S_a = util.make_offsets(num_levels=6, desired_resolution=96, log2_hashmap_size=12, input_dim=D,
                                                align_corners=True)
                emb_a = rng.uniform(-1, 1, (offs_a[-1], C)).astype(dtype)
                out_a, _ = my_grid_forward(T(xa), T(emb_a), T(offs_a), np.float32(S_a), 16, gridtype, align=True)
                want_a, _ = O.grid_encode_forward(xa, emb_a, offs_a, np.float32(S_a), 16, gridtype=gridtype, align_corners=True,
                                                  scale_override=dev_sc)
                assert np.array_equal(N_(out_a), want_a), ("align_corners", D, C, dtype, gridtype)
            g = rng.standard_normal((3000, 6 * C)).astype(dtype)
            ge, _ = my_grid_backward(T(g), T(x), T(offs), int(offs[-1]), C, np.float32(S), 16, gridtype)
            truth = O.grid_encode_backward(g, x, offs, offs[-1], C, np.float32(S), 16, gridtype=gridtype,
                                           scale_override=dev_sc)
            assert util.rel_l2(N_(ge), truth) < 1e-6, (D, C, dtype, gridtype)


def test_grid_unsupported_arguments_are_reported():
    c = cabi()
    lib = c.load()
    x = torch.zeros(8, 3, device=DEV)
    emb = torch.zeros(64, 3, device=DEV)   # C = 3 is not instantiated by the reference either
    offs = torch.tensor([0, 64], dtype=torch.int32, device=DEV)
    out = torch.zeros(8, 3, device=DEV)
    rc = lib.ngp_grid_encode_forward(c.ptr(x), c.ptr(emb), c.ptr(offs), c.ptr(out), 8, 3, 3, 1, 0.0, 16, None, 0, 0, 0, 1, None)
    assert rc == -2
    rc = lib.ngp_grid_encode_forward(None, c.ptr(emb), c.ptr(offs), c.ptr(out), 8, 3, 2, 1, 0.0, 16, None, 0, 0, 0, 1, None)
    assert rc == -1
    with pytest.raises(RuntimeError):
        c.check(rc, "ngp_grid_encode_forward")


def test_grid_encoder_module_autocast_and_autograd(ref_ext):
    from gridencoder import GridEncoder
    torch.manual_seed(0)
    enc = GridEncoder(num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=16, desired_resolution=2048,
                      gridtype='tiled').to(DEV)
    enc.embeddings.data.uniform_(-1, 1)
    x = torch.rand(5000, 3, device=DEV) * 2 - 1
    assert enc.offsets[-1].item() == 903480 and enc.output_dim == 32
    with torch.autocast('cuda', torch.float16):
        out = enc(x, bound=1)
    assert out.dtype == torch.float16 and out.shape == (5000, 32)
    gout = torch.randn_like(out)
    out.backward(gout)
    assert enc.embeddings.grad.dtype == torch.float32 and enc.embeddings.grad.shape == enc.embeddings.shape
    # reference call path (grid.py:138-154 + :22-84) on the same inputs
    x01 = (x + 1) / 2
    ref_out, _, _ = R.grid_encode_forward(ref_ext, x01, enc.embeddings.detach().half(), enc.offsets,
                                          float(np.log2(enc.per_level_scale)), 16, False, 1, False)
    assert torch.equal(out, ref_out)
    ref_ge, _ = R.grid_encode_backward(ref_ext, gout, x01, enc.embeddings.detach().half(), enc.offsets,
                                       float(np.log2(enc.per_level_scale)), 16, None, 1, False)
    assert util.rel_l2(N_(enc.embeddings.grad), N_(ref_ge.float())) < 1e-3
    out32 = enc(x, bound=1)
    assert out32.dtype == torch.float32
    assert (out32 - out.float()).abs().max() < 5e-3


# ---------------------------------------------------------------------------------------------------------
# raymarching: utils
# ---------------------------------------------------------------------------------------------------------
def test_near_far_morton_packbits_vs_oracle_and_reference(ref_ext):
    import raymarching
    rays_o, rays_d = util.look_at_rays(64, radius=1.4)
    rays_d[5] = [0, 0, 1]                       # axis-aligned ray: 1/0 = inf inside the slab test
    rays_o[9] = [3, 3, 3]; rays_d[9] = [1, 0, 0]  # misses
    aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        n0, f0 = O.near_far_from_aabb(rays_o, rays_d, aabb, 0.2)
    n1, f1 = raymarching.near_far_from_aabb(T(rays_o), T(rays_d), T(aabb), 0.2)
    assert np.array_equal(N_(n1), n0) and np.array_equal(N_(f1), f0)
    n2, f2 = R.near_far_from_aabb(ref_ext, T(rays_o), T(rays_d), T(aabb), 0.2)
    assert torch.equal(n1, n2) and torch.equal(f1, f2)

    coords = np.random.default_rng(0).integers(0, 128, (100000, 3)).astype(np.int32)
    m = raymarching.morton3D(T(coords))
    assert np.array_equal(N_(m), O.morton3D(coords))
    assert np.array_equal(N_(raymarching.morton3D_invert(m)), coords)
    mr = torch.empty_like(m)
    ref_ext.march.morton3D(T(coords), coords.shape[0], mr)
    assert torch.equal(m, mr)

    grid = util.blob_density_grid(2, 128, 2.0, 3)
    bits = raymarching.packbits(T(grid), 10.0)
    assert bits.dtype == torch.uint8 and bits.shape[0] == 2 * 128 ** 3 // 8
    assert np.array_equal(N_(bits), O.packbits(grid, 10.0))
    br = torch.empty_like(bits)
    ref_ext.march.packbits(T(grid), bits.shape[0], 10.0, br)
    assert torch.equal(bits, br)

    sph = raymarching.sph_from_ray(T(rays_o), T(rays_d), 1.6)
    sr = torch.empty_like(sph)
    ref_ext.march.sph_from_ray(T(rays_o), T(rays_d), 1.6, rays_o.shape[0], sr)
    assert torch.allclose(sph, sr, rtol=1e-5, atol=1e-6, equal_nan=True)


# ---------------------------------------------------------------------------------------------------------
# raymarching: training march + composite
# ---------------------------------------------------------------------------------------------------------
MARCH_CASES = [
    dict(side=64, cascade=1, bound=1.0, dt_gamma=0.0, max_steps=1024),
    dict(side=48, cascade=1, bound=1.0, dt_gamma=0.0, max_steps=64),          # rays hit the max_steps cap
    dict(side=48, cascade=2, bound=2.0, dt_gamma=1.0 / 128, max_steps=512),   # cascades + adaptive step
    dict(side=32, cascade=3, bound=4.0, dt_gamma=1.0 / 256, max_steps=256),
]


def _march_setup(case, seed=0):
    rays_o, rays_d = util.look_at_rays(case["side"], radius=1.3 * case["bound"], theta_deg=60 + 7 * seed, phi_deg=40 * seed)
    grid = util.blob_density_grid(case["cascade"], 128, case["bound"], seed)
    bits = O.packbits(grid, 10.0)
    aabb = np.array([-case["bound"]] * 3 + [case["bound"]] * 3, np.float32)
    nears, fars = O.near_far_from_aabb(rays_o, rays_d, aabb, 0.2)
    noises = np.random.default_rng(seed + 50).random(rays_o.shape[0]).astype(np.float32)
    return rays_o, rays_d, bits, nears, fars, noises


def _my_march(case, rays_o, rays_d, bits, nears, fars, noises, M=None):
    c = cabi()
    lib = c.load()
    N = rays_o.shape[0]
    M = N * case["max_steps"] if M is None else M
    xyzs = torch.zeros(M, 3, device=DEV); dirs = torch.zeros(M, 3, device=DEV); deltas = torch.zeros(M, 2, device=DEV)
    rays = torch.empty(N, 3, dtype=torch.int32, device=DEV)
    counter = torch.zeros(2, dtype=torch.int32, device=DEV)
    ws = torch.empty(int(lib.ngp_march_rays_train_workspace(N, int(case["max_steps"]))), dtype=torch.uint8, device=DEV)
    ws[:256].zero_()   # a new workspace's head must be zero (include/ngp_b200.h)
    # keep every input tensor alive in a local: c.ptr() only returns an integer address
    ro, rd, bf, ne, fa, nz = T(rays_o), T(rays_d), T(bits), T(nears), T(fars), T(noises)
    c.call("ngp_march_rays_train", xyzs.device, c.ptr(ro), c.ptr(rd), c.ptr(bf), float(case["bound"]),
           float(case["dt_gamma"]), case["max_steps"], N, case["cascade"], 128, M, c.ptr(ne), c.ptr(fa),
           c.ptr(xyzs), c.ptr(dirs), c.ptr(deltas), c.ptr(rays), c.ptr(counter), c.ptr(nz), c.ptr(ws), ws.numel())
    torch.cuda.synchronize()
    return xyzs, dirs, deltas, rays, counter


@pytest.mark.parametrize("case", MARCH_CASES)
def test_march_rays_train_bit_exact_vs_oracle_and_reference(case, ref_ext):
    rays_o, rays_d, bits, nears, fars, noises = _march_setup(case, seed=1)
    xyzs, dirs, deltas, rays, counter = _my_march(case, rays_o, rays_d, bits, nears, fars, noises)
    ox, od, ol, orays, ocounter = O.march_rays_train(rays_o, rays_d, case["bound"], bits, case["cascade"], 128, nears, fars,
                                                     noises, case["dt_gamma"], case["max_steps"])
    total = int(ocounter[0])
    assert total > 0
    assert np.array_equal(N_(counter), ocounter)
    assert np.array_equal(N_(rays), orays)
    assert np.array_equal(N_(xyzs[:total]).view(np.uint32), ox[:total].view(np.uint32))
    assert np.array_equal(N_(deltas[:total]).view(np.uint32), ol[:total].view(np.uint32))
    assert np.array_equal(N_(dirs[:total]), od[:total])
    assert not xyzs[total:].any() and not deltas[total:].any()   # nothing written past the total

    # the reference's own kernel: same per-ray counts and, per ray, the same samples (its row order is arbitrary)
    rx, rd, rl, rrays, rcounter = R.march_rays_train(ref_ext, T(rays_o), T(rays_d), case["bound"], T(bits), case["cascade"],
                                                     128, T(nears), T(fars), T(noises), case["dt_gamma"], case["max_steps"])
    assert torch.equal(rcounter, counter)
    counts, (cx, cdirs, cl) = R.canonical_rays(rrays, rx, rd, rl)
    assert torch.equal(counts.int(), rays[:, 2])
    assert torch.equal(cx, xyzs[:total]) and torch.equal(cl, deltas[:total]) and torch.equal(cdirs, dirs[:total])


def _my_march_packed(case, rays_o, rays_d, bits, nears, fars, noises, M=None, start=0):
    c = cabi()
    N = rays_o.shape[0]
    M = N * case["max_steps"] if M is None else M
    xyzs = torch.zeros(M, 3, device=DEV); dirs = torch.zeros(M, 3, device=DEV); deltas = torch.zeros(M, 2, device=DEV)
    rays = torch.empty(N, 3, dtype=torch.int32, device=DEV)
    counter = torch.tensor([start, 0], dtype=torch.int32, device=DEV)
    ro, rd, bf, ne, fa, nz = T(rays_o), T(rays_d), T(bits), T(nears), T(fars), T(noises)
    c.call("ngp_march_rays_train_packed", xyzs.device, c.ptr(ro), c.ptr(rd), c.ptr(bf), float(case["bound"]),
           float(case["dt_gamma"]), case["max_steps"], N, case["cascade"], 128, M, c.ptr(ne), c.ptr(fa),
           c.ptr(xyzs), c.ptr(dirs), c.ptr(deltas), c.ptr(rays), c.ptr(counter), c.ptr(nz))
    torch.cuda.synchronize()
    return xyzs, dirs, deltas, rays, counter


@pytest.mark.parametrize("case", MARCH_CASES)
def test_march_rays_train_packed_same_samples_per_ray(case):
    """The one-launch marcher of the hand-scheduled step (rows claimed with the reference's atomicAdd, raymarching.cu:405-406):
    same counts and, per ray, bit-identical samples as the oracle; the row blocks tile [start, start + total) exactly."""
    rays_o, rays_d, bits, nears, fars, noises = _march_setup(case, seed=1)
    ox, od, ol, orays, ocounter = O.march_rays_train(rays_o, rays_d, case["bound"], bits, case["cascade"], 128, nears, fars,
                                                     noises, case["dt_gamma"], case["max_steps"])
    total, N = int(ocounter[0]), rays_o.shape[0]
    start = 384
    xyzs, dirs, deltas, rays, counter = _my_march_packed(case, rays_o, rays_d, bits, nears, fars, noises,
                                                         M=N * case["max_steps"] + start, start=start)
    r = N_(rays)
    assert counter[0].item() == start + total and counter[1].item() == N
    assert np.array_equal(r[:, 0], np.arange(N)) and np.array_equal(r[:, 2], orays[:, 2])
    live = r[:, 2] > 0
    order = np.argsort(r[live, 1], kind="stable")
    offs, cnts = r[live, 1][order], r[live, 2][order]
    assert offs[0] == start and np.array_equal(offs[1:], offs[:-1] + cnts[:-1]) and offs[-1] + cnts[-1] == start + total
    # gather every ray's block into ray order and compare bit for bit with the oracle's ray-ordered rows
    idx = np.concatenate([np.arange(o, o + c) for o, c in zip(r[live, 1], r[live, 2])])
    gx, gl, gd = N_(xyzs)[idx], N_(deltas)[idx], N_(dirs)[idx]
    assert np.array_equal(gx.view(np.uint32), ox[:total].view(np.uint32))
    assert np.array_equal(gl.view(np.uint32), ol[:total].view(np.uint32))
    assert np.array_equal(gd, od[:total])
    assert not xyzs[:start].any() and not xyzs[start + total:].any()   # nothing written outside the claimed rows


def test_march_rays_train_packed_overflow_drops_whole_rays():
    """Rays whose claimed block does not fit in M rows are recorded but write nothing (raymarching.cu:415-416)."""
    case = MARCH_CASES[0]
    rays_o, rays_d, bits, nears, fars, noises = _march_setup(case, seed=2)
    full = O.march_rays_train(rays_o, rays_d, 1.0, bits, 1, 128, nears, fars, noises, 0.0, 1024)
    total = int(full[4][0])
    M = total // 3
    xyzs, dirs, deltas, rays, counter = _my_march_packed(case, rays_o, rays_d, bits, nears, fars, noises, M=M)
    r = N_(rays)
    assert counter[0].item() == total and np.array_equal(r[:, 2], full[3][:, 2])
    fits = (r[:, 2] > 0) & (r[:, 1] + r[:, 2] <= M)
    written = np.zeros(M, bool)
    for o, c in r[fits][:, 1:]:
        written[o:o + c] = True
    assert fits.any() and np.array_equal(N_(xyzs).any(axis=1) | written, written)


def test_march_rays_train_overflow_and_python_wrapper():
    import raymarching
    case = MARCH_CASES[0]
    rays_o, rays_d, bits, nears, fars, noises = _march_setup(case, seed=2)
    full = O.march_rays_train(rays_o, rays_d, 1.0, bits, 1, 128, nears, fars, noises, 0.0, 1024)
    total = int(full[4][0])
    M = total // 3
    xyzs, dirs, deltas, rays, counter = _my_march(case, rays_o, rays_d, bits, nears, fars, noises, M=M)
    ox, od, ol, orays, ocounter = O.march_rays_train(rays_o, rays_d, 1.0, bits, 1, 128, nears, fars, noises, 0.0, 1024, M=M)
    assert np.array_equal(N_(rays), orays) and np.array_equal(N_(counter), ocounter)
    assert np.array_equal(N_(xyzs), ox) and np.array_equal(N_(deltas), ol)

    # python surface: shapes, padding rule m += 128 - m % 128 (raymarching.py:225-226), zero pad rows
    torch.manual_seed(3)
    counter = torch.zeros(2, dtype=torch.int32, device=DEV)
    x, d, dl, r = raymarching.march_rays_train(T(rays_o), T(rays_d), 1.0, T(bits), 1, 128, T(nears), T(fars), counter, -1,
                                               False, 128, True, 0, 1024)
    zero = O.march_rays_train(rays_o, rays_d, 1.0, bits, 1, 128, nears, fars, np.zeros_like(noises), 0.0, 1024)
    tot = int(zero[4][0])
    m = tot + 128 - tot % 128
    assert x.shape == (m, 3) and d.shape == (m, 3) and dl.shape == (m, 2) and r.shape == (rays_o.shape[0], 3)
    assert counter[0].item() == tot and counter[1].item() == rays_o.shape[0]
    assert np.array_equal(N_(x[:tot]), zero[0][:tot]) and not x[tot:].any() and not dl[tot:].any()
    # perturb=True consumes exactly one torch.rand(N) like the reference (raymarching.py:213-214)
    torch.manual_seed(4)
    want_noise = torch.rand(rays_o.shape[0], device=DEV)
    torch.manual_seed(4)
    counter.zero_()
    x2, _, dl2, r2 = raymarching.march_rays_train(T(rays_o), T(rays_d), 1.0, T(bits), 1, 128, T(nears), T(fars), counter, -1,
                                                  True, 128, True, 0, 1024)
    pert = O.march_rays_train(rays_o, rays_d, 1.0, bits, 1, 128, nears, fars, N_(want_noise), 0.0, 1024)
    assert np.array_equal(N_(r2), pert[3]) and np.array_equal(N_(x2[:int(pert[4][0])]), pert[0][:int(pert[4][0])])


@pytest.mark.parametrize("case", MARCH_CASES[:3])
def test_composite_train_forward_backward(case, ref_ext):
    import raymarching
    rays_o, rays_d, bits, nears, fars, noises = _march_setup(case, seed=3)
    ox, od, ol, orays, ocounter = O.march_rays_train(rays_o, rays_d, case["bound"], bits, case["cascade"], 128, nears, fars,
                                                     noises, case["dt_gamma"], case["max_steps"])
    total = int(ocounter[0])
    M = total + 128 - total % 128
    sig, rgb = util.pseudo_field(ox[:M])
    sig[total:] = 0
    # shuffle the ray rows: the kernels must index outputs by ray id, not by row
    perm = np.random.default_rng(1).permutation(orays.shape[0])
    rays = np.ascontiguousarray(orays[perm])
    ws0, d0, im0 = O.composite_rays_train_forward(sig, rgb, ol[:M], rays, 1e-4)

    s_t = T(sig).requires_grad_(True)
    c_t = T(rgb).requires_grad_(True)
    ws, depth, image = raymarching.composite_rays_train(s_t, c_t, T(ol[:M]), T(rays), 1e-4)
    for got, want in ((ws, ws0), (depth, d0), (image, im0)):
        np.testing.assert_allclose(N_(got), want, rtol=1e-5, atol=1e-6)
    rws, rdepth, rimage = R.composite_rays_train_forward(ref_ext, T(sig), T(rgb), T(ol[:M]), T(rays), 1e-4)
    for got, want in ((ws, rws), (depth, rdepth), (image, rimage)):
        np.testing.assert_allclose(N_(got), N_(want), rtol=1e-5, atol=1e-6)

    rng = np.random.default_rng(2)
    gws = rng.standard_normal(ws0.shape).astype(np.float32)
    gim = rng.standard_normal(im0.shape).astype(np.float32)
    (ws * T(gws)).sum().add((image * T(gim)).sum()).backward()
    gs0, gc0 = O.composite_rays_train_backward(gws, gim, sig, rgb, ol[:M], rays, ws0, im0, 1e-4)
    rgs, rgc = R.composite_rays_train_backward(ref_ext, T(gws), T(gim), T(sig), T(rgb), T(ol[:M]), T(rays), rws, rimage, 1e-4)
    scale = max(np.abs(gs0).max(), 1e-6)
    assert np.abs(N_(s_t.grad) - gs0).max() < 1e-4 * scale + 1e-6
    assert np.abs(N_(s_t.grad) - N_(rgs)).max() < 1e-4 * scale + 1e-6
    np.testing.assert_allclose(N_(c_t.grad), gc0, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(N_(c_t.grad), N_(rgc), rtol=1e-5, atol=1e-6)
    assert not s_t.grad[total:].any()


# ---------------------------------------------------------------------------------------------------------
# raymarching: inference
# ---------------------------------------------------------------------------------------------------------
def test_inference_march_and_composite_vs_oracle_and_reference(ref_ext):
    import raymarching
    case = dict(side=40, cascade=2, bound=2.0, dt_gamma=1.0 / 128, max_steps=512)
    rays_o, rays_d, bits, nears, fars, noises = _march_setup(case, seed=4)
    N = rays_o.shape[0]
    alive = np.arange(N, dtype=np.int32)[::-1].copy()
    rays_t = nears.copy()
    wsum = np.zeros(N, np.float32); depth = np.zeros(N, np.float32); image = np.zeros((N, 3), np.float32)
    t_alive, t_t = T(alive), T(rays_t)
    t_w, t_d, t_i = T(wsum), T(depth), T(image)
    r_alive, r_t, r_w, r_d, r_i = T(alive), T(rays_t), T(wsum), T(depth), T(image)
    first = True
    for it in range(6):
        n_alive = len(alive)
        if n_alive == 0:
            break
        n_step = max(min(N // n_alive, 8), 1)
        nz = noises[:n_alive] if first else np.zeros(n_alive, np.float32)
        ox, od, ol = O.march_rays(n_alive, n_step, alive, rays_t, rays_o, rays_d, 2.0, bits, 2, 128, nears, fars, nz,
                                  case["dt_gamma"], 512, align=128)
        c = cabi()
        M = ox.shape[0]
        x = torch.zeros(M, 3, device=DEV); d = torch.zeros(M, 3, device=DEV); dl = torch.zeros(M, 2, device=DEV)
        ro, rd, bf, ne, fa, nzt = T(rays_o), T(rays_d), T(bits), T(nears), T(fars), T(nz)
        c.call("ngp_march_rays", x.device, n_alive, n_step, c.ptr(t_alive), c.ptr(t_t), c.ptr(ro), c.ptr(rd),
               2.0, float(case["dt_gamma"]), 512, 2, 128, c.ptr(bf), c.ptr(ne), c.ptr(fa), c.ptr(x), c.ptr(d),
               c.ptr(dl), c.ptr(nzt))
        torch.cuda.synchronize()
        assert np.array_equal(N_(x), ox) and np.array_equal(N_(dl), ol) and np.array_equal(N_(d), od)
        rx, rdd, rl = R.march_rays(ref_ext, n_alive, n_step, r_alive, r_t, T(rays_o), T(rays_d), 2.0, T(bits), 2, 128,
                                   T(nears), T(fars), T(nz), case["dt_gamma"], 512, 128)
        assert torch.equal(rx, x) and torch.equal(rl, dl)

        sig, rgb = util.pseudo_field(ox)
        alive2, rays_t, wsum, depth, image = O.composite_rays(n_alive, n_step, alive, rays_t, sig, rgb, ol, wsum, depth,
                                                              image, 1e-2)
        raymarching.composite_rays(n_alive, n_step, t_alive, t_t, T(sig), T(rgb), dl, t_w, t_d, t_i, 1e-2)
        ref_ext.march.composite_rays(n_alive, n_step, 1e-2, r_alive, r_t, T(sig), T(rgb), rl, r_w, r_d, r_i)
        assert np.array_equal(N_(t_alive[:n_alive]), alive2)
        assert torch.equal(t_alive[:n_alive], r_alive[:n_alive])
        np.testing.assert_allclose(N_(t_w), wsum, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(N_(t_i), image, rtol=1e-5, atol=1e-6)
        assert torch.equal(t_w, r_w) and torch.equal(t_i, r_i) and torch.equal(t_d, r_d) and torch.equal(t_t, r_t)

        alive = O.compact_alive(alive2)
        buf, n_out = raymarching.compact_alive(t_alive, n_alive)
        assert n_out.item() == len(alive) and np.array_equal(N_(buf[:len(alive)]), alive)
        t_alive = buf[:len(alive)].contiguous()
        r_alive = r_alive[:n_alive][r_alive[:n_alive] >= 0].contiguous()
        assert torch.equal(t_alive, r_alive)
        first = False
    assert it >= 3


def test_compact_alive_sizes():
    import raymarching
    rng = np.random.default_rng(0)
    for n in (1, 31, 1024, 1025, 70001, 640000):
        a = rng.integers(-1, 5000, n).astype(np.int32)
        a[rng.random(n) < 0.6] = -1
        buf, n_out = raymarching.compact_alive(T(a))
        want = O.compact_alive(a)
        assert n_out.item() == len(want)
        assert np.array_equal(N_(buf[:len(want)]), want)


# ---------------------------------------------------------------------------------------------------------
# freq encoder, occupancy update
# ---------------------------------------------------------------------------------------------------------
def test_freq_encoder_vs_oracle_and_reference(ref_ext):
    from freqencoder import FreqEncoder
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, (4097, 3)).astype(np.float32)
    enc = FreqEncoder(input_dim=3, degree=6)
    xt = T(x).requires_grad_(True)
    out = enc(xt)
    assert out.shape == (4097, 39)
    assert np.abs(N_(out) - O.freq_encode_forward(x, 6)).max() < 5e-6       # sin.approx vs sinf
    ro = torch.empty(4097, 39, device=DEV)
    ref_ext.freq.freq_encode_forward(T(x), 4097, 3, 6, 39, ro)
    assert torch.equal(out.detach(), ro)
    g = rng.standard_normal((4097, 39)).astype(np.float32)
    out.backward(T(g))
    rgi = torch.zeros(4097, 3, device=DEV)
    ref_ext.freq.freq_encode_backward(T(g), ro, 4097, 3, 6, 39, rgi)
    assert torch.allclose(xt.grad, rgi, rtol=1e-5, atol=1e-5)
    assert np.abs(N_(xt.grad) - O.freq_encode_backward(g, N_(ro), 3, 6)).max() < 1e-3


def test_occupancy_update_vs_oracle():
    c = cabi()
    H = 128
    rng = np.random.default_rng(0)
    for bound in (1.0, 2.0):
        noise = rng.random((H ** 3, 3)).astype(np.float32)
        pts = torch.empty(H ** 3, 3, device=DEV)
        hgs = bound / H
        nt = T(noise)
        c.call("ngp_occupancy_cell_points", pts.device, H, float(bound - hgs), float(hgs), c.ptr(nt), c.ptr(pts))
        assert np.array_equal(N_(pts), O.occupancy_cell_points(H, bound, noise))
    n = 2 * H ** 3
    grid = (rng.random(n).astype(np.float32) * 30)
    grid[rng.random(n) < 0.1] = -1.0
    tmp = rng.random(n).astype(np.float32) * 30
    g_t = T(grid)
    mean = torch.empty(1, device=DEV)
    bits = torch.empty(n // 8, dtype=torch.uint8, device=DEV)
    ws = torch.empty(16, dtype=torch.uint8, device=DEV)
    tmp_t = T(tmp)
    c.call("ngp_update_density_grid", g_t.device, c.ptr(g_t), c.ptr(tmp_t), n, 0.95, 10.0, c.ptr(mean), c.ptr(bits), c.ptr(ws), 16)
    new, omean, obits = O.update_density_grid(grid, tmp, 0.95, 10.0)
    assert np.array_equal(N_(g_t), new)
    assert abs(mean.item() - omean) <= 1e-6 * abs(omean)
    assert np.array_equal(N_(bits), O.packbits(new, min(mean.item(), 10.0)))
    # torch restatement of renderer.py:600-607
    gt = T(grid); valid = gt >= 0
    gt[valid] = torch.maximum(gt[valid] * 0.95, T(tmp)[valid])
    assert torch.equal(gt, g_t)
    assert abs(gt[valid].mean().item() - mean.item()) < 1e-5 * mean.item()


def test_grid_backward_warp_aggregated_on_ray_ordered_samples():
    """Marched samples are ray ordered: neighbouring lanes share coarse cells, which the warp-aggregated scatter
    collapses.  Same result as the per-sample scatter and as the oracle's exact sum."""
    c = cabi()
    case = MARCH_CASES[0]
    rays_o, rays_d, bits, nears, fars, noises = _march_setup(case, seed=5)
    ox, _, _, _, ocounter = O.march_rays_train(rays_o, rays_d, 1.0, bits, 1, 128, nears, fars, noises, 0.0, 1024)
    total = int(ocounter[0])
    x01 = ((ox[:total] + 1) / 2).astype(np.float32)
    x01[7] = [1.5, 0.5, 0.5]                        # an out-of-range sample inside a run
    rng = np.random.default_rng(0)
    offs, S = util.make_offsets(log2_hashmap_size=16)
    dev_sc, _ = device_scales(16, np.float32(S), 16)
    for dtype in (np.float16, np.float32):
        g = rng.standard_normal((total, 32)).astype(dtype)
        truth = O.grid_encode_backward(g, x01, offs, offs[-1], 2, np.float32(S), 16, gridtype=1, scale_override=dev_sc)
        ge_agg, _ = my_grid_backward(T(g), T(x01), T(offs), int(offs[-1]), 2, np.float32(S), 16, 1)
        c.load().ngp_grid_set_option(0, 1)
        try:
            ge_plain, _ = my_grid_backward(T(g), T(x01), T(offs), int(offs[-1]), 2, np.float32(S), 16, 1)
        finally:
            c.load().ngp_grid_set_option(0, 0)
        assert util.rel_l2(N_(ge_agg), truth) < 1e-6
        assert util.rel_l2(N_(ge_plain), truth) < 1e-6
    # hashed table, partial last warp
    offs, S = util.make_offsets(log2_hashmap_size=19)
    n = total - 13
    g = rng.standard_normal((n, 32)).astype(np.float16)
    dev_sc, _ = device_scales(16, np.float32(S), 16)
    truth = O.grid_encode_backward(g, x01[:n], offs, offs[-1], 2, np.float32(S), 16, gridtype=0, scale_override=dev_sc)
    ge, _ = my_grid_backward(T(g), T(x01[:n]), T(offs), int(offs[-1]), 2, np.float32(S), 16, 0)
    assert util.rel_l2(N_(ge), truth) < 1e-6


@pytest.mark.parametrize("max_run", [32, 8, 4])
@pytest.mark.parametrize("gridtype,log2_T", [(1, 16), (0, 19)])
def test_sample_scatter_single_and_split_buffers_vs_exact_sum(gridtype, log2_T, max_run, request):
    """(max_run: the longest run of lanes the kernel sums before its reds, ngp_grid_set_option 2.)
    ngp_grid_scatter_samples (one gradient buffer) and ngp_grid_scatter_samples_split + ngp_grid_fold_odd (even / odd
    row pairs as 16-byte reds into two buffers) against the oracle's exact sum, on ray-ordered marched samples in world
    coordinates, with whole runs of all-zero gradient rows (samples behind a ray's early-termination point: their lanes
    issue nothing) and a device-side row count smaller than the buffers."""
    c = cabi()
    case = MARCH_CASES[0]
    rays_o, rays_d, bits, nears, fars, noises = _march_setup(case, seed=11)
    ox, _, _, orays, ocounter = O.march_rays_train(rays_o, rays_d, 1.0, bits, 1, 128, nears, fars, noises, 0.0, 1024)
    total = int(ocounter[0])
    xyz = ox[:total].astype(np.float32).copy()
    xyz[9] = [1.5, 0.0, 0.0]                         # outside [-bound, bound]
    rng = np.random.default_rng(3)
    g = (rng.standard_normal((total, 32)) * 1e-2).astype(np.float16)
    for n in range(0, orays.shape[0], 3):            # every third ray: the last 60 % of its samples carry no gradient
        o, k = int(orays[n, 1]), int(orays[n, 2])
        g[o + int(0.4 * k):o + k] = 0
    g[100:164] = 0                                   # two entirely dead warps
    g[200] = -0.0
    n_rows = total - 37
    offs, S = util.make_offsets(log2_hashmap_size=log2_T)
    dev_sc, _ = device_scales(16, np.float32(S), 16)
    x01 = ((xyz + np.float32(1.0)) * np.float32(0.5)).astype(np.float32)
    truth = O.grid_encode_backward(g[:n_rows], x01[:n_rows], offs, offs[-1], 2, np.float32(S), 16, gridtype=gridtype,
                                   scale_override=dev_sc)
    n_par = int(offs[-1]) * 2
    g_t, x_t, offs_t = T(g), T(xyz), T(offs)
    count = torch.tensor([n_rows, 0], dtype=torch.int32, device=DEV)
    single = torch.zeros(n_par, device=DEV)
    default_run = 8
    assert c.load().ngp_grid_set_option(2, max_run) == 0
    request.addfinalizer(lambda: c.load().ngp_grid_set_option(2, default_run))
    c.call("ngp_grid_scatter_samples", single.device, c.ptr(g_t), c.ptr(x_t), 1.0, c.ptr(count), total, c.ptr(offs_t), 16, 2,
           float(S), 16, gridtype, 0, c.ptr(single))
    even = torch.zeros(n_par, device=DEV)
    odd = torch.zeros(n_par + 4, device=DEV)[2:2 + n_par]
    assert even.data_ptr() % 16 == 0 and odd.data_ptr() % 16 == 8
    c.call("ngp_grid_scatter_samples_split", even.device, c.ptr(g_t), c.ptr(x_t), 1.0, c.ptr(count), total, c.ptr(offs_t), 16, 2,
           float(S), 16, gridtype, 0, c.ptr(even), c.ptr(odd))
    torch.cuda.synchronize()
    odd_share = odd.abs().sum().item() / (even.abs().sum().item() + odd.abs().sum().item())
    c.call("ngp_grid_fold_odd", even.device, c.ptr(even), c.ptr(odd), n_par)
    torch.cuda.synchronize()
    assert odd.abs().sum().item() == 0
    assert util.rel_l2(N_(single).reshape(-1, 2), truth) < 1e-6
    assert util.rel_l2(N_(even).reshape(-1, 2), truth) < 1e-6
    if gridtype == 1:
        assert 0.2 < odd_share < 0.8                 # tiled levels: about half of the pairs start on an odd row
    # a misaligned twin is refused, not silently mis-added
    bad = torch.zeros(n_par + 4, device=DEV)
    rc = c.load().ngp_grid_scatter_samples_split(c.ptr(g_t), c.ptr(x_t), 1.0, c.ptr(count), total, c.ptr(offs_t), 16, 2, float(S),
                                                  16, gridtype, 0, c.ptr(even), c.ptr(bad), None)
    assert rc != 0


@pytest.mark.parametrize("case", MARCH_CASES)
def test_march_thread_per_ray_variant_is_bit_identical(case):
    """The C ABI keeps the reference's one-thread-per-ray decomposition behind an option; both must agree bit for bit
    with the oracle (the default tests above exercise the warp-per-ray walk)."""
    c = cabi()
    rays_o, rays_d, bits, nears, fars, noises = _march_setup(case, seed=6)
    ox, od, ol, orays, ocounter = O.march_rays_train(rays_o, rays_d, case["bound"], bits, case["cascade"], 128, nears, fars,
                                                     noises, case["dt_gamma"], case["max_steps"])
    total = int(ocounter[0])
    outs = []
    for opt in (1, 0):
        c.load().ngp_march_set_option(0, opt)
        try:
            outs.append(_my_march(case, rays_o, rays_d, bits, nears, fars, noises))
        finally:
            c.load().ngp_march_set_option(0, 0)
    for xyzs, dirs, deltas, rays, counter in outs:
        assert np.array_equal(N_(rays), orays) and np.array_equal(N_(counter), ocounter)
        assert np.array_equal(N_(xyzs[:total]), ox[:total]) and np.array_equal(N_(deltas[:total]), ol[:total])
        assert np.array_equal(N_(dirs[:total]), od[:total])
        assert not xyzs[total:].any()


def test_cfg2_full_size_encoder_vs_reference_extension(ref_ext):
    """BASELINE configs[1] at its REAL size: hash grid 2^19 rows/level, 16 x 2, base 16 -> 2048, 2^22 points, fp16.
    Forward bit-equal to the reference extension; backward (fp32-accumulated here, fp16 atomics there) agrees with it to the
    reference's own rounding, conserves the gradient mass exactly (corner weights of a level sum to 1: the size-independent
    property) and a 2^16-point slice matches the exact fp64 sum of the C oracle."""
    from gridencoder import GridEncoder
    from oracle import oracle as O
    from oracle import ref_ext as R
    torch.manual_seed(0)
    enc = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=2048,
                      gridtype="hash").to(DEV)
    with torch.no_grad():
        enc.embeddings.uniform_(-1, 1)
    B = 1 << 22
    x = torch.rand(B, 3, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1)) * 2 - 1
    with torch.autocast("cuda", torch.float16):
        out = enc(x, bound=1)
    assert out.dtype == torch.half and out.shape == (B, 32)
    emb_h = enc.embeddings.detach().half()
    x01 = ((x + 1) / 2).contiguous()
    S = float(np.log2(enc.per_level_scale))
    ref_out, _, _ = R.grid_encode_forward(ref_ext, x01, emb_h, enc.offsets, S, 16, False, 0, False)
    assert torch.equal(out, ref_out.to(out.dtype))
    g = torch.randn(out.shape, device=DEV, dtype=out.dtype, generator=torch.Generator(device=DEV).manual_seed(2))
    out.backward(g)
    mine = enc.embeddings.grad
    assert mine.dtype == torch.float32
    theirs = R.grid_encode_backward(ref_ext, g.contiguous(), x01, emb_h, enc.offsets, S, 16, None, 0, False)[0].float()
    rel = ((mine - theirs).norm() / mine.norm()).item()
    assert rel < 5e-3, rel                                   # the reference accumulates in fp16
    # mass conservation per level and channel, all 4 M points (fp64 sums on the device)
    offs = enc.offsets.cpu().numpy()
    gl = g.double().view(B, 16, 2).sum(0)
    for l in range(16):
        got = mine[offs[l]:offs[l + 1]].double().sum(0)
        assert torch.allclose(got, gl[l], rtol=1e-5, atol=1e-3), (l, got, gl[l])
    # exact sum on a slice the C oracle finishes in seconds
    n = 1 << 16
    sc, _ = device_scales(16, np.float32(S), 16)
    enc.embeddings.grad = None
    with torch.autocast("cuda", torch.float16):
        o2 = enc(x[:n].contiguous(), bound=1)
    o2.backward(g[:n].contiguous())
    truth = O.grid_encode_backward(N_(g[:n]), N_(x01[:n]), offs, enc.embeddings.shape[0], 2, np.float32(S), 16, gridtype=0,
                                   scale_override=sc)
    assert util.rel_l2(N_(enc.embeddings.grad), truth) < 1e-6


THREAD_MARCH_CASES = [
    MARCH_CASES[0],
    MARCH_CASES[1],
    dict(side=48, cascade=2, bound=2.0, dt_gamma=0.0, max_steps=512),     # two cascades, constant step
    dict(side=40, cascade=3, bound=4.0, dt_gamma=0.0, max_steps=1024),    # t spans five binades (0.2 .. 12)
    dict(side=32, cascade=1, bound=1.0, dt_gamma=0.0, max_steps=4096),    # step 2^-10 * sqrt(3) / 2: other m per binade
]


@pytest.mark.parametrize("case", THREAD_MARCH_CASES)
@pytest.mark.parametrize("seed", [2, 7])
def test_march_thread_walk_with_closed_form_jumps_is_bit_identical(case, seed):
    """ngp_march_set_option(2, n): launches of >= n rays with dt_gamma == 0 run one THREAD per ray and collapse the
    reference's inner `do t += dt; while (t < tt)` into integer arithmetic on the bit pattern of t (exact inside a binade;
    binade crossings and ties take a real float step).  Counts, offsets, positions and both deltas must equal the oracle's
    serial float loop bit for bit - including rays that hit the max_steps cap, several cascades and t ranges over many
    binades - and the warp-per-ray walk's output."""
    c = cabi()
    rays_o, rays_d, bits, nears, fars, noises = _march_setup(case, seed=seed)
    ox, od, ol, orays, ocounter = O.march_rays_train(rays_o, rays_d, case["bound"], bits, case["cascade"], 128, nears, fars,
                                                     noises, case["dt_gamma"], case["max_steps"])
    total = int(ocounter[0])
    assert total > 0
    outs = []
    for min_rays in (1, 0):                       # thread-per-ray walk, then the default warp-per-ray walk
        assert c.load().ngp_march_set_option(2, min_rays) == 0
        try:
            outs.append(_my_march(case, rays_o, rays_d, bits, nears, fars, noises))
        finally:
            assert c.load().ngp_march_set_option(2, 0) == 0
    for xyzs, dirs, deltas, rays, counter in outs:
        assert np.array_equal(N_(rays), orays) and np.array_equal(N_(counter), ocounter)
        assert np.array_equal(N_(xyzs[:total]), ox[:total]) and np.array_equal(N_(deltas[:total]), ol[:total])
        assert np.array_equal(N_(dirs[:total]), od[:total])
        assert not xyzs[total:].any()
