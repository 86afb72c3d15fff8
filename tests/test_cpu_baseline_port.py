"""CPU test: the pure-PyTorch port of the reference's non-cuda-ray renderer (oracle/torch_renderer.py, the reported
CPU baseline) reproduces the fixture that oracle/make_golden_cpu.py generated from the REAL reference."""
import os

import numpy as np
import torch

from oracle import torch_renderer as TR

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cpu_renderer_golden.npz"))


def test_port_matches_reference_fixture():
    torch.manual_seed(0)
    port = TR.VanillaNeRF(bound=1.0, min_near=0.1, bg_radius=1.4)   # same init order as nerf/network.py under seed 0
    port.train()
    assert sum(p.numel() for p in port.parameters()) == int(GOLD["n_params"][0]) == 66567
    rays_o, rays_d = TR.make_view(16, 16, seed=3)
    G = torch.randn(1, 256, 3, generator=torch.Generator().manual_seed(1))
    torch.manual_seed(123)
    out = port.run(rays_o, rays_d, num_steps=64, upsample_steps=32, perturb=True)
    (out["image"] * G).sum().backward()
    for k in ("image", "depth", "weights_sum"):
        np.testing.assert_allclose(out[k].detach().numpy(), GOLD[k], rtol=1e-5, atol=1e-6)
    g = dict(port.named_parameters())
    np.testing.assert_allclose(g["sigma_net.net.4.weight"].grad.numpy(), GOLD["grad_sigma_net_last_w"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(g["bg_net.net.0.dense.weight"].grad.numpy(), GOLD["grad_bg_net_first_w"], rtol=1e-4, atol=1e-6)


def test_near_far_shim_matches_c_oracle():
    from oracle import oracle as O
    import ngp_testutil as util
    ro, rd = util.look_at_rays(24, radius=1.4)
    ro[3] = [3, 3, 3]; rd[3] = [1, 0, 0]
    aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    n0, f0 = O.near_far_from_aabb(ro, rd, aabb, 0.1)
    n1, f1 = TR.near_far_from_aabb(torch.from_numpy(ro), torch.from_numpy(rd), torch.from_numpy(aabb), 0.1)
    assert np.array_equal(n1.numpy(), n0) and np.array_equal(f1.numpy(), f0)


def test_train_step_counts_samples():
    torch.manual_seed(0)
    port = TR.VanillaNeRF()
    port.train()
    ro, rd = TR.make_view(8, 8, seed=1)
    n, out = TR.train_step(port, ro, rd, torch.ones(1, 64, 3), num_steps=16, upsample_steps=8)
    assert n == 64 * 24 and torch.isfinite(out["image"]).all()
    assert all(p.grad is not None for p in port.parameters())
