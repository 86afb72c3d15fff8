"""CPU tests of ray generation: oracle.get_rays and the provider's pose helpers against golden vectors the
REAL reference produced (nerf/utils.py:get_rays, nerf/provider.py:circle_poses / rand_poses; oracle/make_golden_host.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "single-stable-dreamfusion_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "rays_golden.npz"))


def test_oracle_get_rays_matches_the_reference():
    from oracle import oracle as O
    H, W = int(GOLD["H"]), int(GOLD["W"])
    for tag in ("a", "b"):
        ro, rd = O.get_rays(GOLD["poses"], GOLD["intrinsics_" + tag], H, W)
        assert np.array_equal(ro, GOLD["rays_o_" + tag])
        np.testing.assert_allclose(rd, GOLD["rays_d_" + tag], rtol=0, atol=2e-7)   # fp32 matmul summation order
        assert abs(np.linalg.norm(rd, axis=-1) - 1).max() < 1e-6


def test_provider_poses_and_rays_match_the_reference():
    from ngp_b200 import provider
    H, W = int(GOLD["H"]), int(GOLD["W"])
    for k, phi in enumerate(GOLD["circle_phis"]):
        pose = provider.circle_pose(1.8, 60.0, float(phi))
        np.testing.assert_allclose(pose, GOLD["poses"][k], rtol=0, atol=2e-6)       # nerf/provider.py:144-175
        fov = 55.0
        ro, rd = provider.get_rays(GOLD["poses"][k], H, W, fov)
        assert np.allclose(provider.intrinsics_of(H, W, fov), GOLD["intrinsics_a"], rtol=1e-6)
        np.testing.assert_allclose(rd, GOLD["rays_d_a"][k], rtol=0, atol=1e-6)
        np.testing.assert_allclose(ro, GOLD["rays_o_a"][k], rtol=0, atol=0)
    # every reference pose is a rigid look-at frame with up = -y; ours builds the same frame from the same centre
    for pose in GOLD["poses"]:
        mine = provider.look_at(pose[:3, 3].astype(np.float64))
        np.testing.assert_allclose(mine, pose, rtol=0, atol=3e-6)


def test_training_poses_describe_the_training_views():
    from ngp_b200 import provider
    from oracle import oracle as O
    poses, intr = provider.make_training_poses(4, 32, 32, seed=7)
    ro, rd = provider.make_training_views(4, 32, 32, seed=7, pin=False)
    o, d = O.get_rays(poses.numpy(), intr.numpy(), 32, 32)
    assert np.array_equal(o, ro.numpy())
    np.testing.assert_allclose(d, rd.numpy(), rtol=0, atol=3e-7)


def test_orbit_cameras_and_video_writer(tmp_path):
    """The --test orbit (nerf/provider.py:214-222, nerf/utils.py:507-555): camera ring and the mp4 writer."""
    from ngp_b200 import orbit, provider
    poses, intr = orbit.orbit_cameras(8, 40, 40, device="cpu")
    assert poses.shape == (8, 4, 4) and np.allclose(np.linalg.norm(poses[:, :3, 3].numpy(), axis=-1), 1.8, atol=1e-5)
    np.testing.assert_allclose(poses[2].numpy(), provider.circle_pose(1.8, 60.0, 90.0), atol=1e-6)
    assert np.allclose(intr.numpy(), provider.intrinsics_of(40, 40, 55.0))
    frames = (np.random.default_rng(0).random((6, 40, 40, 3)) * 255).astype(np.uint8)
    for arr, nm in ((frames, "t_rgb.mp4"), (frames[..., 0], "t_depth.mp4")):
        p = orbit.write_video(str(tmp_path / nm), arr)
        assert os.path.getsize(p) > 0
