"""Synthetic camera views: random orbit poses + per-pixel rays, host side.

Distributions and conventions follow the reference's data provider so that the synthetic workload
has the shape of a real training view: ``rand_poses`` (nerf/provider.py:72-141: radius U[1,1.5];
with probability ``uniform_sphere_rate`` a direction uniform on the upper hemisphere, else
theta U[0,100] deg, phi U[0,360] deg; look-at with up = -y), fovy U[40,70] (:211) and ``get_rays``
(nerf/utils.py:43-106: pixel centres at +0.5, normalised directions, rays_d = dirs @ R^T).
Everything is generated once, with a seeded numpy Generator, into pinned host memory.
"""
import numpy as np
import torch


def _normalize(v):
    return v / np.maximum(np.linalg.norm(v, axis=-1, keepdims=True), 1e-20)


def rand_pose(rng, radius_range=(1.0, 1.5), theta_range=(0.0, 100.0), phi_range=(0.0, 360.0), uniform_sphere_rate=0.5):
    radius = rng.uniform(*radius_range)
    if rng.random() < uniform_sphere_rate:
        unit = _normalize(np.array([rng.uniform(-1, 1), rng.uniform(0, 1), rng.uniform(-1, 1)]))
        centre = unit * radius
    else:
        theta = np.deg2rad(rng.uniform(*theta_range))
        phi = np.deg2rad(rng.uniform(*phi_range))
        centre = radius * np.array([np.sin(theta) * np.sin(phi), np.cos(theta), np.sin(theta) * np.cos(phi)])
    return look_at(centre)


def circle_pose(radius=1.8, theta_deg=60.0, phi_deg=0.0):
    theta, phi = np.deg2rad(theta_deg), np.deg2rad(phi_deg)
    centre = radius * np.array([np.sin(theta) * np.sin(phi), np.cos(theta), np.sin(theta) * np.cos(phi)])
    return look_at(centre)


def look_at(centre, target=np.zeros(3)):
    forward = _normalize(target - centre)
    up0 = np.array([0.0, -1.0, 0.0])
    right = _normalize(np.cross(forward, up0))
    up = _normalize(np.cross(right, forward))
    pose = np.eye(4, dtype=np.float32)
    pose[:3, :3] = np.stack([right, up, forward], axis=-1)
    pose[:3, 3] = centre
    return pose


def get_rays(pose, H, W, fovy_deg):
    focal = H / (2 * np.tan(np.deg2rad(fovy_deg) / 2))
    cx, cy = H / 2, W / 2
    j, i = np.meshgrid(np.arange(H, dtype=np.float32) + 0.5, np.arange(W, dtype=np.float32) + 0.5, indexing="ij")
    dirs = np.stack([(i - cx) / focal, (j - cy) / focal, np.ones_like(i)], -1).reshape(-1, 3)
    dirs = _normalize(dirs)
    rays_d = (dirs @ pose[:3, :3].T).astype(np.float32)
    rays_o = np.broadcast_to(pose[:3, 3].astype(np.float32), rays_d.shape).copy()
    return rays_o, rays_d


def intrinsics_of(H, W, fovy_deg):
    """(fx, fy, cx, cy) as the reference's dataset builds them (nerf/provider.py:188-189,212-213: focal from the image
    HEIGHT, cx = H / 2, cy = W / 2)."""
    focal = H / (2 * np.tan(np.deg2rad(fovy_deg) / 2))
    return np.array([focal, focal, H / 2, W / 2], np.float32)


def make_training_poses(n_views, H=64, W=64, seed=0, fovy_range=(40.0, 70.0)):
    """The seeded training views as CAMERAS: poses float32 [n_views, 4, 4] (cam2world) and intrinsics [n_views, 4] -
    what a train step needs when the rays are generated on the device (get_rays_device / TrainStep(device_rays=...)).
    Same random stream as make_training_views, so both describe the same views."""
    rng = np.random.default_rng(seed)
    poses = np.empty((n_views, 4, 4), np.float32)
    intr = np.empty((n_views, 4), np.float32)
    for v in range(n_views):
        poses[v] = rand_pose(rng)
        intr[v] = intrinsics_of(H, W, rng.uniform(*fovy_range))
    return torch.from_numpy(poses), torch.from_numpy(intr)


def get_rays_device(poses, intrinsics, H, W, row0=0, row_stride=1, n_rows=None):
    """nerf/utils.py:43-106 get_rays (full image, N = -1) as one kernel on the device: poses [B,4,4] (CUDA, fp32),
    intrinsics [4] or [B,4] -> rays_o, rays_d [B, n_rows * W, 3].  row0 / row_stride / n_rows select the interleaved image
    rows a data-parallel rank renders (parallel.shard_rows)."""
    from . import _cabi
    _cabi.require_cuda(poses, intrinsics)
    poses = poses.contiguous().float()
    intrinsics = intrinsics.contiguous().float()
    B = poses.shape[0]
    if n_rows is None:
        n_rows = (H - row0 + row_stride - 1) // row_stride
    per_view = 1 if intrinsics.dim() == 2 else 0
    if per_view and intrinsics.shape != (B, 4) or (not per_view and intrinsics.shape != (4,)):
        raise RuntimeError("intrinsics [4] or [B, 4] expected")
    rays_o = torch.empty(B, n_rows * W, 3, device=poses.device)
    rays_d = torch.empty(B, n_rows * W, 3, device=poses.device)
    _cabi.call("ngp_get_rays", poses.device, _cabi.ptr(poses), _cabi.ptr(intrinsics), per_view, B, H, W, row0, row_stride, n_rows,
               _cabi.ptr(rays_o), _cabi.ptr(rays_d))
    return rays_o, rays_d


def make_training_views(n_views, H=64, W=64, seed=0, fovy_range=(40.0, 70.0), pin=True):
    """Returns rays_o, rays_d as float32 tensors [n_views, H*W, 3] (pinned when a GPU is present)."""
    rng = np.random.default_rng(seed)
    ro = np.empty((n_views, H * W, 3), np.float32)
    rd = np.empty((n_views, H * W, 3), np.float32)
    for v in range(n_views):
        pose = rand_pose(rng)
        ro[v], rd[v] = get_rays(pose, H, W, rng.uniform(*fovy_range))
    ro_t, rd_t = torch.from_numpy(ro), torch.from_numpy(rd)
    if pin and torch.cuda.is_available():
        ro_t, rd_t = ro_t.pin_memory(), rd_t.pin_memory()
    return ro_t, rd_t


def make_orbit_views(n_frames, H=800, W=800, radius=1.8, theta_deg=60.0, fovy_deg=55.0):
    """The reference's test orbit (provider.py:216-222): circle poses at phi = i * 360 / n."""
    views = []
    for i in range(n_frames):
        pose = circle_pose(radius, theta_deg, 360.0 * i / n_frames)
        views.append(get_rays(pose, H, W, fovy_deg))
    return views
