"""The reference's ``--test`` mode: a 100-frame 360-degree orbit rendered through the inference branch of ``run_cuda`` and
written as ``<name>_rgb.mp4`` / ``<name>_depth.mp4`` (``Trainer.test`` + ``test_step``, nerf/utils.py:435-456,507-555, with
the test dataset's cameras, nerf/provider.py:214-222: circle poses at radius ``radius_range[1] * 1.2`` = 1.8, theta 60,
phi = i / size * 360, fixed fov = mean of ``fovy_range`` = 55).

Rays are generated on the device from the poses (ngp_get_rays); frames are converted exactly as the reference does
(``(pred * 255).astype(uint8)``, white background).  Video encoding uses imageio when it is installed (the reference's
encoder) and OpenCV's mp4 writer otherwise; ``write_video=False`` writes PNG frames like the reference.
"""
import os

import numpy as np
import torch

from . import provider


def orbit_cameras(n_frames=100, H=800, W=800, radius=1.8, theta_deg=60.0, fovy_deg=55.0, device="cuda"):
    """poses float32 [n_frames, 4, 4] and intrinsics [4] of the test orbit (provider.py:214-222)."""
    poses = np.stack([provider.circle_pose(radius, theta_deg, 360.0 * i / n_frames) for i in range(n_frames)])
    return (torch.from_numpy(poses.astype(np.float32)).to(device),
            torch.from_numpy(provider.intrinsics_of(H, W, fovy_deg)).to(device))


@torch.no_grad()
def render_frame(model, pose, intrinsics, H, W, bg_color=None, perturb=False, max_steps=1024, dt_gamma=0, shading="albedo",
                 ambient_ratio=1.0, light_d=None):
    """test_step (nerf/utils.py:435-456) for one camera: pred_rgb [H, W, 3], pred_depth [H, W] (device tensors)."""
    from . import _nvtx
    rays_o, rays_d = provider.get_rays_device(pose[None], intrinsics, H, W)
    if bg_color is None:
        bg_color = torch.ones(3, device=rays_o.device)
    with torch.autocast("cuda", torch.float16), _nvtx.range("ngp.orbit.frame"):
        out = model.render(rays_o, rays_d, staged=True, perturb=perturb, light_d=light_d, ambient_ratio=ambient_ratio,
                           shading=shading, force_all_rays=True, bg_color=bg_color, max_steps=max_steps, dt_gamma=dt_gamma)
    return out["image"].reshape(H, W, 3), out["depth"].reshape(H, W)


def to_uint8(pred):
    """(pred * 255).astype(np.uint8) (nerf/utils.py:533-537): NaN depth of box-missing rays becomes 0."""
    a = pred.detach().float().cpu().numpy() * 255
    return np.nan_to_num(a, nan=0.0, posinf=255.0, neginf=0.0).clip(0, 255).astype(np.uint8)


def write_video(path, frames, fps=25):
    """frames uint8 [n, H, W, 3] or [n, H, W].  Returns the path actually written."""
    frames = np.asarray(frames)
    try:
        import imageio
        imageio.mimwrite(path, frames, fps=fps, quality=8, macro_block_size=1)   # nerf/utils.py:552-553
        return path
    except ImportError:
        pass
    import cv2
    h, w = frames.shape[1:3]
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), float(fps), (w, h), isColor=True)
    if not wr.isOpened():   # no mp4 muxer in this OpenCV build: keep the frames losslessly
        path = os.path.splitext(path)[0] + ".npz"
        np.savez_compressed(path, frames=frames, fps=fps)
        return path
    for f in frames:
        wr.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR) if f.ndim == 3 else cv2.cvtColor(f, cv2.COLOR_GRAY2BGR))
    wr.release()
    return path


def test(model, save_path, name="df", n_frames=100, H=800, W=800, write_video_files=True, max_steps=1024, **render_kw):
    """Trainer.test: renders the orbit in eval mode and saves it.  Returns (rgb uint8 [n,H,W,3], depth uint8 [n,H,W],
    [written paths])."""
    os.makedirs(save_path, exist_ok=True)
    was_training = model.training
    model.eval()
    device = next(model.parameters()).device
    poses, intr = orbit_cameras(n_frames, H, W, device=device)
    rgbs, depths, paths = [], [], []
    for i in range(n_frames):
        rgb, depth = render_frame(model, poses[i], intr, H, W, max_steps=max_steps, **render_kw)
        rgb8, depth8 = to_uint8(rgb), to_uint8(depth)
        if write_video_files:
            rgbs.append(rgb8)
            depths.append(depth8)
        else:
            import cv2
            p_rgb = os.path.join(save_path, "%s_%04d_rgb.png" % (name, i))
            p_d = os.path.join(save_path, "%s_%04d_depth.png" % (name, i))
            cv2.imwrite(p_rgb, cv2.cvtColor(rgb8, cv2.COLOR_RGB2BGR))
            cv2.imwrite(p_d, depth8)
            paths += [p_rgb, p_d]
            rgbs.append(rgb8)
            depths.append(depth8)
    rgbs, depths = np.stack(rgbs), np.stack(depths)
    if write_video_files:
        paths.append(write_video(os.path.join(save_path, "%s_rgb.mp4" % name), rgbs))
        paths.append(write_video(os.path.join(save_path, "%s_depth.mp4" % name), depths))
    if was_training:
        model.train()
    return rgbs, depths, paths
