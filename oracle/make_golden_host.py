#!/usr/bin/env python
"""TEST INFRASTRUCTURE: golden vectors for the HOST-side rows of SURVEY 8f, produced by importing the REAL reference from
/root/reference in this (GPU-less) container:

  * nerf/utils.py:get_rays + nerf/provider.py:circle_poses / rand_poses  -> tests/golden/rays_golden.npz
    (poses, intrinsics and the rays the reference generates for them; pins oracle.get_rays, provider.circle_pose and the
    device ray generator ngp_get_rays);
  * nerf/network_grid.py:NeRFNetwork(opt).state_dict() -> tests/golden/ref_state_dict_manifest.json
    (every key with shape and dtype, plus a checksum of a seeded state: pins checkpoint compatibility - the model is
    constructible on CPU because building it only allocates parameters; the reference's own gridencoder / raymarching /
    freqencoder extension modules from oracle/_ref satisfy the imports).

Run from the repo root:  python oracle/make_golden_host.py
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def import_reference_grid_model():
    """The reference's nerf.network_grid.NeRFNetwork, nerf.utils and nerf.provider, importable on CPU."""
    from oracle.make_golden_cpu import import_reference
    import_reference()                       # dependency stubs + sys.path (its CPU shims are replaced below)
    for name in ("raymarching", "freqencoder", "gridencoder", "encoding", "activation"):
        sys.modules.pop(name, None)
    sys.path.insert(0, os.path.join(HERE, "_ref"))   # _gridencoder / _raymarching / _freqencoder, built by oracle/build_ref.py
    from nerf.network_grid import NeRFNetwork
    import nerf.provider as provider
    import nerf.utils as utils
    return NeRFNetwork, utils, provider


def main():
    NeRFNetwork, utils, provider = import_reference_grid_model()
    os.makedirs(GOLD, exist_ok=True)

    # ---- rays ---------------------------------------------------------------------------------------------------
    import random
    random.seed(3)
    torch.manual_seed(3)
    H = W = 24
    poses = [provider.circle_poses("cpu", radius=1.8, theta=60, phi=p)[0] for p in (0.0, 93.6, 270.0)]
    poses.append(provider.rand_poses(2, "cpu", uniform_sphere_rate=0.0)[0])
    poses.append(provider.rand_poses(2, "cpu", uniform_sphere_rate=1.0)[0])
    poses = torch.cat(poses, 0).float()
    out = {"poses": poses.numpy(), "H": H, "W": W}
    for tag, fov in (("a", 55.0), ("b", 41.7)):
        focal = H / (2 * np.tan(np.deg2rad(fov) / 2))
        intr = np.array([focal, focal, H / 2, W / 2])
        rays = utils.get_rays(poses, intr, H, W, -1)
        out["intrinsics_" + tag] = intr.astype(np.float32)
        out["rays_o_" + tag] = rays["rays_o"].contiguous().numpy()
        out["rays_d_" + tag] = rays["rays_d"].contiguous().numpy()
    out["circle_phis"] = np.array([0.0, 93.6, 270.0], np.float32)
    np.savez_compressed(os.path.join(GOLD, "rays_golden.npz"), **out)

    # ---- state_dict manifest ------------------------------------------------------------------------------------
    opt = argparse.Namespace(bound=1, cuda_ray=True, min_near=0.1, density_thresh=10, bg_radius=1.4)
    torch.manual_seed(0)
    model = NeRFNetwork(opt)
    sd = model.state_dict()
    manifest = {"source": "nerf/network_grid.py:NeRFNetwork(opt=bound 1, cuda_ray, bg_radius 1.4).state_dict()",
                "keys": {k: {"shape": list(v.shape), "dtype": str(v.dtype).replace("torch.", "")} for k, v in sd.items()},
                "n_parameters": int(sum(p.numel() for p in model.parameters())),
                "param_groups": [{"n_tensors": len(list(g["params"])), "lr_mult": g["lr"] / 1e-3} for g in model.get_params(1e-3)],
                "offsets": sd["encoder.offsets"].tolist()}
    with open(os.path.join(GOLD, "ref_state_dict_manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("wrote", os.path.join(GOLD, "rays_golden.npz"), "and ref_state_dict_manifest.json:", len(manifest["keys"]), "keys")


if __name__ == "__main__":
    main()
