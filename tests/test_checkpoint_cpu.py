"""CPU tests of the checkpoint layout: files in the reference Trainer's format (nerf/utils.py:847-968)
round-trip through the B200 model, and a reference-shaped state_dict loads by name."""
import argparse
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "single-stable-dreamfusion_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# the names a reference checkpoint of nerf/network_grid.NeRFNetwork (-O, bg_radius > 0) carries
REFERENCE_KEYS = ["aabb_train", "aabb_infer", "density_grid", "density_bitfield", "step_counter", "encoder.embeddings",
                  "encoder.offsets", "sigma_net.net.0.weight", "sigma_net.net.0.bias", "sigma_net.net.1.weight",
                  "sigma_net.net.1.bias", "sigma_net.net.2.weight", "sigma_net.net.2.bias", "bg_net.net.0.weight",
                  "bg_net.net.0.bias", "bg_net.net.1.weight", "bg_net.net.1.bias"]


def _model(seed):
    from ngp_b200.network_grid import NeRFNetwork
    torch.manual_seed(seed)
    return NeRFNetwork(argparse.Namespace(bound=1, cuda_ray=True, min_near=0.1, density_thresh=10, bg_radius=1.4))


def test_state_dict_has_exactly_the_reference_names_and_shapes():
    sd = _model(0).state_dict()
    assert sorted(sd) == sorted(REFERENCE_KEYS)
    assert sd["encoder.embeddings"].shape == (903480, 2) and sd["encoder.offsets"].shape == (17,)
    assert sd["density_grid"].shape == (1, 128 ** 3) and sd["density_bitfield"].shape == (128 ** 3 // 8,)
    assert sd["density_bitfield"].dtype == torch.uint8 and sd["step_counter"].shape == (16, 2)
    assert sd["sigma_net.net.0.weight"].shape == (64, 32) and sd["sigma_net.net.2.weight"].shape == (4, 64)
    assert sd["bg_net.net.0.weight"].shape == (64, 39) and sd["bg_net.net.1.weight"].shape == (3, 64)


def test_checkpoint_roundtrip_in_the_reference_layout():
    from ngp_b200 import checkpoint as ck
    a, b = _model(0), _model(1)
    with torch.no_grad():
        a.density_grid.uniform_(0, 20)
        a.density_bitfield.random_(0, 255)
        a.step_counter.random_(0, 1000)
    a.mean_count, a.mean_density = 1234, 5.5
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "df_ep0007.pth")
        ck.save_checkpoint(path, a, epoch=7, global_step=700)
        raw = torch.load(path, weights_only=False)
        assert set(raw) == {"epoch", "global_step", "stats", "mean_count", "mean_density", "model"}
        info = ck.load_checkpoint(path, b)
    assert info["epoch"] == 7 and info["global_step"] == 700 and not info["missing_keys"] and not info["unexpected_keys"]
    assert b.mean_count == 1234 and b.mean_density == 5.5
    for (k, va), (_, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(va, vb), k


def test_reference_shaped_files_load():
    """A dict built the way the reference writes it (extra keys, 'best' files that are bare state_dicts)."""
    from ngp_b200 import checkpoint as ck
    a, b, c = _model(2), _model(3), _model(4)
    ref_file = {"epoch": 3, "global_step": 300, "stats": {"loss": [0.1]}, "mean_count": 77, "mean_density": 0.25,
                "model": dict(a.state_dict(), **{"ema_shadow.0": torch.zeros(1)}),      # an unexpected key is tolerated
                "optimizer": {"state": {}, "param_groups": []}, "lr_scheduler": {}, "scaler": {"scale": 65536.0}}
    info = ck.load_checkpoint(ref_file, b, model_only=True)
    assert info["unexpected_keys"] == ["ema_shadow.0"] and b.mean_count == 77
    assert torch.equal(b.encoder.embeddings, a.encoder.embeddings)
    ck.load_checkpoint(a.state_dict(), c)                                               # bare state_dict
    assert torch.equal(c.sigma_net.net[1].weight, a.sigma_net.net[1].weight)

    class _Opt:                                                                         # a foreign optimizer state is reported
        def load_state_dict(self, sd):
            raise RuntimeError("layout mismatch")
    info = ck.load_checkpoint(ref_file, b, optimizer=_Opt())
    assert info["epoch"] == 3 and any("optimizer" in w for w in info["warnings"])
