"""Checkpoints in the reference Trainer's on-disk layout (nerf/utils.py:847-968), so that `.pth` files written by the
reference load into the B200 model and vice versa.

A checkpoint is one ``torch.save``d dict:  ``epoch``, ``global_step``, ``stats``; ``mean_count`` / ``mean_density`` (cuda-ray
models); ``model`` = ``state_dict()`` with the reference's parameter / buffer names (``encoder.embeddings``,
``encoder.offsets``, ``sigma_net.net.{0,1,2}.{weight,bias}``, ``bg_net.net.{0,1}.*``, ``density_grid``, ``density_bitfield``,
``step_counter``, ``aabb_train``, ``aabb_infer``); with ``full=True`` also ``optimizer`` and ``scaler``.  A bare
``state_dict`` (the reference's "best" files without the wrapper) loads too.
"""
import torch


def checkpoint_dict(model, epoch=0, global_step=0, stats=None, optimizer=None, scaler=None, full=False):
    state = {"epoch": int(epoch), "global_step": int(global_step),
             "stats": stats if stats is not None else {"loss": [], "valid_loss": [], "results": [], "checkpoints": [],
                                                       "best_result": None}}
    if getattr(model, "cuda_ray", False):
        state["mean_count"] = model.mean_count
        state["mean_density"] = model.mean_density
    if full:
        if optimizer is not None:
            state["optimizer"] = optimizer.state_dict()
        if scaler is not None and scaler is not optimizer and hasattr(scaler, "state_dict"):
            state["scaler"] = scaler.state_dict()
    state["model"] = model.state_dict()
    return state


def save_checkpoint(path, model, **kw):
    torch.save(checkpoint_dict(model, **kw), path)


def load_checkpoint(path_or_dict, model, optimizer=None, scaler=None, model_only=False, map_location=None):
    """Mirrors Trainer.load_checkpoint (nerf/utils.py:896-968).  Returns dict(epoch, global_step, stats, missing_keys,
    unexpected_keys); optimizer / scaler state is restored when present and requested, failures there are reported in
    the returned dict instead of raised (the reference logs a warning and carries on)."""
    ck = path_or_dict
    if not isinstance(ck, dict):
        ck = torch.load(path_or_dict, map_location=map_location, weights_only=False)
    info = {"epoch": 0, "global_step": 0, "stats": None, "missing_keys": [], "unexpected_keys": [], "warnings": []}
    if "model" not in ck:                      # a bare state_dict
        model.load_state_dict(ck)
        return info
    res = model.load_state_dict(ck["model"], strict=False)
    info["missing_keys"], info["unexpected_keys"] = list(res.missing_keys), list(res.unexpected_keys)
    if getattr(model, "cuda_ray", False):
        if "mean_count" in ck:
            model.mean_count = ck["mean_count"]
        if "mean_density" in ck:
            model.mean_density = ck["mean_density"]
    if model_only:
        return info
    info.update(epoch=ck.get("epoch", 0), global_step=ck.get("global_step", 0), stats=ck.get("stats"))
    for name, obj in (("optimizer", optimizer), ("scaler", scaler)):
        if obj is not None and name in ck:
            try:
                obj.load_state_dict(ck[name])
            except Exception as e:  # noqa: BLE001 - as the reference: warn and continue
                info["warnings"].append("failed to load %s: %r" % (name, e))
    return info
