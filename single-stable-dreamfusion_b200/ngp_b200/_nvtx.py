"""Optional NVTX ranges around the stages of the hot path (tracing).  Enabled with NGP_NVTX=1 (or
``_nvtx.ENABLED = True``); a no-op otherwise, so the graphed step pays nothing.  Ranges: ``ngp.occupancy_refresh``,
``ngp.step.march``, ``ngp.step.compute``, ``ngp.step.optimizer``, ``ngp.infer.loop``, ``ngp.orbit.frame``."""
import contextlib
import os

import torch

ENABLED = os.environ.get("NGP_NVTX", "0") not in ("", "0")


@contextlib.contextmanager
def range(name):
    if not ENABLED:
        yield
        return
    torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()
