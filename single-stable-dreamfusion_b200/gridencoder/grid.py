"""Drop-in replacement for the reference's ``gridencoder/grid.py`` on top of libngp_b200.so.

Same names, arguments, defaults, shapes and dtypes as the reference (grid.py:19-154):
``GridEncoder(...)``, ``grid_encode(inputs, embeddings, offsets, per_level_scale, base_resolution,
calc_grad_inputs, gridtype, align_corners)``.  What changed underneath:

* the kernel writes ``[B, L*C]`` directly - no ``[L,B,C]`` buffer + permute copy (grid.py:42,52,70);
* under autocast the half copy of the table is cached until the parameter is modified
  (the reference re-casts the whole table on every forward, grid.py:38-39);
* the backward accumulates the table gradient in fp32 (the reference uses fp16 atomics under
  autocast and lets autograd up-cast afterwards), so the returned gradient already has the
  parameter's dtype.
"""
import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from ngp_b200 import _cabi

_gridtype_to_id = {'hash': 0, 'tiled': 1}


def _half_table(embeddings):
    from ngp_b200.field import cached_half
    return cached_half(embeddings)


class _grid_encode(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type='cuda')
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False,
                gridtype=0, align_corners=False):
        # inputs: [B, D] float in [0, 1]; embeddings: [sO, C]; offsets: [L + 1] int32; returns [B, L * C]
        _cabi.require_cuda(inputs, embeddings, offsets)
        if not inputs.is_floating_point() or not embeddings.is_floating_point():
            raise RuntimeError("inputs and embeddings must be floating tensors")
        if offsets.dtype != torch.int32:
            raise RuntimeError("offsets must be an int tensor")
        inputs = inputs.contiguous().float() if inputs.dtype != torch.float32 else inputs.contiguous()
        offsets = offsets.contiguous()

        B, D = inputs.shape
        L = offsets.shape[0] - 1
        C = embeddings.shape[1]
        S = float(np.log2(per_level_scale))
        H = int(base_resolution)

        param_dtype = embeddings.dtype
        # manual autocast (grid.py:36-39): half table iff autocast is on and C is even
        if torch.is_autocast_enabled('cuda') and C % 2 == 0 and embeddings.dtype == torch.float32:
            table = _half_table(embeddings)
        else:
            table = embeddings.detach().contiguous()
        dt = _cabi.dtype_code(table.dtype)

        outputs = torch.empty(B, L * C, device=inputs.device, dtype=table.dtype)
        dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=table.dtype) if calc_grad_inputs else None

        _cabi.call("ngp_grid_encode_forward", inputs.device, _cabi.ptr(inputs), _cabi.ptr(table), _cabi.ptr(offsets),
                   _cabi.ptr(outputs), B, D, C, L, S, H, _cabi.ptr(dy_dx), int(gridtype), int(bool(align_corners)), dt,
                   _cabi.LAYOUT_BLC)

        ctx.save_for_backward(inputs, offsets, dy_dx)
        ctx.dims = [B, D, C, L, S, H, gridtype]
        ctx.align_corners = align_corners
        ctx.table_dtype = table.dtype
        ctx.param_dtype = param_dtype
        ctx.n_rows = embeddings.shape[0]
        return outputs

    @staticmethod
    @torch.amp.custom_bwd(device_type='cuda')
    def backward(ctx, grad):
        inputs, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H, gridtype = ctx.dims

        grad = grad.contiguous()  # [B, L * C]; consumed in place of the reference's [L, B, C] permute copy
        if grad.dtype != ctx.table_dtype:
            grad = grad.to(ctx.table_dtype)
        dt = _cabi.dtype_code(grad.dtype)

        grad_embeddings = torch.zeros(ctx.n_rows, C, device=grad.device, dtype=torch.float32)
        grad_inputs = torch.zeros(B, D, device=grad.device, dtype=ctx.table_dtype) if dy_dx is not None else None

        _cabi.call("ngp_grid_encode_backward", grad.device, _cabi.ptr(grad), _cabi.ptr(inputs), None, _cabi.ptr(offsets),
                   _cabi.ptr(grad_embeddings), B, D, C, L, S, H, _cabi.ptr(dy_dx), _cabi.ptr(grad_inputs), int(gridtype),
                   int(bool(ctx.align_corners)), dt, _cabi.LAYOUT_BLC, _cabi.NGP_F32)

        if grad_embeddings.dtype != ctx.param_dtype:
            grad_embeddings = grad_embeddings.to(ctx.param_dtype)
        if grad_inputs is not None:
            grad_inputs = grad_inputs.to(inputs.dtype)
        return grad_inputs, grad_embeddings, None, None, None, None, None, None


grid_encode = _grid_encode.apply


class GridEncoder(nn.Module):
    """Multi-resolution hash / tiled grid encoder; constructor and attributes as grid.py:91-133."""

    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, gridtype='hash', align_corners=False):
        super().__init__()

        # the finest resolution desired at the last level overrides per_level_scale
        if desired_resolution is not None:
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))

        self.input_dim = input_dim
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.per_level_scale = per_level_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype = gridtype
        self.gridtype_id = _gridtype_to_id[gridtype]
        self.align_corners = align_corners

        # rows per level: dense while it fits, capped at 2^log2_hashmap_size, rounded up to 8 (grid.py:110-120)
        offsets = []
        offset = 0
        self.max_params = 2 ** log2_hashmap_size
        for i in range(num_levels):
            resolution = int(np.ceil(base_resolution * per_level_scale ** i))
            params_in_level = min(self.max_params, (resolution if align_corners else resolution + 1) ** input_dim)
            params_in_level = int(np.ceil(params_in_level / 8) * 8)
            offsets.append(offset)
            offset += params_in_level
        offsets.append(offset)
        offsets = torch.from_numpy(np.array(offsets, dtype=np.int32))
        self.register_buffer('offsets', offsets)

        self.n_params = offsets[-1] * level_dim

        self.embeddings = nn.Parameter(torch.empty(offset, level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        std = 1e-4
        self.embeddings.data.uniform_(-std, std)

    def __repr__(self):
        return (f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"resolution={self.base_resolution} -> "
                f"{int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))} "
                f"per_level_scale={self.per_level_scale:.4f} params={tuple(self.embeddings.shape)} "
                f"gridtype={self.gridtype} align_corners={self.align_corners}")

    def forward(self, inputs, bound=1):
        # inputs: [..., input_dim] in [-bound, bound]; returns [..., num_levels * level_dim]
        inputs = (inputs + bound) / (2 * bound)  # map to [0, 1]
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.input_dim)
        outputs = grid_encode(inputs, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution,
                              inputs.requires_grad, self.gridtype_id, self.align_corners)
        return outputs.view(prefix_shape + [self.output_dim])
