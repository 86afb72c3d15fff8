"""Fused field op: grid encode -> MLP(32, 64, 64, 4) -> trunc_exp / sigmoid as ONE kernel per direction
(csrc/field_mlp.cu, tcgen05 tensor cores), wrapped as an autograd Function.

It computes what ``NeRFNetwork.common_forward`` (nerf/network_grid.py:76-87) computes under fp16 autocast
and back-propagates into the embedding table, the three Linear weights and biases.  ``albedo`` is returned
as fp32 holding the fp16-rounded sigmoid (the reference hands a half tensor to compositing, which
immediately widens it: raymarching.py:240).
"""
import weakref

import numpy as np
import torch
from torch.autograd import Function

from . import _cabi

_half_cache = {}
_cache_epoch = [0]


def invalidate_half_cache():
    """Forget every cached fp16 copy (call after parameters were updated behind autograd's back, e.g. by a
    CUDA-graph replay of the optimizer step)."""
    _cache_epoch[0] += 1


_shadows = {}  # id(param) -> (weakref to param, fp16 shadow tensor, [param._version the shadow matches])


def register_half_shadow(param, shadow):
    """Declare `shadow` (fp16, same shape) as the maintained fp16 copy of `param`: the fused optimizer kernel rewrites
    it together with the parameter (optim.py), so no per-forward cast is needed.  In-place edits made through torch
    (load_state_dict, init code) bump param._version and trigger one refresh."""
    _shadows[id(param)] = (weakref.ref(param), shadow, [param._version])


def cached_half(t):
    """fp16 copy of a parameter, refreshed only when the parameter changes (autocast re-casts per forward)."""
    key = id(t)
    s = _shadows.get(key)
    if s is not None and s[0]() is t:
        if t._version != s[2][0]:
            with torch.no_grad():
                s[1].copy_(t)
            s[2][0] = t._version
        return s[1]
    ver = (t.data_ptr(), t._version, tuple(t.shape), _cache_epoch[0])
    hit = _half_cache.get(key)
    # the entry must belong to THIS tensor object (ids and device addresses are recycled once a model is freed), and
    # while a CUDA graph is being captured the cast must become part of the graph (replays do not bump _version)
    if hit is not None and hit[2]() is t and hit[0] == ver and not torch.cuda.is_current_stream_capturing():
        return hit[1]
    h = t.detach().to(torch.half).contiguous()
    if len(_half_cache) > 256:
        _half_cache.clear()
    _half_cache[key] = (ver, h, weakref.ref(t))
    return h


class _FusedField(Function):
    @staticmethod
    def forward(ctx, xyzs, embeddings, w1, b1, w2, b2, w3, b3, offsets, S, H, gridtype, align_corners, bound, count):
        _cabi.require_cuda(xyzs, embeddings, w1, w2, w3, offsets)
        xyzs = xyzs.contiguous()
        if xyzs.dtype != torch.float32:
            xyzs = xyzs.float()
        M = xyzs.shape[0]
        dev = xyzs.device
        table = cached_half(embeddings)
        hw = [cached_half(t) for t in (w1, b1, w2, b2, w3, b3)]
        need_grad = any(ctx.needs_input_grad[1:8])
        sigma = torch.empty(M, device=dev, dtype=torch.float32)
        rgb = torch.empty(M, 3, device=dev, dtype=torch.float32)
        # saves for the backward: whole 128-sample tiles in the kernels' own (tile-major) layout - opaque to the host
        Mp = (M + 127) // 128 * 128
        enc = torch.empty(Mp, 32, device=dev, dtype=torch.half) if need_grad else None
        h1 = torch.empty(Mp, 64, device=dev, dtype=torch.half) if need_grad else None
        h2 = torch.empty(Mp, 64, device=dev, dtype=torch.half) if need_grad else None
        L = offsets.shape[0] - 1
        _cabi.call("ngp_field_forward", dev, _cabi.ptr(xyzs), M, _cabi.ptr(count), _cabi.ptr(table), _cabi.ptr(offsets), L,
                   embeddings.shape[1], float(S), int(H), int(gridtype), int(bool(align_corners)), float(bound),
                   *[_cabi.ptr(t) for t in hw], w1.shape[0], w3.shape[0], _cabi.ptr(sigma), _cabi.ptr(rgb), _cabi.ptr(enc),
                   _cabi.ptr(h1), _cabi.ptr(h2))
        if need_grad:
            ctx.save_for_backward(xyzs, offsets, sigma, rgb, enc, h1, h2, hw[0], hw[2], hw[4], count)
            ctx.meta = (M, L, embeddings.shape, float(S), int(H), int(gridtype), bool(align_corners), float(bound),
                        embeddings.dtype, w1.dtype)
        return sigma, rgb

    @staticmethod
    def backward(ctx, d_sigma, d_rgb):
        xyzs, offsets, sigma, rgb, enc, h1, h2, w1h, w2h, w3h, count = ctx.saved_tensors
        M, L, emb_shape, S, H, gridtype, align, bound, emb_dtype, w_dtype = ctx.meta
        dev = xyzs.device
        d_sigma = d_sigma.contiguous().float() if d_sigma is not None else torch.zeros(M, device=dev)
        d_rgb = d_rgb.contiguous().float() if d_rgb is not None else torch.zeros(M, 3, device=dev)
        # with a device-side row count the tail rows are never written by the kernel: keep them zero for the scatter
        d_enc = (torch.empty if count is None else torch.zeros)(M, 32, device=dev, dtype=torch.half)
        # one zero-filled buffer for all MLP gradients: gw1[64,32] gb1[64] gw2[64,64] gb2[64] gw3[4,64] gb3[4]
        sizes = [64 * 32, 64, 64 * 64, 64, 4 * 64, 4]
        flat = torch.zeros(sum(sizes), device=dev, dtype=torch.float32)
        gw1, gb1, gw2, gb2, gw3, gb3 = torch.split(flat, sizes)
        _cabi.call("ngp_field_backward", dev, M, _cabi.ptr(count), _cabi.ptr(w1h), _cabi.ptr(w2h), _cabi.ptr(w3h), 64, 4,
                   _cabi.ptr(d_sigma), _cabi.ptr(d_rgb), _cabi.ptr(sigma), _cabi.ptr(rgb), _cabi.ptr(enc), _cabi.ptr(h1),
                   _cabi.ptr(h2), _cabi.ptr(d_enc), _cabi.ptr(gw1), _cabi.ptr(gb1), _cabi.ptr(gw2), _cabi.ptr(gb2),
                   _cabi.ptr(gw3), _cabi.ptr(gb3))
        grad_table = None
        if ctx.needs_input_grad[1]:
            x01 = (xyzs + bound) / (2 * bound)
            grad_table = torch.zeros(emb_shape, device=dev, dtype=torch.float32)
            _cabi.call("ngp_grid_encode_backward", dev, _cabi.ptr(d_enc), _cabi.ptr(x01), None, _cabi.ptr(offsets),
                       _cabi.ptr(grad_table), M, 3, emb_shape[1], L, S, H, None, None, gridtype, int(align),
                       _cabi.NGP_F16, _cabi.LAYOUT_BLC, _cabi.NGP_F32)
            if grad_table.dtype != emb_dtype:
                grad_table = grad_table.to(emb_dtype)
        return (None, grad_table, gw1.view(64, 32), gb1, gw2.view(64, 64), gb2, gw3.view(4, 64), gb3, None, None, None, None,
                None, None, None)


def fused_field(xyzs, encoder, sigma_net, bound, count=None):
    """sigma [M] fp32, albedo [M,3] fp32 (fp16-rounded) for xyzs [M,3] in [-bound, bound]."""
    l0, l1, l2 = sigma_net.net
    return _FusedField.apply(xyzs, encoder.embeddings, l0.weight, l0.bias, l1.weight, l1.bias, l2.weight, l2.bias,
                             encoder.offsets, float(np.log2(encoder.per_level_scale)), encoder.base_resolution,
                             encoder.gridtype_id, encoder.align_corners, bound, count)


@torch.no_grad()
def fused_field_into(xyzs, encoder, sigma_net, bound, sigma_out, rgb_out):
    """Inference-only fused field: writes sigma [M] / albedo [M,3] (fp32) into caller-owned buffers, allocates nothing."""
    M = xyzs.shape[0]
    if sigma_out.shape[0] < M or rgb_out.shape[0] < M or not xyzs.is_contiguous() or xyzs.dtype != torch.float32:
        raise RuntimeError("fused_field_into: contiguous fp32 xyzs and output buffers of at least M rows expected")
    l0, l1, l2 = sigma_net.net
    dev = xyzs.device
    hw = [cached_half(t) for t in (l0.weight, l0.bias, l1.weight, l1.bias, l2.weight, l2.bias)]
    table = cached_half(encoder.embeddings)
    _cabi.call("ngp_field_forward", dev, _cabi.ptr(xyzs), M, None, _cabi.ptr(table), _cabi.ptr(encoder.offsets),
               encoder.offsets.shape[0] - 1, encoder.embeddings.shape[1], float(np.log2(encoder.per_level_scale)),
               int(encoder.base_resolution), int(encoder.gridtype_id), int(bool(encoder.align_corners)), float(bound),
               *[_cabi.ptr(t) for t in hw], l0.weight.shape[0], l2.weight.shape[0], _cabi.ptr(sigma_out), _cabi.ptr(rgb_out),
               None, None, None)


def can_fuse(x, encoder, sigma_net):
    """The fused kernels are built for the reference's own field shape under fp16 autocast on a CUDA tensor."""
    try:
        return (x.is_cuda and torch.is_autocast_enabled('cuda') and torch.get_autocast_dtype('cuda') == torch.float16
                and not x.requires_grad and x.dim() == 2 and x.shape[-1] == 3
                and encoder.num_levels == 16 and encoder.level_dim == 2 and encoder.input_dim == 3
                and sigma_net.num_layers == 3 and sigma_net.dim_hidden == 64 and sigma_net.dim_in == 32
                and sigma_net.dim_out == 4 and sigma_net.net[0].bias is not None
                and encoder.embeddings.dtype == torch.float32)
    except AttributeError:
        return False
