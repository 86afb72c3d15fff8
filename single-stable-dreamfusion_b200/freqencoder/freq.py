"""Drop-in replacement for the reference's ``freqencoder/freq.py`` (freq.py:15-77)."""
import torch
import torch.nn as nn
from torch.autograd import Function

from ngp_b200 import _cabi


class _freq_encoder(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type='cuda', cast_inputs=torch.float32)  # fp32 for precision, as freq.py:17
    def forward(ctx, inputs, degree, output_dim):
        # inputs: [B, input_dim] float -> [B, output_dim] float
        if not inputs.is_cuda:
            inputs = inputs.cuda()
        inputs = inputs.contiguous()
        B, input_dim = inputs.shape
        outputs = torch.empty(B, output_dim, dtype=inputs.dtype, device=inputs.device)
        _cabi.call("ngp_freq_encode_forward", inputs.device, _cabi.ptr(inputs), B, input_dim, degree, output_dim,
                   _cabi.ptr(outputs))
        ctx.save_for_backward(inputs, outputs)
        ctx.dims = [B, input_dim, degree, output_dim]
        return outputs

    @staticmethod
    @torch.amp.custom_bwd(device_type='cuda')
    def backward(ctx, grad):
        grad = grad.contiguous()
        inputs, outputs = ctx.saved_tensors
        B, input_dim, degree, output_dim = ctx.dims
        grad_inputs = torch.empty_like(inputs)
        _cabi.call("ngp_freq_encode_backward", grad.device, _cabi.ptr(grad), _cabi.ptr(outputs), B, input_dim, degree,
                   output_dim, _cabi.ptr(grad_inputs))
        return grad_inputs, None, None


freq_encode = _freq_encoder.apply


class FreqEncoder(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim = input_dim
        self.degree = degree
        self.output_dim = input_dim + input_dim * 2 * degree

    def __repr__(self):
        return f"FreqEncoder: input_dim={self.input_dim} degree={self.degree} output_dim={self.output_dim}"

    def forward(self, inputs, **kwargs):
        # inputs: [..., input_dim] -> [..., output_dim]
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.reshape(-1, self.input_dim)
        outputs = freq_encode(inputs, self.degree, self.output_dim)
        return outputs.reshape(prefix_shape + [self.output_dim])
