// Frequency (sin/cos) positional encoding, forward + backward.
//
// Behavioural contract: freqencoder/src/freqencoder.cu:30-94 of the reference.  The reference builds
// that file with -use_fast_math (freqencoder/backend.py:9), so this translation unit - and only this
// one - is built with --use_fast_math too: sin.approx via __sinf, flush-to-zero, and the same FMA
// contraction, which keeps the encoding bit-equal to the reference on the same GPU.
//
// Layout change: one thread produces one OUTPUT element in the forward (coalesced stores, the input
// row is re-read from L1), and one thread owns one input element in the backward.
#include "common.cuh"

namespace ngp {
namespace freq {

__global__ void __launch_bounds__(256) freq_forward_kernel(const float* __restrict__ inputs, uint32_t B, uint32_t D,
                                                           uint32_t deg, uint32_t C, float* __restrict__ outputs) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)B * C) return;
    const uint32_t b = (uint32_t)(t / C);
    const uint32_t c = (uint32_t)(t - (uint64_t)b * C);
    const float* in = inputs + (size_t)b * D;
    if (c < D) {  // identity block (freqencoder.cu:48)
        outputs[t] = in[c];
        return;
    }
    // column blocks after the identity: sin(2^0 x), cos(2^0 x), sin(2^1 x), ... (freqencoder.cu:52-56);
    // cos is sin shifted by pi/2.
    const uint32_t col = c / D - 1;
    const uint32_t d = c % D;
    const uint32_t octave = col / 2;
    const float phase = (col % 2) * (3.141592653589793f / 2);
    outputs[t] = __sinf(scalbnf(in[d], octave) + phase);
}

// d/dx [x, sin(2^f x), cos(2^f x)] = [1, 2^f cos, -2^f sin], read back from the saved outputs
// (freqencoder.cu:81-90).
__global__ void __launch_bounds__(256) freq_backward_kernel(const float* __restrict__ grad, const float* __restrict__ outputs,
                                                            uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
                                                            float* __restrict__ grad_inputs) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)B * D) return;
    const uint32_t b = (uint32_t)(t / D);
    const uint32_t d = (uint32_t)(t - (uint64_t)b * D);
    const float* g = grad + (size_t)b * C;
    const float* o = outputs + (size_t)b * C;
    float acc = g[d];
    g += D;
    o += D;
    for (uint32_t f = 0; f < deg; ++f) {
        acc += scalbnf(1.0f, f) * (g[d] * o[D + d] - g[D + d] * o[d]);
        g += 2 * D;
        o += 2 * D;
    }
    grad_inputs[t] = acc;
}

}  // namespace freq
}  // namespace ngp

using namespace ngp;

extern "C" int ngp_freq_encode_forward(const float* inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
                                       float* outputs, void* stream) {
    if (!inputs || !outputs) return NGP_ERR_BAD_ARG;
    if (C != D + D * 2 * deg) return NGP_ERR_BAD_ARG;
    if (B == 0) return NGP_OK;
    freq::freq_forward_kernel<<<cdiv((uint64_t)B * C, 256), 256, 0, as_stream(stream)>>>(inputs, B, D, deg, C, outputs);
    return launch_status();
}

extern "C" int ngp_freq_encode_backward(const float* grad, const float* outputs, uint32_t B, uint32_t D, uint32_t deg,
                                        uint32_t C, float* grad_inputs, void* stream) {
    if (!grad || !outputs || !grad_inputs) return NGP_ERR_BAD_ARG;
    if (C != D + D * 2 * deg) return NGP_ERR_BAD_ARG;
    if (B == 0) return NGP_OK;
    freq::freq_backward_kernel<<<cdiv((uint64_t)B * D, 256), 256, 0, as_stream(stream)>>>(grad, outputs, B, D, deg, C,
                                                                                          grad_inputs);
    return launch_status();
}
