// L2 roofline micro-benchmarks: the hash-grid encoder is bound by random 4-byte gathers (forward) and
// random 8-byte red.add (backward) into an L2-resident table, and MEASURED_PEAKS.json has no L2 figure.
// These kernels measure the chip's ceiling for exactly those access shapes; bench.py reports the
// encoder against them.
#include "common.cuh"

namespace ngp {
namespace ubench {

NGP_DEVINL uint32_t mix(uint32_t x) {  // xorshift-multiply: cheap, full-period enough for address noise
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

__global__ void __launch_bounds__(256) gather4_kernel(const uint32_t* __restrict__ table, uint32_t mask, uint32_t* sink,
                                                      uint32_t n_threads, uint32_t iters, uint32_t seed) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_threads) return;
    uint32_t s = mix(t ^ seed), acc = 0;
    for (uint32_t it = 0; it < iters; ++it) {
        uint32_t a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { s = mix(s + 0x9e3779b9u); a[k] = s & mask; }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += __ldg(table + a[k]);
    }
    sink[t] = acc;
}

__global__ void __launch_bounds__(256) red8_kernel(float* table, uint32_t mask, uint32_t n_threads, uint32_t iters,
                                                   uint32_t seed) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_threads) return;
    uint32_t s = mix(t ^ seed);
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            s = mix(s + 0x9e3779b9u);
            red_add_f32x2(table + 2 * (size_t)(s & mask), 1.0f, 0.5f);
        }
    }
}

// red.global.add of WIDTH consecutive floats per op; only lanes with (lane % lane_stride == 0) are active
template <int WIDTH>
__global__ void __launch_bounds__(256) red_width_kernel(float* table, uint32_t mask, uint32_t n_threads, uint32_t iters,
                                                        uint32_t seed, uint32_t lane_stride) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_threads || (threadIdx.x % lane_stride) != 0) return;
    uint32_t s = mix(t ^ seed);
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            s = mix(s + 0x9e3779b9u);
            float* dst = table + WIDTH * (size_t)(s & mask);
            if (WIDTH == 1) red_add_f32(dst, 1.0f);
            else if (WIDTH == 2) red_add_f32x2(dst, 1.0f, 0.5f);
            else red_add_f32x4(dst, 1.0f, 0.5f, 0.25f, 2.0f);
        }
    }
}

// one thread, one store: the device's nanosecond clock (%globaltimer) at the moment the stream reaches this node
__global__ void stamp_kernel(unsigned long long* slot) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    *slot = t;
}

}  // namespace ubench
}  // namespace ngp

using namespace ngp;

extern "C" int ngp_bench_gather4(const uint32_t* table, uint32_t table_words, uint32_t* sink, uint32_t n_threads,
                                 uint32_t iters, uint32_t seed, void* stream) {
    if (!table || !sink || table_words == 0 || (table_words & (table_words - 1))) return NGP_ERR_BAD_ARG;
    if (n_threads == 0) return NGP_OK;
    ubench::gather4_kernel<<<cdiv(n_threads, 256), 256, 0, as_stream(stream)>>>(table, table_words - 1, sink, n_threads, iters,
                                                                              seed);
    return launch_status();
}

extern "C" int ngp_bench_red8(float* table, uint32_t table_words, uint32_t n_threads, uint32_t iters, uint32_t seed,
                              void* stream) {
    if (!table || table_words < 2 || (table_words & (table_words - 1))) return NGP_ERR_BAD_ARG;
    if (n_threads == 0) return NGP_OK;
    ubench::red8_kernel<<<cdiv(n_threads, 256), 256, 0, as_stream(stream)>>>(table, table_words / 2 - 1, n_threads, iters, seed);
    return launch_status();
}

extern "C" int ngp_bench_red_width(float* table, uint32_t table_words, uint32_t n_threads, uint32_t iters, uint32_t seed,
                                   uint32_t width, uint32_t lane_stride, void* stream) {
    if (!table || table_words < 4 || (table_words & (table_words - 1)) || lane_stride == 0) return NGP_ERR_BAD_ARG;
    if (n_threads == 0) return NGP_OK;
    const dim3 grid(cdiv(n_threads, 256)), block(256);
    cudaStream_t st = as_stream(stream);
    if (width == 1) ubench::red_width_kernel<1><<<grid, block, 0, st>>>(table, table_words - 1, n_threads, iters, seed, lane_stride);
    else if (width == 2) ubench::red_width_kernel<2><<<grid, block, 0, st>>>(table, table_words / 2 - 1, n_threads, iters, seed, lane_stride);
    else if (width == 4) ubench::red_width_kernel<4><<<grid, block, 0, st>>>(table, table_words / 4 - 1, n_threads, iters, seed, lane_stride);
    else return NGP_ERR_UNSUPPORTED;
    return launch_status();
}

extern "C" int ngp_stamp(uint64_t* slot, void* stream) {
    if (!slot) return NGP_ERR_BAD_ARG;
    ubench::stamp_kernel<<<1, 1, 0, as_stream(stream)>>>(reinterpret_cast<unsigned long long*>(slot));
    return launch_status();
}
