// Occupancy-grid maintenance: jittered cell-centre query points, EMA-max update, mean density and
// bitfield packing.  Behavioural contract: NeRFRenderer.update_extra_state, nerf/renderer.py:562-613
// of the reference, which does this with ~15 eager torch ops, 25-50 MB temporaries and two host syncs.
// Here: one kernel builds the query points directly in Morton order (no coords / indices / scatter),
// one fused kernel does the EMA-max + reduction, one packs the bitfield with the threshold read from
// device memory (no .item()).
#include "common.cuh"

namespace ngp {
namespace occ {

NGP_DEVINL uint32_t compact3(uint32_t x) {
    x &= 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

// Query point of Morton cell m.  Every arithmetic step is a separately rounded fp32 op because the
// reference evaluates it as a chain of eager torch kernels (renderer.py:584-593); torch's CUDA
// division by a python scalar is a multiply by the fp32 reciprocal, hence inv_hm1.
__global__ void __launch_bounds__(256) cell_points_kernel(uint32_t H, float cell_scale, float half_cell,
                                                          const float* __restrict__ noise, float* __restrict__ xyzs) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_cells = H * H * H;
    if (m >= n_cells) return;
    const uint32_t c[3] = {compact3(m), compact3(m >> 1), compact3(m >> 2)};
    const size_t lin = ((size_t)c[0] * H + c[1]) * H + c[2];  // the order torch.rand_like is consumed in
    const float inv_hm1 = __fdiv_rn(1.0f, (float)(H - 1));
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float centre = __fsub_rn(__fmul_rn(__fmul_rn(2.0f, (float)c[a]), inv_hm1), 1.0f);
        const float scaled = __fmul_rn(centre, cell_scale);
        const float jitter = __fmul_rn(__fsub_rn(__fmul_rn(noise[lin * 3 + a], 2.0f), 1.0f), half_cell);
        xyzs[(size_t)m * 3 + a] = __fadd_rn(scaled, jitter);
    }
}

// The same query points for the Morton cells [first, first + count) only, jitter taken from noise[count, 3] indexed by
// m - first: one rank's share of a data-parallel occupancy refresh (every rank draws its own jitter for its own cells).
__global__ void __launch_bounds__(256) cell_points_range_kernel(uint32_t H, float cell_scale, float half_cell,
                                                                const float* __restrict__ noise, uint32_t first, uint32_t count,
                                                                float* __restrict__ xyzs) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const uint32_t m = first + k;
    const uint32_t c[3] = {compact3(m), compact3(m >> 1), compact3(m >> 2)};
    const float inv_hm1 = __fdiv_rn(1.0f, (float)(H - 1));
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float centre = __fsub_rn(__fmul_rn(__fmul_rn(2.0f, (float)c[a]), inv_hm1), 1.0f);
        const float scaled = __fmul_rn(centre, cell_scale);
        const float jitter = __fmul_rn(__fsub_rn(__fmul_rn(noise[(size_t)k * 3 + a], 2.0f), 1.0f), half_cell);
        xyzs[(size_t)k * 3 + a] = __fadd_rn(scaled, jitter);
    }
}

struct Accum {
    double sum;
    unsigned long long count;
};

// grid = where(grid >= 0, maximum(grid * decay, tmp), grid); accumulate sum/count of the valid cells.
__global__ void __launch_bounds__(256) ema_max_kernel(float* __restrict__ grid, const float* __restrict__ tmp, uint32_t n,
                                                      float decay, Accum* __restrict__ acc) {
    double local = 0.0;
    unsigned int cnt = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float g = grid[i];
        if (g >= 0) {
            const float a = __fmul_rn(g, decay), b = tmp[i];
            g = (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b);  // torch.maximum propagates NaN
            grid[i] = g;
            local += (double)g;
            ++cnt;
        }
    }
    // block reduction (double sum, integer count)
    __shared__ double s_sum[8];
    __shared__ unsigned int s_cnt[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        local += __shfl_xor_sync(0xffffffffu, local, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) { s_sum[warp] = local; s_cnt[warp] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double bs = 0.0;
        unsigned long long bc = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { bs += s_sum[w]; bc += s_cnt[w]; }
        atomicAdd(&acc->sum, bs);
        atomicAdd(&acc->count, bc);
    }
}

__global__ void __launch_bounds__(256) threshold_pack_kernel(const float* __restrict__ grid, uint32_t n_bytes,
                                                             float density_thresh, const Accum* __restrict__ acc,
                                                             float* __restrict__ mean_out, uint8_t* __restrict__ bitfield) {
    const float mean = (float)(acc->sum / (double)acc->count);  // 0/0 -> NaN, as torch.mean of an empty selection
    // python: min(mean, density_thresh) returns `mean` unless density_thresh < mean (NaN-preserving)
    const float thresh = (density_thresh < mean) ? density_thresh : mean;
    if (blockIdx.x == 0 && threadIdx.x == 0) *mean_out = mean;
    const uint32_t nb = blockIdx.x * blockDim.x + threadIdx.x;
    if (nb >= n_bytes) return;
    const float4 a = __ldg(reinterpret_cast<const float4*>(grid) + 2 * (size_t)nb);
    const float4 b = __ldg(reinterpret_cast<const float4*>(grid) + 2 * (size_t)nb + 1);
    uint32_t bits = 0;
    bits |= (a.x > thresh) ? 1u : 0u;   bits |= (a.y > thresh) ? 2u : 0u;
    bits |= (a.z > thresh) ? 4u : 0u;   bits |= (a.w > thresh) ? 8u : 0u;
    bits |= (b.x > thresh) ? 16u : 0u;  bits |= (b.y > thresh) ? 32u : 0u;
    bits |= (b.z > thresh) ? 64u : 0u;  bits |= (b.w > thresh) ? 128u : 0u;
    bitfield[nb] = (uint8_t)bits;
}

}  // namespace occ
}  // namespace ngp

using namespace ngp;

extern "C" int ngp_occupancy_cell_points(uint32_t H, float cell_scale, float half_cell, const float* noise, float* xyzs,
                                         void* stream) {
    if (!noise || !xyzs || H < 2 || H > 1024) return NGP_ERR_BAD_ARG;
    const uint32_t n = H * H * H;
    occ::cell_points_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(H, cell_scale, half_cell, noise, xyzs);
    return launch_status();
}

extern "C" int ngp_occupancy_cell_points_range(uint32_t H, float cell_scale, float half_cell, const float* noise, uint32_t first,
                                               uint32_t count, float* xyzs, void* stream) {
    if (!noise || !xyzs || H < 2 || H > 1024) return NGP_ERR_BAD_ARG;
    if ((uint64_t)first + count > (uint64_t)H * H * H) return NGP_ERR_BAD_ARG;
    if (count == 0) return NGP_OK;
    occ::cell_points_range_kernel<<<cdiv(count, 256), 256, 0, as_stream(stream)>>>(H, cell_scale, half_cell, noise, first, count, xyzs);
    return launch_status();
}

extern "C" int ngp_update_density_grid(float* grid, const float* tmp_grid, uint32_t n_cells, float decay,
                                       float density_thresh, float* mean_out, uint8_t* bitfield, void* workspace,
                                       uint64_t workspace_bytes, void* stream) {
    if (!grid || !tmp_grid || !mean_out || !bitfield || (n_cells % 8) != 0) return NGP_ERR_BAD_ARG;
    if (!workspace || workspace_bytes < sizeof(occ::Accum)) return NGP_ERR_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    occ::Accum* acc = static_cast<occ::Accum*>(workspace);
    cudaMemsetAsync(acc, 0, sizeof(occ::Accum), st);
    if (n_cells == 0) return launch_status();
    const int blocks = min(cdiv(n_cells, 256), num_sms() * 8);
    occ::ema_max_kernel<<<blocks, 256, 0, st>>>(grid, tmp_grid, n_cells, decay, acc);
    occ::threshold_pack_kernel<<<cdiv(n_cells / 8, 256), 256, 0, st>>>(grid, n_cells / 8, density_thresh, acc, mean_out, bitfield);
    return launch_status();
}
