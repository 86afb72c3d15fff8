// C-ABI entry points of the grid encoder; kernels live in grid_encode.cuh, one TU per input dimension.
#include "grid_encode.cuh"

using namespace ngp;

namespace ngp { namespace grid {
bool g_disable_warpagg = false;
int g_scatter_max_run = 8;                          // ngp_grid_set_option(2, x): longest run of lanes summed before the reds
bool g_count_reds = false;                          // ngp_grid_set_option(1, x): the counting build of the scatter
__device__ unsigned long long g_red_lane_ops = 0;   // post-aggregation red instructions x active lanes, ngp_grid_red_count
} }

extern "C" int ngp_grid_set_option(int option, int value) {
    if (option == 0) { grid::g_disable_warpagg = (value != 0); return NGP_OK; }
    if (option == 1) { grid::g_count_reds = (value != 0); return NGP_OK; }
    if (option == 2 && (value == 4 || value == 8 || value == 16 || value == 32)) { grid::g_scatter_max_run = value; return NGP_OK; }
    return NGP_ERR_BAD_ARG;
}

extern "C" int ngp_grid_red_count(uint64_t* lane_ops, int reset) {
    if (!lane_ops) return NGP_ERR_BAD_ARG;
    unsigned long long v = 0;
    cudaError_t e = cudaMemcpyFromSymbol(&v, grid::g_red_lane_ops, sizeof(v));   // synchronises the device
    if (e != cudaSuccess) return (int)e;
    *lane_ops = v;
    if (reset) {
        v = 0;
        e = cudaMemcpyToSymbol(grid::g_red_lane_ops, &v, sizeof(v));
        if (e != cudaSuccess) return (int)e;
    }
    return NGP_OK;
}

extern "C" int ngp_grid_encode_forward(const float* inputs, const void* embeddings, const int* offsets, void* outputs,
                                       uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                       void* dy_dx, uint32_t gridtype, int align_corners, int dtype, int out_layout,
                                       void* stream) {
    if (!inputs || !embeddings || !offsets || !outputs) return NGP_ERR_BAD_ARG;
    if (L == 0 || L > grid::kMaxLevels || gridtype > 1) return NGP_ERR_UNSUPPORTED;
    if (out_layout != NGP_LAYOUT_LBC && out_layout != NGP_LAYOUT_BLC) return NGP_ERR_BAD_ARG;
    if (dtype != NGP_F32 && dtype != NGP_F16) return NGP_ERR_UNSUPPORTED;
    if (B == 0) return NGP_OK;
    cudaStream_t st = as_stream(stream);
    const bool al = align_corners != 0;
    switch (D) {  // gridencoder.cu:366-373
        case 1: return grid::forward_for_dim<1>(inputs, embeddings, offsets, outputs, dy_dx, B, C, L, S, H, gridtype, al, dtype, out_layout, st);
        case 2: return grid::forward_for_dim<2>(inputs, embeddings, offsets, outputs, dy_dx, B, C, L, S, H, gridtype, al, dtype, out_layout, st);
        case 3: return grid::forward_for_dim<3>(inputs, embeddings, offsets, outputs, dy_dx, B, C, L, S, H, gridtype, al, dtype, out_layout, st);
        case 4: return grid::forward_for_dim<4>(inputs, embeddings, offsets, outputs, dy_dx, B, C, L, S, H, gridtype, al, dtype, out_layout, st);
        case 5: return grid::forward_for_dim<5>(inputs, embeddings, offsets, outputs, dy_dx, B, C, L, S, H, gridtype, al, dtype, out_layout, st);
        default: return NGP_ERR_UNSUPPORTED;
    }
}

extern "C" int ngp_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings,
                                        const int* offsets, void* grad_embeddings, uint32_t B, uint32_t D, uint32_t C,
                                        uint32_t L, float S, uint32_t H, const void* dy_dx, void* grad_inputs,
                                        uint32_t gridtype, int align_corners, int dtype, int grad_layout,
                                        int grad_emb_dtype, void* stream) {
    (void)embeddings;  // the reference passes the table but its backward never reads it either
    if (!grad || !inputs || !offsets || !grad_embeddings) return NGP_ERR_BAD_ARG;
    if (L == 0 || L > grid::kMaxLevels || gridtype > 1) return NGP_ERR_UNSUPPORTED;
    if (grad_layout != NGP_LAYOUT_LBC && grad_layout != NGP_LAYOUT_BLC) return NGP_ERR_BAD_ARG;
    if (dtype != NGP_F32 && dtype != NGP_F16) return NGP_ERR_UNSUPPORTED;
    if (B == 0) return NGP_OK;
    cudaStream_t st = as_stream(stream);
    const bool al = align_corners != 0;
    switch (D) {
        case 1: return grid::backward_for_dim<1>(grad, inputs, offsets, grad_embeddings, dy_dx, grad_inputs, B, C, L, S, H, gridtype, al, dtype, grad_layout, grad_emb_dtype, st);
        case 2: return grid::backward_for_dim<2>(grad, inputs, offsets, grad_embeddings, dy_dx, grad_inputs, B, C, L, S, H, gridtype, al, dtype, grad_layout, grad_emb_dtype, st);
        case 3: return grid::backward_for_dim<3>(grad, inputs, offsets, grad_embeddings, dy_dx, grad_inputs, B, C, L, S, H, gridtype, al, dtype, grad_layout, grad_emb_dtype, st);
        case 4: return grid::backward_for_dim<4>(grad, inputs, offsets, grad_embeddings, dy_dx, grad_inputs, B, C, L, S, H, gridtype, al, dtype, grad_layout, grad_emb_dtype, st);
        case 5: return grid::backward_for_dim<5>(grad, inputs, offsets, grad_embeddings, dy_dx, grad_inputs, B, C, L, S, H, gridtype, al, dtype, grad_layout, grad_emb_dtype, st);
        default: return NGP_ERR_UNSUPPORTED;
    }
}

extern "C" int ngp_grid_level_params(uint32_t L, float S, uint32_t H, float* scales, uint32_t* resolutions,
                                     void* stream) {
    if (!scales || !resolutions) return NGP_ERR_BAD_ARG;
    if (L == 0) return NGP_OK;
    grid::level_params_kernel<<<cdiv(L, 64), 64, 0, as_stream(stream)>>>(L, S, H, scales, resolutions);
    return launch_status();
}

// Sync-free training path: scatter the MLP's encoding gradients of the first *count_ptr samples.
// grad_table_odd (optional): the odd-frame twin of the gradient table (see warpagg_level) - an fp32 buffer indexed like
// grad_table whose ADDRESS is 8 bytes off a 16-byte boundary; the table gradient is then grad_table + grad_table_odd.
static int scatter_samples(const void* d_enc, const float* xyzs, float bound, const int* count_ptr, uint32_t M_cap,
                           const int* offsets, uint32_t L, uint32_t C, float S, uint32_t H, uint32_t gridtype, int align_corners,
                           float* grad_table, float* grad_table_odd, void* stream) {
    if (!d_enc || !xyzs || !offsets || !grad_table) return NGP_ERR_BAD_ARG;
    if (L == 0 || L > grid::kMaxLevels || gridtype > 1 || C != 2) return NGP_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(grad_table) & 15) != 0) return NGP_ERR_BAD_ARG;
    if (grad_table_odd && (reinterpret_cast<uintptr_t>(grad_table_odd) & 15) != 8) return NGP_ERR_BAD_ARG;
    if (M_cap == 0) return NGP_OK;
    const int blocks = min(cdiv(M_cap, 256), num_sms() * 16);
    if (grid::g_count_reds) {   // measurement only (bench.py's roofline pass): same kernel, plus one counter per lane
        unsigned long long* ctr = nullptr;
        if (cudaGetSymbolAddress(reinterpret_cast<void**>(&ctr), grid::g_red_lane_ops) != cudaSuccess) return launch_status();
        grid::encode_backward_warpagg_kernel<__half, 2, true><<<blocks, 256, 0, as_stream(stream)>>>(
            static_cast<const __half*>(d_enc), xyzs, offsets, grad_table, M_cap, L, S, H, gridtype, align_corners != 0, count_ptr,
            bound, ctr, grad_table_odd, (uint32_t)grid::g_scatter_max_run);
        return launch_status();
    }
    grid::encode_backward_warpagg_kernel<__half, 2><<<blocks, 256, 0, as_stream(stream)>>>(
        static_cast<const __half*>(d_enc), xyzs, offsets, grad_table, M_cap, L, S, H, gridtype, align_corners != 0, count_ptr, bound,
        nullptr, grad_table_odd, (uint32_t)grid::g_scatter_max_run);
    return launch_status();
}

extern "C" int ngp_grid_scatter_samples(const void* d_enc, const float* xyzs, float bound, const int* count_ptr, uint32_t M_cap,
                                        const int* offsets, uint32_t L, uint32_t C, float S, uint32_t H, uint32_t gridtype,
                                        int align_corners, float* grad_table, void* stream) {
    return scatter_samples(d_enc, xyzs, bound, count_ptr, M_cap, offsets, L, C, S, H, gridtype, align_corners, grad_table, nullptr,
                           stream);
}

extern "C" int ngp_grid_scatter_samples_split(const void* d_enc, const float* xyzs, float bound, const int* count_ptr,
                                              uint32_t M_cap, const int* offsets, uint32_t L, uint32_t C, float S, uint32_t H,
                                              uint32_t gridtype, int align_corners, float* grad_table, float* grad_table_odd,
                                              void* stream) {
    if (!grad_table_odd) return NGP_ERR_BAD_ARG;
    return scatter_samples(d_enc, xyzs, bound, count_ptr, M_cap, offsets, L, C, S, H, gridtype, align_corners, grad_table,
                           grad_table_odd, stream);
}

namespace ngp { namespace grid {
// grad_table += grad_table_odd; grad_table_odd = 0   (n floats, n even; float2 granularity: the twin is 8-byte aligned)
__global__ void __launch_bounds__(256) fold_odd_kernel(float2* __restrict__ dst, float2* __restrict__ odd, uint64_t n2) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (uint64_t)gridDim.x * blockDim.x) {
        const float2 o = odd[i];
        if (o.x != 0.f || o.y != 0.f) {
            float2 d = dst[i];
            d.x += o.x; d.y += o.y;
            dst[i] = d;
            odd[i] = make_float2(0.f, 0.f);
        }
    }
}
} }

extern "C" int ngp_grid_fold_odd(float* grad_table, float* grad_table_odd, uint64_t n, void* stream) {
    if (!grad_table || !grad_table_odd || (n & 1)) return NGP_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(grad_table) & 7) != 0 || (reinterpret_cast<uintptr_t>(grad_table_odd) & 7) != 0) return NGP_ERR_BAD_ARG;
    if (n == 0) return NGP_OK;
    const uint64_t n2 = n / 2;
    const uint64_t want = (n2 + 255) / 256, most = (uint64_t)num_sms() * 8;
    const int blocks = (int)(want < most ? want : most);
    grid::fold_odd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<float2*>(grad_table),
                                                               reinterpret_cast<float2*>(grad_table_odd), n2);
    return launch_status();
}
