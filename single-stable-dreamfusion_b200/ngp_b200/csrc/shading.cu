// Shading path of the grid field as a 7-point stencil: finite-difference normals + Lambertian / textureless /
// normal colouring, forward and backward.
//
// Behavioural contract: nerf/network_grid.py:90-144.  The reference evaluates the field SEVEN times per shaded sample -
// common_forward(x) plus common_forward((x +- eps e_i).clamp(-bound, bound)) for the three axes (:93-98), each a separate
// encoder launch + 3 GEMMs + ~10 elementwise kernels - and SIX more for the smoothness regulariser's normal(x + noise)
// (nerf/renderer.py:491).  Here the stencil points of a sample are laid out as K CONSECUTIVE rows of ONE batch
// (K = 7 with the centre, 6 without): one fused-field launch evaluates all of them, the K points of a sample sit within
// 2 eps = 0.02 of each other, so at every level coarser than that they hit the same grid cell - the corner fetches come
// from L1 and the warp-aggregated scatter of the backward merges their updates before they reach L2 - and two small
// kernels here turn the K densities into normal / colour and back-propagate through them.
//
// Arithmetic under fp16 autocast, mirrored: densities are fp32 (trunc_exp returns float); grad = 0.5 * (s+ - s-) / eps
// in fp32 (a division by a python scalar multiplies by the fp32 reciprocal); safe_normalize clamps the squared norm at
// 1e-20, NaNs -> 0; `normal @ l` runs as a half matrix-vector product (operands rounded to half, fp32 accumulate, half
// result); ratio + (1 - ratio) * clamp(., 0) and albedo * lambertian are half operations (each rounds to half).
#include "common.cuh"

namespace ngp {
namespace shade {

constexpr float kRcpEps100 = 100.0f;   // fl(1 / fl(0.01)) in fp32

NGP_DEVINL float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }
NGP_DEVINL float h16(float x) { return __half2float(__float2half_rn(x)); }

// K rows per sample: [x,] x + eps e_x, x - eps e_x, x + eps e_y, ... (the reference's dx_pos, dx_neg, dy_pos, ... order)
__global__ void stencil_points_kernel(const float* __restrict__ xyzs, uint32_t M, float eps, float bound, int with_centre,
                                      float* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const float x[3] = {xyzs[(size_t)i * 3], xyzs[(size_t)i * 3 + 1], xyzs[(size_t)i * 3 + 2]};
    const uint32_t K = with_centre ? 7u : 6u;
    float* o = out + (size_t)i * K * 3;
    if (with_centre) { o[0] = x[0]; o[1] = x[1]; o[2] = x[2]; o += 3; }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            float p[3] = {x[0], x[1], x[2]};
            p[a] = __fadd_rn(p[a], s == 0 ? eps : -eps);
            // (x + offset).clamp(-bound, bound): all three coordinates are clamped (network_grid.py:93-98)
            o[0] = clampf(p[0], -bound, bound); o[1] = clampf(p[1], -bound, bound); o[2] = clampf(p[2], -bound, bound);
            o += 3;
        }
    }
}

struct Normal {
    float g[3];     // -0.5 (s+ - s-) / eps  (finite_difference_normal)
    float n[3];     // safe_normalize(g), NaN -> 0
    float inv;      // 1 / sqrt(max(|g|^2, 1e-20))
    bool clamped;   // |g|^2 < 1e-20: the norm is the constant 1e-10
    bool nan[3];
};
NGP_DEVINL Normal normal_of(const float* __restrict__ s6) {
    Normal r;
#pragma unroll
    for (int a = 0; a < 3; ++a) r.g[a] = -(__fmul_rn(__fmul_rn(0.5f, __fsub_rn(s6[2 * a], s6[2 * a + 1])), kRcpEps100));
    const float nn = __fadd_rn(__fadd_rn(__fmul_rn(r.g[0], r.g[0]), __fmul_rn(r.g[1], r.g[1])), __fmul_rn(r.g[2], r.g[2]));
    r.clamped = !(nn >= 1e-20f);          // (NaN compares false: treated like the clamp for the backward; values become 0)
    r.inv = 1.0f / sqrtf(fmaxf(nn, 1e-20f));
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float v = __fdiv_rn(r.g[a], sqrtf(fmaxf(nn, 1e-20f)));
        r.nan[a] = v != v;
        r.n[a] = r.nan[a] ? 0.f : v;      // normal[torch.isnan(normal)] = 0
    }
    return r;
}

// mode: 0 = lambertian, 1 = textureless, 2 = normal ; K = 7 (centre first) or 6 (normal only: color == nullptr)
__global__ void shade_forward_kernel(const float* __restrict__ sigma_all, const float* __restrict__ rgb_all, uint32_t M, uint32_t K,
                                     const float* __restrict__ light, float ratio, int mode, float* __restrict__ normal_out,
                                     float* __restrict__ color_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const float* s = sigma_all + (size_t)i * K + (K == 7 ? 1 : 0);
    float s6[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) s6[k] = s[k];
    const Normal nm = normal_of(s6);
    normal_out[(size_t)i * 3] = nm.n[0]; normal_out[(size_t)i * 3 + 1] = nm.n[1]; normal_out[(size_t)i * 3 + 2] = nm.n[2];
    if (!color_out) return;
    float* c = color_out + (size_t)i * 3;
    if (mode == 2) {                      // (normal + 1) / 2, fp32
#pragma unroll
        for (int a = 0; a < 3; ++a) c[a] = __fmul_rn(__fadd_rn(nm.n[a], 1.0f), 0.5f);
        return;
    }
    // lambertian = ratio + (1 - ratio) * (normal @ l).clamp(min=0), in half (see the header)
    float dot = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) dot = fmaf(h16(nm.n[a]), h16(light[a]), dot);
    const float dh = fmaxf(h16(dot), 0.f);
    const float lam = h16(ratio + h16((1.0f - ratio) * dh));
    if (mode == 1) { c[0] = lam; c[1] = lam; c[2] = lam; return; }
    const float* alb = rgb_all + (size_t)i * K * 3;    // the centre row's albedo (fp32 holding half values)
#pragma unroll
    for (int a = 0; a < 3; ++a) c[a] = h16(alb[a] * lam);
}

// d_sigma_all [M*K] (the centre's entry, k = 0 of K = 7, is written as 0: its own gradient arrives through the sigma
// output), d_rgb_all [M*K,3] (non-zero only in the centre row, lambertian mode); either may be nullptr.
__global__ void shade_backward_kernel(const float* __restrict__ sigma_all, const float* __restrict__ rgb_all, uint32_t M, uint32_t K,
                                      const float* __restrict__ light, float ratio, int mode, const float* __restrict__ d_normal,
                                      const float* __restrict__ d_color, float* __restrict__ d_sigma_all,
                                      float* __restrict__ d_rgb_all) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const uint32_t c0 = (K == 7 ? 1u : 0u);
    const float* s = sigma_all + (size_t)i * K + c0;
    float s6[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) s6[k] = s[k];
    const Normal nm = normal_of(s6);
    float dn[3] = {0.f, 0.f, 0.f};
    if (d_normal) { dn[0] = d_normal[(size_t)i * 3]; dn[1] = d_normal[(size_t)i * 3 + 1]; dn[2] = d_normal[(size_t)i * 3 + 2]; }
    float d_alb[3] = {0.f, 0.f, 0.f};
    if (d_color) {
        const float* dc = d_color + (size_t)i * 3;
        if (mode == 2) {
#pragma unroll
            for (int a = 0; a < 3; ++a) dn[a] += 0.5f * dc[a];
        } else {
            float dot = 0.f;
#pragma unroll
            for (int a = 0; a < 3; ++a) dot = fmaf(h16(nm.n[a]), h16(light[a]), dot);
            const float dh = h16(dot);
            const float lam = h16(ratio + h16((1.0f - ratio) * fmaxf(dh, 0.f)));
            float d_lam;
            if (mode == 1) {
                d_lam = dc[0] + dc[1] + dc[2];
            } else {
                const float* alb = rgb_all + (size_t)i * K * 3;
                d_lam = dc[0] * alb[0] + dc[1] * alb[1] + dc[2] * alb[2];
#pragma unroll
                for (int a = 0; a < 3; ++a) d_alb[a] = dc[a] * lam;
            }
            if (dh >= 0.f) {                  // clamp(min=0) passes the gradient where input >= 0 (torch: grad * (x >= min))
                const float d_dot = d_lam * (1.0f - ratio);
#pragma unroll
                for (int a = 0; a < 3; ++a) dn[a] += d_dot * light[a];
            }
        }
    }
    // through normal[isnan] = 0 and safe_normalize: n = g * inv, inv = max(|g|^2, 1e-20)^-1/2
    float dg[3];
    {
        float dnm[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) dnm[a] = nm.nan[a] ? 0.f : dn[a];
        if (nm.clamped) {
#pragma unroll
            for (int a = 0; a < 3; ++a) dg[a] = dnm[a] * nm.inv;
        } else {
            const float proj = nm.n[0] * dnm[0] + nm.n[1] * dnm[1] + nm.n[2] * dnm[2];
#pragma unroll
            for (int a = 0; a < 3; ++a) dg[a] = (dnm[a] - nm.n[a] * proj) * nm.inv;
        }
    }
    if (d_sigma_all) {
        float* ds = d_sigma_all + (size_t)i * K;
        if (K == 7) ds[0] = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float v = -0.5f * kRcpEps100 * dg[a];
            if (v != v) v = 0.f;
            ds[c0 + 2 * a] = v;
            ds[c0 + 2 * a + 1] = -v;
        }
    }
    if (d_rgb_all) {
        float* dr = d_rgb_all + (size_t)i * K * 3;
        for (uint32_t k = 0; k < K * 3; ++k) dr[k] = 0.f;
        if (K == 7 && mode == 0) { dr[0] = d_alb[0]; dr[1] = d_alb[1]; dr[2] = d_alb[2]; }
    }
}

}  // namespace shade
}  // namespace ngp

using namespace ngp;

extern "C" int ngp_stencil_points(const float* xyzs, uint32_t M, float eps, float bound, int with_centre, float* out, void* stream) {
    if (!xyzs || !out) return NGP_ERR_BAD_ARG;
    if (M == 0) return NGP_OK;
    shade::stencil_points_kernel<<<cdiv(M, 256), 256, 0, as_stream(stream)>>>(xyzs, M, eps, bound, with_centre, out);
    return launch_status();
}

extern "C" int ngp_shade_forward(const float* sigma_all, const float* rgb_all, uint32_t M, uint32_t K, const float* light, float ratio,
                                 int mode, float* normal_out, float* color_out, void* stream) {
    if (!sigma_all || !normal_out || (K != 6 && K != 7)) return NGP_ERR_BAD_ARG;
    if (color_out && (K != 7 || mode < 0 || mode > 2 || (mode != 2 && !light) || (mode == 0 && !rgb_all))) return NGP_ERR_BAD_ARG;
    if (M == 0) return NGP_OK;
    shade::shade_forward_kernel<<<cdiv(M, 256), 256, 0, as_stream(stream)>>>(sigma_all, rgb_all, M, K, light, ratio, mode, normal_out,
                                                                             color_out);
    return launch_status();
}

extern "C" int ngp_shade_backward(const float* sigma_all, const float* rgb_all, uint32_t M, uint32_t K, const float* light, float ratio,
                                  int mode, const float* d_normal, const float* d_color, float* d_sigma_all, float* d_rgb_all,
                                  void* stream) {
    if (!sigma_all || (K != 6 && K != 7)) return NGP_ERR_BAD_ARG;
    if (d_color && (K != 7 || mode < 0 || mode > 2 || (mode != 2 && !light) || (mode == 0 && !rgb_all))) return NGP_ERR_BAD_ARG;
    if (M == 0) return NGP_OK;
    shade::shade_backward_kernel<<<cdiv(M, 256), 256, 0, as_stream(stream)>>>(sigma_all, rgb_all, M, K, light, ratio, mode, d_normal,
                                                                              d_color, d_sigma_all, d_rgb_all);
    return launch_status();
}
