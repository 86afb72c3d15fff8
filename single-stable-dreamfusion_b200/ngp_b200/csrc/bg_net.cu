// Background colour network of network_grid.NeRFNetwork as ONE kernel per direction:
//   FreqEncoder(degree 6) -> Linear(39, 64) + ReLU -> Linear(64, 3) -> sigmoid      (per ray, not per sample)
// Behavioural contract: nerf/network_grid.py:54-62,158-167 under fp16 autocast (freqencoder.cu:30-58 for the
// encoding; every Linear = half inputs x half weights, fp32 accumulate, half output).  The reference spends ~25 eager
// launches on this per step (encode, 3 casts, 2 GEMMs with K = 39 unaligned, bias/ReLU/sigmoid and their backward);
// the work itself is ~5 kFLOP per ray, so it is plain SIMT here: no tensor cores, weights broadcast from shared
// memory, and in the backward a register-blocked [64 x 40] += gz1^T [enc | 1] product over each 128-ray tile whose
// accumulators stay in registers across all tiles of the CTA (one flush of atomics per CTA).
#include "common.cuh"

namespace ngp {
namespace bg {

constexpr uint32_t kDeg = 6, kEnc = 3 + 3 * 2 * kDeg;  // 39
constexpr uint32_t kEncPad = 48;                      // encoding columns in shared memory: 39 features, a 1, zeros
constexpr uint32_t kHid = 64, kOut = 3, kRays = 128;

NGP_DEVINL float round_h(float v) { return __half2float(__float2half_rn(v)); }

// feature c of the frequency encoding of direction d (same expression as freq_forward_kernel)
NGP_DEVINL float freq_feature(const float (&d)[3], uint32_t c) {
    if (c < 3) return d[c];
    const uint32_t col = c / 3 - 1, ax = c % 3;
    return __sinf(scalbnf(d[ax], (int)(col / 2)) + (col % 2) * (3.141592653589793f / 2));
}

// the same for a column index only known at run time (no dynamic indexing of d: it would move to local memory)
NGP_DEVINL float freq_feature_rt(const float (&d)[3], uint32_t c) {
    const uint32_t ax = c % 3;
    const float dv = ax == 0 ? d[0] : (ax == 1 ? d[1] : d[2]);
    if (c < 3) return dv;
    const uint32_t col = c / 3 - 1;
    return __sinf(scalbnf(dv, (int)(col / 2)) + (col % 2) * (3.141592653589793f / 2));
}

struct Weights {
    const __half *w1, *b1, *w2, *b2;  // fp16 casts of bg_net.net.{0,1}.{weight,bias}: [64,39] [64] [3,64] [3]
};

// shared-memory weights as floats: W1 rows padded to 40, then b1, W2 [3][64], b2
struct SmemW {
    static constexpr uint32_t w1 = 0, b1 = w1 + kHid * 40, w2 = b1 + kHid, b2 = w2 + kOut * kHid, total = b2 + 4;
};

// (128 threads; fixed trip counts and loads gathered into registers before the stores, so the ~23 global loads of a thread are
//  in flight together - as a plain strided loop they were 20 dependent round trips, ~10 us at the head of both kernels)
NGP_DEVINL void load_weights(const Weights& w, float* s) {
    constexpr uint32_t kT = kRays;                       // threads per CTA
    constexpr uint32_t kIt1 = (kHid * 40 + kT - 1) / kT; // 20
    __half v1[kIt1];
#pragma unroll
    for (uint32_t it = 0; it < kIt1; ++it) {
        const uint32_t i = it * kT + threadIdx.x;
        const uint32_t j = i / 40, k = i % 40;
        v1[it] = (i < kHid * 40 && k < kEnc) ? w.w1[j * kEnc + k] : __float2half(0.f);
    }
    constexpr uint32_t kIt2 = (kOut * kHid + kT - 1) / kT;   // 2
    __half v2[kIt2];
#pragma unroll
    for (uint32_t it = 0; it < kIt2; ++it) {
        const uint32_t i = it * kT + threadIdx.x;
        v2[it] = i < kOut * kHid ? w.w2[i] : __float2half(0.f);
    }
    const __half vb1 = threadIdx.x < kHid ? w.b1[threadIdx.x] : __float2half(0.f);
    const __half vb2 = threadIdx.x < kOut ? w.b2[threadIdx.x] : __float2half(0.f);
#pragma unroll
    for (uint32_t it = 0; it < kIt1; ++it) {
        const uint32_t i = it * kT + threadIdx.x;
        if (i < kHid * 40) s[SmemW::w1 + i] = __half2float(v1[it]);
    }
#pragma unroll
    for (uint32_t it = 0; it < kIt2; ++it) {
        const uint32_t i = it * kT + threadIdx.x;
        if (i < kOut * kHid) s[SmemW::w2 + i] = __half2float(v2[it]);
    }
    if (threadIdx.x < kHid) s[SmemW::b1 + threadIdx.x] = __half2float(vb1);
    if (threadIdx.x < kOut) s[SmemW::b2 + threadIdx.x] = __half2float(vb2);
}

// hidden pre-activation j of one ray: half(enc_h . W1[j] + b1[j])
NGP_DEVINL float hidden_pre(const float (&enc)[40], const float* sw, uint32_t j) {
    const float4* row = reinterpret_cast<const float4*>(sw + SmemW::w1 + j * 40);
    float acc = 0.f;
#pragma unroll
    for (uint32_t q = 0; q < 10; ++q) {
        const float4 w = row[q];
        acc = fmaf(enc[4 * q], w.x, acc); acc = fmaf(enc[4 * q + 1], w.y, acc);
        acc = fmaf(enc[4 * q + 2], w.z, acc); acc = fmaf(enc[4 * q + 3], w.w, acc);
    }
    return round_h(acc + sw[SmemW::b1 + j]);
}

__global__ void __launch_bounds__(kRays) bg_forward_kernel(const float* __restrict__ dirs, uint32_t N, const Weights w,
                                                           __half* __restrict__ out) {
    extern __shared__ __align__(16) float sw[];
    load_weights(w, sw);
    __syncthreads();
    for (uint32_t n = blockIdx.x * kRays + threadIdx.x; n < N; n += gridDim.x * kRays) {
        const float d[3] = {dirs[(size_t)n * 3], dirs[(size_t)n * 3 + 1], dirs[(size_t)n * 3 + 2]};
        float enc[40];
#pragma unroll
        for (uint32_t c = 0; c < 40; ++c) enc[c] = c < kEnc ? round_h(freq_feature(d, c)) : 0.f;
        float z2[3] = {0.f, 0.f, 0.f};
        for (uint32_t j = 0; j < kHid; ++j) {
            const float a = fmaxf(hidden_pre(enc, sw, j), 0.f);
#pragma unroll
            for (uint32_t c = 0; c < 3; ++c) z2[c] = fmaf(a, sw[SmemW::w2 + c * kHid + j], z2[c]);
        }
#pragma unroll
        for (uint32_t c = 0; c < 3; ++c) {
            const float z = round_h(z2[c] + sw[SmemW::b2 + c]);
            out[(size_t)n * 3 + c] = __float2half_rn(1.0f / (1.0f + expf(-z)));
        }
    }
}

// Backward: recomputes the forward per ray (cheaper than saving [N,64] activations), then
//   gz2 = half(half(g) * y (1 - y)),  ga1 = half(gz2 . W2),  gz1 = ga1 * (z1 > 0)
//   [gW1 | gb1] += gz1^T [enc | 1],   [gW2 ; gb2] += gz2^T [a1 | 1]
// A CTA of 128 threads owns tiles of kTileB = 32 rays: FOUR threads per ray (a quad of neighbouring lanes), each computing
// ten of the 40 encoding columns and sixteen of the 64 hidden units.  The per-ray work is a long dependent chain (64 dot
// products of length 40), so with one thread per ray and 128-ray tiles a 4096-ray step (one view, or one rank's share of
// eight) kept 32 CTAs busy for 45 us - 85 us beside the field kernels - on the critical path of the step; quads cut the
// chain four-fold and spread the same rays over 128 CTAs.
// Shared-memory tiles per 32 rays: gz1 half [32][64], a1 half [32][64], enc half [32][48], gz2 float [32][4].
constexpr uint32_t kTileB = 32, kQuad = kRays / kTileB;   // rays per tile, threads per ray
static_assert(kQuad == 4 && kHid % kQuad == 0, "the backward assumes four threads per ray");
struct SmemB {
    static constexpr uint32_t gz1 = SmemW::total * 4;                  // bytes
    static constexpr uint32_t a1 = gz1 + kTileB * kHid * 2;
    static constexpr uint32_t enc = a1 + kTileB * kHid * 2;
    static constexpr uint32_t gz2 = enc + kTileB * kEncPad * 2;
    static constexpr uint32_t total = gz2 + kTileB * 4 * 4;
};

__global__ void __launch_bounds__(kRays, 3) bg_backward_kernel(const float* __restrict__ dirs, const float* __restrict__ grad_out,
                                                            uint32_t N, const Weights w, float* __restrict__ gw1,
                                                            float* __restrict__ gb1, float* __restrict__ gw2,
                                                            float* __restrict__ gb2) {
    extern __shared__ __align__(16) float sw[];
    uint8_t* sbytes = reinterpret_cast<uint8_t*>(sw);
    __half* s_gz1 = reinterpret_cast<__half*>(sbytes + SmemB::gz1);
    __half* s_a1 = reinterpret_cast<__half*>(sbytes + SmemB::a1);
    __half* s_enc = reinterpret_cast<__half*>(sbytes + SmemB::enc);
    float* s_gz2 = reinterpret_cast<float*>(sbytes + SmemB::gz2);
    load_weights(w, sw);
    __syncthreads();

    const uint32_t t = threadIdx.x;
    const uint32_t ray_l = t / kQuad, q = t % kQuad;   // phase 1: local ray of the tile, position in its quad
    // phase-2 ownership: threads 0..95 own a [4 j] x [8 k] block of [gW1 | gb1 | 0]; threads 96..111 own [4 j] x
    // [gz2 c = 0..2] of gW2^T; thread 112 owns gb2
    const uint32_t jb = t < 96 ? t / 6 : (t - 96), kb = t % 6;
    float acc[32];
#pragma unroll
    for (uint32_t i = 0; i < 32; ++i) acc[i] = 0.f;

    const uint32_t n_tiles = (N + kTileB - 1) / kTileB;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t n = tile * kTileB + ray_l;
        const bool active = n < N;
        {   // ---- phase 1: four threads per ray ------------------------------------------------------------------
            float d[3] = {0.f, 0.f, 0.f}, g[3] = {0.f, 0.f, 0.f};
            if (active) {
#pragma unroll
                for (uint32_t c = 0; c < 3; ++c) { d[c] = dirs[(size_t)n * 3 + c]; g[c] = round_h(grad_out[(size_t)n * 3 + c]); }
            }
            // [enc | 1 | 0...] row of the tile: this thread's 12 of the 48 columns (inactive rays contribute nothing: their
            // gz are zero)
#pragma unroll
            for (uint32_t i = 0; i < kEncPad / kQuad; ++i) {
                const uint32_t c = q + kQuad * i;
                const float e = c < kEnc ? round_h(freq_feature_rt(d, c)) : (c == kEnc ? 1.f : 0.f);
                s_enc[ray_l * kEncPad + c] = __float2half_rn(e);
            }
            __syncwarp();   // (a quad lives in one warp)
            float enc[40];
#pragma unroll
            for (uint32_t v = 0; v < 5; ++v) {
                const uint4 raw = *reinterpret_cast<const uint4*>(s_enc + ray_l * kEncPad + v * 8);
                const uint32_t wv[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                for (uint32_t u = 0; u < 4; ++u) {
                    const float2 e = __half22float2(*reinterpret_cast<const __half2*>(&wv[u]));
                    enc[v * 8 + 2 * u] = e.x; enc[v * 8 + 2 * u + 1] = e.y;
                }
            }
            enc[39] = 0.f;   // column 39 of the tile is the bias 1; the weight rows are zero-padded there anyway
            // forward recompute of this thread's 16 hidden units: a1 into the tile and registers, partial z2
            constexpr uint32_t kPer = kHid / kQuad;
            float a_reg[kPer];
            float z2[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (uint32_t jj = 0; jj < kPer; ++jj) {
                const uint32_t j = q * kPer + jj;
                const float a = fmaxf(hidden_pre(enc, sw, j), 0.f);
                a_reg[jj] = a;
#pragma unroll
                for (uint32_t c = 0; c < 3; ++c) z2[c] = fmaf(a, sw[SmemW::w2 + c * kHid + j], z2[c]);
            }
            {   // a1 is half-representable (ReLU of a half): 16 values = two 16-byte stores
                uint32_t pk[kPer / 2];
#pragma unroll
                for (uint32_t u = 0; u < kPer / 2; ++u) {
                    const __half2 h = __floats2half2_rn(a_reg[2 * u], a_reg[2 * u + 1]);
                    pk[u] = *reinterpret_cast<const uint32_t*>(&h);
                }
                uint4* dst = reinterpret_cast<uint4*>(s_a1 + ray_l * kHid + q * kPer);
                dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
            float gz2[3];
#pragma unroll
            for (uint32_t c = 0; c < 3; ++c) {
                z2[c] += __shfl_xor_sync(0xffffffffu, z2[c], 1);
                z2[c] += __shfl_xor_sync(0xffffffffu, z2[c], 2);   // every thread of the quad holds the same total
                const float z = round_h(z2[c] + sw[SmemW::b2 + c]);
                const float y = round_h(1.0f / (1.0f + expf(-z)));
                gz2[c] = active ? round_h(g[c] * (1.f - y) * y) : 0.f;
            }
            if (q == 0) *reinterpret_cast<float4*>(s_gz2 + ray_l * 4) = make_float4(gz2[0], gz2[1], gz2[2], 0.f);
            {
                uint32_t pk[kPer / 2];
#pragma unroll
                for (uint32_t u = 0; u < kPer / 2; ++u) {
                    float gv[2];
#pragma unroll
                    for (uint32_t e = 0; e < 2; ++e) {
                        const uint32_t jj = 2 * u + e, j = q * kPer + jj;
                        float ga = 0.f;
#pragma unroll
                        for (uint32_t c = 0; c < 3; ++c) ga = fmaf(gz2[c], sw[SmemW::w2 + c * kHid + j], ga);
                        gv[e] = a_reg[jj] > 0.f ? round_h(ga) : 0.f;
                    }
                    const __half2 h = __floats2half2_rn(gv[0], gv[1]);
                    pk[u] = *reinterpret_cast<const uint32_t*>(&h);
                }
                uint4* dst = reinterpret_cast<uint4*>(s_gz1 + ray_l * kHid + q * kPer);
                dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
        }
        __syncthreads();
        // ---- phase 2: block-wide products over the tile's 32 rays, accumulators in registers ----------------------
        if (t < 96) {
#pragma unroll 4
            for (uint32_t ray = 0; ray < kTileB; ++ray) {
                const uint2 gq = *reinterpret_cast<const uint2*>(s_gz1 + ray * kHid + jb * 4);
                const uint4 eq = *reinterpret_cast<const uint4*>(s_enc + ray * kEncPad + kb * 8);
                const float2 g01 = __half22float2(*reinterpret_cast<const __half2*>(&gq.x));
                const float2 g23 = __half22float2(*reinterpret_cast<const __half2*>(&gq.y));
                const float gj[4] = {g01.x, g01.y, g23.x, g23.y};
                const uint32_t ew[4] = {eq.x, eq.y, eq.z, eq.w};
                float ek[8];
#pragma unroll
                for (uint32_t u = 0; u < 4; ++u) {
                    const float2 e = __half22float2(*reinterpret_cast<const __half2*>(&ew[u]));
                    ek[2 * u] = e.x; ek[2 * u + 1] = e.y;
                }
#pragma unroll
                for (uint32_t a = 0; a < 4; ++a)
#pragma unroll
                    for (uint32_t b = 0; b < 8; ++b) acc[a * 8 + b] = fmaf(gj[a], ek[b], acc[a * 8 + b]);
            }
        } else if (t < 112) {
            for (uint32_t ray = 0; ray < kTileB; ++ray) {
                const float4 gz = *reinterpret_cast<const float4*>(s_gz2 + ray * 4);
                const uint2 aq = *reinterpret_cast<const uint2*>(s_a1 + ray * kHid + jb * 4);
                const float2 a01 = __half22float2(*reinterpret_cast<const __half2*>(&aq.x));
                const float2 a23 = __half22float2(*reinterpret_cast<const __half2*>(&aq.y));
                const float aj[4] = {a01.x, a01.y, a23.x, a23.y};
                const float gc[3] = {gz.x, gz.y, gz.z};
#pragma unroll
                for (uint32_t a = 0; a < 4; ++a)
#pragma unroll
                    for (uint32_t c = 0; c < 3; ++c) acc[a * 3 + c] = fmaf(aj[a], gc[c], acc[a * 3 + c]);
            }
        } else if (t == 112) {
            for (uint32_t ray = 0; ray < kTileB; ++ray) {
                const float4 gz = *reinterpret_cast<const float4*>(s_gz2 + ray * 4);
                acc[0] += gz.x; acc[1] += gz.y; acc[2] += gz.z;
            }
        }
        __syncthreads();  // tiles are rewritten by the next iteration
    }

    // ---- flush ------------------------------------------------------------------------------------------------
    if (t < 96) {
#pragma unroll
        for (uint32_t a = 0; a < 4; ++a) {
            const uint32_t j = jb * 4 + a;
#pragma unroll
            for (uint32_t b = 0; b < 8; ++b) {
                const uint32_t k = kb * 8 + b;
                if (k < kEnc) atomicAdd(gw1 + j * kEnc + k, acc[a * 8 + b]);
                else if (k == kEnc) atomicAdd(gb1 + j, acc[a * 8 + b]);
            }
        }
    } else if (t < 112) {
#pragma unroll
        for (uint32_t a = 0; a < 4; ++a)
#pragma unroll
            for (uint32_t c = 0; c < 3; ++c) atomicAdd(gw2 + c * kHid + jb * 4 + a, acc[a * 3 + c]);
    } else if (t == 112) {
        atomicAdd(gb2 + 0, acc[0]); atomicAdd(gb2 + 1, acc[1]); atomicAdd(gb2 + 2, acc[2]);
    }
}

}  // namespace bg
}  // namespace ngp

using namespace ngp;

extern "C" int ngp_bg_forward(const float* dirs, uint32_t N, const void* w1, const void* b1, const void* w2, const void* b2,
                              uint32_t degree, uint32_t hidden, void* out_rgb, void* stream) {
    if (!dirs || !w1 || !b1 || !w2 || !b2 || !out_rgb) return NGP_ERR_BAD_ARG;
    if (degree != bg::kDeg || hidden != bg::kHid) return NGP_ERR_UNSUPPORTED;
    if (N == 0) return NGP_OK;
    const bg::Weights w = {static_cast<const __half*>(w1), static_cast<const __half*>(b1), static_cast<const __half*>(w2),
                           static_cast<const __half*>(b2)};
    const int blocks = min(cdiv(N, bg::kRays), num_sms() * 4);
    bg::bg_forward_kernel<<<blocks, bg::kRays, bg::SmemW::total * 4, as_stream(stream)>>>(dirs, N, w, static_cast<__half*>(out_rgb));
    return launch_status();
}

extern "C" int ngp_bg_backward(const float* dirs, const float* grad_rgb, uint32_t N, const void* w1, const void* b1,
                               const void* w2, const void* b2, uint32_t degree, uint32_t hidden, float* gw1, float* gb1,
                               float* gw2, float* gb2, void* stream) {
    if (!dirs || !grad_rgb || !w1 || !b1 || !w2 || !b2 || !gw1 || !gb1 || !gw2 || !gb2) return NGP_ERR_BAD_ARG;
    if (degree != bg::kDeg || hidden != bg::kHid) return NGP_ERR_UNSUPPORTED;
    if (N == 0) return NGP_OK;
    const bg::Weights w = {static_cast<const __half*>(w1), static_cast<const __half*>(b1), static_cast<const __half*>(w2),
                           static_cast<const __half*>(b2)};
    static PerDeviceAttr attr;
    const int arc = set_kernel_smem(&attr, reinterpret_cast<const void*>(bg::bg_backward_kernel), (int)bg::SmemB::total);
    if (arc != NGP_OK) return arc;
    // one CTA per SM at most: in the train step this kernel runs beside the (persistent, shared-memory hungry) field
    // kernels on a side stream and should fill their gaps, not evict them
    const int blocks = min(cdiv(N, bg::kTileB), num_sms());
    bg::bg_backward_kernel<<<blocks, bg::kRays, bg::SmemB::total, as_stream(stream)>>>(dirs, grad_rgb, N, w, gw1, gb1, gw2, gb2);
    return launch_status();
}
