// Multi-resolution hash / tiled grid encoding, forward + backward, for sm_100a.
//
// Behavioural contract: gridencoder/src/gridencoder.cu of the reference (kernel_grid :76-223,
// kernel_grid_backward :227-313, kernel_input_backward :317-342, get_grid_index :54-72,
// fast_hash :35-51).  The floating-point expression shapes below (which products feed which
// adds, where a value is rounded to half) are kept identical to the reference so that fp32 AND
// fp16 results are bit-equal to it; the parallel decomposition is not the reference's:
//
//   * one thread owns (point, LPT consecutive levels) and keeps the whole row of features in
//     registers; inputs are read once per LPT levels instead of once per level, and the result is
//     stored straight into the [B, L*C] layout the Python side wants (the reference writes
//     [L,B,C] and then pays a full permute copy, grid.py:52,70);
//   * feature rows are fetched with ONE vector load per corner (half2 for C=2) from the
//     L2/L1-resident table instead of C scalar 2-byte loads;
//   * the backward scatters whole rows with one red.global.add.v2/v4.f32 per corner into an fp32
//     gradient table (no fp16 accumulation), or - for strict drop-in use through the C ABI - with
//     half2 atomics like the reference.
#pragma once
#include "common.cuh"

namespace ngp {
namespace grid {

constexpr uint32_t kMaxLevels = 64;

// test / profiling switch (ngp_grid_set_option): force the per-sample scatter kernel
extern bool g_disable_warpagg;

struct LevelParams {
    float scale;
    uint32_t resolution;
    uint32_t hashmap_size;
    uint32_t offset;  // in rows
};

// gridencoder.cu:124-126 - evaluated on the device, with the reference's operand order.
NGP_DEVINL LevelParams make_level(const int* __restrict__ offsets, uint32_t level, float S, uint32_t H) {
    LevelParams p;
    p.offset = (uint32_t)offsets[level];
    p.hashmap_size = (uint32_t)offsets[level + 1] - (uint32_t)offsets[level];
    p.scale = exp2f(level * S) * H - 1.0f;
    p.resolution = (uint32_t)ceil(p.scale) + 1;
    return p;
}

// ---- per-level addressing, resolved ONCE per level instead of once per corner ------------------------------------
// lattice_row() re-derives, for every corner, which axes enter the index, whether the level is hashed and how the
// index wraps.  All of that depends only on the level, so FastLevel holds the outcome: the linear stride of each
// used axis (0 for dropped axes), hash-or-linear, and the wrap (none / AND mask / modulo).  corner_rows() then turns
// a base lattice point into the 2^D row indices with a handful of adds (linear) or XORs (hash:
// (p + 1) * prime == p * prime + prime in uint32).  Same uint32 arithmetic as gridencoder.cu:54-72, so same rows.
constexpr uint32_t kWrapNone = 0, kWrapMask = 1, kWrapMod = 2;

template <uint32_t D>
struct FastLevel {
    float scale;
    uint32_t offset;     // first row of the level
    uint32_t size;       // rows in the level
    uint32_t mask;       // size - 1 when size is a power of two
    uint32_t wrap;       // kWrapNone / kWrapMask / kWrapMod
    uint32_t hashed;     // 1: spatial hash of all axes; 0: linear index over the leading `used` axes
    uint32_t used;       // axes that enter the index (D when hashed)
    uint32_t stride[D];  // linear stride per axis (0: axis ignored)
};

template <uint32_t D>
NGP_DEVINL FastLevel<D> make_fast_level(const int* __restrict__ offsets, uint32_t level, float S, uint32_t H,
                                        uint32_t gridtype, bool align_corners) {
    const LevelParams lp = make_level(offsets, level, S, H);
    FastLevel<D> f;
    f.scale = lp.scale;
    f.offset = lp.offset;
    f.size = lp.hashmap_size;
    f.mask = lp.hashmap_size - 1;
    uint32_t stride = 1, used = 0;
    bool open = true;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        open = open && (stride <= lp.hashmap_size);
        f.stride[d] = 0;
        if (open) {
            f.stride[d] = stride;
            stride *= align_corners ? lp.resolution : (lp.resolution + 1);
            ++used;
        }
    }
    f.hashed = (gridtype == NGP_GRID_HASH && stride > lp.hashmap_size) ? 1u : 0u;
    f.used = f.hashed ? D : used;
    if (!f.hashed && open && !align_corners && stride <= lp.hashmap_size) f.wrap = kWrapNone;
    else if ((lp.hashmap_size & (lp.hashmap_size - 1)) == 0) f.wrap = kWrapMask;
    else f.wrap = kWrapMod;
    return f;
}

// rows[corner] for all 2^D corners of the cell at `base` (bit d of corner = +1 along axis d).
template <uint32_t D>
NGP_DEVINL void corner_rows(const FastLevel<D>& f, const uint32_t (&base)[D], uint32_t (&rows)[1u << D]) {
    if (f.hashed) {
        // gridencoder.cu:42 - the instant-ngp primes; 1 for the first axis keeps x-neighbours adjacent
        constexpr uint32_t kPrimes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
        rows[0] = 0;
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) {
            const uint32_t h0 = base[d] * kPrimes[d], h1 = h0 + kPrimes[d];
#pragma unroll
            for (uint32_t c = 0; c < (1u << d); ++c) {
                rows[c | (1u << d)] = rows[c] ^ h1;
                rows[c] ^= h0;
            }
        }
    } else {
        uint32_t lin = 0;
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) lin += base[d] * f.stride[d];
        rows[0] = lin;
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) {
#pragma unroll
            for (uint32_t c = 0; c < (1u << d); ++c) rows[c | (1u << d)] = rows[c] + f.stride[d];
        }
    }
    if (f.wrap == kWrapMask) {
#pragma unroll
        for (uint32_t c = 0; c < (1u << D); ++c) rows[c] &= f.mask;
    } else if (f.wrap == kWrapMod) {
#pragma unroll
        for (uint32_t c = 0; c < (1u << D); ++c) rows[c] %= f.size;
    }
}

// Trilinear (2^D-linear) weights in the reference's multiplication order: w = ((1 * a0) * a1) * a2 ...
template <uint32_t D>
NGP_DEVINL void corner_weights(const float (&frac)[D], float (&wts)[1u << D]) {
    wts[0] = 1.0f;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        const float lo = 1 - frac[d], hi = frac[d];
        // process corners top-down so that wts[c] (prefix over axes < d) is still intact when read
#pragma unroll
        for (uint32_t c = 0; c < (1u << d); ++c) {
            const float prefix = wts[c];
            wts[c | (1u << d)] = prefix * hi;
            wts[c] = prefix * lo;
        }
    }
}

// Static-index helpers for the above (dynamic indexing would push the per-corner arrays to local memory).
template <uint32_t D, uint32_t C, uint32_t U>
NGP_DEVINL void replicate_rows_static(float (&rows)[1u << D][C]) {
#pragma unroll
    for (uint32_t corner = (1u << U); corner < (1u << D); ++corner) {
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) rows[corner][c] = rows[corner & ((1u << U) - 1)][c];
    }
}
template <uint32_t D, uint32_t C>
NGP_DEVINL void replicate_rows(float (&rows)[1u << D][C], uint32_t axes_used) {
    if constexpr (D >= 2) { if (axes_used == 1) { replicate_rows_static<D, C, 1>(rows); return; } }
    if constexpr (D >= 3) { if (axes_used == 2) { replicate_rows_static<D, C, 2>(rows); return; } }
    if constexpr (D >= 4) { if (axes_used == 3) { replicate_rows_static<D, C, 3>(rows); return; } }
    if constexpr (D >= 5) { if (axes_used == 4) { replicate_rows_static<D, C, 4>(rows); return; } }
}
// ---- element-type plumbing ----------------------------------------------------------------------
template <typename T> struct ElemOps;
template <> struct ElemOps<float> {
    static NGP_DEVINL float to_f(float v) { return v; }
    static NGP_DEVINL float round(float v) { return v; }
};
template <> struct ElemOps<__half> {
    static NGP_DEVINL float to_f(__half v) { return __half2float(v); }
    static NGP_DEVINL float round(float v) { return __half2float(__float2half_rn(v)); }
};

// Load one row of C features as floats with a single (or, for 32 bytes, two) vector load.
template <typename T, uint32_t C>
NGP_DEVINL void load_row(const T* __restrict__ p, float (&out)[C]) {
    constexpr uint32_t BYTES = sizeof(T) * C;
    if constexpr (BYTES == 2) {
        unsigned short raw = __ldg(reinterpret_cast<const unsigned short*>(p));
        out[0] = __half2float(__ushort_as_half(raw));
    } else if constexpr (BYTES == 4) {
        unsigned int raw = __ldg(reinterpret_cast<const unsigned int*>(p));
        if constexpr (sizeof(T) == 4) {
            out[0] = __uint_as_float(raw);
        } else {
            float2 f = __half22float2(*reinterpret_cast<const __half2*>(&raw));
            out[0] = f.x; out[1] = f.y;
        }
    } else if constexpr (BYTES == 8) {
        uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
        if constexpr (sizeof(T) == 4) {
            out[0] = __uint_as_float(raw.x); out[1] = __uint_as_float(raw.y);
        } else {
            float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
            float2 b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
            out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
        }
    } else if constexpr (BYTES == 16) {
        uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
        const unsigned int w[4] = {raw.x, raw.y, raw.z, raw.w};
        if constexpr (sizeof(T) == 4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) out[i] = __uint_as_float(w[i]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float2 a = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
                out[2 * i] = a.x; out[2 * i + 1] = a.y;
            }
        }
    } else {  // 32 bytes: fp32, C = 8
        static_assert(BYTES == 32 && sizeof(T) == 4, "unexpected row size");
        uint4 r0 = __ldg(reinterpret_cast<const uint4*>(p));
        uint4 r1 = __ldg(reinterpret_cast<const uint4*>(p) + 1);
        out[0] = __uint_as_float(r0.x); out[1] = __uint_as_float(r0.y);
        out[2] = __uint_as_float(r0.z); out[3] = __uint_as_float(r0.w);
        out[4] = __uint_as_float(r1.x); out[5] = __uint_as_float(r1.y);
        out[6] = __uint_as_float(r1.z); out[7] = __uint_as_float(r1.w);
    }
}

template <typename T, uint32_t C>
NGP_DEVINL void store_row(T* p, const float (&v)[C]) {
    if constexpr (sizeof(T) == 4) {
        if constexpr (C == 1) {
            p[0] = v[0];
        } else if constexpr (C == 2) {
            *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
        } else {
#pragma unroll
            for (uint32_t i = 0; i < C; i += 4)
                *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
    } else {
        if constexpr (C == 1) {
            p[0] = __float2half_rn(v[0]);
        } else if constexpr (C == 2) {
            *reinterpret_cast<__half2*>(p) = __floats2half2_rn(v[0], v[1]);
        } else if constexpr (C == 4) {
            __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
            uint2 raw;
            raw.x = *reinterpret_cast<unsigned int*>(&a); raw.y = *reinterpret_cast<unsigned int*>(&b);
            *reinterpret_cast<uint2*>(p) = raw;
        } else {
            __half2 h[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
            uint4 raw;
            raw.x = *reinterpret_cast<unsigned int*>(&h[0]); raw.y = *reinterpret_cast<unsigned int*>(&h[1]);
            raw.z = *reinterpret_cast<unsigned int*>(&h[2]); raw.w = *reinterpret_cast<unsigned int*>(&h[3]);
            *reinterpret_cast<uint4*>(p) = raw;
        }
    }
}

// Fractional position + base lattice point for one level (gridencoder.cu:132-137).
template <uint32_t D>
NGP_DEVINL void locate(const float (&x)[D], float scale, bool align_corners, float (&frac)[D], uint32_t (&base)[D]) {
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        float pos = x[d] * scale + (align_corners ? 0.0f : 0.5f);
        base[d] = floorf(pos);
        frac[d] = pos - (float)base[d];
    }
}

template <uint32_t D>
NGP_DEVINL bool out_of_unit_cube(const float (&x)[D]) {
    bool oob = false;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) oob = oob || (x[d] < 0 || x[d] > 1);
    return oob;
}

// -------------------------------------------------------------------------------------------------
// Forward
// -------------------------------------------------------------------------------------------------
// Packed-half accumulation of one corner for the C = 2 half table: `acc += w * row` with c10::Half semantics
// (gridencoder.cu:165: the float product is rounded to half, then the half + half sum is rounded again).  HADD2
// rounds each lane's exact sum once, which equals rounding the fp32 sum of two halves (that fp32 sum is exact
// whenever it could land on a half tie).
NGP_DEVINL __half2 half2_axpy(__half2 acc, float w, uint32_t raw_row) {
    const float2 r = __half22float2(*reinterpret_cast<const __half2*>(&raw_row));
    return __hadd2(acc, __floats2half2_rn(w * r.x, w * r.y));
}

template <typename T, uint32_t D, uint32_t C, uint32_t LPT, bool OUT_BLC, bool WITH_DYDX>
__global__ void __launch_bounds__(256) encode_forward_kernel(
    const float* __restrict__ inputs, const T* __restrict__ table, const int* __restrict__ offsets,
    T* __restrict__ outputs, T* __restrict__ dy_dx, uint32_t B, uint32_t L, float S, uint32_t H,
    uint32_t gridtype, bool align_corners) {
    __shared__ FastLevel<D> s_levels[kMaxLevels];
    for (uint32_t l = threadIdx.x; l < L; l += blockDim.x) s_levels[l] = make_fast_level<D>(offsets, l, S, H, gridtype, align_corners);
    __syncthreads();

    const uint32_t groups = (L + LPT - 1) / LPT;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t b = (uint32_t)(tid / groups);
    const uint32_t g = (uint32_t)(tid - (uint64_t)b * groups);
    if (b >= B) return;

    float x[D];
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) x[d] = __ldg(inputs + (size_t)b * D + d);
    const bool oob = out_of_unit_cube<D>(x);

#pragma unroll
    for (uint32_t li = 0; li < LPT; ++li) {
        const uint32_t level = g * LPT + li;
        if (level >= L) break;
        T* out = OUT_BLC ? outputs + ((size_t)b * L + level) * C : outputs + ((size_t)level * B + b) * C;
        T* dout = WITH_DYDX ? dy_dx + ((size_t)b * L + level) * D * C : nullptr;
        float acc[C];
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) acc[c] = 0.f;

        if (oob) {  // gridencoder.cu:106-122
            store_row<T, C>(out, acc);
            if (WITH_DYDX) {
#pragma unroll
                for (uint32_t d = 0; d < D; ++d) store_row<T, C>(dout + d * C, acc);
            }
            continue;
        }

        const FastLevel<D>& lp = s_levels[level];
        const T* __restrict__ tbl = table + (size_t)lp.offset * C;
        float frac[D];
        uint32_t base[D];
        locate<D>(x, lp.scale, align_corners, frac, base);
        uint32_t ridx[1u << D];
        corner_rows<D>(lp, base, ridx);
        float wts[1u << D];
        corner_weights<D>(frac, wts);

        // Issue all row gathers first (independent loads in flight), then blend.  Corners that differ only in an
        // axis the level ignores share a row: it is fetched once and reused (same values, same arithmetic).
        const uint32_t distinct = 1u << lp.used;
        if constexpr (sizeof(T) == 2 && C == 2) {
            uint32_t raw[1u << D];
#pragma unroll
            for (uint32_t corner = 0; corner < (1u << D); ++corner)
                if (corner < distinct) raw[corner] = __ldg(reinterpret_cast<const uint32_t*>(tbl) + ridx[corner]);
#pragma unroll
            for (uint32_t corner = 0; corner < (1u << D); ++corner)
                if (corner >= distinct) raw[corner] = raw[corner & (distinct - 1)];
            __half2 acc2 = __floats2half2_rn(0.f, 0.f);
#pragma unroll
            for (uint32_t corner = 0; corner < (1u << D); ++corner) acc2 = half2_axpy(acc2, wts[corner], raw[corner]);
            *reinterpret_cast<__half2*>(out) = acc2;
            if (WITH_DYDX) {
#pragma unroll
                for (uint32_t gd = 0; gd < D; ++gd) {
                    __half2 g2 = __floats2half2_rn(0.f, 0.f);
#pragma unroll
                    for (uint32_t corner = 0; corner < (1u << (D - 1)); ++corner) {
                        // spread the D-1 bits of `corner` around axis gd; weight in the reference's order (:186-196)
                        float w = lp.scale;
                        uint32_t left = 0;
#pragma unroll
                        for (uint32_t nd = 0; nd < D - 1; ++nd) {
                            const uint32_t d = (nd >= gd) ? (nd + 1) : nd;
                            if ((corner & (1u << nd)) == 0) w *= 1 - frac[d];
                            else { w *= frac[d]; left |= 1u << d; }
                        }
                        const __half2 lo = *reinterpret_cast<const __half2*>(&raw[left]);
                        const __half2 hi = *reinterpret_cast<const __half2*>(&raw[left | (1u << gd)]);
                        const float2 diff = __half22float2(__hsub2(hi, lo));
                        g2 = __hadd2(g2, __floats2half2_rn(w * diff.x, w * diff.y));
                    }
                    *reinterpret_cast<__half2*>(dout + gd * C) = g2;
                }
            }
        } else {
            float rows[1u << D][C];
#pragma unroll
            for (uint32_t corner = 0; corner < (1u << D); ++corner)
                if (corner < distinct) load_row<T, C>(tbl + (size_t)ridx[corner] * C, rows[corner]);
            replicate_rows<D, C>(rows, lp.used);
#pragma unroll
            for (uint32_t corner = 0; corner < (1u << D); ++corner) {
#pragma unroll
                for (uint32_t c = 0; c < C; ++c) {
                    if constexpr (sizeof(T) == 4) {
                        acc[c] += wts[corner] * rows[corner][c];  // contracts to one FFMA, as in the reference
                    } else {
                        const float prod = ElemOps<T>::round(wts[corner] * rows[corner][c]);
                        acc[c] = ElemOps<T>::round(acc[c] + prod);
                    }
                }
            }
            store_row<T, C>(out, acc);

            if (WITH_DYDX) {  // gridencoder.cu:179-222
#pragma unroll
                for (uint32_t gd = 0; gd < D; ++gd) {
                    float gacc[C];
#pragma unroll
                    for (uint32_t c = 0; c < C; ++c) gacc[c] = 0.f;
#pragma unroll
                    for (uint32_t corner = 0; corner < (1u << (D - 1)); ++corner) {
                        float w = lp.scale;
                        uint32_t left = 0;
#pragma unroll
                        for (uint32_t nd = 0; nd < D - 1; ++nd) {
                            const uint32_t d = (nd >= gd) ? (nd + 1) : nd;
                            if ((corner & (1u << nd)) == 0) w *= 1 - frac[d];
                            else { w *= frac[d]; left |= 1u << d; }
                        }
                        const uint32_t right = left | (1u << gd);
#pragma unroll
                        for (uint32_t c = 0; c < C; ++c) {
                            if constexpr (sizeof(T) == 4) {
                                gacc[c] += w * (rows[right][c] - rows[left][c]);
                            } else {
                                const float diff = ElemOps<T>::round(rows[right][c] - rows[left][c]);
                                const float prod = ElemOps<T>::round(w * diff);
                                gacc[c] = ElemOps<T>::round(gacc[c] + prod);
                            }
                        }
                    }
                    store_row<T, C>(dout + gd * C, gacc);
                }
            }
        }
    }
}

// -------------------------------------------------------------------------------------------------
// Backward: scatter-add of w * grad rows into the gradient table.
//   GT = float : fp32 table, one red.global.add (v2/v4 where C allows) per corner.
//   GT = __half: reference-compatible half2 atomics (gridencoder.cu:298-304); needs even C.
// -------------------------------------------------------------------------------------------------
template <typename GT, uint32_t C>
NGP_DEVINL void scatter_row(GT* dst, float w, const float (&g)[C]) {
    if constexpr (sizeof(GT) == 4) {
        if constexpr (C == 1) {
            red_add_f32(dst, w * g[0]);
        } else if constexpr (C == 2) {
            red_add_f32x2(dst, w * g[0], w * g[1]);
        } else {
#pragma unroll
            for (uint32_t c = 0; c < C; c += 4) red_add_f32x4(dst + c, w * g[c], w * g[c + 1], w * g[c + 2], w * g[c + 3]);
        }
    } else {
        static_assert(C % 2 == 0, "half gradient tables need an even feature count");
#pragma unroll
        for (uint32_t c = 0; c < C; c += 2) {
            __half2 v = __halves2half2(__float2half_rn(w * g[c]), __float2half_rn(w * g[c + 1]));
            atomicAdd(reinterpret_cast<__half2*>(dst + c), v);
        }
    }
}

template <typename T, typename GT, uint32_t D, uint32_t C, uint32_t LPT, bool GRAD_BLC>
__global__ void __launch_bounds__(256) encode_backward_kernel(
    const T* __restrict__ grad, const float* __restrict__ inputs, const int* __restrict__ offsets,
    GT* __restrict__ grad_table, uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype,
    bool align_corners) {
    __shared__ FastLevel<D> s_levels[kMaxLevels];
    for (uint32_t l = threadIdx.x; l < L; l += blockDim.x) s_levels[l] = make_fast_level<D>(offsets, l, S, H, gridtype, align_corners);
    __syncthreads();

    const uint32_t groups = (L + LPT - 1) / LPT;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t b = (uint32_t)(tid / groups);
    const uint32_t g = (uint32_t)(tid - (uint64_t)b * groups);
    if (b >= B) return;

    float x[D];
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) x[d] = __ldg(inputs + (size_t)b * D + d);
    if (out_of_unit_cube<D>(x)) return;  // gridencoder.cu:253-258

#pragma unroll
    for (uint32_t li = 0; li < LPT; ++li) {
        const uint32_t level = g * LPT + li;
        if (level >= L) break;
        const FastLevel<D>& lp = s_levels[level];
        const T* gsrc = GRAD_BLC ? grad + ((size_t)b * L + level) * C : grad + ((size_t)level * B + b) * C;
        float gr[C];
        load_row<T, C>(gsrc, gr);

        float frac[D];
        uint32_t base[D];
        locate<D>(x, lp.scale, align_corners, frac, base);
        uint32_t ridx[1u << D];
        corner_rows<D>(lp, base, ridx);
        float wts[1u << D];
        corner_weights<D>(frac, wts);
        GT* tbl = grad_table + (size_t)lp.offset * C;
#pragma unroll
        for (uint32_t corner = 0; corner < (1u << D); ++corner) scatter_row<GT, C>(tbl + (size_t)ridx[corner] * C, wts[corner], gr);
    }
}

// -------------------------------------------------------------------------------------------------
// Backward, warp-aggregated (D = 3, fp32 table, [B, L*C] gradients): one thread per sample, one warp per 32
// CONSECUTIVE samples.  Marched samples arrive in ray order, so at coarse levels long runs of neighbouring
// lanes fall into the same grid cell and would hit the same 8 table rows.  Per level the warp finds those
// runs (cell key compared with the previous lane), sums the 8 corner contributions of each run with a
// segmented shuffle reduction, and only the first lane of a run issues the 8 red.global.add - instead of
// every sample issuing its own atomics (gridencoder.cu:298-310).  The SM issues reds at ~1.3 cycles per
// active LANE, so fewer active lanes is a direct speed-up; levels whose cells are finer than the sample
// spacing have runs of length 1 and pay only two shuffles and a ballot.
// -------------------------------------------------------------------------------------------------
// The reds of ONE cell of one level: v[corner][c] = the summed contributions to the cell's 2^USED distinct rows.
template <uint32_t USED, bool HASHED, uint32_t C, bool COUNT>
NGP_DEVINL void issue_cell_reds(const FastLevel<3>& lp, const uint32_t (&base)[USED], const float (&v)[1u << USED][C],
                                float* __restrict__ grad_table, float* __restrict__ grad_odd, uint32_t* reds) {
    constexpr uint32_t NC = 1u << USED;
    uint32_t rows[NC];
    if constexpr (HASHED) {
        constexpr uint32_t kPrimes[3] = {1u, 2654435761u, 805459861u};
        rows[0] = 0;
#pragma unroll
        for (uint32_t d = 0; d < USED; ++d) {
            const uint32_t h0 = base[d] * kPrimes[d], h1 = h0 + kPrimes[d];
#pragma unroll
            for (uint32_t c = 0; c < (1u << d); ++c) { rows[c | (1u << d)] = rows[c] ^ h1; rows[c] ^= h0; }
        }
    } else {
        uint32_t lin = 0;
#pragma unroll
        for (uint32_t d = 0; d < USED; ++d) lin += base[d] * lp.stride[d];
        rows[0] = lin;
#pragma unroll
        for (uint32_t d = 0; d < USED; ++d) {
#pragma unroll
            for (uint32_t c = 0; c < (1u << d); ++c) rows[c | (1u << d)] = rows[c] + lp.stride[d];
        }
    }
    if (lp.wrap == kWrapMask) {
#pragma unroll
        for (uint32_t c = 0; c < NC; ++c) rows[c] &= lp.mask;
    } else if (lp.wrap == kWrapMod) {
#pragma unroll
        for (uint32_t c = 0; c < NC; ++c) rows[c] %= lp.size;
    }
    float* tbl = grad_table + (size_t)lp.offset * C;
    if constexpr (C == 2 && !HASHED && USED >= 1) {
        // The SM issues reds at a fixed rate per LANE-op whatever their width (profiles/: ~217 G ops/s for 4-, 8- and
        // 16-byte reds alike), so the two x-neighbours of a corner pair - adjacent rows of a linear level - go out
        // as ONE 16-byte red whenever the pair is 16-byte aligned.  Pairs that start on an odd row are not - unless the
        // caller supplies the ODD-FRAME TWIN of the table (`grad_odd`: same indexing, but the buffer starts 8 bytes off a
        // 16-byte boundary, so exactly the odd-row pairs are aligned there).  The sum of the two buffers is the gradient
        // (ngp_grid_fold_odd, or the optimizer's finite-check pass); every pair then costs one lane-op instead of 1.5.
        float* tbl_odd = grad_odd ? grad_odd + (size_t)lp.offset * C : nullptr;
#pragma unroll
        for (uint32_t corner = 0; corner < NC; corner += 2) {
            const uint32_t lo = rows[corner], hi = rows[corner + 1];
            float* dst = tbl + (size_t)lo * 2;
            const bool even = ((lp.offset + lo) & 1u) == 0;
            if (hi == lo + 1 && (even || tbl_odd)) {
                red_add_f32x4(even ? dst : tbl_odd + (size_t)lo * 2, v[corner][0], v[corner][1], v[corner + 1][0], v[corner + 1][1]);
                if constexpr (COUNT) *reds += 1u;
            } else {
                red_add_f32x2(dst, v[corner][0], v[corner][1]);
                red_add_f32x2(tbl + (size_t)hi * 2, v[corner + 1][0], v[corner + 1][1]);
                if constexpr (COUNT) *reds += 2u;
            }
        }
    } else {
#pragma unroll
        for (uint32_t corner = 0; corner < NC; ++corner) {
            float* dst = tbl + (size_t)rows[corner] * C;
            if constexpr (C == 1) red_add_f32(dst, v[corner][0]);
            else if constexpr (C == 2) red_add_f32x2(dst, v[corner][0], v[corner][1]);
            else {
#pragma unroll
                for (uint32_t c = 0; c < C; c += 4) red_add_f32x4(dst + c, v[corner][c], v[corner][c + 1], v[corner][c + 2], v[corner][c + 3]);
            }
            if constexpr (COUNT) *reds += (C <= 2 ? 1u : C / 4);
        }
    }
}

// One level of the warp-aggregated scatter, specialised at compile time on how the level addresses its rows:
// USED = axes that enter the row index (a 'tiled' level whose table is smaller than (res+1)^2 drops z: 4 distinct
// rows per cell, and since the weights of the two z corners sum to 1 only the (x, y) bilinear weights are needed),
// HASHED = spatial hash of all three axes.
// COUNT (measurement builds of the kernel only): `reds` accumulates the red instructions THIS lane issued, i.e. the
// post-aggregation atomic lane-ops that bench.py's roofline divides by the measured red issue ceiling.
template <uint32_t USED, bool HASHED, uint32_t C, bool COUNT = false>
NGP_DEVINL void warpagg_level(const FastLevel<3>& lp, const float (&x)[3], bool align_corners, bool valid, const float (&g)[C],
                              uint32_t lane, float* __restrict__ grad_table, float* __restrict__ grad_odd, uint32_t* reds = nullptr,
                              uint32_t run_cap = 32u) {
    constexpr uint32_t NC = 1u << USED;  // distinct rows per cell
    float frac[USED];
    uint32_t base[USED];
    {
        float xs[USED];
#pragma unroll
        for (uint32_t d = 0; d < USED; ++d) xs[d] = x[d];
        locate<USED>(xs, lp.scale, align_corners, frac, base);
    }
    // run detection: same cell (over the axes the level uses) as the previous lane, both valid
    uint32_t kxy = base[0], kz = 0u;
    if constexpr (USED >= 2) kxy |= base[1] << 16;
    if constexpr (USED >= 3) kz = base[2];
    if (!valid) { kxy = 0xffffffffu; kz = 0x80000000u | lane; }
    const uint32_t pxy = __shfl_up_sync(0xffffffffu, kxy, 1);
    uint32_t pz = kz;
    if constexpr (USED >= 3) pz = __shfl_up_sync(0xffffffffu, kz, 1);  // every lane shuffles (no short-circuit around it)
    // run_cap (a power of two <= 32) additionally starts a run every run_cap lanes: fewer reduction steps below (they
    // are ~1/3 of this kernel's instructions, which is what bounds it) for a few more reds at the coarse levels
    const bool head = ((lane & (run_cap - 1u)) == 0) || !valid || (pxy != kxy) || (pz != kz);
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    // last lane of my run = (next head above me) - 1
    const uint32_t above = (lane == 31) ? 0u : (heads >> (lane + 1));
    const uint32_t run_end = above ? (lane + (uint32_t)__ffs(above) - 1u) : 31u;
    // longest run in the warp bounds the number of reduction steps (warp-uniform)
    const uint32_t my_len = head ? (run_end - lane + 1u) : 0u;
    const uint32_t max_len = __reduce_max_sync(0xffffffffu, my_len);

    float wts[NC];
    corner_weights<USED>(frac, wts);
    float v[NC][C];
#pragma unroll
    for (uint32_t corner = 0; corner < NC; ++corner) {
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) v[corner][c] = wts[corner] * g[c];
    }
    for (uint32_t off = 1; off < max_len; off <<= 1) {
        const bool take = (lane + off) <= run_end;
#pragma unroll
        for (uint32_t corner = 0; corner < NC; ++corner) {
#pragma unroll
            for (uint32_t c = 0; c < C; ++c) {
                const float o = __shfl_down_sync(0xffffffffu, v[corner][c], off);
                if (take) v[corner][c] += o;
            }
        }
    }
    if (head && valid) issue_cell_reds<USED, HASHED, C, COUNT>(lp, base, v, grad_table, grad_odd, reds);
}

template <typename T, uint32_t C, bool COUNT = false>
__global__ void __launch_bounds__(256) encode_backward_warpagg_kernel(
    const T* __restrict__ grad, const float* __restrict__ inputs, const int* __restrict__ offsets,
    float* __restrict__ grad_table, uint32_t B_cap, uint32_t L, float S, uint32_t H, uint32_t gridtype,
    bool align_corners, const int* __restrict__ count_ptr, float bound, unsigned long long* red_lane_ops = nullptr,
    float* __restrict__ grad_odd = nullptr, uint32_t run_cap = 32u) {
    constexpr uint32_t D = 3;
    __shared__ FastLevel<D> s_levels[kMaxLevels];
    for (uint32_t l = threadIdx.x; l < L; l += blockDim.x) s_levels[l] = make_fast_level<D>(offsets, l, S, H, gridtype, align_corners);
    __syncthreads();

    // optional device-side row count (sync-free training path) and optional [-bound, bound] -> [0, 1] mapping
    const uint32_t B = count_ptr ? min((uint32_t)max(*count_ptr, 0), B_cap) : B_cap;
    const float inv_2b = bound > 0.f ? __fdiv_rn(1.0f, 2 * bound) : 1.0f;
    const uint32_t lane = threadIdx.x & 31;
    constexpr bool kPacked = (sizeof(T) == 2 && C == 2);
    uint32_t reds = 0;

    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; (b & ~31u) < B; b += gridDim.x * blockDim.x) {
        // whole warps stay alive (shuffles below); out-of-range / out-of-cube samples just contribute nothing
        bool valid = b < B;
        float x[D] = {0.f, 0.f, 0.f};
        if (valid) {
#pragma unroll
            for (uint32_t d = 0; d < D; ++d) {
                x[d] = __ldg(inputs + (size_t)b * D + d);
                if (bound > 0.f) x[d] = __fmul_rn(__fadd_rn(x[d], bound), inv_2b);  // GridEncoder.forward's mapping (grid.py:142)
            }
            valid = !out_of_unit_cube<D>(x);  // gridencoder.cu:253-258
        }
        // half2 gradient rows: a sample's [L*C] row is 2 sectors; for the 16-level field all four 16-byte quarters are
        // fetched up front (four loads in flight instead of one per level group).  Samples behind their ray's
        // early-termination point carry an all-zero row; a warp whose 32 rows are all zero has nothing to add and
        // skips the whole level loop (adding exact zeros is the identity, so the result is unchanged).
        uint32_t graw[4] = {0u, 0u, 0u, 0u};
        uint4 gq[4];
        constexpr bool kRow16 = kPacked;
        const bool row16 = kRow16 && L == 16;
        if constexpr (kPacked) {
            if (row16) {
                uint32_t any = 0u;
#pragma unroll
                for (uint32_t q = 0; q < 4; ++q) {
                    gq[q] = make_uint4(0u, 0u, 0u, 0u);
                    if (valid) gq[q] = __ldg(reinterpret_cast<const uint4*>(grad + (size_t)b * 32) + q);
                    any |= (gq[q].x | gq[q].y | gq[q].z | gq[q].w) & 0x7fff7fffu;     // (-0 is zero too)
                }
                if (!__any_sync(0xffffffffu, any != 0u)) continue;
                // ... and a LANE whose row is all zero issues no reds of its own: it drops out of the run detection like an
                // out-of-cube sample (its neighbours' runs simply end / start around it)
                valid = valid && (any != 0u);
            }
        }
        // One level.  The loops over the levels below are deliberately ROLLED: with the 16 levels unrolled the kernel held
        // 64 copies of warpagg_level (27.7 k instructions, 440 KB of code) and every 32 samples walked 16 different copies
        // through the instruction cache; rolled it is one copy per addressing class.
        auto process_level = [&](uint32_t level, uint32_t g_packed) {
            const FastLevel<D>& lp = s_levels[level];
            float g[C];
#pragma unroll
            for (uint32_t c = 0; c < C; ++c) g[c] = 0.f;
            if constexpr (kPacked) {
                const float2 gf = __half22float2(*reinterpret_cast<const __half2*>(&g_packed));
                g[0] = gf.x; g[1] = gf.y;
            } else {
                if (valid) load_row<T, C>(grad + ((size_t)b * L + level) * C, g);
            }
            // warp-uniform dispatch on the level's addressing class
            if (lp.hashed)          warpagg_level<3, true, C, COUNT>(lp, x, align_corners, valid, g, lane, grad_table, grad_odd, &reds, run_cap);
            else if (lp.used == 3)  warpagg_level<3, false, C, COUNT>(lp, x, align_corners, valid, g, lane, grad_table, grad_odd, &reds, run_cap);
            else if (lp.used == 2)  warpagg_level<2, false, C, COUNT>(lp, x, align_corners, valid, g, lane, grad_table, grad_odd, &reds, run_cap);
            else                    warpagg_level<1, false, C, COUNT>(lp, x, align_corners, valid, g, lane, grad_table, grad_odd, &reds, run_cap);
        };
        auto pick = [](const uint4& v, uint32_t j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); };
        if (row16) {
#pragma unroll 1
            for (uint32_t level = 0; level < 16; ++level) {
                const uint32_t q = level >> 2;
                const uint4 cur = q == 0 ? gq[0] : (q == 1 ? gq[1] : (q == 2 ? gq[2] : gq[3]));
                process_level(level, pick(cur, level & 3u));
            }
        } else {
#pragma unroll 1
            for (uint32_t level0 = 0; level0 < L; level0 += 4) {
                if constexpr (kPacked) {
                    if ((L % 4) == 0) {
                        uint4 q = make_uint4(0u, 0u, 0u, 0u);
                        if (valid) q = __ldg(reinterpret_cast<const uint4*>(grad + ((size_t)b * L + level0) * C));
                        graw[0] = q.x; graw[1] = q.y; graw[2] = q.z; graw[3] = q.w;
                    } else {
#pragma unroll
                        for (uint32_t j = 0; j < 4; ++j)
                            graw[j] = (valid && level0 + j < L) ? __ldg(reinterpret_cast<const uint32_t*>(grad + ((size_t)b * L + level0 + j) * C)) : 0u;
                    }
                }
                const uint4 cur = make_uint4(graw[0], graw[1], graw[2], graw[3]);
#pragma unroll 1
                for (uint32_t lj = 0; lj < 4 && level0 + lj < L; ++lj) process_level(level0 + lj, pick(cur, lj));
            }
        }
    }
    if constexpr (COUNT) {
        const int total = warp_sum_i((int)reds);
        if (lane == 0 && red_lane_ops && total) atomicAdd(red_lane_ops, (unsigned long long)total);
    }
}

// grad_inputs[b,d] = sum_{l,c} grad[l,b,c] * dy_dx[b,l,d,c]   (gridencoder.cu:317-342).  In half mode the
// reference's running sum AND each product are at::Half, i.e. rounded to half after every operation.
template <typename T, uint32_t D, uint32_t C, bool GRAD_BLC>
__global__ void __launch_bounds__(256) input_backward_kernel(const T* __restrict__ grad, const T* __restrict__ dy_dx,
                                                             T* __restrict__ grad_inputs, uint32_t B, uint32_t L) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)B * D) return;
    const uint32_t b = (uint32_t)(t / D), d = (uint32_t)(t - (uint64_t)b * D);
    float acc = 0.f;
    for (uint32_t l = 0; l < L; ++l) {
        const T* gsrc = GRAD_BLC ? grad + ((size_t)b * L + l) * C : grad + ((size_t)l * B + b) * C;
        const T* jsrc = dy_dx + (((size_t)b * L + l) * D + d) * C;
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) {
            const float gv = ElemOps<T>::to_f(gsrc[c]), jv = ElemOps<T>::to_f(jsrc[c]);
            if constexpr (sizeof(T) == 4) {
                acc += gv * jv;
            } else {
                acc = ElemOps<T>::round(acc + ElemOps<T>::round(gv * jv));
            }
        }
    }
    if constexpr (sizeof(T) == 4) grad_inputs[t] = acc; else grad_inputs[t] = __float2half_rn(acc);
}

static __global__ void level_params_kernel(uint32_t L, float S, uint32_t H, float* scales, uint32_t* resolutions) {
    const uint32_t level = blockIdx.x * blockDim.x + threadIdx.x;
    if (level >= L) return;
    const float scale = exp2f(level * S) * H - 1.0f;
    scales[level] = scale;
    resolutions[level] = (uint32_t)ceil(scale) + 1;
}

// ---- host-side dispatch -------------------------------------------------------------------------
// Levels per thread: 4 keeps 16-byte (C=2, half) output stores and 32 gathers in flight per thread
// while giving 4x more threads than a thread-per-point mapping (matters for the ~3e5-sample
// batches of a 64x64 training view).
template <uint32_t C> struct Lpt { static constexpr uint32_t value = (C <= 2) ? 4 : (C == 4 ? 2 : 1); };

template <typename T, uint32_t D, uint32_t C>
int forward_dispatch(const float* inputs, const T* table, const int* offsets, T* outputs, T* dy_dx, uint32_t B,
                     uint32_t L, float S, uint32_t H, uint32_t gridtype, bool align, int layout, cudaStream_t st) {
    constexpr uint32_t LPT = Lpt<C>::value;
    const uint32_t groups = (L + LPT - 1) / LPT;
    const uint64_t threads = (uint64_t)B * groups;
    const dim3 grid((unsigned)((threads + 255) / 256)), block(256);
    const bool blc = layout == NGP_LAYOUT_BLC;
#define NGP_FWD(OUT_BLC, DYDX)                                                                              \
    encode_forward_kernel<T, D, C, LPT, OUT_BLC, DYDX><<<grid, block, 0, st>>>(inputs, table, offsets, outputs, dy_dx, \
                                                                            B, L, S, H, gridtype, align)
    if (dy_dx) { if (blc) NGP_FWD(true, true); else NGP_FWD(false, true); }
    else       { if (blc) NGP_FWD(true, false); else NGP_FWD(false, false); }
#undef NGP_FWD
    return launch_status();
}

template <typename T, uint32_t D>
int forward_dispatch_c(const float* inputs, const void* table, const int* offsets, void* outputs, void* dy_dx,
                       uint32_t B, uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype, bool align,
                       int layout, cudaStream_t st) {
    const T* t = static_cast<const T*>(table);
    T* o = static_cast<T*>(outputs);
    T* j = static_cast<T*>(dy_dx);
    switch (C) {
        case 1: return forward_dispatch<T, D, 1>(inputs, t, offsets, o, j, B, L, S, H, gridtype, align, layout, st);
        case 2: return forward_dispatch<T, D, 2>(inputs, t, offsets, o, j, B, L, S, H, gridtype, align, layout, st);
        case 4: return forward_dispatch<T, D, 4>(inputs, t, offsets, o, j, B, L, S, H, gridtype, align, layout, st);
        case 8: return forward_dispatch<T, D, 8>(inputs, t, offsets, o, j, B, L, S, H, gridtype, align, layout, st);
        default: return NGP_ERR_UNSUPPORTED;  // gridencoder.cu:354
    }
}

template <typename T, typename GT, uint32_t D, uint32_t C>
int backward_launch(const T* grad, const float* inputs, const int* offsets, GT* grad_table, const T* dy_dx,
                    T* grad_inputs, uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype, bool align,
                    int layout, cudaStream_t st) {
    constexpr uint32_t LPT = Lpt<C>::value;
    const uint32_t groups = (L + LPT - 1) / LPT;
    const uint64_t threads = (uint64_t)B * groups;
    const dim3 grid((unsigned)((threads + 255) / 256)), block(256);
    const bool blc = layout == NGP_LAYOUT_BLC;
    bool launched = false;
    if constexpr (D == 3 && sizeof(GT) == 4 && C <= 2) {
        // hot path: ray-ordered samples, fp32 table -> warp-aggregated scatter
        if (blc && !g_disable_warpagg) {
            const dim3 g1((unsigned)(((uint64_t)B + 255) / 256));
            encode_backward_warpagg_kernel<T, C><<<g1, block, 0, st>>>(grad, inputs, offsets, grad_table, B, L, S, H, gridtype, align, nullptr, 0.f);
            launched = true;
        }
    }
    if (!launched) {
        if (blc) encode_backward_kernel<T, GT, D, C, LPT, true><<<grid, block, 0, st>>>(grad, inputs, offsets, grad_table, B, L, S, H, gridtype, align);
        else     encode_backward_kernel<T, GT, D, C, LPT, false><<<grid, block, 0, st>>>(grad, inputs, offsets, grad_table, B, L, S, H, gridtype, align);
    }
    int rc = launch_status();
    if (rc != NGP_OK) return rc;
    if (dy_dx && grad_inputs) {
        const dim3 g2((unsigned)(((uint64_t)B * D + 255) / 256));
        if (blc) input_backward_kernel<T, D, C, true><<<g2, block, 0, st>>>(grad, dy_dx, grad_inputs, B, L);
        else     input_backward_kernel<T, D, C, false><<<g2, block, 0, st>>>(grad, dy_dx, grad_inputs, B, L);
        rc = launch_status();
    }
    return rc;
}

template <typename T, uint32_t D, uint32_t C>
int backward_dispatch_gt(const T* grad, const float* inputs, const int* offsets, void* grad_table, const T* dy_dx,
                         T* grad_inputs, uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype, bool align,
                         int layout, int gt, cudaStream_t st) {
    if (gt == NGP_F32)
        return backward_launch<T, float, D, C>(grad, inputs, offsets, static_cast<float*>(grad_table), dy_dx, grad_inputs, B, L, S, H, gridtype, align, layout, st);
    if constexpr (C % 2 == 0) {
        if (gt == NGP_F16)
            return backward_launch<T, __half, D, C>(grad, inputs, offsets, static_cast<__half*>(grad_table), dy_dx, grad_inputs, B, L, S, H, gridtype, align, layout, st);
    }
    return NGP_ERR_UNSUPPORTED;  // half table with C == 1: the reference's path is a body-less stub (gridencoder.cu:22-26)
}

template <typename T, uint32_t D>
int backward_dispatch_c(const void* grad, const float* inputs, const int* offsets, void* grad_table, const void* dy_dx,
                        void* grad_inputs, uint32_t B, uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                        bool align, int layout, int gt, cudaStream_t st) {
    const T* g = static_cast<const T*>(grad);
    const T* j = static_cast<const T*>(dy_dx);
    T* gi = static_cast<T*>(grad_inputs);
    switch (C) {
        case 1: return backward_dispatch_gt<T, D, 1>(g, inputs, offsets, grad_table, j, gi, B, L, S, H, gridtype, align, layout, gt, st);
        case 2: return backward_dispatch_gt<T, D, 2>(g, inputs, offsets, grad_table, j, gi, B, L, S, H, gridtype, align, layout, gt, st);
        case 4: return backward_dispatch_gt<T, D, 4>(g, inputs, offsets, grad_table, j, gi, B, L, S, H, gridtype, align, layout, gt, st);
        case 8: return backward_dispatch_gt<T, D, 8>(g, inputs, offsets, grad_table, j, gi, B, L, S, H, gridtype, align, layout, gt, st);
        default: return NGP_ERR_UNSUPPORTED;
    }
}


// Per-dimension entry points: each D is instantiated in its own translation unit
// (grid_encode_d<D>.cu) so the 2 dtypes x 4 feature widths x layouts compile in parallel.
template <uint32_t D>
int forward_for_dim(const float* inputs, const void* table, const int* offsets, void* outputs, void* dy_dx, uint32_t B,
                    uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype, bool align, int dtype, int layout,
                    cudaStream_t st);
template <uint32_t D>
int backward_for_dim(const void* grad, const float* inputs, const int* offsets, void* grad_table, const void* dy_dx,
                     void* grad_inputs, uint32_t B, uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                     bool align, int dtype, int layout, int gt, cudaStream_t st);

#define NGP_GRID_INSTANTIATE_DIM(DIM)                                                                                  \
    template <>                                                                                                        \
    int forward_for_dim<DIM>(const float* inputs, const void* table, const int* offsets, void* outputs, void* dy_dx,   \
                             uint32_t B, uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype, bool align,   \
                             int dtype, int layout, cudaStream_t st) {                                                 \
        if (dtype == NGP_F32)                                                                                          \
            return forward_dispatch_c<float, DIM>(inputs, table, offsets, outputs, dy_dx, B, C, L, S, H, gridtype,     \
                                                  align, layout, st);                                                  \
        if (dtype == NGP_F16)                                                                                          \
            return forward_dispatch_c<__half, DIM>(inputs, table, offsets, outputs, dy_dx, B, C, L, S, H, gridtype,    \
                                                   align, layout, st);                                                 \
        return NGP_ERR_UNSUPPORTED;                                                                                    \
    }                                                                                                                  \
    template <>                                                                                                        \
    int backward_for_dim<DIM>(const void* grad, const float* inputs, const int* offsets, void* grad_table,             \
                              const void* dy_dx, void* grad_inputs, uint32_t B, uint32_t C, uint32_t L, float S,       \
                              uint32_t H, uint32_t gridtype, bool align, int dtype, int layout, int gt,                \
                              cudaStream_t st) {                                                                       \
        if (dtype == NGP_F32)                                                                                          \
            return backward_dispatch_c<float, DIM>(grad, inputs, offsets, grad_table, dy_dx, grad_inputs, B, C, L, S,  \
                                                   H, gridtype, align, layout, gt, st);                                \
        if (dtype == NGP_F16)                                                                                          \
            return backward_dispatch_c<__half, DIM>(grad, inputs, offsets, grad_table, dy_dx, grad_inputs, B, C, L, S, \
                                                    H, gridtype, align, layout, gt, st);                               \
        return NGP_ERR_UNSUPPORTED;                                                                                    \
    }

}  // namespace grid
}  // namespace ngp
