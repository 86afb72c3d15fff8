// Fused NeRF field for sm_100a: tiled-grid encode -> MLP(32 -> 64 -> 64 -> 4) -> trunc_exp / sigmoid, forward and
// backward, with the three GEMMs on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// Behavioural contract: nerf/network_grid.py:76-87 (common_forward) under fp16 autocast, i.e. what the reference
// evaluates as GridEncoder -> 3 x (cuBLAS half GEMM + bias [+ ReLU]) -> exp / sigmoid with ~10 elementwise kernels
// and the 64-wide activations round-tripping HBM.  Here one CTA owns a tile of 128 samples:
//   forward : each thread encodes one sample (32 fp16 features, same arithmetic as grid_encode.cuh) straight into a
//             shared-memory operand tile; one elected thread issues the UMMAs; every thread then reads ITS row of the
//             fp32 accumulator from TMEM (tcgen05.ld 32x32b: lane == row), adds the bias, rounds to fp16 (as the
//             reference's half GEMM output is), applies ReLU and writes the row as the next layer's operand.
//   backward: dgrad GEMMs (G x W) and weight-gradient GEMMs (X^T x G, K = the 128 samples of the tile) run on the
//             same operand tiles - the core-matrix layout reads as K-major for one and MN-major for the other
//             (tcgen05.cuh) - and the weight gradients stay resident in TMEM across all tiles of the CTA.
// The hidden activations are saved in fp16 by the forward for the backward (h1, h2, encoding: 320 B / sample).
#include "grid_encode.cuh"
#include "tcgen05.cuh"

namespace ngp {
namespace field {

constexpr uint32_t kTile = 128;   // samples per CTA tile == threads per CTA == TMEM lanes
constexpr uint32_t kLevels = 16;  // grid levels (x 2 features = 32 MLP inputs)
constexpr uint32_t kIn = 32, kHid = 64, kOut = 4, kOutPad = 16;
constexpr uint32_t kRg32 = (32 / 8) * 128;  // row-group stride of a 32-column tile  (512 B)
constexpr uint32_t kRg64 = (64 / 8) * 128;  // 64-column tile (1024 B)
constexpr uint32_t kRg16 = (16 / 8) * 128;  // 16-column tile (256 B)

struct Weights {
    const __half *w1, *b1, *w2, *b2, *w3, *b3;  // fp16 copies, row-major [out, in] as nn.Linear stores them
};

struct GridDesc {
    const __half* table;
    const int* offsets;
    float S;
    uint32_t H, gridtype;
    int align_corners;
    float bound;
};

// global row-major [R x C] fp16 weights -> core-matrix tile; rows >= R_valid are zero-filled
NGP_DEVINL void load_weight_tile(const __half* g, uint32_t R_valid, uint32_t R, uint32_t C, uint8_t* smem, uint32_t rg) {
    const uint32_t chunks = C / 8;
    for (uint32_t i = threadIdx.x; i < R * chunks; i += blockDim.x) {
        const uint32_t r = i / chunks, cc = i % chunks;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r < R_valid) v = *reinterpret_cast<const uint4*>(g + (size_t)r * C + cc * 8);
        *reinterpret_cast<uint4*>(smem + tc::tile_chunk_off(r, cc, rg)) = v;
    }
}

NGP_DEVINL uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
NGP_DEVINL float2 unpack_half2(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }

// One level of one sample's encoding as a packed half2 - the arithmetic of encode_forward_kernel<__half, 3, 2>.
NGP_DEVINL uint32_t encode_level(const float (&x01)[3], bool oob, const GridDesc& gd, const grid::LevelParams& lp) {
    float acc[2] = {0.f, 0.f};
    if (!oob) {
        const __half* tbl = gd.table + (size_t)lp.offset * 2;
        float frac[3];
        uint32_t base[3];
        grid::locate<3>(x01, lp.scale, gd.align_corners != 0, frac, base);
        float rows[8][2], wts[8];
        // levels that ignore trailing axes (tiled, resolution >= 256: z) have only 4 distinct rows per cell
        const uint32_t axes_used = grid::level_axes_used<3>(gd.gridtype, gd.align_corners != 0, lp.hashmap_size, lp.resolution);
        const uint32_t distinct = 1u << axes_used;
#pragma unroll
        for (uint32_t corner = 0; corner < 8; ++corner) {
            float w = 1;
            uint32_t p[3];
#pragma unroll
            for (uint32_t d = 0; d < 3; ++d) {
                if ((corner & (1u << d)) == 0) { w *= 1 - frac[d]; p[d] = base[d]; }
                else                           { w *= frac[d];     p[d] = base[d] + 1; }
            }
            wts[corner] = w;
            if (corner < distinct) {
                const uint32_t row = grid::lattice_row<3>(gd.gridtype, gd.align_corners != 0, lp.hashmap_size, lp.resolution, p);
                grid::load_row<__half, 2>(tbl + (size_t)row * 2, rows[corner]);
            }
        }
        grid::replicate_rows<3, 2>(rows, axes_used);
#pragma unroll
        for (uint32_t corner = 0; corner < 8; ++corner) {
#pragma unroll
            for (uint32_t c = 0; c < 2; ++c) {
                const float prod = grid::ElemOps<__half>::round(wts[corner] * rows[corner][c]);
                acc[c] = grid::ElemOps<__half>::round(acc[c] + prod);
            }
        }
    }
    return pack_half2(acc[0], acc[1]);
}

// shared-memory carve-up of the forward kernel
struct FwdSmem {
    static constexpr uint32_t a0 = 0;                       // [128 x 32] encodings
    static constexpr uint32_t a1 = a0 + 16 * kRg32;         // [128 x 64] hidden activations
    static constexpr uint32_t w1 = a1 + 16 * kRg64;         // [64 x 32]
    static constexpr uint32_t w2 = w1 + 8 * kRg32;          // [64 x 64]
    static constexpr uint32_t w3 = w2 + 8 * kRg64;          // [16 x 64] (rows 4.. zero)
    static constexpr uint32_t bias = w3 + 2 * kRg64;        // b1[64] b2[64] b3[4] as float
    static constexpr uint32_t total = bias + (64 + 64 + 4) * 4;
};

struct FwdArgs {
    const float* xyzs;
    uint32_t M;
    const int* count_ptr;  // optional device-side row count (rows >= *count_ptr are skipped)
    GridDesc gd;
    Weights w;
    float* sigma;   // [M] fp32
    float* rgb;     // [M, 3] fp32 holding the fp16-rounded sigmoid
    __half* enc;    // optional saves for the backward
    __half* h1;
    __half* h2;
};

// bias + fp16 rounding + ReLU of one accumulator row segment, written as 16-byte chunks of the next operand tile
template <uint32_t NCOLS>
NGP_DEVINL void hidden_epilogue(const uint32_t (&acc)[NCOLS], const float* bias, uint32_t col0, uint32_t r, uint8_t* tile,
                                __half* save_row) {
#pragma unroll
    for (uint32_t q = 0; q < NCOLS / 8; ++q) {
        uint32_t w[4];
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            const uint32_t c = q * 8 + 2 * j;
            // the reference's Linear under autocast returns half(acc + bias); ReLU on the half value
            const float v0 = fmaxf(__half2float(__float2half_rn(__uint_as_float(acc[c]) + bias[col0 + c])), 0.f);
            const float v1 = fmaxf(__half2float(__float2half_rn(__uint_as_float(acc[c + 1]) + bias[col0 + c + 1])), 0.f);
            w[j] = pack_half2(v0, v1);
        }
        const uint4 v = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(tile + tc::tile_chunk_off(r, col0 / 8 + q, kRg64)) = v;
        if (save_row) *reinterpret_cast<uint4*>(save_row + col0 + q * 8) = v;
    }
}

__global__ void __launch_bounds__(kTile, 4) field_forward_kernel(const FwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ grid::LevelParams s_levels[kLevels];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t r = threadIdx.x, warp = r >> 5;

    if (r < kLevels) s_levels[r] = grid::make_level(a.gd.offsets, r, a.gd.S, a.gd.H);
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 64);
    if (r == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    load_weight_tile(a.w.w1, kHid, kHid, kIn, smem + FwdSmem::w1, kRg32);
    load_weight_tile(a.w.w2, kHid, kHid, kHid, smem + FwdSmem::w2, kRg64);
    load_weight_tile(a.w.w3, kOut, kOutPad, kHid, smem + FwdSmem::w3, kRg64);
    float* s_bias = reinterpret_cast<float*>(smem + FwdSmem::bias);
    if (r < 64) { s_bias[r] = __half2float(a.w.b1[r]); s_bias[64 + r] = __half2float(a.w.b2[r]); }
    if (r < 4) s_bias[128 + r] = __half2float(a.w.b3[r]);
    tc::fence_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t tmem_row = tc::tmem_addr(tmem, warp * 32, 0);

    const uint32_t sa0 = tc::smem_u32(smem + FwdSmem::a0), sa1 = tc::smem_u32(smem + FwdSmem::a1);
    const uint32_t sw1 = tc::smem_u32(smem + FwdSmem::w1), sw2 = tc::smem_u32(smem + FwdSmem::w2);
    const uint32_t sw3 = tc::smem_u32(smem + FwdSmem::w3);
    constexpr uint32_t idesc_h = tc::instr_desc(128, kHid, false, false);
    constexpr uint32_t idesc_o = tc::instr_desc(128, kOutPad, false, false);

    const uint32_t M = a.count_ptr ? min((uint32_t)max(*a.count_ptr, 0), a.M) : a.M;
    const uint32_t n_tiles = (M + kTile - 1) / kTile;
    const float inv_2b = __fdiv_rn(1.0f, 2 * a.gd.bound);
    uint32_t phase = 0;

    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t m = tile * kTile + r;
        const bool active = m < M;
        float x[3] = {0.f, 0.f, 0.f};
        if (active) { x[0] = __ldg(a.xyzs + (size_t)m * 3); x[1] = __ldg(a.xyzs + (size_t)m * 3 + 1); x[2] = __ldg(a.xyzs + (size_t)m * 3 + 2); }
        // GridEncoder.forward maps [-bound, bound] -> [0, 1] as (x + bound) * (1 / (2 bound)) (grid.py:142)
        const float x01[3] = {__fmul_rn(__fadd_rn(x[0], a.gd.bound), inv_2b), __fmul_rn(__fadd_rn(x[1], a.gd.bound), inv_2b),
                              __fmul_rn(__fadd_rn(x[2], a.gd.bound), inv_2b)};
        const bool oob = grid::out_of_unit_cube<3>(x01);
        for (uint32_t cc = 0; cc < 4; ++cc) {  // 4 levels = one 16-byte chunk of the operand row
            uint32_t e[4];
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j) e[j] = encode_level(x01, oob, a.gd, s_levels[cc * 4 + j]);
            const uint4 v = make_uint4(e[0], e[1], e[2], e[3]);
            *reinterpret_cast<uint4*>(smem + FwdSmem::a0 + tc::tile_chunk_off(r, cc, kRg32)) = v;
            if (a.enc && active) *reinterpret_cast<uint4*>(a.enc + (size_t)m * kIn + cc * 8) = v;
        }
        tc::fence_async_smem();
        tc::tc_fence_before_sync();
        __syncthreads();

        // ---- layer 1: [128 x 32] x W1^T -> TMEM [128 x 64] ----
        if (r == 0) {
            tc::tc_fence_after_sync();
#pragma unroll
            for (uint32_t k = 0; k < kIn / 16; ++k)
                tc::umma_f16(tmem, tc::smem_desc(sa0 + k * 256, 128, kRg32), tc::smem_desc(sw1 + k * 256, 128, kRg32), idesc_h, k > 0);
            tc::umma_commit(&bar);
        }
        tc::mbar_wait(&bar, phase); phase ^= 1;
        tc::tc_fence_after_sync();
#pragma unroll
        for (uint32_t half = 0; half < 2; ++half) {
            uint32_t acc[32];
            tc::tmem_ld_x32(tmem_row + half * 32, acc);
            tc::tmem_ld_wait();
            hidden_epilogue<32>(acc, s_bias, half * 32, r, smem + FwdSmem::a1, (a.h1 && active) ? a.h1 + (size_t)m * kHid : nullptr);
        }
        tc::fence_async_smem();
        tc::tc_fence_before_sync();
        __syncthreads();

        // ---- layer 2: [128 x 64] x W2^T ----
        if (r == 0) {
            tc::tc_fence_after_sync();
#pragma unroll
            for (uint32_t k = 0; k < kHid / 16; ++k)
                tc::umma_f16(tmem, tc::smem_desc(sa1 + k * 256, 128, kRg64), tc::smem_desc(sw2 + k * 256, 128, kRg64), idesc_h, k > 0);
            tc::umma_commit(&bar);
        }
        tc::mbar_wait(&bar, phase); phase ^= 1;
        tc::tc_fence_after_sync();
#pragma unroll
        for (uint32_t half = 0; half < 2; ++half) {
            uint32_t acc[32];
            tc::tmem_ld_x32(tmem_row + half * 32, acc);
            tc::tmem_ld_wait();
            hidden_epilogue<32>(acc, s_bias + 64, half * 32, r, smem + FwdSmem::a1, (a.h2 && active) ? a.h2 + (size_t)m * kHid : nullptr);
        }
        tc::fence_async_smem();
        tc::tc_fence_before_sync();
        __syncthreads();

        // ---- layer 3: [128 x 64] x W3^T (4 outputs padded to 16) ----
        if (r == 0) {
            tc::tc_fence_after_sync();
#pragma unroll
            for (uint32_t k = 0; k < kHid / 16; ++k)
                tc::umma_f16(tmem, tc::smem_desc(sa1 + k * 256, 128, kRg64), tc::smem_desc(sw3 + k * 256, 128, kRg64), idesc_o, k > 0);
            tc::umma_commit(&bar);
        }
        tc::mbar_wait(&bar, phase); phase ^= 1;
        tc::tc_fence_after_sync();
        {
            uint32_t acc[16];
            tc::tmem_ld_x16(tmem_row, acc);
            tc::tmem_ld_wait();
            if (active) {
                float h[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) h[j] = __half2float(__float2half_rn(__uint_as_float(acc[j]) + s_bias[128 + j]));
                // sigma = trunc_exp(h0 + 5 exp(-|x|^2 / (2 * 0.2^2))) in fp32 (network_grid.py:66-84, activation.py:8)
                const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(x[0], x[0]), __fmul_rn(x[1], x[1])), __fmul_rn(x[2], x[2]));
                const float blob = 5 * expf(-d2 * 12.5f);
                a.sigma[m] = expf(h[0] + blob);
                // albedo = sigmoid(h1..3) computed on the half tensor -> rounded to half
#pragma unroll
                for (int j = 0; j < 3; ++j) a.rgb[(size_t)m * 3 + j] = __half2float(__float2half_rn(1.0f / (1.0f + expf(-h[j + 1]))));
            }
        }
        tc::tc_fence_before_sync();  // the next tile's first MMA overwrites these TMEM columns
    }
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 64);
}

// -------------------------------------------------------------------------------------------------------------
// backward
// -------------------------------------------------------------------------------------------------------------
struct BwdSmem {
    static constexpr uint32_t x = 0;                     // [128 x 64] activations (h2, then h1, then the 32-wide encoding)
    static constexpr uint32_t g = x + 16 * kRg64;        // [128 x 64] upstream grads of the current layer (dh2, then dh1)
    static constexpr uint32_t g3 = g + 16 * kRg64;       // [128 x 16] dh_out (4 used)
    static constexpr uint32_t w1 = g3 + 16 * kRg16;
    static constexpr uint32_t w2 = w1 + 8 * kRg32;
    static constexpr uint32_t w3 = w2 + 8 * kRg64;
    static constexpr uint32_t total = w3 + 2 * kRg64;
};
// TMEM columns
constexpr uint32_t kColD = 0, kColW3 = 64, kColW1 = 96, kColW2 = 128, kTmemColsBwd = 256;

struct BwdArgs {
    uint32_t M;
    const int* count_ptr;
    Weights w;
    const float* d_sigma;   // [M]
    const float* d_rgb;     // [M, 3]
    const float* sigma;     // forward outputs
    const float* rgb;
    const __half* enc;      // forward saves
    const __half* h1;
    const __half* h2;
    __half* d_enc;          // [M, 32] out: gradient wrt the encoding (feeds the grid scatter)
    float *gw1, *gb1, *gw2, *gb2, *gw3, *gb3;  // fp32 accumulators (+=, atomics)
};

NGP_DEVINL void load_row_to_tile(const __half* src_row, uint32_t ncols, uint32_t r, uint8_t* tile, uint32_t rg, bool active) {
#pragma unroll 8
    for (uint32_t cc = 0; cc < ncols / 8; ++cc) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (active) v = __ldg(reinterpret_cast<const uint4*>(src_row) + cc);
        *reinterpret_cast<uint4*>(tile + tc::tile_chunk_off(r, cc, rg)) = v;
    }
}

// dh = relu'(h) * half(acc): reads this thread's activation row from the X tile, writes its row of the G tile
NGP_DEVINL void relu_backward_epilogue(uint32_t tmem_row, uint32_t r, const uint8_t* xtile, uint8_t* gtile) {
#pragma unroll
    for (uint32_t half = 0; half < 2; ++half) {
        uint32_t acc[32];
        tc::tmem_ld_x32(tmem_row + kColD + half * 32, acc);
        tc::tmem_ld_wait();
#pragma unroll
        for (uint32_t q = 0; q < 4; ++q) {
            const uint32_t cc = half * 4 + q;
            const uint4 hv = *reinterpret_cast<const uint4*>(xtile + tc::tile_chunk_off(r, cc, kRg64));
            const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
            uint32_t w[4];
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j) {
                const float2 h = unpack_half2(hw[j]);
                const float g0 = h.x > 0.f ? __uint_as_float(acc[q * 8 + 2 * j]) : 0.f;
                const float g1 = h.y > 0.f ? __uint_as_float(acc[q * 8 + 2 * j + 1]) : 0.f;
                w[j] = pack_half2(g0, g1);
            }
            *reinterpret_cast<uint4*>(gtile + tc::tile_chunk_off(r, cc, kRg64)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

// column sum over the 128 rows of a gradient tile (bias gradient); thread j owns column j
NGP_DEVINL float tile_column_sum(const uint8_t* tile, uint32_t rg, uint32_t col) {
    float s = 0.f;
    const uint8_t* base = tile + (col >> 3) * 128u + (col & 7u) * 2u;
#pragma unroll 8
    for (uint32_t row = 0; row < kTile; ++row)
        s += __half2float(*reinterpret_cast<const __half*>(base + (row >> 3) * rg + (row & 7u) * 16u));
    return s;
}

__global__ void __launch_bounds__(kTile, 2) field_backward_kernel(const BwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t r = threadIdx.x, warp = r >> 5, lane = r & 31;

    if (warp == 0) tc::tmem_alloc(&tmem_base_s, kTmemColsBwd);
    if (r == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    load_weight_tile(a.w.w1, kHid, kHid, kIn, smem + BwdSmem::w1, kRg32);
    load_weight_tile(a.w.w2, kHid, kHid, kHid, smem + BwdSmem::w2, kRg64);
    load_weight_tile(a.w.w3, kOut, kOutPad, kHid, smem + BwdSmem::w3, kRg64);
    tc::fence_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t tmem_row = tc::tmem_addr(tmem, warp * 32, 0);

    const uint32_t sx = tc::smem_u32(smem + BwdSmem::x), sg = tc::smem_u32(smem + BwdSmem::g), sg3 = tc::smem_u32(smem + BwdSmem::g3);
    const uint32_t sw1 = tc::smem_u32(smem + BwdSmem::w1), sw2 = tc::smem_u32(smem + BwdSmem::w2), sw3 = tc::smem_u32(smem + BwdSmem::w3);
    // data gradients: A = upstream grads (K-major), B = weights [out, in] read MN-major (K = out)
    constexpr uint32_t id_dgrad64 = tc::instr_desc(128, 64, false, true);
    constexpr uint32_t id_dgrad32 = tc::instr_desc(128, 32, false, true);
    // weight gradients: both operands MN-major, K = samples
    constexpr uint32_t id_w3 = tc::instr_desc(64, 16, true, true);   // D[i, o] = sum_m h2[m, i] dh3[m, o]
    constexpr uint32_t id_w2 = tc::instr_desc(64, 64, true, true);   // D[i, o] = sum_m h1[m, i] dh2[m, o]
    constexpr uint32_t id_w1 = tc::instr_desc(64, 32, true, true);   // D[o, i] = sum_m dh1[m, o] enc[m, i]

    const uint32_t M = a.count_ptr ? min((uint32_t)max(*a.count_ptr, 0), a.M) : a.M;
    const uint32_t n_tiles = (M + kTile - 1) / kTile;
    uint32_t phase = 0;
    float gb_acc2 = 0.f, gb_acc1 = 0.f, gb_acc3 = 0.f;  // thread j accumulates column j of the bias gradients
    bool first = true;

    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t m = tile * kTile + r;
        const bool active = m < M;

        // ---- output layer: dh_out from (d_sigma, d_rgb) through trunc_exp / sigmoid -----------------------------
        float dh[4] = {0.f, 0.f, 0.f, 0.f};
        if (active) {
            // trunc_exp backward: g * exp(clamp(x, -15, 15)) with sigma = exp(x) (activation.py:13-15)
            const float s = fminf(fmaxf(a.sigma[m], 3.0590232e-7f), 3269017.4f);
            dh[0] = __half2float(__float2half_rn(a.d_sigma[m] * s));
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const float y = a.rgb[(size_t)m * 3 + j];
                const float gh = __half2float(__float2half_rn(a.d_rgb[(size_t)m * 3 + j]));
                dh[j + 1] = __half2float(__float2half_rn(gh * (1.f - y) * y));
            }
        }
        *reinterpret_cast<uint4*>(smem + BwdSmem::g3 + tc::tile_chunk_off(r, 0, kRg16)) =
            make_uint4(pack_half2(dh[0], dh[1]), pack_half2(dh[2], dh[3]), 0u, 0u);
        *reinterpret_cast<uint4*>(smem + BwdSmem::g3 + tc::tile_chunk_off(r, 1, kRg16)) = make_uint4(0u, 0u, 0u, 0u);
        load_row_to_tile(a.h2 + (size_t)m * kHid, kHid, r, smem + BwdSmem::x, kRg64, active);
        tc::fence_async_smem();
        tc::tc_fence_before_sync();
        __syncthreads();
        if (r == 0) {
            tc::tc_fence_after_sync();
            // dh2_pre[128 x 64] = dh3[128 x 16] x W3[16(out) x 64(in)]
            tc::umma_f16(tmem + kColD, tc::smem_desc(sg3, 128, kRg16), tc::smem_desc(sw3, kRg64, 128), id_dgrad64, 0);
            // gW3^T[64(in) x 16(out)] += h2^T dh3
#pragma unroll
            for (uint32_t k = 0; k < kTile / 16; ++k)
                tc::umma_f16(tmem + kColW3, tc::smem_desc(sx + k * 2 * kRg64, kRg64, 128), tc::smem_desc(sg3 + k * 2 * kRg16, kRg16, 128),
                             id_w3, (!first || k > 0) ? 1u : 0u);
            tc::umma_commit(&bar);
        }
        if (r < 4) gb_acc3 += tile_column_sum(smem + BwdSmem::g3, kRg16, r);
        tc::mbar_wait(&bar, phase); phase ^= 1;
        tc::tc_fence_after_sync();
        relu_backward_epilogue(tmem_row, r, smem + BwdSmem::x, smem + BwdSmem::g);      // dh2 -> G
        load_row_to_tile(a.h1 + (size_t)m * kHid, kHid, r, smem + BwdSmem::x, kRg64, active);  // X <- h1 (after reading h2 above)
        tc::fence_async_smem();
        tc::tc_fence_before_sync();
        __syncthreads();

        // ---- hidden layer 2 ------------------------------------------------------------------------------------
        if (r == 0) {
            tc::tc_fence_after_sync();
#pragma unroll
            for (uint32_t k = 0; k < kHid / 16; ++k)   // dh1_pre = dh2[128 x 64] x W2[64(out) x 64(in)]
                tc::umma_f16(tmem + kColD, tc::smem_desc(sg + k * 256, 128, kRg64), tc::smem_desc(sw2 + k * 2 * kRg64, kRg64, 128), id_dgrad64, k > 0);
#pragma unroll
            for (uint32_t k = 0; k < kTile / 16; ++k)  // gW2^T[64(in) x 64(out)] += h1^T dh2
                tc::umma_f16(tmem + kColW2, tc::smem_desc(sx + k * 2 * kRg64, kRg64, 128), tc::smem_desc(sg + k * 2 * kRg64, kRg64, 128),
                             id_w2, (!first || k > 0) ? 1u : 0u);
            tc::umma_commit(&bar);
        }
        if (r < 64) gb_acc2 += tile_column_sum(smem + BwdSmem::g, kRg64, r);
        tc::mbar_wait(&bar, phase); phase ^= 1;
        __syncthreads();  // the column sums above read every row of G; the epilogue below rewrites it
        tc::tc_fence_after_sync();
        relu_backward_epilogue(tmem_row, r, smem + BwdSmem::x, smem + BwdSmem::g);      // dh1 -> G
        load_row_to_tile(a.enc + (size_t)m * kIn, kIn, r, smem + BwdSmem::x, kRg32, active);   // X <- encoding (32 wide)
        tc::fence_async_smem();
        tc::tc_fence_before_sync();
        __syncthreads();

        // ---- hidden layer 1 ------------------------------------------------------------------------------------
        if (r == 0) {
            tc::tc_fence_after_sync();
#pragma unroll
            for (uint32_t k = 0; k < kHid / 16; ++k)   // d_enc[128 x 32] = dh1[128 x 64] x W1[64(out) x 32(in)]
                tc::umma_f16(tmem + kColD, tc::smem_desc(sg + k * 256, 128, kRg64), tc::smem_desc(sw1 + k * 2 * kRg32, kRg32, 128), id_dgrad32, k > 0);
#pragma unroll
            for (uint32_t k = 0; k < kTile / 16; ++k)  // gW1[64(out) x 32(in)] += dh1^T enc
                tc::umma_f16(tmem + kColW1, tc::smem_desc(sg + k * 2 * kRg64, kRg64, 128), tc::smem_desc(sx + k * 2 * kRg32, kRg32, 128),
                             id_w1, (!first || k > 0) ? 1u : 0u);
            tc::umma_commit(&bar);
        }
        if (r < 64) gb_acc1 += tile_column_sum(smem + BwdSmem::g, kRg64, r);
        tc::mbar_wait(&bar, phase); phase ^= 1;
        tc::tc_fence_after_sync();
        {
            uint32_t acc[32];
            tc::tmem_ld_x32(tmem_row + kColD, acc);
            tc::tmem_ld_wait();
            if (active) {
#pragma unroll
                for (uint32_t cc = 0; cc < 4; ++cc) {
                    uint32_t w[4];
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j) w[j] = pack_half2(__uint_as_float(acc[cc * 8 + 2 * j]), __uint_as_float(acc[cc * 8 + 2 * j + 1]));
                    *reinterpret_cast<uint4*>(a.d_enc + (size_t)m * kIn + cc * 8) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
        tc::tc_fence_before_sync();
        __syncthreads();  // X / G / G3 tiles are rewritten by the next tile
        first = false;
    }

    // ---- flush the weight gradients held in TMEM (M = 64 accumulators: rows 16w..16w+15 live in lanes 32w..32w+15) ----
    tc::tc_fence_after_sync();
    if (!first) {
        const uint32_t row = warp * 16 + lane;  // valid for lane < 16
        {   // gW3[o, i] = D[i, o]
            uint32_t acc[16];
            tc::tmem_ld_x16(tmem_row + kColW3, acc);
            tc::tmem_ld_wait();
            if (lane < 16)
                for (uint32_t o = 0; o < kOut; ++o) atomicAdd(a.gw3 + o * kHid + row, __uint_as_float(acc[o]));
        }
        for (uint32_t c0 = 0; c0 < 64; c0 += 32) {  // gW2[o, i] = D[i, o]
            uint32_t acc[32];
            tc::tmem_ld_x32(tmem_row + kColW2 + c0, acc);
            tc::tmem_ld_wait();
            if (lane < 16)
                for (uint32_t o = 0; o < 32; ++o) atomicAdd(a.gw2 + (c0 + o) * kHid + row, __uint_as_float(acc[o]));
        }
        {   // gW1[o, i] = D[o, i]
            uint32_t acc[32];
            tc::tmem_ld_x32(tmem_row + kColW1, acc);
            tc::tmem_ld_wait();
            if (lane < 16)
                for (uint32_t i = 0; i < kIn; ++i) atomicAdd(a.gw1 + row * kIn + i, __uint_as_float(acc[i]));
        }
        if (r < 64) { atomicAdd(a.gb2 + r, gb_acc2); atomicAdd(a.gb1 + r, gb_acc1); }
        if (r < 4) atomicAdd(a.gb3 + r, gb_acc3);
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, kTmemColsBwd);
}

}  // namespace field
}  // namespace ngp

using namespace ngp;

static int check_field_dims(uint32_t L, uint32_t C, uint32_t D, uint32_t hidden, uint32_t out) {
    if (L != field::kLevels || C != 2 || D != 3 || hidden != field::kHid || out != field::kOut) return NGP_ERR_UNSUPPORTED;
    return NGP_OK;
}

extern "C" int ngp_field_forward(const float* xyzs, uint32_t M, const int* count_ptr, const void* table, const int* offsets,
                                 uint32_t L, uint32_t C, float S, uint32_t H, uint32_t gridtype, int align_corners,
                                 float bound, const void* w1, const void* b1, const void* w2, const void* b2,
                                 const void* w3, const void* b3, uint32_t hidden, uint32_t out_dim, float* sigma,
                                 float* rgb, void* enc_save, void* h1_save, void* h2_save, void* stream) {
    if (!xyzs || !table || !offsets || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !sigma || !rgb) return NGP_ERR_BAD_ARG;
    int rc = check_field_dims(L, C, 3, hidden, out_dim);
    if (rc != NGP_OK) return rc;
    if (M == 0) return NGP_OK;
    field::FwdArgs a;
    a.xyzs = xyzs; a.M = M; a.count_ptr = count_ptr;
    a.gd = {static_cast<const __half*>(table), offsets, S, H, gridtype, align_corners, bound};
    a.w = {static_cast<const __half*>(w1), static_cast<const __half*>(b1), static_cast<const __half*>(w2),
           static_cast<const __half*>(b2), static_cast<const __half*>(w3), static_cast<const __half*>(b3)};
    a.sigma = sigma; a.rgb = rgb;
    a.enc = static_cast<__half*>(enc_save); a.h1 = static_cast<__half*>(h1_save); a.h2 = static_cast<__half*>(h2_save);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(field::field_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)field::FwdSmem::total);
        attr_set = true;
    }
    const int tiles = cdiv(M, field::kTile);
    const int grid = min(tiles, num_sms() * 4);
    field::field_forward_kernel<<<grid, field::kTile, field::FwdSmem::total, as_stream(stream)>>>(a);
    return launch_status();
}

extern "C" int ngp_field_backward(uint32_t M, const int* count_ptr, const void* w1, const void* w2, const void* w3,
                                  uint32_t hidden, uint32_t out_dim, const float* d_sigma, const float* d_rgb,
                                  const float* sigma, const float* rgb, const void* enc_save, const void* h1_save,
                                  const void* h2_save, void* d_enc, float* gw1, float* gb1, float* gw2, float* gb2,
                                  float* gw3, float* gb3, void* stream) {
    if (!w1 || !w2 || !w3 || !d_sigma || !d_rgb || !sigma || !rgb || !enc_save || !h1_save || !h2_save || !d_enc || !gw1 ||
        !gb1 || !gw2 || !gb2 || !gw3 || !gb3)
        return NGP_ERR_BAD_ARG;
    int rc = check_field_dims(field::kLevels, 2, 3, hidden, out_dim);
    if (rc != NGP_OK) return rc;
    if (M == 0) return NGP_OK;
    field::BwdArgs a;
    a.M = M; a.count_ptr = count_ptr;
    a.w = {static_cast<const __half*>(w1), nullptr, static_cast<const __half*>(w2), nullptr, static_cast<const __half*>(w3), nullptr};
    a.d_sigma = d_sigma; a.d_rgb = d_rgb; a.sigma = sigma; a.rgb = rgb;
    a.enc = static_cast<const __half*>(enc_save); a.h1 = static_cast<const __half*>(h1_save); a.h2 = static_cast<const __half*>(h2_save);
    a.d_enc = static_cast<__half*>(d_enc);
    a.gw1 = gw1; a.gb1 = gb1; a.gw2 = gw2; a.gb2 = gb2; a.gw3 = gw3; a.gb3 = gb3;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(field::field_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)field::BwdSmem::total);
        attr_set = true;
    }
    const int tiles = cdiv(M, field::kTile);
    const int grid = min(tiles, num_sms() * 2);
    field::field_backward_kernel<<<grid, field::kTile, field::BwdSmem::total, as_stream(stream)>>>(a);
    return launch_status();
}
