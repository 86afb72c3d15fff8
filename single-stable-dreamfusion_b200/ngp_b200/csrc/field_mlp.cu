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

int g_fwd_ctas_per_sm = 0;   // ngp_field_set_option(0, n): resident CTAs per SM the grid is sized for (0 = 8 / groups)
int g_fwd_carveout = -1;     // ngp_field_set_option(1, percent); -1 = driver default
int g_fwd_groups = 2;        // ngp_field_set_option(2, g): 128-thread groups per CTA (1, 2 or 4)

constexpr uint32_t kTile = 128;   // samples per CTA tile == threads per CTA == TMEM lanes
constexpr uint32_t kLevels = 16;  // grid levels (x 2 features = 32 MLP inputs)
constexpr uint32_t kIn = 32, kHid = 64, kOut = 4, kOutPad = 16;
// chunk strides (bytes between consecutive 8-column chunks) of the core-matrix tiles, = rows * 16 (tcgen05.cuh)
constexpr uint32_t kCs128 = kTile * 16;  // activation / gradient tiles: 128 samples
constexpr uint32_t kCs64 = kHid * 16;    // W1 [64 x 32], W2 [64 x 64]
constexpr uint32_t kCs16 = kOutPad * 16; // W3 padded to [16 x 64]
constexpr uint32_t kEncTileBytes = kTile * kIn * 2;   // 8 KB: one saved tile of encodings
constexpr uint32_t kHidTileBytes = kTile * kHid * 2;  // 16 KB: one saved tile of hidden activations

struct Weights {
    const __half *w1, *b1, *w2, *b2, *w3, *b3;  // fp16 copies, row-major [out, in] as nn.Linear stores them
};

struct GridDesc {
    const __half* table;
    const uint4* quads;   // optional quad table (ngp_grid_quad_table): row r of a linear level = the four (x, y) corners at r
    const int* offsets;
    float S;
    uint32_t H, gridtype;
    int align_corners;
    float bound;
};

// global row-major [R x C] fp16 weights -> core-matrix tile; rows >= R_valid are zero-filled
NGP_DEVINL void load_weight_tile(const __half* g, uint32_t R_valid, uint32_t R, uint32_t C, uint8_t* smem) {
    const uint32_t chunks = C / 8;
    for (uint32_t i = threadIdx.x; i < R * chunks; i += blockDim.x) {
        const uint32_t r = i / chunks, cc = i % chunks;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r < R_valid) v = *reinterpret_cast<const uint4*>(g + (size_t)r * C + cc * 8);
        *reinterpret_cast<uint4*>(smem + tc::tile_chunk_off(r, cc, R * 16)) = v;
    }
}

NGP_DEVINL uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
NGP_DEVINL float2 unpack_half2(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }

// One level of one sample's encoding as a packed half2 - the arithmetic of encode_forward_kernel<__half, 3, 2>
// (bit-equal to the reference's half kernel): rows resolved by FastLevel, one 4-byte gather per distinct corner,
// packed-half accumulate in the reference's corner order.
NGP_DEVINL uint32_t encode_level(const float (&x01)[3], const uint32_t* __restrict__ table_u32, bool align_corners,
                                 const grid::FastLevel<3>& lp) {
    float frac[3];
    uint32_t base[3];
    grid::locate<3>(x01, lp.scale, align_corners, frac, base);
    uint32_t ridx[8];
    grid::corner_rows<3>(lp, base, ridx);
    const uint32_t* __restrict__ tbl = table_u32 + lp.offset;
    uint32_t raw[8];
    // levels that ignore trailing axes (tiled, resolution >= 256: z) have only 4 distinct rows per cell
    if (lp.used == 2) {
#pragma unroll
        for (uint32_t c = 0; c < 4; ++c) { raw[c] = __ldg(tbl + ridx[c]); raw[c + 4] = raw[c]; }
    } else {
#pragma unroll
        for (uint32_t c = 0; c < 8; ++c) raw[c] = __ldg(tbl + ridx[c]);  // (used == 1 reloads equal rows: same values)
    }
    float wts[8];
    grid::corner_weights<3>(frac, wts);
    __half2 acc = __floats2half2_rn(0.f, 0.f);
#pragma unroll
    for (uint32_t c = 0; c < 8; ++c) acc = grid::half2_axpy(acc, wts[c], raw[c]);
    return *reinterpret_cast<uint32_t*>(&acc);
}

// The same level from the QUAD table (linear / tiled levels only): row r holds the features of rows r, r + 1, r + stride_y
// and r + stride_y + 1 (each wrapped like corner_rows wraps them), i.e. the cell's four corners of one z slice in corner
// order - ONE 16-byte gather per slice instead of four 4-byte gathers and their address arithmetic.  Same feature values,
// same accumulation order: bit-equal to encode_level.
NGP_DEVINL uint32_t wrap_row(const grid::FastLevel<3>& lp, uint32_t row) {
    if (lp.wrap == grid::kWrapMask) return row & lp.mask;
    if (lp.wrap == grid::kWrapMod) return row % lp.size;
    return row;
}
NGP_DEVINL uint32_t encode_level_quads(const float (&x01)[3], const uint4* __restrict__ quads, bool align_corners,
                                       const grid::FastLevel<3>& lp) {
    float frac[3];
    uint32_t base[3];
    grid::locate<3>(x01, lp.scale, align_corners, frac, base);
    const uint32_t lin = base[0] * lp.stride[0] + base[1] * lp.stride[1] + base[2] * lp.stride[2];
    const uint4* __restrict__ q = quads + lp.offset;
    const uint4 lo = __ldg(q + wrap_row(lp, lin));
    uint4 hi = lo;                                   // levels that ignore z: both slices are the same rows
    if (lp.used == 3) hi = __ldg(q + wrap_row(lp, lin + lp.stride[2]));
    const uint32_t raw[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    float wts[8];
    grid::corner_weights<3>(frac, wts);
    __half2 acc = __floats2half2_rn(0.f, 0.f);
#pragma unroll
    for (uint32_t c = 0; c < 8; ++c) acc = grid::half2_axpy(acc, wts[c], raw[c]);
    return *reinterpret_cast<uint32_t*>(&acc);
}

// shared-memory carve-up of the forward kernel.  A CTA holds G independent GROUPS of 128 threads; the weights, biases and
// level descriptors are shared by the groups (they were 14.5 of the 38.5 KB of the one-group CTA, replicated five times per
// SM), and each group owns ONE [128 x 64] operand tile: the encodings occupy its first four chunks, h1 overwrites them once
// the layer-1 MMAs have read them, h2 overwrites h1 once the layer-2 MMAs have.  G = 2: 46.5 KB per 8 warps -> 4 CTAs = 32
// warps per SM (the register file's limit at 64 registers per thread) instead of 20.
struct FwdSmem {
    static constexpr uint32_t w1 = 0;                       // [64 x 32]
    static constexpr uint32_t w2 = w1 + 4 * kCs64;          // [64 x 64]
    static constexpr uint32_t w3 = w2 + 8 * kCs64;          // [16 x 64] (rows 4.. zero)
    static constexpr uint32_t bias = w3 + 8 * kCs16;        // b1[64] b2[64] b3[4] as float
    static constexpr uint32_t tiles = (bias + (64 + 64 + 4) * 4 + 127) / 128 * 128;   // G x [128 x 64]
    static constexpr uint32_t total(uint32_t G) { return tiles + G * kHidTileBytes; }
};

struct FwdArgs {
    const float* xyzs;
    uint32_t M;
    const int* count_ptr;  // optional device-side row count (rows >= *count_ptr are skipped)
    GridDesc gd;
    Weights w;
    float* sigma;   // [M] fp32
    float* rgb;     // [M, 3] fp32 holding the fp16-rounded sigmoid
    __half* enc;    // optional saves for the backward, TILE-MAJOR: tile t = the smem image of its 128 rows
    __half* h1;
    __half* h2;
};

// barrier among the 128 threads of one group (ids 1.. : 0 is __syncthreads)
NGP_DEVINL void group_sync(uint32_t grp) { asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory"); }

// bias + fp16 rounding + ReLU of one accumulator row segment, written as 16-byte chunks of the next operand tile and - for
// the backward - to the same offsets of the tile's image in global memory (a warp's 32 rows of one chunk are 512
// contiguous bytes either way).
// The reference's Linear under autocast returns half(acc + bias) and applies ReLU to the half value: two fp32 adds, ONE
// packed conversion (cvt.rn.f16x2.f32) and one packed max per column pair - same bits as converting, widening, comparing
// and re-packing each value (the earlier form, ~2x the instructions of this epilogue).
template <uint32_t NCOLS>
NGP_DEVINL void hidden_epilogue(const uint32_t (&acc)[NCOLS], const float* bias, uint32_t col0, uint32_t r, uint8_t* tile,
                                uint8_t* gsave) {
    const __half2 zero = __float2half2_rn(0.f);
#pragma unroll
    for (uint32_t q = 0; q < NCOLS / 8; ++q) {
        uint32_t w[4];
        const float4 b_lo = *reinterpret_cast<const float4*>(bias + col0 + q * 8);       // (broadcast reads, 16 bytes each)
        const float4 b_hi = *reinterpret_cast<const float4*>(bias + col0 + q * 8 + 4);
        const float bb[8] = {b_lo.x, b_lo.y, b_lo.z, b_lo.w, b_hi.x, b_hi.y, b_hi.z, b_hi.w};
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            const uint32_t c = q * 8 + 2 * j;
            __half2 h = __floats2half2_rn(__uint_as_float(acc[c]) + bb[2 * j], __uint_as_float(acc[c + 1]) + bb[2 * j + 1]);
            h = __hmax2(h, zero);
            w[j] = *reinterpret_cast<uint32_t*>(&h);
        }
        const uint32_t off = tc::tile_chunk_off(r, col0 / 8 + q, kCs128);
        const uint4 v = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(tile + off) = v;
        if (gsave) __stcs(reinterpret_cast<uint4*>(gsave + off), v);
    }
}

// Forward.  Per tile and group three phases, each: the group's threads write the operand tile -> fence + group barrier ->
// the group's thread 0 issues the layer's MMAs -> everyone waits for them on the group's mbarrier and reads its accumulator
// row from TMEM (the group's 64 columns).  Hazards on the one operand tile: every overwrite (h1 over the encodings, h2 over
// h1, the next tile's encodings over h2) happens after ALL threads of the group have observed the completion of the MMAs
// that read the old contents; the saves for the backward go from registers straight to global memory, so nothing else
// ever reads the tile.  TMEM: the next tile's first MMA is issued behind a group barrier that every thread reaches after
// its last tcgen05.ld has completed.
template <uint32_t G>
__global__ void __launch_bounds__(kTile * G, 8 / G) field_forward_kernel(const FwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ grid::FastLevel<3> s_levels[kLevels];
    __shared__ uint64_t bars[G];
    __shared__ uint32_t tmem_base_s;
    const uint32_t grp = threadIdx.x >> 7, r = threadIdx.x & 127u, warp = r >> 5;

    if (threadIdx.x < kLevels)
        s_levels[threadIdx.x] = grid::make_fast_level<3>(a.gd.offsets, threadIdx.x, a.gd.S, a.gd.H, a.gd.gridtype, a.gd.align_corners != 0);
    if (threadIdx.x < 32) tc::tmem_alloc(&tmem_base_s, 64 * G);
    if (r == 0) { tc::mbar_init(&bars[grp], 1); tc::fence_mbar_init(); }
    load_weight_tile(a.w.w1, kHid, kHid, kIn, smem + FwdSmem::w1);
    load_weight_tile(a.w.w2, kHid, kHid, kHid, smem + FwdSmem::w2);
    load_weight_tile(a.w.w3, kOut, kOutPad, kHid, smem + FwdSmem::w3);
    float* s_bias = reinterpret_cast<float*>(smem + FwdSmem::bias);
    if (threadIdx.x < 64) { s_bias[threadIdx.x] = __half2float(a.w.b1[threadIdx.x]); s_bias[64 + threadIdx.x] = __half2float(a.w.b2[threadIdx.x]); }
    if (threadIdx.x < 4) s_bias[128 + threadIdx.x] = __half2float(a.w.b3[threadIdx.x]);
    tc::fence_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    uint64_t* bar = &bars[grp];
    const uint32_t tmem = tmem_base_s + grp * 64;                      // this group's accumulator columns
    const uint32_t tmem_row = tc::tmem_addr(tmem, warp * 32, 0);

    const uint32_t sw1 = tc::smem_u32(smem + FwdSmem::w1), sw2 = tc::smem_u32(smem + FwdSmem::w2);
    const uint32_t sw3 = tc::smem_u32(smem + FwdSmem::w3);
    constexpr uint32_t idesc_h = tc::instr_desc(128, kHid, false, false);
    constexpr uint32_t idesc_o = tc::instr_desc(128, kOutPad, false, false);

    const uint32_t M = a.count_ptr ? min((uint32_t)max(*a.count_ptr, 0), a.M) : a.M;
    const uint32_t n_tiles = (M + kTile - 1) / kTile;
    const float inv_2b = __fdiv_rn(1.0f, 2 * a.gd.bound);
    const bool save = a.enc != nullptr;
    const uint32_t* table_u32 = reinterpret_cast<const uint32_t*>(a.gd.table);
    const uint4* quads = a.gd.quads;
    const bool align = a.gd.align_corners != 0;
    uint32_t phase = 0;
    uint8_t* buf = smem + FwdSmem::tiles + grp * kHidTileBytes;   // encodings (chunks 0..3), then h1, then h2
    const uint32_t sT = tc::smem_u32(buf);

    for (uint32_t tile = blockIdx.x * G + grp; tile < n_tiles; tile += gridDim.x * G) {
        const uint32_t m = tile * kTile + r;
        const bool active = m < M;
        float x[3] = {0.f, 0.f, 0.f};
        if (active) { x[0] = __ldg(a.xyzs + (size_t)m * 3); x[1] = __ldg(a.xyzs + (size_t)m * 3 + 1); x[2] = __ldg(a.xyzs + (size_t)m * 3 + 2); }
        // GridEncoder.forward maps [-bound, bound] -> [0, 1] as (x + bound) * (1 / (2 bound)) (grid.py:142)
        const float x01[3] = {__fmul_rn(__fadd_rn(x[0], a.gd.bound), inv_2b), __fmul_rn(__fadd_rn(x[1], a.gd.bound), inv_2b),
                              __fmul_rn(__fadd_rn(x[2], a.gd.bound), inv_2b)};
        const bool oob = grid::out_of_unit_cube<3>(x01);
        uint8_t* g_enc = save ? reinterpret_cast<uint8_t*>(a.enc) + (size_t)tile * kEncTileBytes : nullptr;
        for (uint32_t cc = 0; cc < 4; ++cc) {  // 4 levels = one 16-byte chunk of the operand row
            uint32_t e[4] = {0u, 0u, 0u, 0u};  // out-of-cube samples encode to zeros (gridencoder.cu:106-122)
            if (!oob) {
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j) {
                    const grid::FastLevel<3>& lp = s_levels[cc * 4 + j];
                    e[j] = (quads && !lp.hashed) ? encode_level_quads(x01, quads, align, lp) : encode_level(x01, table_u32, align, lp);
                }
            }
            const uint32_t off = tc::tile_chunk_off(r, cc, kCs128);
            const uint4 v = make_uint4(e[0], e[1], e[2], e[3]);
            *reinterpret_cast<uint4*>(buf + off) = v;
            if (save) __stcs(reinterpret_cast<uint4*>(g_enc + off), v);
        }
        tc::fence_async_smem();
        tc::tc_fence_before_sync();
        group_sync(grp);

        // ---- layer 1: [128 x 32] x W1^T -> TMEM [128 x 64] ----
        if (r == 0) {
            tc::tc_fence_after_sync();
#pragma unroll
            for (uint32_t k = 0; k < kIn / 16; ++k)
                tc::umma_f16(tmem, tc::desc_k_major(sT, kCs128, k), tc::desc_k_major(sw1, kCs64, k), idesc_h, k > 0);
            tc::umma_commit(bar);
        }
        tc::mbar_wait(bar, phase); phase ^= 1;
        tc::tc_fence_after_sync();
        {
            uint8_t* g_h1 = save ? reinterpret_cast<uint8_t*>(a.h1) + (size_t)tile * kHidTileBytes : nullptr;
#pragma unroll
            for (uint32_t half = 0; half < 2; ++half) {
                uint32_t acc[32];
                tc::tmem_ld_x32(tmem_row + half * 32, acc);
                tc::tmem_ld_wait();
                hidden_epilogue<32>(acc, s_bias, half * 32, r, buf, g_h1);
            }
        }
        tc::fence_async_smem();
        tc::tc_fence_before_sync();
        group_sync(grp);

        // ---- layer 2: [128 x 64] x W2^T ----
        if (r == 0) {
            tc::tc_fence_after_sync();
#pragma unroll
            for (uint32_t k = 0; k < kHid / 16; ++k)
                tc::umma_f16(tmem, tc::desc_k_major(sT, kCs128, k), tc::desc_k_major(sw2, kCs64, k), idesc_h, k > 0);
            tc::umma_commit(bar);
        }
        tc::mbar_wait(bar, phase); phase ^= 1;
        tc::tc_fence_after_sync();
        {
            uint8_t* g_h2 = save ? reinterpret_cast<uint8_t*>(a.h2) + (size_t)tile * kHidTileBytes : nullptr;
#pragma unroll
            for (uint32_t half = 0; half < 2; ++half) {
                uint32_t acc[32];
                tc::tmem_ld_x32(tmem_row + half * 32, acc);
                tc::tmem_ld_wait();
                hidden_epilogue<32>(acc, s_bias + 64, half * 32, r, buf, g_h2);
            }
        }
        tc::fence_async_smem();
        tc::tc_fence_before_sync();
        group_sync(grp);

        // ---- layer 3: [128 x 64] x W3^T (4 outputs padded to 16) ----
        if (r == 0) {
            tc::tc_fence_after_sync();
#pragma unroll
            for (uint32_t k = 0; k < kHid / 16; ++k)
                tc::umma_f16(tmem, tc::desc_k_major(sT, kCs128, k), tc::desc_k_major(sw3, kCs16, k), idesc_o, k > 0);
            tc::umma_commit(bar);
        }
        tc::mbar_wait(bar, phase); phase ^= 1;
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
        tc::tc_fence_before_sync();  // the next tile's first MMA (behind the next group barrier) overwrites these TMEM columns
    }
    __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc(tmem_base_s, 64 * G);
}

// -------------------------------------------------------------------------------------------------------------
// backward
// -------------------------------------------------------------------------------------------------------------
// Shared memory: the three saved activation tiles of the current sample tile arrive by bulk (TMA) loads, each into
// its own buffer with its own mbarrier, and are re-armed for the NEXT tile as soon as their last reader is done, so
// the loads of tile t+1 overlap the GEMMs of tile t.  The h1 and encoding tiles are followed by a constant chunk
// whose first column is 1: read as a [128 x (C+8)] MN-major operand, column C makes the weight-gradient GEMM
// deliver the bias gradient as well (sum over samples of dh x 1).
struct BwdSmem {
    static constexpr uint32_t xh2 = 0;                          // [128 x 64] h2
    static constexpr uint32_t xh1 = xh2 + kHidTileBytes;        // [128 x 64] h1
    static constexpr uint32_t one1 = xh1 + kHidTileBytes;       // [128 x 8] ones chunk (column 64 of the h1 operand)
    static constexpr uint32_t xe = one1 + kCs128;               // [128 x 32] encodings
    static constexpr uint32_t onee = xe + kEncTileBytes;        // [128 x 8] ones chunk (column 32 of the encoding operand)
    static constexpr uint32_t g = onee + kCs128;                // [128 x 64] upstream grads of the current layer (dh2, then dh1)
    static constexpr uint32_t g3 = g + kHidTileBytes;           // [128 x 16] dh_out (4 used)
    static constexpr uint32_t w1 = g3 + 2 * kCs128;
    static constexpr uint32_t w2 = w1 + 4 * kCs64;
    static constexpr uint32_t w3 = w2 + 8 * kCs64;
    static constexpr uint32_t total = w3 + 8 * kCs16;
};
// TMEM columns: D = data gradients [128 lanes x 64]; W3 = [64(in) x 16(out)]; W1 = [64(out) x 32(in) + bias col];
// W2 = [64(out) x 64(in) + bias col]
constexpr uint32_t kColD = 0, kColW3 = 64, kColW1 = 80, kColW2 = 128, kTmemColsBwd = 256;
constexpr uint32_t kNW1 = kIn + 8, kNW2 = kHid + 8;

struct BwdArgs {
    uint32_t M;
    const int* count_ptr;
    Weights w;
    const float* d_sigma;   // [M]
    const float* d_rgb;     // [M, 3]
    const float* sigma;     // forward outputs
    const float* rgb;
    const __half* enc;      // forward saves (tile-major)
    const __half* h1;
    const __half* h2;
    __half* d_enc;          // [M, 32] row-major out: gradient wrt the encoding (feeds the grid scatter)
    float *gw1, *gb1, *gw2, *gb2, *gw3, *gb3;  // fp32 accumulators (+=, atomics)
};

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
            const uint4 hv = *reinterpret_cast<const uint4*>(xtile + tc::tile_chunk_off(r, cc, kCs128));
            const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
            uint32_t w[4];
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j) {
                const float2 h = unpack_half2(hw[j]);
                const float g0 = h.x > 0.f ? __uint_as_float(acc[q * 8 + 2 * j]) : 0.f;
                const float g1 = h.y > 0.f ? __uint_as_float(acc[q * 8 + 2 * j + 1]) : 0.f;
                w[j] = pack_half2(g0, g1);
            }
            *reinterpret_cast<uint4*>(gtile + tc::tile_chunk_off(r, cc, kCs128)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

__global__ void __launch_bounds__(kTile, 2) field_backward_kernel(const BwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_mma, bar_h2, bar_h1, bar_e;
    __shared__ uint32_t tmem_base_s;
    const uint32_t r = threadIdx.x, warp = r >> 5, lane = r & 31;

    const uint32_t M = a.count_ptr ? min((uint32_t)max(*a.count_ptr, 0), a.M) : a.M;
    const uint32_t n_tiles = (M + kTile - 1) / kTile;

    if (warp == 0) tc::tmem_alloc(&tmem_base_s, kTmemColsBwd);
    if (r == 0) {
        tc::mbar_init(&bar_mma, 1); tc::mbar_init(&bar_h2, 1); tc::mbar_init(&bar_h1, 1); tc::mbar_init(&bar_e, 1);
        tc::fence_mbar_init();
    }
    load_weight_tile(a.w.w1, kHid, kHid, kIn, smem + BwdSmem::w1);
    load_weight_tile(a.w.w2, kHid, kHid, kHid, smem + BwdSmem::w2);
    load_weight_tile(a.w.w3, kOut, kOutPad, kHid, smem + BwdSmem::w3);
    {   // the two constant chunks: row r = (1.0h, 0, 0, 0, 0, 0, 0, 0)
        const uint4 one = make_uint4(0x00003C00u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(smem + BwdSmem::one1 + tc::tile_chunk_off(r, 0, kCs128)) = one;
        *reinterpret_cast<uint4*>(smem + BwdSmem::onee + tc::tile_chunk_off(r, 0, kCs128)) = one;
    }
    tc::fence_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t tmem_row = tc::tmem_addr(tmem, warp * 32, 0);

    auto load_h2 = [&](uint32_t t) { tc::mbar_expect_tx(&bar_h2, kHidTileBytes); tc::bulk_load(smem + BwdSmem::xh2, a.h2 + (size_t)t * (kTile * kHid), kHidTileBytes, &bar_h2); };
    auto load_h1 = [&](uint32_t t) { tc::mbar_expect_tx(&bar_h1, kHidTileBytes); tc::bulk_load(smem + BwdSmem::xh1, a.h1 + (size_t)t * (kTile * kHid), kHidTileBytes, &bar_h1); };
    auto load_e = [&](uint32_t t) { tc::mbar_expect_tx(&bar_e, kEncTileBytes); tc::bulk_load(smem + BwdSmem::xe, a.enc + (size_t)t * (kTile * kIn), kEncTileBytes, &bar_e); };
    if (r == 0 && blockIdx.x < n_tiles) { load_h2(blockIdx.x); load_h1(blockIdx.x); load_e(blockIdx.x); }

    const uint32_t sx2 = tc::smem_u32(smem + BwdSmem::xh2), sx1 = tc::smem_u32(smem + BwdSmem::xh1), sxe = tc::smem_u32(smem + BwdSmem::xe);
    const uint32_t sg = tc::smem_u32(smem + BwdSmem::g), sg3 = tc::smem_u32(smem + BwdSmem::g3);
    const uint32_t sw1 = tc::smem_u32(smem + BwdSmem::w1), sw2 = tc::smem_u32(smem + BwdSmem::w2), sw3 = tc::smem_u32(smem + BwdSmem::w3);
    // data gradients: A = upstream grads (K-major), B = weights [out, in] read MN-major (K = out)
    constexpr uint32_t id_dgrad64 = tc::instr_desc(128, 64, false, true);
    constexpr uint32_t id_dgrad32 = tc::instr_desc(128, 32, false, true);
    // weight gradients: both operands MN-major, K = samples
    constexpr uint32_t id_w3 = tc::instr_desc(64, 16, true, true);     // D[i, o]  = sum_m h2[m, i] dh3[m, o]
    constexpr uint32_t id_w2 = tc::instr_desc(64, kNW2, true, true);   // D[o, i'] = sum_m dh2[m, o] [h1 | 1][m, i']
    constexpr uint32_t id_w1 = tc::instr_desc(64, kNW1, true, true);   // D[o, i'] = sum_m dh1[m, o] [enc | 1][m, i']

    uint32_t ph_mma = 0, ph_ld = 0;
    float gb3_acc[4] = {0.f, 0.f, 0.f, 0.f};  // this thread's share of the output-layer bias gradient
    bool first = true;

    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ph_ld ^= 1) {
        const uint32_t m = tile * kTile + r;
        const bool active = m < M;
        const uint32_t next = tile + gridDim.x;
        const bool has_next = next < n_tiles;

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
#pragma unroll
        for (int j = 0; j < 4; ++j) gb3_acc[j] += dh[j];
        // (the previous tile's d_enc bulk store read the G tile, not G3; G3's last readers were awaited MMAs)
        *reinterpret_cast<uint4*>(smem + BwdSmem::g3 + tc::tile_chunk_off(r, 0, kCs128)) =
            make_uint4(pack_half2(dh[0], dh[1]), pack_half2(dh[2], dh[3]), 0u, 0u);
        *reinterpret_cast<uint4*>(smem + BwdSmem::g3 + tc::tile_chunk_off(r, 1, kCs128)) = make_uint4(0u, 0u, 0u, 0u);
        tc::mbar_wait(&bar_h2, ph_ld);   // h2 tile has landed (async-proxy write, visible after the wait)
        if (r == 0) tc::bulk_store_wait_read();   // the previous tile's d_enc store has finished reading G
        tc::fence_async_smem();
        tc::tc_fence_before_sync();
        __syncthreads();
        if (r == 0) {
            tc::tc_fence_after_sync();
            // dh2_pre[128 x 64] = dh3[128 x 16] x W3[16(out) x 64(in)]
            tc::umma_f16(tmem + kColD, tc::desc_k_major(sg3, kCs128, 0), tc::desc_mn_major(sw3, kCs16, 0), id_dgrad64, 0);
            // gW3^T[64(in) x 16(out)] += h2^T dh3
#pragma unroll
            for (uint32_t k = 0; k < kTile / 16; ++k)
                tc::umma_f16(tmem + kColW3, tc::desc_mn_major(sx2, kCs128, k), tc::desc_mn_major(sg3, kCs128, k), id_w3,
                             (!first || k > 0) ? 1u : 0u);
            tc::umma_commit(&bar_mma);
        }
        tc::mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1;
        tc::tc_fence_after_sync();
        relu_backward_epilogue(tmem_row, r, smem + BwdSmem::xh2, smem + BwdSmem::g);      // dh2 -> G
        tc::mbar_wait(&bar_h1, ph_ld);
        tc::fence_async_smem();
        tc::tc_fence_before_sync();
        __syncthreads();

        // ---- hidden layer 2 ------------------------------------------------------------------------------------
        if (r == 0) {
            tc::tc_fence_after_sync();
            if (has_next) load_h2(next);   // h2's readers (the weight-gradient MMAs above, the epilogue) are done
#pragma unroll
            for (uint32_t k = 0; k < kHid / 16; ++k)   // dh1_pre = dh2[128 x 64] x W2[64(out) x 64(in)]
                tc::umma_f16(tmem + kColD, tc::desc_k_major(sg, kCs128, k), tc::desc_mn_major(sw2, kCs64, k), id_dgrad64, k > 0);
#pragma unroll
            for (uint32_t k = 0; k < kTile / 16; ++k)  // [gW2 | gb2][64(out) x 72] += dh2^T [h1 | 1]
                tc::umma_f16(tmem + kColW2, tc::desc_mn_major(sg, kCs128, k), tc::desc_mn_major(sx1, kCs128, k), id_w2,
                             (!first || k > 0) ? 1u : 0u);
            tc::umma_commit(&bar_mma);
        }
        tc::mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1;
        tc::tc_fence_after_sync();
        relu_backward_epilogue(tmem_row, r, smem + BwdSmem::xh1, smem + BwdSmem::g);      // dh1 -> G
        tc::mbar_wait(&bar_e, ph_ld);
        tc::fence_async_smem();
        tc::tc_fence_before_sync();
        __syncthreads();

        // ---- hidden layer 1 ------------------------------------------------------------------------------------
        if (r == 0) {
            tc::tc_fence_after_sync();
            if (has_next) load_h1(next);
#pragma unroll
            for (uint32_t k = 0; k < kHid / 16; ++k)   // d_enc[128 x 32] = dh1[128 x 64] x W1[64(out) x 32(in)]
                tc::umma_f16(tmem + kColD, tc::desc_k_major(sg, kCs128, k), tc::desc_mn_major(sw1, kCs64, k), id_dgrad32, k > 0);
#pragma unroll
            for (uint32_t k = 0; k < kTile / 16; ++k)  // [gW1 | gb1][64(out) x 40] += dh1^T [enc | 1]
                tc::umma_f16(tmem + kColW1, tc::desc_mn_major(sg, kCs128, k), tc::desc_mn_major(sxe, kCs128, k), id_w1,
                             (!first || k > 0) ? 1u : 0u);
            tc::umma_commit(&bar_mma);
        }
        tc::mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1;
        tc::tc_fence_after_sync();
        {
            uint32_t acc[32];
            tc::tmem_ld_x32(tmem_row + kColD, acc);
            tc::tmem_ld_wait();
            uint32_t w[16];
#pragma unroll
            for (uint32_t j = 0; j < 16; ++j) w[j] = pack_half2(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1]));
            const bool full_tile = (tile + 1) * kTile <= M;
            if (full_tile) {
                // stage the tile's [128 x 32] d_enc rows row-major in the (now idle) G tile: one 8 KB bulk store
                uint4* dst = reinterpret_cast<uint4*>(smem + BwdSmem::g + r * (kIn * 2));
#pragma unroll
                for (uint32_t cc = 0; cc < 4; ++cc) dst[cc] = make_uint4(w[4 * cc], w[4 * cc + 1], w[4 * cc + 2], w[4 * cc + 3]);
            } else if (active) {
#pragma unroll
                for (uint32_t cc = 0; cc < 4; ++cc)
                    *reinterpret_cast<uint4*>(a.d_enc + (size_t)m * kIn + cc * 8) = make_uint4(w[4 * cc], w[4 * cc + 1], w[4 * cc + 2], w[4 * cc + 3]);
            }
            tc::fence_async_smem();
            tc::tc_fence_before_sync();
            __syncthreads();   // staged rows complete; every thread is done with TMEM D and (generic) reads of X tiles
            if (r == 0) {
                if (full_tile) tc::bulk_store(a.d_enc + (size_t)tile * (kTile * kIn), smem + BwdSmem::g, kEncTileBytes);
                if (has_next) load_e(next);
            }
        }
        first = false;
    }

    // ---- flush the weight gradients held in TMEM (M = 64 accumulators: rows 16w..16w+15 live in lanes 32w..32w+15) ----
    tc::tc_fence_after_sync();
    if (!first) {
        const uint32_t row = warp * 16 + lane;  // valid for lane < 16
        // 16-byte vector reductions (4x fewer atomic operations per CTA flush) when the accumulators are aligned
        const bool vec_ok = ((reinterpret_cast<uintptr_t>(a.gw1) | reinterpret_cast<uintptr_t>(a.gw2)) & 15) == 0;
        {   // gW3[o, i] = D[i, o]
            uint32_t acc[16];
            tc::tmem_ld_x16(tmem_row + kColW3, acc);
            tc::tmem_ld_wait();
            if (lane < 16)
                for (uint32_t o = 0; o < kOut; ++o) atomicAdd(a.gw3 + o * kHid + row, __uint_as_float(acc[o]));
        }
        for (uint32_t c0 = 0; c0 < kHid; c0 += 32) {  // gW2[o, i] = D[o, i]
            uint32_t acc[32];
            tc::tmem_ld_x32(tmem_row + kColW2 + c0, acc);
            tc::tmem_ld_wait();
            if (lane < 16) {
                float* dst = a.gw2 + row * kHid + c0;
                if (vec_ok) {
#pragma unroll
                    for (uint32_t i = 0; i < 32; i += 4)
                        red_add_f32x4(dst + i, __uint_as_float(acc[i]), __uint_as_float(acc[i + 1]), __uint_as_float(acc[i + 2]),
                                      __uint_as_float(acc[i + 3]));
                } else {
                    for (uint32_t i = 0; i < 32; ++i) atomicAdd(dst + i, __uint_as_float(acc[i]));
                }
            }
        }
        {   // gb2[o] = D[o, 64]
            uint32_t acc[8];
            tc::tmem_ld_x8(tmem_row + kColW2 + kHid, acc);
            tc::tmem_ld_wait();
            if (lane < 16) atomicAdd(a.gb2 + row, __uint_as_float(acc[0]));
        }
        {   // gW1[o, i] = D[o, i]
            uint32_t acc[32];
            tc::tmem_ld_x32(tmem_row + kColW1, acc);
            tc::tmem_ld_wait();
            if (lane < 16) {
                float* dst = a.gw1 + row * kIn;
                if (vec_ok) {
#pragma unroll
                    for (uint32_t i = 0; i < kIn; i += 4)
                        red_add_f32x4(dst + i, __uint_as_float(acc[i]), __uint_as_float(acc[i + 1]), __uint_as_float(acc[i + 2]),
                                      __uint_as_float(acc[i + 3]));
                } else {
                    for (uint32_t i = 0; i < kIn; ++i) atomicAdd(dst + i, __uint_as_float(acc[i]));
                }
            }
        }
        {   // gb1[o] = D[o, 32]
            uint32_t acc[8];
            tc::tmem_ld_x8(tmem_row + kColW1 + kIn, acc);
            tc::tmem_ld_wait();
            if (lane < 16) atomicAdd(a.gb1 + row, __uint_as_float(acc[0]));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float s = warp_sum(gb3_acc[j]);
            if (lane == 0) atomicAdd(a.gb3 + j, s);
        }
    }
    if (r == 0) tc::bulk_store_wait_all();
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, kTmemColsBwd);
}

// quads[g] for every row g of the table (see encode_level_quads); rows of hashed levels are copied plainly (never read)
__global__ void __launch_bounds__(256) quad_table_kernel(const uint32_t* __restrict__ table, const int* __restrict__ offsets,
                                                         uint32_t L, float S, uint32_t H, uint32_t gridtype, bool align_corners,
                                                         uint4* __restrict__ quads) {
    __shared__ grid::FastLevel<3> s_levels[grid::kMaxLevels];
    if (threadIdx.x < L) s_levels[threadIdx.x] = grid::make_fast_level<3>(offsets, threadIdx.x, S, H, gridtype, align_corners);
    __syncthreads();
    const uint32_t rows = s_levels[L - 1].offset + s_levels[L - 1].size;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < rows; g += gridDim.x * blockDim.x) {
        uint32_t l = 0;
        while (l + 1 < L && g >= s_levels[l + 1].offset) ++l;
        const grid::FastLevel<3>& lp = s_levels[l];
        const uint32_t r = g - lp.offset;
        const uint32_t* __restrict__ t = table + lp.offset;
        uint32_t n[3] = {r + 1u, r + lp.stride[1], r + lp.stride[1] + 1u};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            n[k] = wrap_row(lp, n[k]);
            if (n[k] >= lp.size) n[k] -= lp.size;   // (unwrapped levels: rows past the end are nobody's neighbours)
            if (lp.hashed) n[k] = r;
        }
        quads[g] = make_uint4(__ldg(t + r), __ldg(t + n[0]), __ldg(t + n[1]), __ldg(t + n[2]));
    }
}

}  // namespace field
}  // namespace ngp

using namespace ngp;

extern "C" int ngp_grid_quad_table(const void* table, const int* offsets, uint32_t L, uint32_t total_rows, float S, uint32_t H,
                                   uint32_t gridtype, int align_corners, void* quads, void* stream) {
    if (!table || !offsets || !quads) return NGP_ERR_BAD_ARG;
    if (L == 0 || L > grid::kMaxLevels || gridtype > 1) return NGP_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(quads) & 15) != 0 || (reinterpret_cast<uintptr_t>(table) & 3) != 0) return NGP_ERR_BAD_ARG;
    if (total_rows == 0) return NGP_OK;
    const int blocks = min(cdiv(total_rows, 256 * 4), num_sms() * 8);
    field::quad_table_kernel<<<blocks, 256, 0, as_stream(stream)>>>(static_cast<const uint32_t*>(table), offsets, L, S, H, gridtype,
                                                                   align_corners != 0, static_cast<uint4*>(quads));
    return launch_status();
}

static int check_field_dims(uint32_t L, uint32_t C, uint32_t D, uint32_t hidden, uint32_t out) {
    if (L != field::kLevels || C != 2 || D != 3 || hidden != field::kHid || out != field::kOut) return NGP_ERR_UNSUPPORTED;
    return NGP_OK;
}

static int field_forward_impl(const float* xyzs, uint32_t M, const int* count_ptr, const void* table, const void* quads,
                              const int* offsets, uint32_t L, uint32_t C, float S, uint32_t H, uint32_t gridtype, int align_corners,
                              float bound, const void* w1, const void* b1, const void* w2, const void* b2,
                              const void* w3, const void* b3, uint32_t hidden, uint32_t out_dim, float* sigma,
                              float* rgb, void* enc_save, void* h1_save, void* h2_save, void* stream) {
    if (!xyzs || !table || !offsets || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !sigma || !rgb) return NGP_ERR_BAD_ARG;
    int rc = check_field_dims(L, C, 3, hidden, out_dim);
    if (rc != NGP_OK) return rc;
    if (M == 0) return NGP_OK;
    field::FwdArgs a;
    a.xyzs = xyzs; a.M = M; a.count_ptr = count_ptr;
    if (quads && (reinterpret_cast<uintptr_t>(quads) & 15) != 0) return NGP_ERR_BAD_ARG;
    a.gd = {static_cast<const __half*>(table), static_cast<const uint4*>(quads), offsets, S, H, gridtype, align_corners, bound};
    a.w = {static_cast<const __half*>(w1), static_cast<const __half*>(b1), static_cast<const __half*>(w2),
           static_cast<const __half*>(b2), static_cast<const __half*>(w3), static_cast<const __half*>(b3)};
    a.sigma = sigma; a.rgb = rgb;
    a.enc = static_cast<__half*>(enc_save); a.h1 = static_cast<__half*>(h1_save); a.h2 = static_cast<__half*>(h2_save);
    static PerDeviceAttr attr[3];
    const int G = field::g_fwd_groups, slot = G == 1 ? 0 : (G == 2 ? 1 : 2);
    const void* fn = G == 1 ? reinterpret_cast<const void*>(field::field_forward_kernel<1>)
                   : G == 2 ? reinterpret_cast<const void*>(field::field_forward_kernel<2>)
                            : reinterpret_cast<const void*>(field::field_forward_kernel<4>);
    const int smem = (int)field::FwdSmem::total(G);
    rc = set_kernel_smem(&attr[slot], fn, smem, field::g_fwd_carveout);
    if (rc != NGP_OK) return rc;
    const int tiles = cdiv(M, field::kTile);
    const int per_sm = field::g_fwd_ctas_per_sm > 0 ? field::g_fwd_ctas_per_sm : (G == 1 ? 7 : 8 / G);
    const int grid = min(cdiv(tiles, G), num_sms() * per_sm);
    if (G == 1) field::field_forward_kernel<1><<<grid, field::kTile, smem, as_stream(stream)>>>(a);
    else if (G == 2) field::field_forward_kernel<2><<<grid, field::kTile * 2, smem, as_stream(stream)>>>(a);
    else field::field_forward_kernel<4><<<grid, field::kTile * 4, smem, as_stream(stream)>>>(a);
    return launch_status();
}

extern "C" int ngp_field_forward(const float* xyzs, uint32_t M, const int* count_ptr, const void* table, const int* offsets,
                                 uint32_t L, uint32_t C, float S, uint32_t H, uint32_t gridtype, int align_corners,
                                 float bound, const void* w1, const void* b1, const void* w2, const void* b2,
                                 const void* w3, const void* b3, uint32_t hidden, uint32_t out_dim, float* sigma,
                                 float* rgb, void* enc_save, void* h1_save, void* h2_save, void* stream) {
    return field_forward_impl(xyzs, M, count_ptr, table, nullptr, offsets, L, C, S, H, gridtype, align_corners, bound, w1, b1, w2, b2,
                              w3, b3, hidden, out_dim, sigma, rgb, enc_save, h1_save, h2_save, stream);
}

extern "C" int ngp_field_forward_quads(const float* xyzs, uint32_t M, const int* count_ptr, const void* table, const void* quads,
                                       const int* offsets, uint32_t L, uint32_t C, float S, uint32_t H, uint32_t gridtype,
                                       int align_corners, float bound, const void* w1, const void* b1, const void* w2,
                                       const void* b2, const void* w3, const void* b3, uint32_t hidden, uint32_t out_dim,
                                       float* sigma, float* rgb, void* enc_save, void* h1_save, void* h2_save, void* stream) {
    if (!quads) return NGP_ERR_BAD_ARG;
    return field_forward_impl(xyzs, M, count_ptr, table, quads, offsets, L, C, S, H, gridtype, align_corners, bound, w1, b1, w2, b2,
                              w3, b3, hidden, out_dim, sigma, rgb, enc_save, h1_save, h2_save, stream);
}

// tuning switches (profiles/kbench.py): 0 = forward CTAs per SM (grid size; 0 = as many as fit), 1 = forward smem carveout
// percent (-1 = driver default), 2 = 128-thread groups per forward CTA (1, 2, 4).  More resident CTAs mean more gathers in flight but a smaller L1 for the table.
extern "C" int ngp_field_set_option(int option, int value) {
    if (option == 0 && value >= 0 && value <= 8) { field::g_fwd_ctas_per_sm = value; return NGP_OK; }
    if (option == 2 && (value == 1 || value == 2 || value == 4)) { field::g_fwd_groups = value; return NGP_OK; }
    if (option == 1 && value >= -1 && value <= 100) { field::g_fwd_carveout = value; return NGP_OK; }
    return NGP_ERR_BAD_ARG;
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
    static PerDeviceAttr attr;
    rc = set_kernel_smem(&attr, reinterpret_cast<const void*>(field::field_backward_kernel), (int)field::BwdSmem::total);
    if (rc != NGP_OK) return rc;
    const int tiles = cdiv(M, field::kTile);
    const int grid = min(tiles, num_sms() * 2);
    field::field_backward_kernel<<<grid, field::kTile, field::BwdSmem::total, as_stream(stream)>>>(a);
    return launch_status();
}
