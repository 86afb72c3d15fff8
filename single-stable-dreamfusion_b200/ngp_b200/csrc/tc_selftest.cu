// Hardware self-test for the hand-written tcgen05 path (tcgen05.cuh): one CTA, one small GEMM per mode,
// covering exactly the operand readings the field-MLP kernels rely on.  Called from tests/ on a GPU box.
//   mode 0: D[128 x N] = A[128 x K] * B[N x K]^T          A, B K-major             (forward GEMM)
//   mode 1: D[M   x N] = sum_s A[s, m] * B[s, n], s < 128  A, B MN-major, M in {64,128}, N % 8 == 0 (weight-gradient GEMM)
//   mode 2: D[128 x N] = A[128 x K] * B[K x N]             A K-major, B MN-major    (data-gradient GEMM)
#include "tcgen05.cuh"

namespace ngp {
namespace tcst {

// global row-major [R x C] fp16 -> core-matrix tile in shared memory
NGP_DEVINL void load_tile(const __half* g, uint32_t R, uint32_t C, uint8_t* smem, uint32_t chunk_stride) {
    const uint32_t chunks = C / 8;
    for (uint32_t i = threadIdx.x; i < R * chunks; i += blockDim.x) {
        const uint32_t r = i / chunks, cc = i % chunks;
        const uint4 v = *reinterpret_cast<const uint4*>(g + (size_t)r * C + cc * 8);
        *reinterpret_cast<uint4*>(smem + tc::tile_chunk_off(r, cc, chunk_stride)) = v;
    }
}

__global__ void __launch_bounds__(128) selftest_kernel(int mode, const __half* A, const __half* B, float* D, uint32_t M,
                                                       uint32_t N, uint32_t K) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // tile geometry (rows x cols of the row-major sources)
    uint32_t RA, CA, RB, CB;
    if (mode == 0) { RA = 128; CA = K; RB = N; CB = K; }
    else if (mode == 1) { RA = 128; CA = M; RB = 128; CB = N; }
    else { RA = 128; CA = K; RB = K; CB = N; }
    const uint32_t csA = RA * 16, csB = RB * 16;  // chunk strides
    uint8_t* sA = smem;
    uint8_t* sB = smem + ((CA / 8) * csA + 1023) / 1024 * 1024;

    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 128);
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    load_tile(A, RA, CA, sA, csA);
    load_tile(B, RB, CB, sB, csB);
    tc::fence_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    if (threadIdx.x == 0) {
        const uint32_t a0 = tc::smem_u32(sA), b0 = tc::smem_u32(sB);
        if (mode == 0) {
            const uint32_t idesc = tc::instr_desc(128, N, false, false);
            for (uint32_t k = 0; k < K / 16; ++k)
                tc::umma_f16(tmem, tc::desc_k_major(a0, csA, k), tc::desc_k_major(b0, csB, k), idesc, k > 0);
        } else if (mode == 1) {
            const uint32_t idesc = tc::instr_desc(M, N, true, true);
            for (uint32_t k = 0; k < 128 / 16; ++k)
                tc::umma_f16(tmem, tc::desc_mn_major(a0, csA, k), tc::desc_mn_major(b0, csB, k), idesc, k > 0);
        } else {
            const uint32_t idesc = tc::instr_desc(128, N, false, true);
            for (uint32_t k = 0; k < K / 16; ++k)
                tc::umma_f16(tmem, tc::desc_k_major(a0, csA, k), tc::desc_mn_major(b0, csB, k), idesc, k > 0);
        }
        tc::umma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::tc_fence_after_sync();

    // read back: M = 128 -> TMEM lane == row; M = 64 -> rows 16w..16w+15 sit in lanes 32w..32w+15
    const uint32_t rowsM = (mode == 1) ? M : 128;
    for (uint32_t c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        tc::tmem_ld_x8(tc::tmem_addr(tmem, warp * 32, c0), v);
        tc::tmem_ld_wait();
        int row = -1;
        if (rowsM == 128) row = (int)(warp * 32 + lane);
        else if (lane < 16) row = (int)(warp * 16 + lane);
        if (row >= 0) {
            for (int j = 0; j < 8; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

}  // namespace tcst
}  // namespace ngp

using namespace ngp;

extern "C" int ngp_tc_selftest(int mode, const void* A, const void* B, float* D, uint32_t M, uint32_t N, uint32_t K,
                               void* stream) {
    if (!A || !B || !D) return NGP_ERR_BAD_ARG;
    if (mode < 0 || mode > 2 || N % 8 || N > 128 || N < 8 || (mode != 1 && N % 16)) return NGP_ERR_BAD_ARG;
    if (mode == 1 && M != 64 && M != 128) return NGP_ERR_BAD_ARG;
    if (mode != 1 && (K % 16 || K > 128 || K < 16)) return NGP_ERR_BAD_ARG;
    const size_t smem = 64 * 1024;
    cudaFuncSetAttribute(tcst::selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tcst::selftest_kernel<<<1, 128, smem, as_stream(stream)>>>(mode, static_cast<const __half*>(A),
                                                               static_cast<const __half*>(B), D, M, N, K);
    return launch_status();
}
