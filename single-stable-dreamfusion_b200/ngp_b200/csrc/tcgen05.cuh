// Minimal hand-written wrappers for the sm_100a tensor-core path: TMEM allocation, UMMA shared-memory /
// instruction descriptors, tcgen05.mma / commit / ld, mbarrier and the proxy fences they need.
// Only what the 64-wide field MLP uses: kind::f16 (fp16 operands, fp32 accumulate), cta_group::1,
// no-swizzle ("interleaved" 8x16-byte core matrix) operand tiles.
#pragma once
#include "common.cuh"

namespace ngp {
namespace tc {

NGP_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -----------------------------------------------------------------------------------
NGP_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
NGP_DEVINL void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
NGP_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}

// ---- fences ---------------------------------------------------------------------------------------
// generic-proxy smem writes -> visible to the async proxy (the tensor core reads operands through it)
NGP_DEVINL void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
NGP_DEVINL void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
NGP_DEVINL void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM -------------------------------------------------------------------------------------------
// Executed by ONE full warp.  ncols: power of two in [32, 512].  The base address lands in *smem_dst.
NGP_DEVINL void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
NGP_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- descriptors ------------------------------------------------------------------------------------
// Operand tiles use ONE physical layout everywhere ("core-matrix tile"): element (r, c) of a [R x C] fp16 tile
// lives at byte  (c/8)*chunk_stride + (r/8)*128 + (r%8)*16 + (c%8)*2  with chunk_stride = R*16, i.e. 8x8 core
// matrices of 128 contiguous bytes; the R/8 core matrices of one 8-column chunk are contiguous (one chunk = the
// 16-byte slices of all R rows = R*16 bytes), chunks follow each other.
//   * read as a K-major operand  (MN index = r, K index = c):  LBO = chunk_stride (next core matrix along K),
//                                                               SBO = 128 (next 8 rows);
//   * read as an MN-major operand (K index = r, MN index = c):  LBO = 128 (next 8 K-rows),
//                                                               SBO = chunk_stride (next 8 MN-elements).
// The second reading is what lets the SAME tile feed the forward / dgrad GEMM (K = features) and the
// weight-gradient GEMM (K = samples) without a transposed copy.  Chunk-major order also means a tile can grow by
// whole column chunks (the constant "ones" chunk that turns a weight-gradient GEMM into weight + bias gradient),
// that a thread-per-row write of one chunk is 512 contiguous bytes per warp (conflict-free), and that a whole tile
// is ONE contiguous block - saved to / restored from global memory with a single bulk (TMA) copy.
// Field layout (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
// base_offset [49,52), lbo_mode [52], layout_type [61,64) (0 = no swizzle).
NGP_DEVINL uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D fmt [4,6) (1 = f32), A fmt [7,10), B fmt [10,13)
// (0 = f16), A major [15], B major [16] (0 = K, 1 = MN), N>>3 [17,23), M>>4 [24,29).
constexpr uint32_t instr_desc(uint32_t M, uint32_t N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues on behalf of the CTA.
NGP_DEVINL void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// All previously issued MMAs of this thread arrive on `bar` when they have completed (implies
// tcgen05.fence::before_thread_sync).
NGP_DEVINL void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM -> registers: thread t of warp w reads lane 32*(w%4)+t, N consecutive 32-bit columns -----------------
NGP_DEVINL void tmem_ld_x8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
NGP_DEVINL void tmem_ld_x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
NGP_DEVINL void tmem_ld_x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, "
        "[%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
NGP_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// TMEM address = (lane << 16) | column, relative to the allocation base
NGP_DEVINL uint32_t tmem_addr(uint32_t base, uint32_t lane, uint32_t col) { return base + (lane << 16) + col; }

// ---- the core-matrix tile ------------------------------------------------------------------------------
// byte offset of the 16-byte chunk holding columns [8*cc, 8*cc+8) of row r; chunk_stride = rows of the tile * 16
NGP_DEVINL uint32_t tile_chunk_off(uint32_t r, uint32_t cc, uint32_t chunk_stride) {
    return cc * chunk_stride + (r >> 3) * 128u + (r & 7u) * 16u;
}
// descriptors of the tile at `addr` for the k-th K=16 step of an MMA
NGP_DEVINL uint64_t desc_k_major(uint32_t addr, uint32_t chunk_stride, uint32_t k) {   // K = columns: 2 chunks per step
    return smem_desc(addr + k * 2u * chunk_stride, chunk_stride, 128u);
}
NGP_DEVINL uint64_t desc_mn_major(uint32_t addr, uint32_t chunk_stride, uint32_t k) {  // K = rows: 2 row groups per step
    return smem_desc(addr + k * 256u, 128u, chunk_stride);
}

// ---- bulk (TMA, non-tensor) copies of whole tiles ----------------------------------------------------------
// global -> shared, completion counted in bytes on an mbarrier (the caller arms it with mbar_expect_tx first)
NGP_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
NGP_DEVINL void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
NGP_DEVINL void bulk_store(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// wait until the bulk stores of this thread have finished READING shared memory (the source may be rewritten)
NGP_DEVINL void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// wait until they are complete (writes performed)
NGP_DEVINL void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace tc
}  // namespace ngp
