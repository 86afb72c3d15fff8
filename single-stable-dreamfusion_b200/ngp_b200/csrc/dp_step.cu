// The optimisation step of the data-parallel `-O` train step as ONE cooperative kernel (ngp_adam_step_fused):
//
//   world == 1:  GradScaler inf/nan check -> grid barrier -> Adam + LambdaLR + GradScaler.update + fp16 shadow refresh
//                + zero-fill of the gradient bucket                       (what ngp_check_finite + ngp_adam_step do in two)
//   world  > 1:  the gradient all-reduce is fused in, over NVLink peer memory (no NCCL call on the step):
//                  B0  cross-GPU barrier: every rank's backward has finished writing its gradient bucket
//                  P1  reduce-scatter by P2P loads: rank r sums slice r of all `world` buckets (fixed rank order),
//                      keeps the sum in its own bucket and checks it for inf/nan
//                  B1  cross-GPU barrier that also ORs the found_inf flags, so every rank takes the same branch
//                  P2  Adam on slice r only (the moments are sharded: rank r owns slice r of exp_avg / exp_avg_sq),
//                      then the new fp32 parameters and their fp16 shadow are written to EVERY rank's replica by P2P
//                      stores (the all-gather half of the all-reduce moves parameters instead of gradients, so the
//                      optimizer arithmetic is done once per element, not `world` times); zero-fill of the local bucket
//                  B2  cross-GPU barrier: all replicas are complete before anybody's next forward reads them
//
// Reference semantics: `scaler.step(optimizer); scaler.update()` on Adam(betas (0.9, 0.99), eps 1e-15) with per-group lr
// and LambdaLR (main.py:128-131, network_grid.py:170-181, nerf/utils.py:708-713), gradients averaged over ranks as
// DistributedDataParallel (the reference's dormant wrap, nerf/utils.py:200-202) would.
//
// Cross-GPU barriers are flag exchanges in peer memory (st.release.sys / ld.acquire.sys), one slot per (barrier, block,
// source rank), carrying a monotonically increasing epoch kept on the device, so the kernel is CUDA-graph replayable.
// Every spin is bounded by a wall-clock timeout (default 20 s, ngp_dp_set_option(0, milliseconds)): a lost peer raises
// the sticky error flag state[5] instead of hanging the GPU, and the step becomes a NO-OP: the flag travels with the
// found_inf exchange of B1, a rank that saw it (locally or from a peer) neither runs Adam, nor writes to any replica,
// nor clears its bucket, and every later launch returns at once.  TrainStep polls the flag and raises.
#include <cooperative_groups.h>
#include <math.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace ngp {
namespace dp {

constexpr uint32_t kMaxWorld = NGP_DP_MAX_WORLD;
constexpr uint32_t kMaxBlocks = NGP_DP_MAX_BLOCKS;
uint64_t g_spin_timeout_ns = 20ull * 1000 * 1000 * 1000;  // 20 s; ngp_dp_set_option(0, ms)
int g_blocks = 0;                                          // 0 = one block per SM; ngp_dp_set_option(1, blocks)

struct Args {
    float* p;          // parameters (flat fp32), this rank's replica
    float* g;          // gradient bucket (flat fp32), scaled by state[0]
    float* m;          // exp_avg      (only this rank's slice is maintained when world > 1)
    float* v;          // exp_avg_sq
    __half* h;         // optional fp16 shadow of p
    uint64_t n;        // elements; a multiple of 4
    uint32_t n_seg;
    uint64_t seg_end[NGP_ADAM_MAX_SEGMENTS];
    float seg_lr[NGP_ADAM_MAX_SEGMENTS];
    float beta1, beta2, eps;
    float grad_div;
    float lr_decay_ln, lr_decay_steps;
    float growth, backoff;
    uint32_t growth_interval;
    float* state;          // [0] scale [1] growth tracker [2] steps [3] found_inf [4] skipped [5] error flag [6] armed
    int deferred;          // 1: a launch with state[6] == 0 only arms (sets state[6] = 1) and applies nothing
    uint32_t* sync;        // [0] blocks_done [1] epoch
    // data parallel
    uint32_t rank, world;
    float* peer_g[kMaxWorld];
    float* peer_p[kMaxWorld];
    __half* peer_h[kMaxWorld];
    uint32_t* peer_flags[kMaxWorld];  // each: [3][kMaxBlocks][kMaxWorld] uint32, zero-initialised
    // optional NVLink SHARP (NVLS) multicast views of the same three buffers: one multimem.ld_reduce returns the sum of an
    // element over all replicas (added inside the switch), one multimem.st writes an element to every replica
    float* mc_g;
    float* mc_p;
    __half* mc_h;
    uint64_t timeout_ns;
};

NGP_DEVINL void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
NGP_DEVINL uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
NGP_DEVINL float4 ld_relaxed_sys_f4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
NGP_DEVINL float4 multimem_ld_reduce_add_f4(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
NGP_DEVINL void multimem_st_f4(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
NGP_DEVINL void multimem_st_b64(void* mc, uint2 v) {  // 8 bytes moved as two f32 lanes (a store: the bit patterns pass through)
    asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(mc), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)) : "memory");
}
NGP_DEVINL uint64_t now_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

NGP_DEVINL uint32_t* flag_slot(uint32_t* pad, uint32_t barrier, uint32_t block, uint32_t src) {
    return pad + ((size_t)barrier * kMaxBlocks + block) * kMaxWorld + src;
}

// Block-level barrier across ranks: block b of every rank meets block b of every other rank.  `payload` (2 bits) is
// OR-ed over the ranks and returned to every thread of the block.  A timed-out wait sets the sticky error flag state[5]
// and reports payload bit 1 (kPayloadError).
constexpr uint32_t kPayloadInf = 1u, kPayloadError = 2u;
// coef (optional, shared float[3]): while the first `world` threads exchange flags, the block's last thread evaluates the
// step's Adam bias corrections and learning-rate factor into it (double-precision pow: microseconds on one thread, hidden
// behind the wait for the peers)
NGP_DEVINL uint32_t xrank_barrier(const Args& a, uint32_t barrier, uint32_t block, uint32_t epoch, uint32_t payload,
                                  float* coef = nullptr, float step0 = 0.f) {
    __shared__ uint32_t s_or;
    if (threadIdx.x == 0) s_or = 0u;
    __syncthreads();
    if (coef && threadIdx.x == blockDim.x - 1) {
        const float t = step0 + 1.f;
        coef[0] = (float)(1.0 - pow((double)a.beta1, (double)t));
        coef[1] = (float)sqrt(1.0 - pow((double)a.beta2, (double)t));
        coef[2] = a.lr_decay_ln != 0.f ? expf(a.lr_decay_ln * fminf(step0, a.lr_decay_steps)) : 1.f;
    }
    if (threadIdx.x < a.world) {
        const uint32_t q = threadIdx.x;
        uint32_t got = 0;
        __threadfence_system();
        st_release_sys(flag_slot(a.peer_flags[q], barrier, block, a.rank), epoch * 4 + (payload & 3u));
        const uint32_t* mine = flag_slot(a.peer_flags[a.rank], barrier, block, q);
        const uint64_t t0 = now_ns();
        for (;;) {
            const uint32_t f = ld_acquire_sys(mine);
            if ((int32_t)((f >> 2) - epoch) >= 0) { got = f & 3u; break; }
            if (now_ns() - t0 > a.timeout_ns) { *(volatile float*)(a.state + 5) = 1.f; got = kPayloadError; break; }
        }
        if (got) atomicOr(&s_or, got);
    }
    __syncthreads();
    const uint32_t r = s_or;
    __syncthreads();   // s_or may be re-initialised by the next barrier of this block
    return r;
}

__global__ void __launch_bounds__(512) adam_step_fused_kernel(const Args a) {
    cg::grid_group grid = cg::this_grid();
    // a cross-GPU wait timed out in an earlier launch: the data-parallel job is broken, apply nothing (sticky; read by
    // every block before any barrier of this launch could set it)
    if (a.world > 1 && *(volatile float*)(a.state + 5) != 0.f) return;
    if (a.deferred && *(volatile float*)(a.state + 6) == 0.f) {
        // deferred mode (the step is applied at the START of the next train step, overlapped with its ray marching):
        // nothing is pending yet - every block has read the flag before block 0 raises it
        grid.sync();
        if (blockIdx.x == 0 && threadIdx.x == 0) a.state[6] = 1.f;
        return;
    }
    const float scale = a.state[0];
    const float step0 = a.state[2];
    const uint32_t epoch = a.sync[1] + 1;
    const uint64_t n4 = a.n / 4;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (uint64_t)gridDim.x * blockDim.x;

    // this rank's slice of the flat buffers, in float4 units
    uint64_t lo = 0, hi = n4;
    if (a.world > 1) {
        const uint64_t per = (n4 + a.world - 1) / a.world;
        lo = per * a.rank < n4 ? per * a.rank : n4;
        hi = lo + per < n4 ? lo + per : n4;
        xrank_barrier(a, 0, blockIdx.x, epoch, 0);   // B0
    }

    // ---- P1: (reduce the slice across ranks,) look for inf / nan -----------------------------------------------------
    bool bad = false;
    for (uint64_t i = lo + tid; i < hi; i += nthreads) {
        float4 s;
        if (a.world > 1 && a.mc_g) {
            s = multimem_ld_reduce_add_f4(a.mc_g + i * 4);   // summed in the NVSwitch: 1/world of the P2P ingress
            *reinterpret_cast<float4*>(a.g + i * 4) = s;
        } else if (a.world > 1) {
            // all `world` loads are issued before the first add: one NVLink round trip per element, not `world` of them
            float4 v[kMaxWorld];
#pragma unroll
            for (uint32_t q = 0; q < kMaxWorld; ++q)
                if (q < a.world) v[q] = ld_relaxed_sys_f4(a.peer_g[q] + i * 4);
            s = v[0];
#pragma unroll
            for (uint32_t q = 1; q < kMaxWorld; ++q)
                if (q < a.world) { s.x += v[q].x; s.y += v[q].y; s.z += v[q].z; s.w += v[q].w; }   // fixed rank order
            *reinterpret_cast<float4*>(a.g + i * 4) = s;  // nobody else reads slice `rank` of this rank's bucket
        } else {
            s = *reinterpret_cast<const float4*>(a.g + i * 4);
        }
        // x - x is 0 for finite x and NaN for +-inf / NaN
        bad = bad || (((s.x - s.x) + (s.y - s.y) + (s.z - s.z) + (s.w - s.w)) != 0.f);
    }
    if (__syncthreads_or((int)bad) && threadIdx.x == 0) a.state[3] = 1.0f;
    __threadfence();
    grid.sync();
    if (a.world > 1) {   // B1: OR the flags of all ranks (block 0 talks to the peers, the grid barrier spreads the result)
        if (blockIdx.x == 0) {
            // bit 0: found_inf; bit 1: a B0 wait of this rank timed out (any block) - then NOBODY may apply the step
            const uint32_t mine = (*(volatile float*)(a.state + 3) != 0.f ? kPayloadInf : 0u) |
                                  (*(volatile float*)(a.state + 5) != 0.f ? kPayloadError : 0u);
            const uint32_t any = xrank_barrier(a, 1, 0, epoch, mine);
            if (threadIdx.x == 0) {
                if (any & kPayloadInf) a.state[3] = 1.0f;
                if (any & kPayloadError) a.state[5] = 1.0f;
            }
            __threadfence();
        }
        grid.sync();
        // broken exchange: leave parameters, moments, buckets and the scaler state exactly as they are (on every rank
        // that learned of it); all blocks take this branch together, nobody waits in B2 or the election below
        if (*(volatile float*)(a.state + 5) != 0.f) return;
    }
    const bool skip = *(volatile float*)(a.state + 3) != 0.f;

    // ---- P2: Adam on the slice; new parameters to every replica; zero the whole local bucket --------------------------
    const float t = step0 + 1.f;
    // same expressions as torch's fused Adam (bias corrections evaluated in double, then narrowed); once per block
    __shared__ float s_coef[3];
    if (threadIdx.x == 0) {
        s_coef[0] = (float)(1.0 - pow((double)a.beta1, (double)t));
        s_coef[1] = (float)sqrt(1.0 - pow((double)a.beta2, (double)t));
        s_coef[2] = a.lr_decay_ln != 0.f ? expf(a.lr_decay_ln * fminf(step0, a.lr_decay_steps)) : 1.f;
    }
    __syncthreads();
    const float bc1 = s_coef[0], bc2_sqrt = s_coef[1], lr_mult = s_coef[2];
    const float inv = 1.0f / (scale * a.grad_div);
    const float w1 = 1.f - a.beta1, w2 = 1.f - a.beta2;
    if (!skip) {
        for (uint64_t i = lo + tid; i < hi; i += nthreads) {
            const uint64_t e0 = i * 4;
            const float4 g4 = *reinterpret_cast<const float4*>(a.g + e0);
            const float4 p4 = *reinterpret_cast<const float4*>(a.p + e0);
            const float4 m4 = *reinterpret_cast<const float4*>(a.m + e0);
            const float4 v4 = *reinterpret_cast<const float4*>(a.v + e0);
            float g[4] = {g4.x, g4.y, g4.z, g4.w}, p[4] = {p4.x, p4.y, p4.z, p4.w};
            float m[4] = {m4.x, m4.y, m4.z, m4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j) {
                float lr = a.seg_lr[0];
#pragma unroll
                for (uint32_t s = 1; s < NGP_ADAM_MAX_SEGMENTS; ++s)
                    if (s < a.n_seg && e0 + j >= a.seg_end[s - 1]) lr = a.seg_lr[s];
                const float step_size = lr * lr_mult / bc1;
                const float gr = g[j] * inv;
                m[j] = fmaf(w1, gr - m[j], m[j]);                 // exp_avg.lerp_(grad, 1 - beta1)
                v[j] = a.beta2 * v[j] + w2 * gr * gr;
                const float denom = sqrtf(v[j]) / bc2_sqrt + a.eps;
                p[j] -= step_size * m[j] / denom;
            }
            *reinterpret_cast<float4*>(a.m + e0) = make_float4(m[0], m[1], m[2], m[3]);
            *reinterpret_cast<float4*>(a.v + e0) = make_float4(v[0], v[1], v[2], v[3]);
            const float4 pn = make_float4(p[0], p[1], p[2], p[3]);
            __half2 h0 = __floats2half2_rn(p[0], p[1]), h1 = __floats2half2_rn(p[2], p[3]);
            uint2 raw;
            raw.x = *reinterpret_cast<uint32_t*>(&h0); raw.y = *reinterpret_cast<uint32_t*>(&h1);
            if (a.world > 1 && a.mc_p) {
                multimem_st_f4(a.mc_p + e0, pn);             // one store, replicated by the switch to every rank
                if (a.h) multimem_st_b64(a.mc_h + e0, raw);
            } else if (a.world > 1) {
                for (uint32_t q = 0; q < a.world; ++q) {
                    const uint32_t dst = (a.rank + q) % a.world;   // start with the own replica, spread the link load
                    *reinterpret_cast<float4*>(a.peer_p[dst] + e0) = pn;
                    if (a.h) *reinterpret_cast<uint2*>(a.peer_h[dst] + e0) = raw;
                }
            } else {
                *reinterpret_cast<float4*>(a.p + e0) = pn;
                if (a.h) *reinterpret_cast<uint2*>(a.h + e0) = raw;
            }
            *reinterpret_cast<float4*>(a.g + e0) = make_float4(0.f, 0.f, 0.f, 0.f);  // by the thread that consumed it
        }
    } else {
        for (uint64_t i = lo + tid; i < hi; i += nthreads) *reinterpret_cast<float4*>(a.g + i * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // the rest of the bucket: every rank has finished reading it (B1 was a full barrier), clear it for the next backward.
    // (The own slice was cleared above by the threads that read it - a different thread mapping here would race them.)
    for (uint64_t i = tid; i < n4; i += nthreads)
        if (i < lo || i >= hi) *reinterpret_cast<float4*>(a.g + i * 4) = make_float4(0.f, 0.f, 0.f, 0.f);

    if (a.world > 1) xrank_barrier(a, 2, blockIdx.x, epoch, 0);   // B2

    // ---- last block: GradScaler.update() + step count + epoch ---------------------------------------------------------
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(&a.sync[0], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        float tracker = a.state[1];
        float new_scale = scale;
        if (skip) {
            new_scale = scale * a.backoff;
            tracker = 0.f;
            a.state[4] += 1.f;
        } else {
            a.state[2] = t;
            tracker += 1.f;
            if (tracker >= (float)a.growth_interval) {
                const float grown = scale * a.growth;
                if (isfinite(grown)) new_scale = grown;
                tracker = 0.f;
            }
        }
        a.state[0] = new_scale;
        a.state[1] = tracker;
        a.state[3] = 0.f;
        a.sync[0] = 0u;
        a.sync[1] = epoch;
    }
}


// -------------------------------------------------------------------------------------------------------------------
// world > 1, second generation (ngp_adam_step_dp): the same step WITHOUT grid-wide barriers, as a plain (non-cooperative)
// launch of independent blocks.  What made the cooperative kernel need two grid barriers and the B1 exchange was the
// GradScaler's inf / nan check on the REDUCED gradient.  A sum is non-finite iff an addend is (finite fp32 gradients cannot
// overflow a sum of <= 8), so every rank checks its OWN bucket first (ngp_check_finite, one small launch before this one)
// and the verdict rides on B0's flag exchange: after B0 every block of every rank knows the global found_inf.  Block b then
// owns sub-slice b of its rank's slice end to end:
//     B0(b)  flags of block b of all ranks (payload: local found_inf | error)
//     P      per element: `world` P2P loads (or one multimem.ld_reduce) -> sum -> Adam -> new parameter + fp16 shadow
//            stored to every replica (or one multimem.st each)
//     B2(b)  every rank's block b is done: its reads of MY bucket and its writes to MY replica have completed
//     Z      zero sub-slice b of ALL `world` slices of my bucket (exactly what the peers' blocks b have finished reading)
// No block ever waits for another block of its own grid, so the grid can be small (it runs beside the next step's ray
// marching in pipelined mode) and needs no co-residency guarantee.
// -------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) adam_dp_kernel(const Args a) {
    __shared__ bool s_last;
    if (*(volatile float*)(a.state + 5) != 0.f) return;                 // sticky: a peer was lost in an earlier launch
    const bool arm_only = a.deferred && *(volatile float*)(a.state + 6) == 0.f;   // deferred mode, nothing pending yet
    const float scale = a.state[0];
    const float step0 = a.state[2];
    const uint32_t epoch = a.sync[1] + 1;
    bool skip = false, broken = false;
    if (!arm_only) {
        const uint64_t n4 = a.n / 4;
        const uint64_t per = (n4 + a.world - 1) / a.world;
        const uint64_t lo = per * a.rank < n4 ? per * a.rank : n4;
        const uint64_t hi = lo + per < n4 ? lo + per : n4;
        const uint64_t chunk = (per + gridDim.x - 1) / gridDim.x;       // float4s of a slice that one block owns
        const uint64_t b_lo = lo + chunk * blockIdx.x < hi ? lo + chunk * blockIdx.x : hi;
        const uint64_t b_hi = b_lo + chunk < hi ? b_lo + chunk : hi;

        const uint32_t mine = (*(volatile float*)(a.state + 3) != 0.f ? kPayloadInf : 0u);
        __shared__ float s_coef[3];
        const uint32_t any = xrank_barrier(a, 0, blockIdx.x, epoch, mine, s_coef, step0);      // B0 (+ the step's coefficients)
        broken = (any & kPayloadError) != 0u;
        skip = (any & kPayloadInf) != 0u;
        if (!broken) {
            if (!skip) {
                const float bc1 = s_coef[0], bc2_sqrt = s_coef[1], lr_mult = s_coef[2];
                const float inv = 1.0f / (scale * a.grad_div);
                const float w1 = 1.f - a.beta1, w2 = 1.f - a.beta2;
                for (uint64_t i = b_lo + threadIdx.x; i < b_hi; i += blockDim.x) {
                    const uint64_t e0 = i * 4;
                    float4 s;
                    if (a.mc_g) {
                        s = multimem_ld_reduce_add_f4(a.mc_g + e0);
                    } else {
                        float4 v[kMaxWorld];
#pragma unroll
                        for (uint32_t q = 0; q < kMaxWorld; ++q)
                            if (q < a.world) v[q] = ld_relaxed_sys_f4(a.peer_g[q] + e0);
                        s = v[0];
#pragma unroll
                        for (uint32_t q = 1; q < kMaxWorld; ++q)
                            if (q < a.world) { s.x += v[q].x; s.y += v[q].y; s.z += v[q].z; s.w += v[q].w; }   // fixed rank order
                    }
                    const float4 p4 = *reinterpret_cast<const float4*>(a.p + e0);
                    const float4 m4 = *reinterpret_cast<const float4*>(a.m + e0);
                    const float4 v4 = *reinterpret_cast<const float4*>(a.v + e0);
                    float g[4] = {s.x, s.y, s.z, s.w}, p[4] = {p4.x, p4.y, p4.z, p4.w};
                    float m[4] = {m4.x, m4.y, m4.z, m4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j) {
                        float lr = a.seg_lr[0];
#pragma unroll
                        for (uint32_t sg = 1; sg < NGP_ADAM_MAX_SEGMENTS; ++sg)
                            if (sg < a.n_seg && e0 + j >= a.seg_end[sg - 1]) lr = a.seg_lr[sg];
                        const float step_size = lr * lr_mult / bc1;
                        const float gr = g[j] * inv;
                        m[j] = fmaf(w1, gr - m[j], m[j]);
                        v[j] = a.beta2 * v[j] + w2 * gr * gr;
                        const float denom = sqrtf(v[j]) / bc2_sqrt + a.eps;
                        p[j] -= step_size * m[j] / denom;
                    }
                    *reinterpret_cast<float4*>(a.m + e0) = make_float4(m[0], m[1], m[2], m[3]);
                    *reinterpret_cast<float4*>(a.v + e0) = make_float4(v[0], v[1], v[2], v[3]);
                    const float4 pn = make_float4(p[0], p[1], p[2], p[3]);
                    __half2 h0 = __floats2half2_rn(p[0], p[1]), h1 = __floats2half2_rn(p[2], p[3]);
                    uint2 raw;
                    raw.x = *reinterpret_cast<uint32_t*>(&h0); raw.y = *reinterpret_cast<uint32_t*>(&h1);
                    if (a.mc_p) {
                        multimem_st_f4(a.mc_p + e0, pn);
                        if (a.h) multimem_st_b64(a.mc_h + e0, raw);
                    } else {
                        for (uint32_t q = 0; q < a.world; ++q) {
                            const uint32_t dst = (a.rank + q) % a.world;
                            *reinterpret_cast<float4*>(a.peer_p[dst] + e0) = pn;
                            if (a.h) *reinterpret_cast<uint2*>(a.peer_h[dst] + e0) = raw;
                        }
                    }
                }
            }
            const uint32_t late = xrank_barrier(a, 2, blockIdx.x, epoch, 0);     // B2
            broken = (late & kPayloadError) != 0u;
            if (!broken) {
                // every rank's block b has finished reading sub-slice b of "its" slice of my bucket: clear those
                for (uint32_t q = 0; q < a.world; ++q) {
                    const uint64_t q_lo = per * q < n4 ? per * q : n4;
                    const uint64_t q_hi = q_lo + per < n4 ? q_lo + per : n4;
                    const uint64_t z_lo = q_lo + chunk * blockIdx.x < q_hi ? q_lo + chunk * blockIdx.x : q_hi;
                    const uint64_t z_hi = z_lo + chunk < q_hi ? z_lo + chunk : q_hi;
                    for (uint64_t i = z_lo + threadIdx.x; i < z_hi; i += blockDim.x)
                        *reinterpret_cast<float4*>(a.g + i * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
    }
    // ---- last block: GradScaler.update() + step count + epoch (or: arm the deferred mode) ---------------------------------
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(&a.sync[0], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        a.sync[0] = 0u;
        if (arm_only) { a.state[6] = 1.f; return; }
        if (*(volatile float*)(a.state + 5) != 0.f) return;      // broken exchange: scaler state and epoch stay as they are
        float tracker = a.state[1];
        float new_scale = scale;
        if (skip) {
            new_scale = scale * a.backoff;
            tracker = 0.f;
            a.state[4] += 1.f;
        } else {
            a.state[2] = step0 + 1.f;
            tracker += 1.f;
            if (tracker >= (float)a.growth_interval) {
                const float grown = scale * a.growth;
                if (isfinite(grown)) new_scale = grown;
                tracker = 0.f;
            }
        }
        a.state[0] = new_scale;
        a.state[1] = tracker;
        a.state[3] = 0.f;
        a.sync[1] = epoch;
    }
}

}  // namespace dp
}  // namespace ngp

using namespace ngp;

extern "C" uint64_t ngp_dp_flags_bytes(void) { return (uint64_t)3 * dp::kMaxBlocks * dp::kMaxWorld * sizeof(uint32_t); }

extern "C" int ngp_enable_peer_access(int peer_device) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (peer_device == dev) return NGP_OK;
    int can = 0;
    e = cudaDeviceCanAccessPeer(&can, dev, peer_device);
    if (e != cudaSuccess) return (int)e;
    if (!can) return NGP_ERR_UNSUPPORTED;
    e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return NGP_OK; }
    return e == cudaSuccess ? NGP_OK : (int)e;
}

// 0: spin timeout of the cross-GPU waits in milliseconds (>= 1); 1: blocks of the cooperative launch (0 = one per SM)
extern "C" int ngp_dp_set_option(int option, int value) {
    if (option == 0 && value >= 1) { dp::g_spin_timeout_ns = (uint64_t)value * 1000000ull; return NGP_OK; }
    if (option == 1 && value >= 0 && value <= (int)dp::kMaxBlocks) { dp::g_blocks = value; return NGP_OK; }
    return NGP_ERR_BAD_ARG;
}

static int fill_dp_args(dp::Args& a, float* params, float* grads, float* exp_avg, float* exp_avg_sq, void* half_shadow, uint64_t n,
                        uint32_t n_segments, const uint64_t* seg_end, const float* seg_lr, float beta1, float beta2, float eps,
                        float grad_div, float lr_decay_ln, float lr_decay_steps, float growth_factor, float backoff_factor,
                        uint32_t growth_interval, int deferred, float* state, uint32_t* sync, uint32_t rank, uint32_t world,
                        const uint64_t* peer_grads, const uint64_t* peer_params, const uint64_t* peer_half,
                        const uint64_t* peer_flags, const uint64_t* multicast);

// The data-parallel step without grid-wide barriers (see adam_dp_kernel): world > 1 only, and the caller must have run
// ngp_check_finite(grads, n, state + 3) on the same stream first.  Same arguments as ngp_adam_step_fused.
extern "C" int ngp_adam_step_dp(float* params, float* grads, float* exp_avg, float* exp_avg_sq, void* half_shadow, uint64_t n,
                                uint32_t n_segments, const uint64_t* seg_end, const float* seg_lr, float beta1, float beta2,
                                float eps, float grad_div, float lr_decay_ln, float lr_decay_steps, float growth_factor,
                                float backoff_factor, uint32_t growth_interval, int deferred, float* state, uint32_t* sync,
                                uint32_t rank, uint32_t world, const uint64_t* peer_grads, const uint64_t* peer_params,
                                const uint64_t* peer_half, const uint64_t* peer_flags, const uint64_t* multicast, void* stream) {
    if (world < 2) return NGP_ERR_BAD_ARG;
    dp::Args a;
    const int rc = fill_dp_args(a, params, grads, exp_avg, exp_avg_sq, half_shadow, n, n_segments, seg_end, seg_lr, beta1, beta2, eps,
                                grad_div, lr_decay_ln, lr_decay_steps, growth_factor, backoff_factor, growth_interval, deferred, state,
                                sync, rank, world, peer_grads, peer_params, peer_half, peer_flags, multicast);
    if (rc != NGP_OK) return rc;
    // a rank's slice is n / (4 world) float4s; ~4 per thread keeps `world` x 4 peer loads in flight per thread
    const uint64_t per = (n / 4 + world - 1) / world;
    int blocks = dp::g_blocks > 0 ? dp::g_blocks : (int)((per + 512 * 4 - 1) / (512 * 4));
    if (blocks < 8) blocks = 8;
    if (blocks > num_sms()) blocks = num_sms();
    if (blocks > (int)dp::kMaxBlocks) blocks = (int)dp::kMaxBlocks;
    dp::adam_dp_kernel<<<blocks, 512, 0, as_stream(stream)>>>(a);
    return launch_status();
}

extern "C" int ngp_adam_step_fused(float* params, float* grads, float* exp_avg, float* exp_avg_sq, void* half_shadow,
                                   uint64_t n, uint32_t n_segments, const uint64_t* seg_end, const float* seg_lr, float beta1,
                                   float beta2, float eps, float grad_div, float lr_decay_ln, float lr_decay_steps,
                                   float growth_factor, float backoff_factor, uint32_t growth_interval, int deferred,
                                   float* state, uint32_t* sync, uint32_t rank, uint32_t world, const uint64_t* peer_grads,
                                   const uint64_t* peer_params, const uint64_t* peer_half, const uint64_t* peer_flags,
                                   const uint64_t* multicast, void* stream) {
    dp::Args a;
    const int frc = fill_dp_args(a, params, grads, exp_avg, exp_avg_sq, half_shadow, n, n_segments, seg_end, seg_lr, beta1, beta2, eps,
                                 grad_div, lr_decay_ln, lr_decay_steps, growth_factor, backoff_factor, growth_interval, deferred, state,
                                 sync, rank, world, peer_grads, peer_params, peer_half, peer_flags, multicast);
    if (frc != NGP_OK) return frc;
    int blocks = dp::g_blocks > 0 ? dp::g_blocks : num_sms();
    if (blocks > num_sms()) blocks = num_sms();
    if (blocks > (int)dp::kMaxBlocks) blocks = (int)dp::kMaxBlocks;
    void* kargs[] = {&a};
    cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(dp::adam_step_fused_kernel), dim3(blocks), dim3(512),
                                                kargs, 0, as_stream(stream));
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    return launch_status();
}

static int fill_dp_args(dp::Args& a, float* params, float* grads, float* exp_avg, float* exp_avg_sq, void* half_shadow, uint64_t n,
                        uint32_t n_segments, const uint64_t* seg_end, const float* seg_lr, float beta1, float beta2, float eps,
                        float grad_div, float lr_decay_ln, float lr_decay_steps, float growth_factor, float backoff_factor,
                        uint32_t growth_interval, int deferred, float* state, uint32_t* sync, uint32_t rank, uint32_t world,
                        const uint64_t* peer_grads, const uint64_t* peer_params, const uint64_t* peer_half,
                        const uint64_t* peer_flags, const uint64_t* multicast) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || !state || !sync || !seg_end || !seg_lr) return NGP_ERR_BAD_ARG;
    if (n_segments == 0 || n_segments > NGP_ADAM_MAX_SEGMENTS || seg_end[n_segments - 1] != n) return NGP_ERR_BAD_ARG;
    if (n == 0 || (n & 3) != 0) return NGP_ERR_BAD_ARG;
    const uintptr_t al = reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) |
                         reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq);
    if ((al & 15) != 0 || (reinterpret_cast<uintptr_t>(half_shadow) & 7) != 0) return NGP_ERR_BAD_ARG;
    if (world == 0 || world > dp::kMaxWorld || rank >= world) return NGP_ERR_BAD_ARG;
    if (world > 1 && (!peer_grads || !peer_params || !peer_flags || (half_shadow && !peer_half))) return NGP_ERR_BAD_ARG;
    a.p = params; a.g = grads; a.m = exp_avg; a.v = exp_avg_sq; a.h = static_cast<__half*>(half_shadow); a.n = n;
    a.n_seg = n_segments;
    for (uint32_t s = 0; s < NGP_ADAM_MAX_SEGMENTS; ++s) {
        a.seg_end[s] = s < n_segments ? seg_end[s] : n;
        a.seg_lr[s] = s < n_segments ? seg_lr[s] : 0.f;
    }
    a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.grad_div = grad_div > 0.f ? grad_div : 1.f;
    a.lr_decay_ln = lr_decay_ln; a.lr_decay_steps = lr_decay_steps;
    a.growth = growth_factor; a.backoff = backoff_factor; a.growth_interval = growth_interval;
    a.state = state; a.sync = sync; a.rank = rank; a.world = world; a.deferred = deferred;
    for (uint32_t q = 0; q < dp::kMaxWorld; ++q) {
        const bool on = world > 1 && q < world;
        a.peer_g[q] = on ? reinterpret_cast<float*>(peer_grads[q]) : nullptr;
        a.peer_p[q] = on ? reinterpret_cast<float*>(peer_params[q]) : nullptr;
        a.peer_h[q] = on && peer_half ? reinterpret_cast<__half*>(peer_half[q]) : nullptr;
        a.peer_flags[q] = on ? reinterpret_cast<uint32_t*>(peer_flags[q]) : nullptr;
        if (on && (!a.peer_g[q] || !a.peer_p[q] || !a.peer_flags[q])) return NGP_ERR_BAD_ARG;
    }
    const bool mc = world > 1 && multicast && multicast[0] && multicast[1] && (!half_shadow || multicast[2]);
    a.mc_g = mc ? reinterpret_cast<float*>(multicast[0]) : nullptr;
    a.mc_p = mc ? reinterpret_cast<float*>(multicast[1]) : nullptr;
    a.mc_h = mc && half_shadow ? reinterpret_cast<__half*>(multicast[2]) : nullptr;
    a.timeout_ns = dp::g_spin_timeout_ns;
    return NGP_OK;
}
