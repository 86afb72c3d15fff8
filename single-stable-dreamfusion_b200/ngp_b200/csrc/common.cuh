// Shared helpers for the sm_100a NeRF hot-path kernels.  No torch headers anywhere in csrc/.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../../include/ngp_b200.h"

#define NGP_DEVINL __device__ __forceinline__

namespace ngp {

constexpr int kNumSMsB200 = 148;

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Every extern "C" entry point funnels its launch status through here.
inline int launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? NGP_OK : (int)e;
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

int num_sms();  // cached cudaDevAttrMultiProcessorCount of the current device (cabi.cu)

// Kernel attributes (dynamic shared memory limit, carve-out) are per device: a per-kernel cache of "what was last set
// on device d" (cabi.cu), thread-safe, errors returned.  `slot` = one static PerDeviceAttr per kernel.
struct PerDeviceAttr {
    int smem[64];      // 0 = never set on that device
    int carveout[64];  // -2 = never set
};
int set_kernel_smem(PerDeviceAttr* slot, const void* func, int smem_bytes, int carveout_percent = -1);

// ---- warp primitives -------------------------------------------------------------------------
NGP_DEVINL float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
NGP_DEVINL int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// inclusive prefix sum across the warp
NGP_DEVINL int warp_incl_scan_i(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// ---- vector reductions to global memory (sm_90+: red.global.add.v2.f32) -------------------------
NGP_DEVINL void red_add_f32x2(float* addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
NGP_DEVINL void red_add_f32x4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}
NGP_DEVINL void red_add_f32(float* addr, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}

// streaming (evict-first) stores / loads for tensors that are written once and read once
NGP_DEVINL void st_cs_u32(void* p, uint32_t v) { __stcs(reinterpret_cast<unsigned int*>(p), v); }

}  // namespace ngp
