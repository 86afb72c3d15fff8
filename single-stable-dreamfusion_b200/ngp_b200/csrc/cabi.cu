// Library-level entry points of the C ABI (version, error strings, device info).
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace ngp {
int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return kNumSMsB200;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMsB200;
        cached[dev] = n;
    }
    return cached[dev];
}

int set_kernel_smem(PerDeviceAttr* slot, const void* func, int smem_bytes, int carveout_percent) {
    static std::mutex mu;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    std::lock_guard<std::mutex> lock(mu);
    const bool tracked = dev >= 0 && dev < 64;
    if (tracked && slot->smem[dev] == smem_bytes && slot->carveout[dev] == carveout_percent + 2) return NGP_OK;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    if (carveout_percent >= 0) {
        e = cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_percent);
        if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    }
    if (tracked) { slot->smem[dev] = smem_bytes; slot->carveout[dev] = carveout_percent + 2; }
    return NGP_OK;
}
}  // namespace ngp

extern "C" int ngp_version(void) { return 100; }

extern "C" const char* ngp_error_string(int code) {
    switch (code) {
        case NGP_OK: return "ok";
        case NGP_ERR_BAD_ARG: return "bad argument (null pointer or inconsistent sizes)";
        case NGP_ERR_UNSUPPORTED: return "unsupported D / C / dtype combination";
        case NGP_ERR_WORKSPACE: return "workspace missing or too small";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown error";
}

extern "C" int ngp_device_info(char* name, int name_len, int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return (int)e;
    if (name && name_len > 0) {
        strncpy(name, prop.name, (size_t)name_len - 1);
        name[name_len - 1] = 0;
    }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return NGP_OK;
}

extern "C" int ngp_graph_launch(void* graph_exec, void* stream) {
    if (!graph_exec) return NGP_ERR_BAD_ARG;
    const cudaError_t e = cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(graph_exec), ngp::as_stream(stream));
    return e == cudaSuccess ? NGP_OK : (int)e;
}
