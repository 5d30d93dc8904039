// capi.cu - error plumbing and device facts for the C ABI in include/radnet_b200.h.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace radnet {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return RADNET_E_CUDA;
}

}  // namespace radnet

extern "C" int radnet_version(void) { return RADNET_ABI_VERSION; }

extern "C" const char *radnet_last_error_string(void) { return radnet::g_err; }

extern "C" const char *radnet_error_name(int code) {
    switch (code) {
        case RADNET_OK: return "RADNET_OK";
        case RADNET_E_INVALID: return "RADNET_E_INVALID";
        case RADNET_E_CUDA: return "RADNET_E_CUDA";
        case RADNET_E_WORKSPACE: return "RADNET_E_WORKSPACE";
        case RADNET_E_UNSUPPORTED: return "RADNET_E_UNSUPPORTED";
        default: return "RADNET_E_UNKNOWN";
    }
}

extern "C" int radnet_device_info(int *h_out3) {
    RADNET_CHECK_ARG(h_out3, "device_info: null pointer");
    int dev = 0, sms = 0, smem = 0, major = 0, minor = 0;
    RADNET_CUDA(cudaGetDevice(&dev));
    RADNET_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    RADNET_CUDA(cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    RADNET_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    RADNET_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    h_out3[0] = sms;
    h_out3[1] = smem;
    h_out3[2] = major * 10 + minor;
    return RADNET_OK;
}
