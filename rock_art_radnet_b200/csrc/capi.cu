// capi.cu - error plumbing and device facts for the C ABI in include/radnet_b200.h.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

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

// ---- options ---------------------------------------------------------------------------------------
static const char *const kOptionNames[kOptCount] = {
    "nms_cluster", "nms_cluster_size", "nms_cluster_ranks", "nms_sel_target",
    "nms_lookahead", "roipool_force_direct", "targets_hit_cap", "roipool_form", "sampler_force_exact",
    "targets_compute_ctas", "targets_two_launches", "roipool_bands", "roipool_lanes", "roipool_cluster", "roipool_sync_every", "roipool_ctas", "targets_fill_bulk"};
static const long long kOptionDefaults[kOptCount] = {-1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, -1, 0, 0, 0};

struct OptionTable {
    std::atomic<long long> v[kOptCount];
    OptionTable() {
        for (int i = 0; i < kOptCount; ++i) {
            long long val = kOptionDefaults[i];
            char env[64] = "RADNET_";
            size_t n = strlen(env);
            for (const char *c = kOptionNames[i]; *c && n + 1 < sizeof(env); ++c)
                env[n++] = (char)((*c >= 'a' && *c <= 'z') ? *c - 32 : *c);
            env[n] = 0;
            if (const char *e = getenv(env)) val = atoll(e);
            v[i].store(val, std::memory_order_relaxed);
        }
    }
};
static OptionTable g_options;       // built when the library is loaded

long long get_option(int opt) { return g_options.v[opt].load(std::memory_order_relaxed); }

static int option_index(const char *name) {
    if (!name) return -1;
    for (int i = 0; i < kOptCount; ++i)
        if (strcmp(name, kOptionNames[i]) == 0) return i;
    return -1;
}

// ---- per-device facts ------------------------------------------------------------------------------
static std::atomic<int> g_smem_optin[64];
static std::atomic<int> g_sm_count[64];

static int cached_attr(std::atomic<int> *table, cudaDeviceAttr attr, int dev) {
    if (dev < 0 || dev >= 64) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, attr, dev) != cudaSuccess) return -1;
        return v;
    }
    int v = table[dev].load(std::memory_order_relaxed);
    if (v > 0) return v;
    cudaError_t e = cudaDeviceGetAttribute(&v, attr, dev);
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaDeviceGetAttribute");
        return -1;
    }
    table[dev].store(v, std::memory_order_relaxed);
    return v;
}

int device_smem_optin(int dev) { return cached_attr(g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev); }
int device_sm_count(int dev) { return cached_attr(g_sm_count, cudaDevAttrMultiProcessorCount, dev); }

static std::mutex g_attr_mutex;
struct AttrEntry {
    const void *func;
    int dev;
    size_t smem;
    bool nonportable;
};
static AttrEntry g_attr[256];
static int g_attr_n = 0;

static AttrEntry *attr_entry(const void *func, int dev) {       // caller holds g_attr_mutex
    for (int i = 0; i < g_attr_n; ++i)
        if (g_attr[i].func == func && g_attr[i].dev == dev) return &g_attr[i];
    if (g_attr_n >= 256) return nullptr;                        // table full: fall back to setting every time
    g_attr[g_attr_n] = AttrEntry{func, dev, 0, false};
    return &g_attr[g_attr_n++];
}

int ensure_dynamic_smem(const void *func, int dev, size_t bytes) {
    if (bytes <= 48 * 1024) return RADNET_OK;
    std::lock_guard<std::mutex> lock(g_attr_mutex);
    AttrEntry *e = attr_entry(func, dev);
    if (e && e->smem >= bytes) return RADNET_OK;
    cudaError_t err = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (err != cudaSuccess) return cuda_fail(err, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    if (e) e->smem = bytes;
    return RADNET_OK;
}

int ensure_nonportable_clusters(const void *func, int dev) {
    std::lock_guard<std::mutex> lock(g_attr_mutex);
    AttrEntry *e = attr_entry(func, dev);
    if (e && e->nonportable) return RADNET_OK;
    cudaError_t err = cudaFuncSetAttribute(func, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (err != cudaSuccess) return cuda_fail(err, "cudaFuncSetAttribute(NonPortableClusterSizeAllowed)");
    if (e) e->nonportable = true;
    return RADNET_OK;
}

}  // namespace radnet

extern "C" int radnet_set_option(const char *name, long long value) {
    const int i = radnet::option_index(name);
    RADNET_CHECK_ARG(i >= 0, "set_option: unknown option '%s'", name ? name : "(null)");
    radnet::g_options.v[i].store(value, std::memory_order_relaxed);
    return RADNET_OK;
}

extern "C" int radnet_get_option(const char *name, long long *h_value) {
    const int i = radnet::option_index(name);
    RADNET_CHECK_ARG(i >= 0 && h_value, "get_option: unknown option '%s'", name ? name : "(null)");
    *h_value = radnet::get_option(i);
    return RADNET_OK;
}

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
