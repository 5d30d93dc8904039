// common.cuh - shared helpers for the sm_100a kernels of libradnet_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/radnet_b200.h"

namespace radnet {

constexpr int kMaxAnchors = 64;
constexpr int kSmCountB200 = 148;

// thread-local error text, set by every failing entry point
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define RADNET_CHECK_ARG(cond, ...)                  \
    do {                                             \
        if (!(cond)) {                               \
            ::radnet::set_error(__VA_ARGS__);        \
            return RADNET_E_INVALID;                 \
        }                                            \
    } while (0)

#define RADNET_CUDA(call)                                                    \
    do {                                                                     \
        cudaError_t e__ = (call);                                            \
        if (e__ != cudaSuccess) return ::radnet::cuda_fail(e__, #call);      \
    } while (0)

// launch-error check that does not synchronise
static inline int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, what);
    return RADNET_OK;
}

struct AnchorTable {
    double wh[kMaxAnchors][2];
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- process-wide tuning options (radnet_set_option / radnet_get_option) ---------------------------
// Atomics, seeded once from the environment (RADNET_<NAME>) when the library is loaded; launches only
// read them.  Any value gives the same results - they select between equivalent code paths.
enum Option {
    kOptNmsCluster = 0,      // nms_cluster: -1 auto, 0 never, 1 when it fits
    kOptNmsClusterSize,      // nms_cluster_size: 0 auto, else 2/4/8/16
    kOptNmsClusterRanks,     // nms_cluster_ranks: 0 auto
    kOptNmsSelTarget,        // nms_sel_target: 0 auto
    kOptNmsLookahead,        // nms_lookahead: 0 auto
    kOptRoipoolForceDirect,  // roipool_force_direct: 0/1
    kOptTargetsHitCap,       // targets_hit_cap: 0 auto
    kOptRoipoolForm,         // roipool_form: 0 auto, 1 whole-map slices, 2 row bands
    kOptSamplerForceExact,   // sampler_force_exact: 0/1 (always the serial cumsum chain)
    kOptTargetsComputeCtas,  // targets_compute_ctas: 0 auto (43 % of the SMs)
    kOptTargetsTwoLaunches,  // targets_two_launches: 0/1 (fill and panels as two launches)
    kOptRoipoolBands,        // roipool_bands: 0 auto, else the number of row bands of the band form
    kOptRoipoolLanes,        // roipool_lanes: 0 auto (8), else 8 / 16 / 32 float4 lanes per pixel in the band form
    kOptRoipoolCluster,      // roipool_cluster: -1 tuned at the first call, 0/1 none, 2 / 4 / 8 CTAs (neighbouring slices) per cluster
    kOptRoipoolSyncEvery,    // roipool_sync_every: with roipool_cluster >= 0, CTA / cluster barrier every this many column rounds
    kOptRoipoolCtas,         // roipool_ctas: whole-map form with this many persistent CTAs (0 = one CTA per work item)
    kOptTargetsFillBulk,     // targets_fill_bulk: 0 plain stores, 1 regression zeros by TMA bulk copies from a zeroed buffer, else buffer bytes
    kOptCount
};
long long get_option(int opt);

// max opt-in shared memory per block of device `dev`, cached per device; -1 (error set) on failure
int device_smem_optin(int dev);
// SM count of device `dev`, cached per device
int device_sm_count(int dev);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) only when a launch of `func` on device `dev` needs more
// than was already granted (a process-wide table keyed by kernel and device, mutex-protected)
int ensure_dynamic_smem(const void *func, int dev, size_t bytes);
// cudaFuncSetAttribute(NonPortableClusterSizeAllowed) once per kernel and device
int ensure_nonportable_clusters(const void *func, int dev);

// ------------------------------------------------------------------ device helpers
#ifdef __CUDACC__

// Order-preserving uint32 image of a float32 score.  0 is reserved for "deleted".
// -0.0 is folded onto +0.0 (NumPy compares them equal) and every NaN sorts last
// (= largest), like np.argsort.
__device__ __forceinline__ uint32_t score_to_key(float s) {
    uint32_t b = __float_as_uint(s);
    if (s != s) return 0xFFFFFFFFu;
    if (b == 0x80000000u) b = 0u;
    uint32_t k = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return k == 0u ? 1u : k;
}
__device__ __forceinline__ float key_to_score(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    return __uint_as_float(b);
}
__device__ __forceinline__ uint64_t score_to_key64(double s) {
    uint64_t b = (uint64_t)__double_as_longlong(s);
    if (s != s) return ~0ull;
    if (b == 0x8000000000000000ull) b = 0ull;
    uint64_t k = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    return k == 0ull ? 1ull : k;
}

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk, SASS UBLKCP) -----------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// wait bounded by ELAPSED TIME (2 s of %globaltimer), not by a number of attempts: under time-slicing, MPS,
// compute-sanitizer or a debugger a healthy partner can be arbitrarily slow per attempt, and a spurious give-up
// would turn a correct run into an error.  Each try_wait suspends the warp in hardware up to its own time limit;
// the clock is only read every 1024 failed attempts.  false = the partner never arrived (results are flagged
// invalid by the caller, which still arrives on its own barrier so that its successors do not hang).
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
    long long t_start = 0;
    for (unsigned tries = 0;; ++tries) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return true;
        if ((tries & 1023u) == 1023u) {
            long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t_start == 0) t_start = now;
            else if (now - t_start > 2000000000LL) return false;
        }
    }
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void tma_bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                             uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---- thread-block cluster helpers (distributed shared memory) -------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
// address of the same shared-memory location in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t cluster_map_shared(const void *p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// split cluster barrier: every thread of every CTA arrives once and waits once per phase
__device__ __forceinline__ void cluster_arrive() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier of another CTA of the cluster and announce `bytes` of bulk-copy traffic
__device__ __forceinline__ void mbar_remote_arrive_expect_tx(uint32_t remote_bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(remote_bar),
                 "r"(bytes)
                 : "memory");
}
// shared memory of this CTA -> shared memory of another CTA of the cluster, completion on ITS mbarrier
__device__ __forceinline__ void bulk_s2s_cluster(uint32_t remote_dst, const void *local_src, uint32_t bytes,
                                                 uint32_t remote_bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(remote_dst),
        "r"(smem_u32(local_src)), "r"(bytes), "r"(remote_bar)
        : "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (bulk copies)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// pacing barrier of a cluster (no data is exchanged): may be executed by partly diverged warps
__device__ __forceinline__ void cluster_sync_relaxed() {
    asm volatile("barrier.cluster.arrive.relaxed;\nbarrier.cluster.wait;" ::: "memory");
}
// all threads of all CTAs of the cluster; release/acquire makes the DSMEM stores visible
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// volatile 64-bit shared-memory accesses by 32-bit shared address (no generic-address arithmetic on the
// polling path)
__device__ __forceinline__ unsigned long long lds_volatile_u64(uint32_t addr) {
    unsigned long long v;
    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_volatile_u64(uint32_t addr, unsigned long long v) {
    asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}

// streaming 16-byte store: written once, never re-read by this kernel
__device__ __forceinline__ void st_stream_f4(float4 *p, const float4 &v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}

#endif  // __CUDACC__

}  // namespace radnet
