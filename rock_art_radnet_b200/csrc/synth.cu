// synth.cu - device-side generator of synthetic panels for the archive sweep (BASELINE configs[4], SURVEY.md
// 8(d) config 5: "generated on device from the seed to avoid H2D dominating").  Bench / test utility: counter-
// based (every value is a hash of (seed, panel id, tensor, element index)), so any panel can be regenerated
// anywhere - on another rank, in another batch split, or copied to the host and handed to the oracle.
//   cls  (B,H,W,A)   uniform scores in (0,1), 24 bits (a handful of ties per panel; ties are counted by K2)
//   regr (B,H,W,4A)  0.5 * N(0,1)     
//   feat (B,H,W,C)   N(0,1)
#include "common.cuh"

namespace radnet {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {          // splitmix64 finaliser
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ float u01(uint32_t bits) { return ((bits >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// two standard normals from one 64-bit hash (Box-Muller)
__device__ __forceinline__ float2 normal2(uint64_t h) {
    const float u1 = u01((uint32_t)h), u2 = u01((uint32_t)(h >> 32));
    const float r = sqrtf(-2.0f * __logf(u1));
    float sn, cs;
    __sincosf(6.28318530717958647692f * u2, &sn, &cs);
    return make_float2(r * cs, r * sn);
}

struct SynthParams {
    uint64_t seed;
    long long first_panel, panel_stride;
    int B;
    long long n_cls, n_regr, n_feat;       // elements per panel
    float *cls, *regr, *feat;
};

__global__ void __launch_bounds__(256) synth_panels_kernel(SynthParams p) {
    const int b = blockIdx.y;
    const uint64_t key = mix64(p.seed ^ mix64((uint64_t)(p.first_panel + (long long)b * p.panel_stride)));
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    const long long t0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    // feature map and regression map: four normals per thread and step (16-byte stores; sizes are multiples of 4
    // or the tail is written element-wise)
    for (int which = 0; which < 2; ++which) {
        float *dst = (which ? p.regr : p.feat) + (size_t)b * (which ? p.n_regr : p.n_feat);
        const long long n = which ? p.n_regr : p.n_feat;
        const float scale = which ? 0.5f : 1.0f;
        const uint64_t tk = key ^ (which ? 0xA5A5A5A5A5A5A5A5ull : 0x5A5A5A5A5A5A5A5Aull);
        for (long long i = t0; i < n; i += stride) {
            const float2 x = normal2(mix64(tk + (uint64_t)i)), y = normal2(mix64(tk + (uint64_t)i + 2));
            if (i + 3 < n && ((reinterpret_cast<uintptr_t>(dst + i) & 15) == 0)) {
                *reinterpret_cast<float4 *>(dst + i) = make_float4(x.x * scale, x.y * scale, y.x * scale, y.y * scale);
            } else {
                const float v[4] = {x.x * scale, x.y * scale, y.x * scale, y.y * scale};
                for (int k = 0; k < 4 && i + k < n; ++k) dst[i + k] = v[k];
            }
        }
    }
    float *cls = p.cls + (size_t)b * p.n_cls;
    for (long long i = t0 / 4; i < p.n_cls; i += stride / 4)
        cls[i] = u01((uint32_t)mix64(key ^ 0x3C3C3C3C3C3C3C3Cull ^ ((uint64_t)i << 1)));
}

}  // namespace radnet

using namespace radnet;

extern "C" int radnet_synth_panels(unsigned long long seed, long long first_panel, long long panel_stride, int B, int H,
                                   int W, int A, int C, float *cls, float *regr, float *feat, void *stream) {
    RADNET_CHECK_ARG(cls && regr && feat && B >= 1 && B <= 65535 && H >= 1 && W >= 1 && A >= 1 && C >= 1,
                     "synth_panels: bad arguments");
    SynthParams p{};
    p.seed = seed; p.first_panel = first_panel; p.panel_stride = panel_stride; p.B = B;
    p.n_cls = (long long)H * W * A; p.n_regr = 4 * p.n_cls; p.n_feat = (long long)H * W * C;
    p.cls = cls; p.regr = regr; p.feat = feat;
    long long blocks = (p.n_feat / 4 + 255) / 256;
    if (blocks > 296) blocks = 296;
    if (blocks < 1) blocks = 1;
    synth_panels_kernel<<<dim3((unsigned)blocks, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("synth_panels_kernel");
}
