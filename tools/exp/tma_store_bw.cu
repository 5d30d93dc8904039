// Experiment: HBM write bandwidth of K4-shaped output streams on B200, STG vs TMA stores.
//   out = [P pixels][C = 1024 floats]; a CTA owns a 32-channel slice (128 B per pixel) of a panel of PP pixels.
//   mode 0: st.global.cs.v4 (8 lanes x 16 B per pixel piece), mode 1: cp.async.bulk.tensor.2d store of a
//   (32 ch x ROWS px) tile from shared memory, NBUF tiles in flight per CTA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/tma_store_bw tools/exp/tma_store_bw.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>

constexpr int C = 1024;
constexpr int PIX_PER_PANEL = 300 * 196;

__global__ void __launch_bounds__(512, 1) stg_kernel(float4 *out, int n_work, int slices) {
    const int q = threadIdx.x % 8, g = threadIdx.x / 8;
    for (int work = blockIdx.x; work < n_work; work += gridDim.x) {
        const int b = work / slices, s = work % slices;
        float4 *base = out + (size_t)b * PIX_PER_PANEL * (C / 4) + s * 8 + q;
        const float4 v = make_float4(1.f, 2.f, 3.f, (float)work);
        // like K4: a group walks 14 rows of one column (px stride 1 pixel, py stride 14 pixels)
        for (int col = g; col < 300 * 14; col += 64) {
            const int r = col / 14, px = col % 14;
            float4 *d = base + (size_t)(r * 196 + px) * (C / 4);
#pragma unroll
            for (int py = 0; py < 14; ++py) {
                asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                d += 14 * (C / 4);
            }
        }
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NBUF>
__global__ void __launch_bounds__(512) tma_kernel(const __grid_constant__ CUtensorMap tmap, int n_work, int slices, int rows,
                                                      int fill_threads_work, int chw) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tile_bytes = rows * chw * 4;
    const int tiles_per_panel = PIX_PER_PANEL / rows;
    int it = 0;
    for (int work = blockIdx.x; work < n_work; work += gridDim.x) {
        const int b = work / slices, s = work % slices;
        for (int t = 0; t < tiles_per_panel; ++t, ++it) {
            unsigned char *buf = smem + (size_t)(it % NBUF) * tile_bytes;
            if (threadIdx.x == 0 && it >= NBUF) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");
            __syncthreads();
            if (fill_threads_work > 1) {      // stand-in for the bilinear work: ~fill_threads_work dependent FMAs per 16 B
                float x = (float)threadIdx.x;
                for (int i = threadIdx.x; i < tile_bytes / 16; i += blockDim.x)
                    for (int k = 0; k < fill_threads_work; ++k) x = x * 1.0001f + 0.5f;
                if (x == 12345.f) reinterpret_cast<float *>(buf)[0] = x;
            }
            if (fill_threads_work) {      // the SM-side cost of producing a tile: 16 B per thread per store
                for (int i = threadIdx.x; i < tile_bytes / 16; i += blockDim.x)
                    reinterpret_cast<float4 *>(buf)[i] = make_float4(1.f, 2.f, (float)t, (float)work);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&tmap),
                             "r"(s * chw), "r"(b * PIX_PER_PANEL + t * rows), "r"(smem_u32(buf))
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 37;
    const int slices = C / 32;
    const size_t bytes = (size_t)B * PIX_PER_PANEL * C * 4;
    float *out;
    if (cudaMalloc(&out, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    void *flush; cudaMalloc(&flush, 256u << 20);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto run = [&](const char *name, auto launch) {
        std::vector<float> ts;
        for (int it = 0; it < 7; ++it) {
            cudaMemsetAsync(flush, it, 256u << 20);
            cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); ts.push_back(ms);
        }
        std::sort(ts.begin(), ts.end());
        cudaError_t e = cudaGetLastError();
        printf("%-52s %8.3f ms  %7.0f GB/s  %s\n", name, ts[ts.size() / 2], bytes / ts[ts.size() / 2] / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
        fflush(stdout);
    };
    run("cudaMemsetAsync", [&] { cudaMemsetAsync(out, 0, bytes); });
    char name[128];
    {
        const int n_work = B * slices;
        for (int grid : {n_work, 148, 128}) {
            snprintf(name, sizeof name, "STG K4-like columns, grid %d of %d", grid, n_work);
            run(name, [&] { stg_kernel<<<grid, 512>>>((float4 *)out, n_work, slices); });
        }
    }
    struct Cfg { int chw, rows, nbuf, threads, work; };
    const Cfg cfgs[] = {{32, 196, 1, 512, 1}, {32, 196, 2, 512, 1}, {32, 196, 4, 512, 1}, {32, 196, 1, 512, 24}, {32, 196, 2, 512, 24},
                        {32, 98, 2, 512, 1}, {32, 98, 3, 512, 1}, {32, 98, 3, 512, 24}, {32, 98, 4, 512, 24},
                        {16, 196, 4, 512, 1}, {16, 196, 4, 512, 24}, {16, 196, 8, 512, 24}, {64, 196, 2, 512, 24}, {64, 98, 4, 512, 24}};
    for (const Cfg &c : cfgs) {
        const int sl = C / c.chw, n_work = B * sl;
        alignas(64) CUtensorMap tmap;
        const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)B * PIX_PER_PANEL};
        const cuuint64_t strides[1] = {(cuuint64_t)C * 4};
        const cuuint32_t box[2] = {(cuuint32_t)c.chw, (cuuint32_t)c.rows};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
        for (int grid : {n_work, 148, 128, 296}) {
            if (n_work % grid != 0) continue;
            // shared memory: the tiles, plus (grid <= 148) a dummy "resident map" so that one CTA fits per SM
            const size_t smem = (size_t)c.nbuf * c.rows * c.chw * 4 + (grid <= 148 ? 120 * 1024 : (grid == 296 ? 60 * 1024 : 0));
            snprintf(name, sizeof name, "TMA ch=%d rows=%d nbuf=%d thr=%d work=%d grid %d/%d", c.chw, c.rows, c.nbuf, c.threads, c.work, grid, n_work);
            auto go = [&](auto kern) {
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                run(name, [&] { kern<<<grid, c.threads, smem>>>(tmap, n_work, sl, c.rows, c.work, c.chw); });
            };
            if (c.nbuf == 1) go(tma_kernel<1>); else if (c.nbuf == 2) go(tma_kernel<2>); else if (c.nbuf == 3) go(tma_kernel<3>);
            else if (c.nbuf == 4) go(tma_kernel<4>); else go(tma_kernel<8>);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
