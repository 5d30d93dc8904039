// Experiment: how fast can N CTAs stream zeros to HBM on B200 (cold L2), vs cudaMemsetAsync?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/fill_bw tools/exp/fill_bw.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>

__global__ void __launch_bounds__(1024, 1) fill_kernel(double2 *dst, long long n_items, int cs) {
    const long long share = ((n_items + gridDim.x - 1) / gridDim.x + 63) & ~63LL;
    long long lo = share * blockIdx.x, hi = lo + share;
    if (hi > n_items) hi = n_items;
    const double2 z = make_double2(0.0, 0.0);
    if (cs) {
        for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x)
            asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(dst + i), "d"(0.0), "d"(0.0) : "memory");
    } else {
#pragma unroll 4
        for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) dst[i] = z;
    }
}

__global__ void __launch_bounds__(256, 4) fill_kernel_small(double2 *dst, long long n_items) {
    const long long share = ((n_items + gridDim.x - 1) / gridDim.x + 63) & ~63LL;
    long long lo = share * blockIdx.x, hi = lo + share;
    if (hi > n_items) hi = n_items;
    const double2 z = make_double2(0.0, 0.0);
#pragma unroll 4
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) dst[i] = z;
}

int main() {
    const size_t flush_bytes = 256u << 20;
    void *flush; cudaMalloc(&flush, flush_bytes);
    for (size_t bytes : {(size_t)66580480, (size_t)532643840}) {
        double2 *buf; cudaMalloc(&buf, bytes);
        const long long n_items = bytes / 16;
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        auto run = [&](const char *name, auto fn) {
            std::vector<float> ts;
            for (int it = 0; it < 12; ++it) {
                cudaMemsetAsync(flush, it, flush_bytes);
                cudaEventRecord(a);
                fn();
                cudaEventRecord(b);
                cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b);
                if (it >= 2) ts.push_back(ms);
            }
            std::sort(ts.begin(), ts.end());
            float p50 = ts[ts.size() / 2];
            printf("%-34s %9zu B  p50 %7.2f us  min %7.2f us  -> %6.0f GB/s\n", name, bytes, p50 * 1e3, ts[0] * 1e3, bytes / (p50 * 1e-3) / 1e9);
        };
        run("cudaMemsetAsync", [&] { cudaMemsetAsync(buf, 0, bytes); });
        for (int n : {32, 64, 84, 100, 120, 148, 296}) {
            char nm[64]; snprintf(nm, sizeof nm, "fill 1024thr x %d CTAs", n);
            run(nm, [&] { fill_kernel<<<n, 1024>>>(buf, n_items, 0); });
        }
        run("fill 1024thr x 148 CTAs st.cs", [&] { fill_kernel<<<148, 1024>>>(buf, n_items, 1); });
        for (int n : {148, 296, 592, 1184}) {
            char nm[64]; snprintf(nm, sizeof nm, "fill 256thr x %d CTAs (4/SM)", n);
            run(nm, [&] { fill_kernel_small<<<n, 256>>>(buf, n_items); });
        }
        cudaFree(buf);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
