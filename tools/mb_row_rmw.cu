// mb_row_rmw.cu -- microbenchmark: what HBM bandwidth can a read-modify-write of scattered 3 KB rows
// reach on this GPU, compared with a streaming copy?  (The feature-accumulate kernel's access pattern
// without its arithmetic.)   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mb_row_rmw mb_row_rmw.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <numeric>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int C = 768;

__device__ __forceinline__ float4 ldna(const float4* p) { float4 v; asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)); return v; }
__device__ __forceinline__ void stna(float4* p, float4 v) { asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }

// warp per row, VPW rows in flight per warp (registers)
template <int VPW>
__global__ void __launch_bounds__(512) rmw_rows(float* data, const uint32_t* idx, uint32_t n, float a)
{
    const int lane = threadIdx.x & 31;
    const uint32_t nw = gridDim.x * 16, gw = blockIdx.x * 16 + (threadIdx.x >> 5);
    for (uint64_t i0 = gw; i0 < n; i0 += (uint64_t)nw * VPW) {
        float4 v[VPW][6]; uint32_t r[VPW]; bool act[VPW];
#pragma unroll
        for (int k = 0; k < VPW; ++k) { uint64_t i = i0 + (uint64_t)k * nw; act[k] = i < n; r[k] = act[k] ? idx[i] : 0; }
#pragma unroll
        for (int k = 0; k < VPW; ++k) if (act[k]) { const float4* p = (const float4*)(data + (size_t)r[k] * C);
#pragma unroll
            for (int j = 0; j < 6; ++j) v[k][j] = ldna(p + j * 32 + lane); }
#pragma unroll
        for (int k = 0; k < VPW; ++k) if (act[k]) { float4* p = (float4*)(data + (size_t)r[k] * C);
#pragma unroll
            for (int j = 0; j < 6; ++j) { float4 x = v[k][j]; x.x = x.x * a + 1.f; x.y = x.y * a + 1.f; x.z = x.z * a + 1.f; x.w = x.w * a + 1.f; stna(p + j * 32 + lane, x); } }
    }
}

// read-only and write-only variants
__global__ void __launch_bounds__(512) read_rows(const float* data, const uint32_t* idx, uint32_t n, float* sink)
{
    const int lane = threadIdx.x & 31; float acc = 0;
    const uint32_t nw = gridDim.x * 16, gw = blockIdx.x * 16 + (threadIdx.x >> 5);
    for (uint64_t i = gw; i < n; i += nw) { const float4* p = (const float4*)(data + (size_t)idx[i] * C);
#pragma unroll
        for (int j = 0; j < 6; ++j) { float4 x = ldna(p + j * 32 + lane); acc += x.x + x.y + x.z + x.w; } }
    if (acc == 123.456f) *sink = acc;
}
__global__ void __launch_bounds__(512) write_rows(float* data, const uint32_t* idx, uint32_t n, float a)
{
    const int lane = threadIdx.x & 31;
    const uint32_t nw = gridDim.x * 16, gw = blockIdx.x * 16 + (threadIdx.x >> 5);
    for (uint64_t i = gw; i < n; i += nw) { float4* p = (float4*)(data + (size_t)idx[i] * C);
#pragma unroll
        for (int j = 0; j < 6; ++j) stna(p + j * 32 + lane, make_float4(a, a, a, a)); }
}
__global__ void copy_stream(const float4* __restrict__ src, float4* __restrict__ dst, size_t n4)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

int main(int argc, char** argv)
{
    const size_t rows = argc > 1 ? atoll(argv[1]) : 8000000;   // 24.6 GB
    const uint32_t n = argc > 2 ? atoi(argv[2]) : 65536;        // rows touched per launch (~200 MB)
    float *data, *sink; uint32_t* didx;
    CK(cudaMalloc(&data, rows * C * sizeof(float))); CK(cudaMemset(data, 0, rows * C * sizeof(float)));
    CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&didx, (size_t)n * 4 * 64));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::mt19937_64 rng(1);
    auto bench = [&](const char* name, int pattern, int kind) {
        // 64 different index sets so consecutive launches never hit L2-resident rows
        std::vector<uint32_t> all;
        for (int rep = 0; rep < 64; ++rep) {
            std::vector<uint32_t> v(n);
            if (pattern == 0) { for (auto& x : v) x = rng() % rows; }
            else if (pattern == 1) { for (auto& x : v) x = rng() % rows; std::sort(v.begin(), v.end()); }
            else if (pattern == 2) { for (uint32_t i = 0; i < n; i += 4) { uint32_t b = rng() % (rows - 4); for (int k = 0; k < 4 && i + k < n; ++k) v[i + k] = b + k; } std::sort(v.begin(), v.end()); }
            else { uint32_t b = rng() % (rows - n); std::iota(v.begin(), v.end(), b); }
            all.insert(all.end(), v.begin(), v.end());
        }
        CK(cudaMemcpy(didx, all.data(), all.size() * 4, cudaMemcpyHostToDevice));
        float best = 1e9, tot = 0;
        for (int it = 0; it < 64; ++it) {
            const uint32_t* ix = didx + (size_t)it * n;
            CK(cudaEventRecord(e0));
            if (kind == 0) rmw_rows<1><<<148, 512>>>(data, ix, n, 0.5f);
            else if (kind == 1) rmw_rows<2><<<148, 512>>>(data, ix, n, 0.5f);
            else if (kind == 2) rmw_rows<2><<<296, 512>>>(data, ix, n, 0.5f);
            else if (kind == 3) read_rows<<<296, 512>>>(data, ix, n, sink);
            else if (kind == 4) write_rows<<<296, 512>>>(data, ix, n, 0.5f);
            else rmw_rows<4><<<148, 512>>>(data, ix, n, 0.5f);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (it >= 4) { best = std::min(best, ms); tot += ms; }
        }
        const double bytes = (double)n * C * 4 * ((kind == 3 || kind == 4) ? 1 : 2);
        printf("%-28s pattern %d  avg %7.1f us  %7.1f GB/s   best %7.1f GB/s\n", name, pattern, tot / 60 * 1e3, bytes / (tot / 60 * 1e-3) / 1e9, bytes / (best * 1e-3) / 1e9);
    };
    // streaming copy reference (read + write bytes)
    { size_t n4 = (size_t)1 << 28; // 4 GiB each way
      float4* a = (float4*)data; float4* b = (float4*)data + n4; float best = 1e9;
      for (int it = 0; it < 6; ++it) { CK(cudaEventRecord(e0)); copy_stream<<<148 * 16, 512>>>(a, b, n4); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms); }
      printf("streaming copy 2 x 4 GiB: %.1f GB/s (read+write)\n", 2.0 * n4 * 16 / (best * 1e-3) / 1e9); }
    const char* pn[] = {"random", "random sorted", "runs of 4 sorted", "contiguous"};
    for (int pattern = 0; pattern < 4; ++pattern) {
        printf("-- %s rows, %u rows of %d B per launch\n", pn[pattern], n, C * 4);
        bench("rmw 148x512 VPW1", pattern, 0);
        bench("rmw 148x512 VPW2", pattern, 1);
        bench("rmw 296x512 VPW2", pattern, 2);
        bench("rmw 148x512 VPW4", pattern, 5);
        bench("read only 296x512", pattern, 3);
        bench("write only 296x512", pattern, 4);
    }
    return 0;
}
