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

// + bilinear mix of 4 rows of a 35 x 768 table held in shared memory (the feature kernel's arithmetic)
// SMALL: + the per-voxel scattered 4-byte read-modify-writes (weight, 3 x rgb, one label counter)
template <bool SMALL>
__global__ void __launch_bounds__(512, 1) rmw_rows_table(float* data, const uint32_t* idx, uint32_t n, const float* table,
                                                        int* weight, float* rgb, int* labels)
{
    extern __shared__ float tab[];
    for (int e = threadIdx.x; e < 35 * C; e += 512) tab[e] = table[e];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t nw = gridDim.x * 16, gw = blockIdx.x * 16 + (threadIdx.x >> 5);
    const float4* tab4 = (const float4*)tab;
    for (uint64_t i = gw; i < n; i += nw) {
        const uint32_t r = idx[i];
        float4 v[6];
        const float4* p = (const float4*)(data + (size_t)r * C);
#pragma unroll
        for (int j = 0; j < 6; ++j) v[j] = ldna(p + j * 32 + lane);
        int w = 1;
        if (SMALL) w = weight[r];
        const float a = 1.0f / (float)(w + 1), b = (float)w * a;
        const int r0 = r % 27, r1 = r0 + 1, r2 = r0 + 7, r3 = r0 + 8;   // 5 x 7 patch grid neighbours
        const float w0 = 0.3f, w1 = 0.2f, w2 = 0.4f, w3 = 0.1f;
        float4* q = (float4*)(data + (size_t)r * C);
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int col = j * 32 + lane;
            const float4 t0 = tab4[r0 * 192 + col], t1 = tab4[r1 * 192 + col], t2 = tab4[r2 * 192 + col], t3 = tab4[r3 * 192 + col];
            float4 x;
            x.x = fmaf(t3.x, w3, fmaf(t2.x, w2, fmaf(t1.x, w1, t0.x * w0))) * a + v[j].x * b;
            x.y = fmaf(t3.y, w3, fmaf(t2.y, w2, fmaf(t1.y, w1, t0.y * w0))) * a + v[j].y * b;
            x.z = fmaf(t3.z, w3, fmaf(t2.z, w2, fmaf(t1.z, w1, t0.z * w0))) * a + v[j].z * b;
            x.w = fmaf(t3.w, w3, fmaf(t2.w, w2, fmaf(t1.w, w1, t0.w * w0))) * a + v[j].w * b;
            stna(q + col, x);
        }
        if (SMALL) {
            if (lane < 3) rgb[(size_t)r * 3 + lane] = rgb[(size_t)r * 3 + lane] * b + a;
            else if (lane == 3) labels[(size_t)r * 143 + (r % 133)] += 1;
            else if (lane == 4) weight[r] = w + 1;
        }
    }
}

// the same work organised like the feature kernel: per batch of 32 rows, lane l does row l's small state
// (all 32 in parallel), then the warp streams the 32 rows
__global__ void __launch_bounds__(512, 1) rmw_rows_table_batched(float* data, const uint32_t* idx, uint32_t n, const float* table,
                                                                int* weight, float* rgb, int* labels)
{
    extern __shared__ float tab[];
    for (int e = threadIdx.x; e < 35 * C; e += 512) tab[e] = table[e];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t nw = gridDim.x * 16, gw = blockIdx.x * 16 + (threadIdx.x >> 5);
    const float4* tab4 = (const float4*)tab;
    const uint32_t k_total = n > gw ? (n - gw + nw - 1) / nw : 0;
    for (uint32_t kb = 0; kb < k_total; kb += 32) {
        const uint32_t cnt = min(32u, k_total - kb);
        uint32_t my_r = 0; int my_w = 0;
        if (lane < cnt) {
            my_r = idx[gw + (kb + lane) * nw];
            my_w = weight[my_r];
            const float a = 1.0f / (float)(my_w + 1), b = (float)my_w * a;
            float* c = rgb + (size_t)my_r * 3;
            const float c0 = c[0], c1 = c[1], c2 = c[2];
            int* lab = labels + (size_t)my_r * 143 + (my_r % 133);
            const int l0 = *lab;
            c[0] = c0 * b + a; c[1] = c1 * b + a; c[2] = c2 * b + a;
            *lab = l0 + 1;
            weight[my_r] = my_w + 1;
        }
        for (uint32_t q = 0; q < cnt; ++q) {
            const uint32_t r = __shfl_sync(0xffffffffu, my_r, q);
            const int w = __shfl_sync(0xffffffffu, my_w, q);
            float4 v[6];
            const float4* p = (const float4*)(data + (size_t)r * C);
#pragma unroll
            for (int j = 0; j < 6; ++j) v[j] = ldna(p + j * 32 + lane);
            const float a = 1.0f / (float)(w + 1), b = (float)w * a;
            const int r0 = r % 27, r1 = r0 + 1, r2 = r0 + 7, r3 = r0 + 8;
            const float w0 = 0.3f, w1 = 0.2f, w2 = 0.4f, w3 = 0.1f;
            float4* qp = (float4*)(data + (size_t)r * C);
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const int col = j * 32 + lane;
                const float4 t0 = tab4[r0 * 192 + col], t1 = tab4[r1 * 192 + col], t2 = tab4[r2 * 192 + col], t3 = tab4[r3 * 192 + col];
                float4 x;
                x.x = fmaf(t3.x, w3, fmaf(t2.x, w2, fmaf(t1.x, w1, t0.x * w0))) * a + v[j].x * b;
                x.y = fmaf(t3.y, w3, fmaf(t2.y, w2, fmaf(t1.y, w1, t0.y * w0))) * a + v[j].y * b;
                x.z = fmaf(t3.z, w3, fmaf(t2.z, w2, fmaf(t1.z, w1, t0.z * w0))) * a + v[j].z * b;
                x.w = fmaf(t3.w, w3, fmaf(t2.w, w2, fmaf(t1.w, w1, t0.w * w0))) * a + v[j].w * b;
                stna(qp + col, x);
            }
        }
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
    const size_t rows = argc > 1 ? atoll(argv[1]) : 6000000;   // 18.4 GB (+ 3.4 GB label counters)
    const uint32_t n = argc > 2 ? atoi(argv[2]) : 65536;        // rows touched per launch (~200 MB)
    float *data, *sink, *table, *rgb; uint32_t* didx; int *weight, *labels;
    CK(cudaMalloc(&table, 35 * C * 4)); CK(cudaMemset(table, 0, 35 * C * 4));
    CK(cudaMalloc(&rgb, rows * 12)); CK(cudaMemset(rgb, 0, rows * 12));
    CK(cudaMalloc(&weight, rows * 4)); CK(cudaMemset(weight, 0, rows * 4));
    CK(cudaMalloc(&labels, rows * 143 * 4)); CK(cudaMemset(labels, 0, rows * 143 * 4));
    CK(cudaFuncSetAttribute(rmw_rows_table<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 35 * C * 4));
    CK(cudaFuncSetAttribute(rmw_rows_table<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 35 * C * 4));
    CK(cudaFuncSetAttribute(rmw_rows_table_batched, cudaFuncAttributeMaxDynamicSharedMemorySize, 35 * C * 4));
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
            else if (kind == 6) rmw_rows_table<false><<<148, 512, 35 * C * 4>>>(data, ix, n, table, weight, rgb, labels);
            else if (kind == 7) rmw_rows_table<true><<<148, 512, 35 * C * 4>>>(data, ix, n, table, weight, rgb, labels);
            else if (kind == 8) rmw_rows_table_batched<<<148, 512, 35 * C * 4>>>(data, ix, n, table, weight, rgb, labels);
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
        bench("rmw + smem table mix", pattern, 6);
        bench("rmw + table + small state", pattern, 7);
        bench("same, small state batched", pattern, 8);
        bench("read only 296x512", pattern, 3);
        bench("write only 296x512", pattern, 4);
    }
    return 0;
}
