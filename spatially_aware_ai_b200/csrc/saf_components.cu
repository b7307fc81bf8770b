// saf_components.cu -- object labelling of the fused class grid: the flood fill of flood_fill_3d
// (/root/reference/handy_utils.py:295-480) as a device connected-component labelling.
//
// The reference walks every voxel in a Python triple loop and flood-fills (26-neighbourhood, equal class id) from
// each unvisited voxel whose class is neither -1 (unobserved) nor the null class (133); objects with fewer than 3
// voxels are rejected; accepted objects get the indices -2, -3, ... in the order the scan meets them, i.e. in
// ascending order of their first (smallest flat index) voxel; voxel_obj_ids holds the index per voxel, -1 elsewhere.
// Here: union-find over the 13 forward neighbours (roots are the smallest flat index of their component, so the
// scan order falls out of a prefix sum over the roots), sizes by atomics, ranks by a two-level scan.
#include <cuda_runtime.h>

#include <algorithm>

#include "saf_internal.cuh"

namespace saf {
namespace {

constexpr int kCcThreads = 256;
constexpr uint32_t kNone = 0xffffffffu;

struct CcParams {
    const long long* labels;   // [n] class ids (argmax_with_check_2d_efficient's output: -1 = unobserved)
    int nx, ny, nz;
    int64_t n;
    int null_class, min_voxels;
    uint32_t nblk;
    uint32_t* parent;   // [n]
    uint32_t* size;     // [n]  component size at the root; later the root's rank
    uint32_t* blk;      // [nblk + 1]
    uint32_t* total;    // [1]
    int32_t* out;
};

__device__ __forceinline__ uint32_t cc_find(uint32_t* parent, uint32_t x)
{
    for (;;) {
        const uint32_t p = ((volatile uint32_t*)parent)[x];
        if (p == x) return x;
        x = p;
    }
}

__device__ __forceinline__ void cc_unite(uint32_t* parent, uint32_t a, uint32_t b)
{
    for (;;) {
        a = cc_find(parent, a);
        b = cc_find(parent, b);
        if (a == b) return;
        if (a > b) {
            const uint32_t t = a;
            a = b;
            b = t;
        }
        const uint32_t old = atomicMin(&parent[b], a);   // link the larger root below the smaller one
        if (old == b) return;
        b = old;
    }
}

__device__ __forceinline__ bool cc_foreground(const CcParams& p, long long c) { return c != -1 && c != p.null_class; }

__global__ void __launch_bounds__(kCcThreads) cc_init_kernel(const CcParams p)
{
    const int64_t v = (int64_t)blockIdx.x * kCcThreads + threadIdx.x;
    if (v >= p.n) return;
    p.parent[v] = cc_foreground(p, __ldg(p.labels + v)) ? (uint32_t)v : kNone;
    p.size[v] = 0;
}

__global__ void __launch_bounds__(kCcThreads) cc_merge_kernel(const CcParams p)
{
    const int64_t v = (int64_t)blockIdx.x * kCcThreads + threadIdx.x;
    if (v >= p.n) return;
    const long long c = __ldg(p.labels + v);
    if (!cc_foreground(p, c)) return;
    const int z = (int)(v % p.nz);
    const int y = (int)((v / p.nz) % p.ny);
    const int x = (int)(v / ((int64_t)p.nz * p.ny));
    // the 13 neighbours that follow v in flat order (the other 13 are handled from their side)
#pragma unroll
    for (int k = 14; k < 27; ++k) {
        const int dx = k / 9 - 1, dy = (k / 3) % 3 - 1, dz = k % 3 - 1;
        const int xx = x + dx, yy = y + dy, zz = z + dz;
        if (xx < 0 || xx >= p.nx || yy < 0 || yy >= p.ny || zz < 0 || zz >= p.nz) continue;
        const int64_t u = ((int64_t)xx * p.ny + yy) * p.nz + zz;
        if (__ldg(p.labels + u) == c) cc_unite(p.parent, (uint32_t)v, (uint32_t)u);
    }
}

__global__ void __launch_bounds__(kCcThreads) cc_flatten_kernel(const CcParams p)
{
    const int64_t v = (int64_t)blockIdx.x * kCcThreads + threadIdx.x;
    if (v >= p.n || p.parent[v] == kNone) return;
    const uint32_t r = cc_find(p.parent, (uint32_t)v);
    p.parent[v] = r;   // r <= v; other threads following this link land on a root either way
    atomicAdd(&p.size[r], 1u);
}

__device__ __forceinline__ uint32_t cc_cta_exclusive(uint32_t val, uint32_t* s_warp, uint32_t& total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = val;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t nb = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += nb;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t base = 0;
    total = 0;
#pragma unroll
    for (int w = 0; w < kCcThreads / 32; ++w) {
        const uint32_t s = s_warp[w];
        if (w < warp) base += s;
        total += s;
    }
    __syncthreads();
    return base + inc - val;
}

__device__ __forceinline__ bool cc_accepted_root(const CcParams& p, int64_t v)
{
    return v < p.n && p.parent[v] == (uint32_t)v && p.size[v] >= (uint32_t)p.min_voxels;
}

__global__ void __launch_bounds__(kCcThreads) cc_count_kernel(const CcParams p)
{
    __shared__ uint32_t s_warp[kCcThreads / 32];
    const int64_t v = (int64_t)blockIdx.x * kCcThreads + threadIdx.x;
    uint32_t total;
    cc_cta_exclusive(cc_accepted_root(p, v) ? 1u : 0u, s_warp, total);
    if (threadIdx.x == 0) p.blk[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) cc_scan_kernel(const CcParams p)
{
    __shared__ uint32_t s_part[1024];
    const uint32_t per = (p.nblk + 1023u) / 1024u;
    const uint32_t lo = min(p.nblk, threadIdx.x * per), hi = min(p.nblk, lo + per);
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; ++i) sum += p.blk[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int i = 0; i < 1024; ++i) {
            const uint32_t t = s_part[i];
            s_part[i] = run;
            run += t;
        }
        *p.total = run;
    }
    __syncthreads();
    uint32_t run = s_part[threadIdx.x];
    for (uint32_t i = lo; i < hi; ++i) {
        const uint32_t t = p.blk[i];
        p.blk[i] = run;
        run += t;
    }
}

__global__ void __launch_bounds__(kCcThreads) cc_rank_kernel(const CcParams p)
{
    __shared__ uint32_t s_warp[kCcThreads / 32];
    const int64_t v = (int64_t)blockIdx.x * kCcThreads + threadIdx.x;
    const bool acc = cc_accepted_root(p, v);
    const bool root = v < p.n && p.parent[v] == (uint32_t)v;
    uint32_t total;
    const uint32_t rank = p.blk[blockIdx.x] + cc_cta_exclusive(acc ? 1u : 0u, s_warp, total);
    if (root) p.size[v] = acc ? rank : kNone;   // size is no longer needed at this root
}

__global__ void __launch_bounds__(kCcThreads) cc_write_kernel(const CcParams p)
{
    const int64_t v = (int64_t)blockIdx.x * kCcThreads + threadIdx.x;
    if (v >= p.n) return;
    const uint32_t r = p.parent[v];
    int32_t o = -1;
    if (r != kNone) {
        const uint32_t rank = p.size[r];
        if (rank != kNone) o = -2 - (int32_t)rank;   // handy_utils.py:356, 445-448
    }
    p.out[v] = o;
}

uint64_t cc_bytes(int64_t n)
{
    const uint64_t nblk = (uint64_t)((n + kCcThreads - 1) / kCcThreads);
    auto up = [](uint64_t v) { return (v + 255ull) & ~255ull; };
    return 256 + up(4ull * (nblk + 1)) + 2 * up(4ull * (uint64_t)n);
}

}  // namespace
}  // namespace saf

using namespace saf;

extern "C" {

int saf_label_components_workspace_bytes(int64_t n_voxels, uint64_t* bytes_out)
{
    if (!bytes_out) return SAF_ERR_NULL;
    if (n_voxels <= 0 || n_voxels >= (1ll << 32) - 1) return SAF_ERR_GRID;
    *bytes_out = cc_bytes(n_voxels);
    return 0;
}

int saf_label_components(const int64_t* labels, int32_t nx, int32_t ny, int32_t nz, int32_t null_class,
                         int32_t min_voxels, int32_t* out_obj, void* ws, uint64_t ws_bytes, uint32_t* n_objects_out,
                         void* stream)
{
    int sms = 0;
    int rc = device_sm_count(&sms, nullptr);
    if (rc) return rc;
    if (!labels || !out_obj || !ws || !n_objects_out) return SAF_ERR_NULL;
    if (nx <= 0 || ny <= 0 || nz <= 0) return SAF_ERR_GRID;
    const int64_t n = (int64_t)nx * ny * nz;
    if (n >= (1ll << 32) - 1) return SAF_ERR_GRID;
    if (ws_bytes < cc_bytes(n)) return SAF_ERR_WORKSPACE;
    if (((uintptr_t)ws & 255u) != 0) return SAF_ERR_ALIGNMENT;
    auto up = [](uint64_t v) { return (v + 255ull) & ~255ull; };
    CcParams p;
    p.labels = (const long long*)labels;
    p.nx = nx;
    p.ny = ny;
    p.nz = nz;
    p.n = n;
    p.null_class = null_class;
    p.min_voxels = min_voxels;
    p.nblk = (uint32_t)((n + kCcThreads - 1) / kCcThreads);
    unsigned char* base = (unsigned char*)ws;
    p.total = (uint32_t*)base;
    p.blk = (uint32_t*)(base + 256);
    p.parent = (uint32_t*)(base + 256 + up(4ull * (p.nblk + 1)));
    p.size = (uint32_t*)((unsigned char*)p.parent + up(4ull * (uint64_t)n));
    p.out = out_obj;
    cudaStream_t st = (cudaStream_t)stream;
    cc_init_kernel<<<p.nblk, kCcThreads, 0, st>>>(p);
    SAF_CHECK_LAUNCH("cc_init_kernel", st);
    cc_merge_kernel<<<p.nblk, kCcThreads, 0, st>>>(p);
    SAF_CHECK_LAUNCH("cc_merge_kernel", st);
    cc_flatten_kernel<<<p.nblk, kCcThreads, 0, st>>>(p);
    SAF_CHECK_LAUNCH("cc_flatten_kernel", st);
    cc_count_kernel<<<p.nblk, kCcThreads, 0, st>>>(p);
    SAF_CHECK_LAUNCH("cc_count_kernel", st);
    cc_scan_kernel<<<1, 1024, 0, st>>>(p);
    SAF_CHECK_LAUNCH("cc_scan_kernel", st);
    cc_rank_kernel<<<p.nblk, kCcThreads, 0, st>>>(p);
    SAF_CHECK_LAUNCH("cc_rank_kernel", st);
    cc_write_kernel<<<p.nblk, kCcThreads, 0, st>>>(p);
    SAF_CHECK_LAUNCH("cc_write_kernel", st);
    SAF_CUDA_TRY(cudaMemcpyAsync(n_objects_out, p.total, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SAF_CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
