// saf_mesh.cu -- extract_mesh (clip_seem_fusion.py:824-888, clipfusion.py:723-763) on the device:
// marching cubes over the TSDF with unobserved voxels masked to NaN, the reference's face / vertex
// filtering folded into the emission, and the per-vertex grid_sample calls.
//
// The reference copies the whole TSDF to the host, calls skimage.measure.marching_cubes there, filters
// NaN faces with numpy and then runs four torch grid_sample calls.  Here:
//   mc_classify_kernel     cell -> case index, surviving triangles (no NaN vertex), marks the grid edges
//                          they use; per-CTA triangle totals
//   mc_count_verts_kernel  per-CTA totals of marked edges (= vertices the reference keeps)
//   mc_scan_kernel         one CTA: exclusive prefix of both per-CTA arrays, grand totals
//   mc_emit_verts_kernel   vertex positions in edge order (voxel-major, then axis) + edge -> vertex index
//   mc_emit_faces_kernel   faces in cell order (then case-table order) through the edge -> index map
//   mesh_sample_kernel     one warp per vertex: ATen's 3-D bilinear / nearest grid_sample, every fp32
//                          operation rounded as the CPU kernel rounds it
// Vertices sit on the same grid edges, at the same linearly interpolated positions, as any marching
// cubes; the case table is our own (tools/gen_mc_tables.py), see oracle/mc.py for the parity status.
#include <cuda_runtime.h>

#include <math.h>
#include <string.h>

#include <algorithm>

#include "saf_internal.cuh"
#include "saf_mc_tables.h"

namespace saf {
namespace {

constexpr int kMcThreads = 256;

struct McParams {
    const float* tsdf;
    const int32_t* weight;
    const float* halo_tsdf;       // optional plane x_end of the grid (the next slab's first plane), [ny*nz]
    const int32_t* halo_weight;
    int64_t n_own;                // voxels of the slab itself; voxels n_own.. are the halo plane
    int nxs, ny, nz;       // planes meshed (slab + halo plane when present), grid extents
    int x_begin;
    int64_t n;             // voxels of the slab
    uint32_t nblk;         // CTAs of the per-voxel kernels
    uint32_t* edge_idx;    // [n*3]  0 = unused, 1 = used (after classify), vertex index + 1 (after emit_verts)
    uint32_t* blk_tri;     // [nblk+1]
    uint32_t* blk_vert;    // [nblk+1]
    uint64_t* totals;      // [2] vertices, faces
    float voxel_size;
    float origin[3];
};

// corner c of a cell: offset (c & 1, (c >> 1) & 1, (c >> 2) & 1); edge e = axis*4 + j (tools/gen_mc_tables.py)
__device__ __forceinline__ int edge_base_corner(int e)
{
    const int axis = e >> 2, j = e & 3;
    const int o0 = axis == 0 ? 1 : 0, o1 = axis == 2 ? 1 : 2;
    return ((j & 1) << o0) | ((j >> 1) << o1);
}

__device__ __forceinline__ float masked_tsdf(const McParams& p, int64_t v)
{
    // clip_seem_fusion.py:826: tsdf.masked_fill(weight == 0, nan)
    if (v >= p.n_own) return __ldg(p.halo_weight + (v - p.n_own)) == 0 ? __int_as_float(0x7fc00000) : __ldg(p.halo_tsdf + (v - p.n_own));
    return __ldg(p.weight + v) == 0 ? __int_as_float(0x7fc00000) : __ldg(p.tsdf + v);
}

// case index and NaN mask of the cell whose minimum corner is voxel v = (x,y,z); false outside the cell range
__device__ __forceinline__ bool load_cell(const McParams& p, int64_t v, int x, int y, int z, uint32_t& cs, uint32_t& nan)
{
    cs = nan = 0;
    if (v >= p.n || x >= p.nxs - 1 || y >= p.ny - 1 || z >= p.nz - 1) return false;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int64_t o = (int64_t)(c & 1) * p.ny * p.nz + (int64_t)((c >> 1) & 1) * p.nz + ((c >> 2) & 1);
        const float f = masked_tsdf(p, v + o);
        cs |= (uint32_t)(!(f > 0.0f)) << c;   // NaN counts as "not above the level"
        nan |= (uint32_t)(f != f) << c;
    }
    return cs != 0 && cs != 255;
}

__device__ __forceinline__ bool triangle_survives(uint32_t cs, uint32_t nan, int t)
{
    // clip_seem_fusion.py:832: faces with a NaN vertex are dropped; a vertex is NaN iff an end of its edge is
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int e = kMcTris[cs][3 * t + k];
        const int c0 = edge_base_corner(e), c1 = c0 | (1 << (e >> 2));
        ok &= !(((nan >> c0) | (nan >> c1)) & 1u);
    }
    return ok;
}

__device__ __forceinline__ int64_t edge_slot(const McParams& p, int64_t v, int e)
{
    const int c0 = edge_base_corner(e);
    const int64_t o = (int64_t)(c0 & 1) * p.ny * p.nz + (int64_t)((c0 >> 1) & 1) * p.nz + ((c0 >> 2) & 1);
    return (v + o) * 3 + (e >> 2);
}

__device__ __forceinline__ void decode(const McParams& p, int64_t v, int& x, int& y, int& z)
{
    z = (int)(v % p.nz);
    const int64_t q = v / p.nz;
    y = (int)(q % p.ny);
    x = (int)(q / p.ny);
}

// CTA-wide exclusive prefix of one value per thread (kMcThreads threads); returns the prefix, total in `total`
__device__ __forceinline__ uint32_t cta_exclusive(uint32_t val, uint32_t* s_warp, uint32_t& total)
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
    for (int w = 0; w < kMcThreads / 32; ++w) {
        const uint32_t s = s_warp[w];
        if (w < warp) base += s;
        total += s;
    }
    __syncthreads();
    return base + inc - val;
}

__global__ void __launch_bounds__(kMcThreads) mc_classify_kernel(const McParams p)
{
    __shared__ uint32_t s_warp[kMcThreads / 32];
    const int64_t v = (int64_t)blockIdx.x * kMcThreads + threadIdx.x;
    int x, y, z;
    decode(p, v, x, y, z);
    uint32_t cs, nan, ntri = 0;
    if (load_cell(p, v, x, y, z, cs, nan)) {
        const int nt = kMcNumTris[cs];
        for (int t = 0; t < nt; ++t) {
            if (!triangle_survives(cs, nan, t)) continue;
            ++ntri;
#pragma unroll
            for (int k = 0; k < 3; ++k) p.edge_idx[edge_slot(p, v, kMcTris[cs][3 * t + k])] = 1u;
        }
    }
    uint32_t total;
    cta_exclusive(ntri, s_warp, total);
    if (threadIdx.x == 0) p.blk_tri[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kMcThreads) mc_count_verts_kernel(const McParams p)
{
    __shared__ uint32_t s_warp[kMcThreads / 32];
    const int64_t v = (int64_t)blockIdx.x * kMcThreads + threadIdx.x;
    uint32_t nv = 0;
    if (v < p.n) nv = (p.edge_idx[v * 3] != 0) + (p.edge_idx[v * 3 + 1] != 0) + (p.edge_idx[v * 3 + 2] != 0);
    uint32_t total;
    cta_exclusive(nv, s_warp, total);
    if (threadIdx.x == 0) p.blk_vert[blockIdx.x] = total;
}

// one CTA: in-place exclusive prefix of blk_tri / blk_vert, totals to p.totals
__global__ void __launch_bounds__(1024) mc_scan_kernel(const McParams p)
{
    __shared__ uint64_t s_part[1024];
    for (int which = 0; which < 2; ++which) {
        uint32_t* a = which ? p.blk_tri : p.blk_vert;
        const uint32_t per = (p.nblk + 1023u) / 1024u;
        const uint32_t lo = min(p.nblk, threadIdx.x * per), hi = min(p.nblk, lo + per);
        uint64_t sum = 0;
        for (uint32_t i = lo; i < hi; ++i) sum += a[i];
        s_part[threadIdx.x] = sum;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint64_t run = 0;
            for (int i = 0; i < 1024; ++i) {
                const uint64_t t = s_part[i];
                s_part[i] = run;
                run += t;
            }
            p.totals[which] = run;
        }
        __syncthreads();
        uint64_t run = s_part[threadIdx.x];
        for (uint32_t i = lo; i < hi; ++i) {
            const uint32_t t = a[i];
            a[i] = (uint32_t)run;
            run += t;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kMcThreads) mc_emit_verts_kernel(const McParams p, float* __restrict__ verts,
                                                                   float* __restrict__ verts_world,
                                                                   long long* __restrict__ edge_ids)
{
    __shared__ uint32_t s_warp[kMcThreads / 32];
    const int64_t v = (int64_t)blockIdx.x * kMcThreads + threadIdx.x;
    uint32_t used = 0;
    if (v < p.n) {
#pragma unroll
        for (int a = 0; a < 3; ++a) used |= (uint32_t)(p.edge_idx[v * 3 + a] != 0) << a;
    }
    uint32_t total;
    uint32_t idx = p.blk_vert[blockIdx.x] + cta_exclusive(__popc(used), s_warp, total);
    if (!used) return;
    int x, y, z;
    decode(p, v, x, y, z);
    const float f0 = masked_tsdf(p, v);
    const int64_t stride[3] = {(int64_t)p.ny * p.nz, (int64_t)p.nz, 1};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (!((used >> a) & 1u)) continue;
        const float f1 = masked_tsdf(p, v + stride[a]);
        const float t = __fdiv_rn(f0, __fsub_rn(f0, f1));   // zero crossing of the linear interpolant
        float pos[3] = {(float)(x + p.x_begin), (float)y, (float)z};
        pos[a] = __fadd_rn(pos[a], t);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            verts[(size_t)idx * 3 + k] = pos[k];
            // clip_seem_fusion.py:880: verts * voxel_size + origin
            if (verts_world) verts_world[(size_t)idx * 3 + k] = __fadd_rn(__fmul_rn(pos[k], p.voxel_size), p.origin[k]);
        }
        // identity of the vertex: its grid edge in GLOBAL numbering (slabs weld their cut-plane copies by it)
        if (edge_ids) edge_ids[idx] = ((((long long)(x + p.x_begin) * p.ny + y) * p.nz + z) * 3) + a;
        p.edge_idx[v * 3 + a] = idx + 1u;
        ++idx;
    }
}

__global__ void __launch_bounds__(kMcThreads) mc_emit_faces_kernel(const McParams p, long long* __restrict__ faces)
{
    __shared__ uint32_t s_warp[kMcThreads / 32];
    const int64_t v = (int64_t)blockIdx.x * kMcThreads + threadIdx.x;
    int x, y, z;
    decode(p, v, x, y, z);
    uint32_t cs, nan, keep = 0;
    const bool active = load_cell(p, v, x, y, z, cs, nan);
    int nt = 0;
    if (active) {
        nt = kMcNumTris[cs];
        for (int t = 0; t < nt; ++t) keep |= (uint32_t)triangle_survives(cs, nan, t) << t;
    }
    uint32_t total;
    uint32_t idx = p.blk_tri[blockIdx.x] + cta_exclusive(__popc(keep), s_warp, total);
    for (int t = 0; t < nt; ++t) {
        if (!((keep >> t) & 1u)) continue;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            faces[(size_t)idx * 3 + k] = (long long)p.edge_idx[edge_slot(p, v, kMcTris[cs][3 * t + k])] - 1ll;
        ++idx;
    }
}

// ---------------------------------------------------------------------------------------------
// vertex sampling: torch.nn.functional.grid_sample on a [1,C,nx,ny,nz] view, align_corners=False,
// zeros padding (clip_seem_fusion.py:843-877)
// ---------------------------------------------------------------------------------------------

struct SampleParams {
    const float* verts;    // [V,3] index coordinates (global x)
    int64_t n_verts;
    const float* field;    // [n, C] rows of the slab
    const float* halo_field;   // optional [ny*nz, C] rows of plane x_end (the next slab's first plane)
    int64_t n_own;
    float* out;            // [V, C]
    int C;
    int nvox[3];           // global grid
    int x_begin, x_end;    // x_end includes the halo plane when halo_field is set
    int nearest;
    int clamp01;
};

// (verts + 0.5) / nvox * 2 - 1 with nvox an int tensor on the right-hand side: torch evaluates
// reciprocal(nvox) * (verts + 0.5); then ATen's unnormalisation ((g + 1) * size - 1) / 2
__device__ __forceinline__ float source_index(float v, int size)
{
    const float n = (float)size;
    const float g = __fsub_rn(__fmul_rn(__fmul_rn(__frcp_rn(n), __fadd_rn(v, 0.5f)), 2.0f), 1.0f);
    return __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.0f), n), 1.0f), 2.0f);
}

template <int VEC>
__global__ void __launch_bounds__(256) mesh_sample_kernel(const SampleParams p)
{
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x / 32);
    const int ny = p.nvox[1], nz = p.nvox[2];
    for (int64_t i = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5); i < p.n_verts; i += nwarps) {
        float s[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) s[k] = source_index(__ldg(p.verts + i * 3 + k), p.nvox[k]);
        float* out = p.out + (size_t)i * p.C;
        int64_t row[8];
        float w[8];
        int ntap;
        if (p.nearest) {
            const int x = (int)nearbyintf(s[0]), y = (int)nearbyintf(s[1]), z = (int)nearbyintf(s[2]);
            const bool ok = x >= p.x_begin && x < p.x_end && y >= 0 && y < ny && z >= 0 && z < nz && s[0] == s[0] &&
                            s[1] == s[1] && s[2] == s[2];
            row[0] = ok ? ((int64_t)(x - p.x_begin) * ny + y) * nz + z : -1;
            w[0] = 1.0f;
            ntap = 1;
        } else {
            float lo[3], hi[3];
            int i0[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float f = floorf(s[k]);
                i0[k] = (int)f;
                hi[k] = __fsub_rn(__fadd_rn(f, 1.0f), s[k]);   // weight of the lower tap
                lo[k] = __fsub_rn(s[k], f);                    // weight of the upper tap
            }
            // ATen order tnw,tne,tsw,tse,bnw,bne,bsw,bse: D (= x) slowest, W (= z) fastest; weight = (wz*wy)*wx
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int dx = t >> 2, dy = (t >> 1) & 1, dz = t & 1;
                const int x = i0[0] + dx, y = i0[1] + dy, z = i0[2] + dz;
                w[t] = __fmul_rn(__fmul_rn(dz ? lo[2] : hi[2], dy ? lo[1] : hi[1]), dx ? lo[0] : hi[0]);
                const bool ok = x >= p.x_begin && x < p.x_end && y >= 0 && y < ny && z >= 0 && z < nz && w[t] != 0.0f;
                row[t] = ok ? ((int64_t)(x - p.x_begin) * ny + y) * nz + z : -1;
            }
            ntap = 8;
        }
        if (VEC == 4) {
            for (int col = lane; col < p.C / 4; col += 32) {
                float4 val[8];
#pragma unroll
                for (int t = 0; t < 8; ++t)
                    if (t < ntap && row[t] >= 0) val[t] = __ldg(reinterpret_cast<const float4*>(row[t] >= p.n_own ? p.halo_field + (row[t] - p.n_own) * p.C
                                                                                        : p.field + row[t] * p.C) + col);
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    if (t < ntap && row[t] >= 0) {
                        acc.x = __fadd_rn(acc.x, __fmul_rn(val[t].x, w[t]));
                        acc.y = __fadd_rn(acc.y, __fmul_rn(val[t].y, w[t]));
                        acc.z = __fadd_rn(acc.z, __fmul_rn(val[t].z, w[t]));
                        acc.w = __fadd_rn(acc.w, __fmul_rn(val[t].w, w[t]));
                    }
                }
                if (p.clamp01) {
                    acc.x = fminf(fmaxf(acc.x, 0.f), 1.f);
                    acc.y = fminf(fmaxf(acc.y, 0.f), 1.f);
                    acc.z = fminf(fmaxf(acc.z, 0.f), 1.f);
                    acc.w = fminf(fmaxf(acc.w, 0.f), 1.f);
                }
                st_stream_f4(reinterpret_cast<float4*>(out) + col, acc);
            }
        } else {
            for (int c = lane; c < p.C; c += 32) {
                float acc = 0.0f;
#pragma unroll
                for (int t = 0; t < 8; ++t)
                    if (t < ntap && row[t] >= 0) acc = __fadd_rn(acc, __fmul_rn(__ldg((row[t] >= p.n_own ? p.halo_field + (row[t] - p.n_own) * p.C
                                                                                  : p.field + row[t] * p.C) + c), w[t]));
                if (p.clamp01) acc = fminf(fmaxf(acc, 0.f), 1.f);
                out[c] = acc;
            }
        }
    }
}

struct McLayout {
    uint64_t off_totals, off_blk_tri, off_blk_vert, off_edges, bytes;
    uint32_t nblk;
    int64_t n;
};

int mc_layout(const saf_grid_desc* g, McLayout* L)
{
    if (!g) return SAF_ERR_NULL;
    if (g->nvox[0] <= 0 || g->nvox[1] <= 0 || g->nvox[2] <= 0 || g->x_begin < 0 || g->x_end > g->nvox[0] ||
        g->x_begin >= g->x_end)
        return SAF_ERR_GRID;
    if (g->x_span != 0 || g->y_ranks > 1) return SAF_ERR_UNSUPPORTED;   // cyclic layouts: contiguous slabs only
    L->n = (int64_t)(g->x_end - g->x_begin + 1) * g->nvox[1] * g->nvox[2];   // the slab plus one halo plane
    if (L->n * 3 >= (1ll << 32)) return SAF_ERR_GRID;   // edge -> vertex map is 32-bit
    L->nblk = (uint32_t)((L->n + kMcThreads - 1) / kMcThreads);
    auto up = [](uint64_t v) { return (v + 255ull) & ~255ull; };
    L->off_totals = 0;
    L->off_blk_tri = 256;
    L->off_blk_vert = up(L->off_blk_tri + 4ull * (L->nblk + 1));
    L->off_edges = up(L->off_blk_vert + 4ull * (L->nblk + 1));
    L->bytes = up(L->off_edges + 12ull * (uint64_t)L->n);
    return 0;
}

int mc_params(const saf_grid_desc* g, const float* tsdf, const int32_t* weight, const float* halo_tsdf,
              const int32_t* halo_weight, void* ws, uint64_t ws_bytes, McParams* p)
{
    McLayout L;
    int rc = mc_layout(g, &L);
    if (rc) return rc;
    if (!tsdf || !weight || !ws) return SAF_ERR_NULL;
    if ((halo_tsdf == nullptr) != (halo_weight == nullptr)) return SAF_ERR_NULL;
    if (halo_tsdf && g->x_end >= g->nvox[0]) return SAF_ERR_GRID;   // the last slab has no plane beyond it
    if (ws_bytes < L.bytes) return SAF_ERR_WORKSPACE;
    if (((uintptr_t)ws & 255u) != 0) return SAF_ERR_ALIGNMENT;
    unsigned char* base = (unsigned char*)ws;
    p->tsdf = tsdf;
    p->weight = weight;
    p->halo_tsdf = halo_tsdf;
    p->halo_weight = halo_weight;
    p->nxs = g->x_end - g->x_begin + (halo_tsdf ? 1 : 0);
    p->ny = g->nvox[1];
    p->nz = g->nvox[2];
    p->x_begin = g->x_begin;
    p->n_own = (int64_t)(g->x_end - g->x_begin) * g->nvox[1] * g->nvox[2];
    p->n = (int64_t)p->nxs * g->nvox[1] * g->nvox[2];
    p->nblk = (uint32_t)((p->n + kMcThreads - 1) / kMcThreads);
    p->totals = (uint64_t*)(base + L.off_totals);
    p->blk_tri = (uint32_t*)(base + L.off_blk_tri);
    p->blk_vert = (uint32_t*)(base + L.off_blk_vert);
    p->edge_idx = (uint32_t*)(base + L.off_edges);
    p->voxel_size = g->voxel_size;
    for (int k = 0; k < 3; ++k) p->origin[k] = g->origin[k];
    return 0;
}

}  // namespace
}  // namespace saf

using namespace saf;

extern "C" {

int saf_mesh_workspace_bytes(const saf_grid_desc* grid, uint64_t* bytes_out)
{
    if (!bytes_out) return SAF_ERR_NULL;
    McLayout L;
    int rc = mc_layout(grid, &L);
    if (rc) return rc;
    *bytes_out = L.bytes;
    return 0;
}

int saf_mesh_count(const saf_grid_desc* grid, const float* tsdf, const int32_t* weight, const float* halo_tsdf,
                   const int32_t* halo_weight, void* ws, uint64_t ws_bytes, uint64_t* n_verts_out,
                   uint64_t* n_faces_out, void* stream)
{
    int sms = 0;
    int rc = device_sm_count(&sms, nullptr);
    if (rc) return rc;
    if (!n_verts_out || !n_faces_out) return SAF_ERR_NULL;
    McParams p;
    rc = mc_params(grid, tsdf, weight, halo_tsdf, halo_weight, ws, ws_bytes, &p);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    SAF_CUDA_TRY(cudaMemsetAsync(p.edge_idx, 0, 12ull * (uint64_t)p.n, st));
    mc_classify_kernel<<<p.nblk, kMcThreads, 0, st>>>(p);
    SAF_CHECK_LAUNCH("mc_classify_kernel", st);
    mc_count_verts_kernel<<<p.nblk, kMcThreads, 0, st>>>(p);
    SAF_CHECK_LAUNCH("mc_count_verts_kernel", st);
    mc_scan_kernel<<<1, 1024, 0, st>>>(p);
    SAF_CHECK_LAUNCH("mc_scan_kernel", st);
    uint64_t totals[2];
    SAF_CUDA_TRY(cudaMemcpyAsync(totals, p.totals, sizeof(totals), cudaMemcpyDeviceToHost, st));
    SAF_CUDA_TRY(cudaStreamSynchronize(st));
    *n_verts_out = totals[0];
    *n_faces_out = totals[1];
    return 0;
}

int saf_mesh_emit(const saf_grid_desc* grid, const float* tsdf, const int32_t* weight, const float* halo_tsdf,
                  const int32_t* halo_weight, void* ws, uint64_t ws_bytes, float* verts_out, float* verts_world_out,
                  int64_t* edge_ids_out, int64_t* faces_out, void* stream)
{
    int sms = 0;
    int rc = device_sm_count(&sms, nullptr);
    if (rc) return rc;
    if (!verts_out || !faces_out) return SAF_ERR_NULL;
    McParams p;
    rc = mc_params(grid, tsdf, weight, halo_tsdf, halo_weight, ws, ws_bytes, &p);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    mc_emit_verts_kernel<<<p.nblk, kMcThreads, 0, st>>>(p, verts_out, verts_world_out, (long long*)edge_ids_out);
    SAF_CHECK_LAUNCH("mc_emit_verts_kernel", st);
    mc_emit_faces_kernel<<<p.nblk, kMcThreads, 0, st>>>(p, (long long*)faces_out);
    SAF_CHECK_LAUNCH("mc_emit_faces_kernel", st);
    return 0;
}

int saf_mesh_sample(const saf_grid_desc* grid, const float* verts, int64_t n_verts, const float* field,
                    const float* halo_field, int32_t channels, int32_t mode, int32_t clamp01, float* out, void* stream)
{
    int sms = 0;
    int rc = device_sm_count(&sms, nullptr);
    if (rc) return rc;
    McLayout L;
    rc = mc_layout(grid, &L);
    if (rc) return rc;
    if (n_verts < 0 || channels <= 0) return SAF_ERR_SHAPE;
    if (mode != SAF_SAMPLE_TRILINEAR && mode != SAF_SAMPLE_NEAREST) return SAF_ERR_UNSUPPORTED;
    if (n_verts == 0) return 0;
    if (!verts || !field || !out) return SAF_ERR_NULL;
    SampleParams p;
    p.verts = verts;
    p.n_verts = n_verts;
    p.field = field;
    p.halo_field = halo_field;
    p.n_own = (int64_t)(grid->x_end - grid->x_begin) * grid->nvox[1] * grid->nvox[2];
    p.out = out;
    p.C = channels;
    for (int k = 0; k < 3; ++k) p.nvox[k] = grid->nvox[k];
    p.x_begin = grid->x_begin;
    p.x_end = grid->x_end + (halo_field ? 1 : 0);
    p.nearest = mode == SAF_SAMPLE_NEAREST;
    p.clamp01 = clamp01;
    const int64_t want = (n_verts + 7) / 8;
    const int blocks = (int)std::min<int64_t>(want, (int64_t)sms * 8);
    const bool vec4 = (channels % 4 == 0) && (((uintptr_t)field & 15u) == 0) && (((uintptr_t)out & 15u) == 0) &&
                      (((uintptr_t)halo_field & 15u) == 0);
    cudaStream_t st = (cudaStream_t)stream;
    if (vec4)
        mesh_sample_kernel<4><<<blocks, 256, 0, st>>>(p);
    else
        mesh_sample_kernel<1><<<blocks, 256, 0, st>>>(p);
    SAF_CHECK_LAUNCH("mesh_sample_kernel", st);
    return 0;
}

}  // extern "C"
