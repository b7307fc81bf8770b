// saf_fusion.cu -- per-frame RGB-D integration on sm_100a.
//
// Replaces ClipSeemFusion.integrate (/root/reference/clip_seem_fusion.py:676-822) and
// ClipFusion.integrate (/root/reference/clipfusion.py:627-721).  Three kernels per call:
//   K1 frame_setup_kernel      conservative frustum test of 8^3 voxel blocks -> compact block list
//                              (+ repack of channel-major feature images to [R,C] rows)
//   K2 tsdf_update_kernel      exact per-voxel projection / depth sample / masks / TSDF running
//                              average for the listed blocks; emits the list of `valid` voxels
//   K3 feature_accumulate_*    one warp per listed voxel: 128-bit streaming read-modify-write of
//                              the C-float feature row, table rows read from shared memory
//                              (staged by TMA bulk copies), rgb + label counter + weight
// All decisions that feed masks use explicitly rounded intrinsics in the reference's op order;
// this file is compiled with -fmad=false so nothing is contracted behind our back.
#include <limits.h>
#include <math.h>
#include <algorithm>
#include <string.h>

#include "saf_internal.cuh"

namespace saf {

struct FusionParams {
    saf_grid_desc grid;
    saf_volume vol;
    saf_frame frames[SAF_MAX_BATCH];
    int32_t batch, H, W;
    float trunc;
    int32_t rgb_mode;
    int32_t frame_index;       // K3: which frame's list
    uint32_t pack_mask;        // frames whose feature image is repacked into the workspace
    uint32_t nb[3];
    uint32_t nblocks_total;
    uint32_t nxs;              // slab extent in x
    uint64_t nslab;
    uint64_t list_cap;
    uint64_t max_table_elems;
    WsHeader* hdr;
    uint32_t* block_list;
    ValidEntry* lists;
    float* tables;
    uint8_t* valid_out;
    uint8_t* tsdf_valid_out;
};

// ---------------------------------------------------------------------------------------------
// exact geometry (clip_seem_fusion.py:664-669, 698-712)
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ float voxel_centre(int idx, float vs, float o)
{
    return __fadd_rn(__fmul_rn(__int2float_rn(idx), vs), o);
}

// Pose (first three rows) and intrinsics of one frame, from the by-value copy in the launch
// parameters or from the caller's device tensors (saf_frame.pose_device / K_device).
struct Geom {
    float P[12];
    float K[9];
};

__device__ __forceinline__ void load_geom(const saf_frame& f, Geom& g)
{
    if (f.pose_device) {
#pragma unroll
        for (int i = 0; i < 12; ++i) g.P[i] = __ldg(f.pose_device + i);
#pragma unroll
        for (int i = 0; i < 9; ++i) g.K[i] = __ldg(f.K_device + i);
    } else {
#pragma unroll
        for (int i = 0; i < 12; ++i) g.P[i] = f.pose[i];
#pragma unroll
        for (int i = 0; i < 9; ++i) g.K[i] = f.K[i];
    }
}

// bmm with K=3 is the left-to-right fma chain (multi-threaded MKL and cuBLAS alike).
__device__ __forceinline__ void project(const float* __restrict__ P, const float* __restrict__ K, float xw, float yw,
                                        float zw, float fW, float fH, float& gx, float& gy, float& z)
{
    const float d0 = __fsub_rn(xw, P[3]), d1 = __fsub_rn(yw, P[7]), d2 = __fsub_rn(zw, P[11]);
    float xc[3], q[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) xc[k] = __fmaf_rn(P[8 + k], d2, __fmaf_rn(P[4 + k], d1, __fmul_rn(P[k], d0)));
#pragma unroll
    for (int k = 0; k < 3; ++k)
        q[k] = __fmaf_rn(K[3 * k + 2], xc[2], __fmaf_rn(K[3 * k + 1], xc[1], __fmul_rn(K[3 * k], xc[0])));
    z = q[2];
    const float u = __fdiv_rn(q[0], z), v = __fdiv_rn(q[1], z);
    gx = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn(u, 0.5f), fW), 2.0f), 1.0f);
    gy = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn(v, 0.5f), fH), 2.0f), 1.0f);
}

// ATen grid_sampler_unnormalize (align_corners=False), evaluated with one rounding after the add.
__device__ __forceinline__ float unnormalize(float g, int size)
{
    return __fmaf_rn(__fadd_rn(g, 1.0f), __fmul_rn(__int2float_rn(size), 0.5f), -0.5f);
}

// nearest tap (round half to even) or -1 when outside the image / NaN (zeros padding)
__device__ __forceinline__ int nearest_index(float g, int size)
{
    const float r = rintf(unnormalize(g, size));
    return (r >= 0.0f && r < __int2float_rn(size)) ? __float2int_rz(r) : -1;
}

struct Taps {
    int idx[4];   // flat y*W+x or -1 (dropped), order nw ne sw se
    float w[4];
};

__device__ __forceinline__ void bilinear_setup(float gx, float gy, int W, int H, Taps& t)
{
    const float x = unnormalize(gx, W), y = unnormalize(gy, H);
    const float xw = floorf(x), yn = floorf(y);
    const float w = __fsub_rn(x, xw), e = __fsub_rn(1.0f, w);
    const float n = __fsub_rn(y, yn), s = __fsub_rn(1.0f, n);
    t.w[0] = __fmul_rn(s, e);
    t.w[1] = __fmul_rn(s, w);
    t.w[2] = __fmul_rn(n, e);
    t.w[3] = __fmul_rn(n, w);
    const bool ok = (xw >= -2.0f) && (xw <= __int2float_rn(W) + 1.0f) && (yn >= -2.0f) && (yn <= __int2float_rn(H) + 1.0f);
    const int ixw = ok ? __float2int_rz(xw) : -2, iyn = ok ? __float2int_rz(yn) : -2;
    const int ixe = ixw + 1, iys = iyn + 1;
    const bool wm = ixw > -1 && ixw < W, em = ixe > -1 && ixe < W;
    const bool nm = iyn > -1 && iyn < H, sm = iys > -1 && iys < H;
    t.idx[0] = (nm && wm) ? iyn * W + ixw : -1;
    t.idx[1] = (nm && em) ? iyn * W + ixe : -1;
    t.idx[2] = (sm && wm) ? iys * W + ixw : -1;
    t.idx[3] = (sm && em) ? iys * W + ixe : -1;
}

__device__ __forceinline__ float bilinear_mix(float v0, float v1, float v2, float v3, const float (&w)[4])
{
    return __fmaf_rn(v3, w[3], __fmaf_rn(v2, w[2], __fmaf_rn(v1, w[1], __fmul_rn(v0, w[0]))));
}

__device__ __forceinline__ float load_class_id(const void* seg, int dtype, int pix)
{
    switch (dtype) {
        case SAF_SEG_U8: return (float)((const uint8_t*)seg)[pix];
        case SAF_SEG_I16: return (float)((const int16_t*)seg)[pix];
        case SAF_SEG_I32: return (float)((const int32_t*)seg)[pix];
        case SAF_SEG_I64: return (float)((const long long*)seg)[pix];
        default: return ((const float*)seg)[pix];
    }
}

// ---------------------------------------------------------------------------------------------
// K1: frame set-up (frustum cull of voxel blocks + feature-image repack)
// ---------------------------------------------------------------------------------------------

constexpr int kK1Threads = 256;

__device__ __forceinline__ bool block_maybe_visible(const Geom& g, float cx, float cy, float cz, float r, float fW,
                                                    float fH)
{
    const float* P = g.P;
    const float* K = g.K;
    const float d0 = cx - P[3], d1 = cy - P[7], d2 = cz - P[11];
    float pc[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) pc[k] = P[k] * d0 + P[4 + k] * d1 + P[8 + k] * d2;
    const float plen = sqrtf(pc[0] * pc[0] + pc[1] * pc[1] + pc[2] * pc[2]) + r;
    // half-spaces n.p >= 0 that every voxel with `_valid` satisfies (clip_seem_fusion.py:709-726):
    //   z > 0;  -0.5 <= u <= W-0.5;  -0.5 <= v <= H-0.5   with (u,v,1) z = K p
    float n[5][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        n[0][k] = K[6 + k];
        n[1][k] = K[k] + 0.5f * K[6 + k];
        n[2][k] = (fW - 0.5f) * K[6 + k] - K[k];
        n[3][k] = K[3 + k] + 0.5f * K[6 + k];
        n[4][k] = (fH - 0.5f) * K[6 + k] - K[3 + k];
    }
    bool cull = false;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const float dot = n[i][0] * pc[0] + n[i][1] * pc[1] + n[i][2] * pc[2];
        const float nlen = sqrtf(n[i][0] * n[i][0] + n[i][1] * n[i][1] + n[i][2] * n[i][2]);
        // largest value of n.p over the block's bounding sphere, plus slack for fp32 rounding in
        // K2's exact evaluation (1e-3 relative is ~4 orders above it).  NaN compares false -> kept.
        cull |= (dot + nlen * r + 1e-3f * nlen * plen + 1e-6f) < 0.0f;
    }
    return !cull;
}

__global__ void __launch_bounds__(kK1Threads) frame_setup_kernel(const FusionParams p, uint32_t cull_ctas)
{
    if (blockIdx.x >= cull_ctas) {
        // repack channel-major feature images into [R,C] rows (only frames in pack_mask)
        const uint32_t pack_ctas = gridDim.x - cull_ctas;
        const int C = p.vol.feature_dim;
        for (int b = 0; b < p.batch; ++b) {
            if (!((p.pack_mask >> b) & 1u)) continue;
            const saf_frame& f = p.frames[b];
            const int64_t R = (int64_t)f.npy * f.npx;
            float* dst = p.tables + (uint64_t)b * p.max_table_elems;
            for (int64_t e = (int64_t)(blockIdx.x - cull_ctas) * kK1Threads + threadIdx.x; e < R * C;
                 e += (int64_t)pack_ctas * kK1Threads) {
                const int64_t r = e / C, c = e - r * C;
                dst[e] = f.table[c * f.table_stride_c + r * f.table_stride_r];
            }
        }
        return;
    }
    const uint32_t blk = blockIdx.x * kK1Threads + threadIdx.x;
    if (blk == 0) {
#pragma unroll
        for (int b = 0; b < SAF_MAX_BATCH; ++b) {
            p.hdr->n_valid[b] = 0;
            p.hdr->n_tsdf_valid[b] = 0;
        }
    }
    bool vis = false;
    if (blk < p.nblocks_total) {
        const uint32_t bz = blk % p.nb[2];
        const uint32_t by = (blk / p.nb[2]) % p.nb[1];
        const uint32_t bx = blk / (p.nb[2] * p.nb[1]);
        const int x0 = (int)bx * kBlockEdge, y0 = (int)by * kBlockEdge, z0 = (int)bz * kBlockEdge;
        const int ex = min(kBlockEdge, (int)p.nxs - x0), ey = min(kBlockEdge, p.grid.nvox[1] - y0),
                  ez = min(kBlockEdge, p.grid.nvox[2] - z0);
        const float vs = p.grid.voxel_size;
        const float hx = 0.5f * vs * (float)(ex - 1), hy = 0.5f * vs * (float)(ey - 1), hz = 0.5f * vs * (float)(ez - 1);
        const float cx = p.grid.origin[0] + vs * (float)(p.grid.x_begin + x0) + hx;
        const float cy = p.grid.origin[1] + vs * (float)y0 + hy;
        const float cz = p.grid.origin[2] + vs * (float)z0 + hz;
        const float r = sqrtf(hx * hx + hy * hy + hz * hz) + 0.01f * vs;
        const float fW = (float)p.W, fH = (float)p.H;
        for (int b = 0; b < p.batch; ++b) {
            Geom g;
            load_geom(p.frames[b], g);
            vis |= block_maybe_visible(g, cx, cy, cz, r, fW, fH);
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, vis);
    if (m) {
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(m) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&p.hdr->n_blocks, (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (vis) p.block_list[base + __popc(m & ((1u << lane) - 1u))] = blk;
    }
}

// ---------------------------------------------------------------------------------------------
// K2: TSDF update over the visible blocks (clip_seem_fusion.py:698-744)
// ---------------------------------------------------------------------------------------------

constexpr int kK2Threads = 128;

template <int BATCH1>
__global__ void __launch_bounds__(kK2Threads) tsdf_update_kernel(const FusionParams p)
{
    WsHeader* hdr = p.hdr;
    const uint32_t n_blocks = hdr->n_blocks;
    const int lane = threadIdx.x & 31;
    const int B = BATCH1 ? 1 : p.batch;
    const float fW = (float)p.W, fH = (float)p.H;
    const int ny = p.grid.nvox[1], nz = p.grid.nvox[2];
    uint32_t tv_count[BATCH1 ? 1 : SAF_MAX_BATCH];
#pragma unroll
    for (int b = 0; b < (BATCH1 ? 1 : SAF_MAX_BATCH); ++b) tv_count[b] = 0;
    Geom g;
    if (BATCH1) load_geom(p.frames[0], g);

    for (uint32_t bi = blockIdx.x; bi < n_blocks; bi += gridDim.x) {
        const uint32_t blk = p.block_list[bi];
        const uint32_t bz = blk % p.nb[2];
        const uint32_t by = (blk / p.nb[2]) % p.nb[1];
        const uint32_t bx = blk / (p.nb[2] * p.nb[1]);
#pragma unroll
        for (int j = 0; j < kBlockVoxels / kK2Threads; ++j) {
            const int local = threadIdx.x + j * kK2Threads;
            const int lx = (int)bx * kBlockEdge + (local >> 6);          // slab-local x
            const int iy = (int)by * kBlockEdge + ((local >> 3) & 7);
            const int iz = (int)bz * kBlockEdge + (local & 7);
            const bool inside = lx < (int)p.nxs && iy < ny && iz < nz;
            const uint32_t v = inside ? (uint32_t)(((uint64_t)lx * ny + iy) * nz + iz) : 0u;
            const float xw = voxel_centre(lx + p.grid.x_begin, p.grid.voxel_size, p.grid.origin[0]);
            const float yw = voxel_centre(iy, p.grid.voxel_size, p.grid.origin[1]);
            const float zw = voxel_centre(iz, p.grid.voxel_size, p.grid.origin[2]);
            int bw = 0;
            float bt = 0.0f;
            for (int b = 0; b < B; ++b) {
                const saf_frame& f = p.frames[b];
                if (!BATCH1) load_geom(f, g);
                float gx, gy, z;
                project(g.P, g.K, xw, yw, zw, fW, fH, gx, gy, z);
                const int px = nearest_index(gx, p.W), py = nearest_index(gy, p.H);
                const float d = (inside && px >= 0 && py >= 0) ? __ldg(f.depth + (size_t)py * p.W + px) : 0.0f;
                const float sdf = __fdiv_rn(__fsub_rn(d, z), p.trunc);
                const bool in_view = inside && (fabsf(gx) <= 1.0f) && (fabsf(gy) <= 1.0f) && (z > 0.0f);
                const bool valid = in_view && (fabsf(sdf) <= 1.0f);
                const bool tv = in_view && (sdf > -1.0f);
                if (tv) {
                    bw += 1;
                    bt = __fadd_rn(bt, fminf(fmaxf(sdf, -1.0f), 1.0f));
                    tv_count[BATCH1 ? 0 : b] += 1;
                }
                if (p.valid_out && inside) {
                    if (valid) p.valid_out[(uint64_t)b * p.nslab + v] = 1;
                    if (tv) p.tsdf_valid_out[(uint64_t)b * p.nslab + v] = 1;
                }
                // append (voxel, gx, gy) to frame b's list: one atomic per warp
                const unsigned m = __ballot_sync(0xffffffffu, valid);
                if (m) {
                    const int leader = __ffs(m) - 1;
                    uint32_t base = 0;
                    if (lane == leader) base = atomicAdd(&hdr->n_valid[b], (uint32_t)__popc(m));
                    base = __shfl_sync(0xffffffffu, base, leader);
                    if (valid) {
                        ValidEntry e;
                        e.voxel = v;
                        e.gx = gx;
                        e.gy = gy;
                        e.pad = 0;
                        p.lists[(uint64_t)b * p.list_cap + base + __popc(m & ((1u << lane) - 1u))] = e;
                    }
                }
            }
            if (bw > 0) {
                // clip_seem_fusion.py:736-744: three separately rounded fp32 ops
                const int tw = p.vol.tsdf_weight[v];
                const int nw = tw + bw;
                const float fnw = __int2float_rn(nw);
                const float t_old = p.vol.tsdf[v];
                p.vol.tsdf[v] = __fadd_rn(__fdiv_rn(bt, fnw), __fmul_rn(t_old, __fdiv_rn(__int2float_rn(tw), fnw)));
                p.vol.tsdf_weight[v] = nw;
            }
        }
    }
    // per-frame tsdf_valid totals: one atomic per warp
#pragma unroll
    for (int b = 0; b < (BATCH1 ? 1 : SAF_MAX_BATCH); ++b) {
        if (b >= B) break;
        uint32_t c = tv_count[b];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0 && c) atomicAdd(&hdr->n_tsdf_valid[b], c);
    }
    // last CTA folds this call into the totals and re-arms the block counter for the next K1
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&hdr->k2_done, 1u) == gridDim.x - 1);
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        __threadfence();
        unsigned long long sv = 0, stv = 0;
        for (int b = 0; b < B; ++b) {
            sv += atomicAdd(&hdr->n_valid[b], 0u);
            stv += atomicAdd(&hdr->n_tsdf_valid[b], 0u);
        }
        hdr->total_valid += sv;
        hdr->total_tsdf_valid += stv;
        hdr->total_blocks += n_blocks;
        hdr->total_frames += (unsigned long long)B;
        hdr->last_blocks = n_blocks;
        hdr->n_blocks = 0;
        hdr->k2_done = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// K3: per-voxel feature / rgb / label accumulation (clip_seem_fusion.py:751-822)
// ---------------------------------------------------------------------------------------------

constexpr int kK3Threads = 512;

struct VoxelScalars {
    float a, b;        // a = 1/(w+1), b = w * a
    int w;
};

// rgb (lanes 0-2), label counter (lane 3) and weight (lane 4) of one voxel
__device__ __forceinline__ void update_small_state(const FusionParams& p, const saf_frame& f, const ValidEntry& e,
                                                   const VoxelScalars& s, int lane)
{
    if (lane < 3) {
        float smp;
        if (p.rgb_mode == SAF_RGB_NEAREST) {
            const int px = nearest_index(e.gx, p.W), py = nearest_index(e.gy, p.H);
            smp = (px >= 0 && py >= 0) ? __ldg(f.rgb + ((size_t)py * p.W + px) * 3 + lane) : 0.0f;
        } else {
            Taps t;
            bilinear_setup(e.gx, e.gy, p.W, p.H, t);
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = t.idx[k] >= 0 ? __ldg(f.rgb + (size_t)t.idx[k] * 3 + lane) : 0.0f;
            smp = bilinear_mix(v[0], v[1], v[2], v[3], t.w);
        }
        float* dst = p.vol.rgb + (size_t)e.voxel * 3 + lane;
        *dst = __fadd_rn(__fmul_rn(smp, s.a), __fmul_rn(*dst, s.b));
    } else if (lane == 3) {
        if (p.vol.labels_one_hot && f.seg) {
            const int px = nearest_index(e.gx, p.W), py = nearest_index(e.gy, p.H);
            const float lf = (px >= 0 && py >= 0) ? load_class_id(f.seg, f.seg_dtype, py * p.W + px) : 0.0f;
            const long long id = (long long)lf;
            if (id >= 0 && id < p.vol.n_classes)
                p.vol.labels_one_hot[(size_t)e.voxel * p.vol.n_classes + id] += 1;
            else
                atomicOr(&p.hdr->error_flags, SAF_FLAG_BAD_CLASS_ID);
        }
    } else if (lane == 4) {
        p.vol.weight[e.voxel] = s.w + 1;
    }
}

__device__ __forceinline__ float4 mix4(const float4& t0, const float4& t1, const float4& t2, const float4& t3,
                                       const float (&w)[4])
{
    float4 r;
    r.x = bilinear_mix(t0.x, t1.x, t2.x, t3.x, w);
    r.y = bilinear_mix(t0.y, t1.y, t2.y, t3.y, w);
    r.z = bilinear_mix(t0.z, t1.z, t2.z, t3.z, w);
    r.w = bilinear_mix(t0.w, t1.w, t2.w, t3.w, w);
    return r;
}

__device__ __forceinline__ float4 blend4(const float4& smp, const float4& old, float a, float b)
{
    float4 r;
    r.x = __fadd_rn(__fmul_rn(smp.x, a), __fmul_rn(old.x, b));
    r.y = __fadd_rn(__fmul_rn(smp.y, a), __fmul_rn(old.y, b));
    r.z = __fadd_rn(__fmul_rn(smp.z, a), __fmul_rn(old.z, b));
    r.w = __fadd_rn(__fmul_rn(smp.w, a), __fmul_rn(old.w, b));
    return r;
}

// Stage the [R,C] table into shared memory: one TMA bulk copy per row when 16-byte rules allow,
// plain loads otherwise.  Returns after the copies are ISSUED; wait with mbar_wait(bar, 0).
__device__ __forceinline__ void stage_table(float* tab, const float* src, int64_t src_stride_r, int R, int C,
                                            uint64_t* bar, bool use_tma)
{
    if (use_tma) {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            if (threadIdx.x == 0) mbar_arrive_expect_tx(bar, (uint32_t)R * (uint32_t)C * 4u);
            __syncwarp();
            for (int r = threadIdx.x; r < R; r += 32)
                tma_bulk_g2s(tab + (size_t)r * C, src + (size_t)r * src_stride_r, (uint32_t)C * 4u, bar);
        }
    } else {
        for (int64_t e = threadIdx.x; e < (int64_t)R * C; e += blockDim.x) {
            const int64_t r = e / C, c = e - r * C;
            tab[e] = src[r * src_stride_r + c];
        }
        __syncthreads();
    }
}

// CHUNKS = C/128 float4 per lane (compile time), VPW voxels in flight per warp.
template <int CHUNKS, int VPW, bool TABLE_SMEM>
__global__ void __launch_bounds__(kK3Threads, 1)
feature_accumulate_kernel(const FusionParams p, const float* __restrict__ table, int64_t table_stride_r, int use_tma)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    float* tab = reinterpret_cast<float*>(smem_raw);
    constexpr int C = CHUNKS * 128;
    const saf_frame& f = p.frames[p.frame_index];
    const int R = f.npy * f.npx;
    const int lane = threadIdx.x & 31;
    const uint32_t n = p.hdr->n_valid[p.frame_index];
    if (n == 0) return;

    if (TABLE_SMEM) stage_table(tab, table, table_stride_r, R, C, &bar, use_tma != 0);

    const uint32_t nwarps = gridDim.x * (kK3Threads / 32);
    const uint32_t gwarp = blockIdx.x * (kK3Threads / 32) + (threadIdx.x >> 5);
    const ValidEntry* __restrict__ list = p.lists + (uint64_t)p.frame_index * p.list_cap;
    const float4* tab4 = reinterpret_cast<const float4*>(TABLE_SMEM ? tab : table);
    const int64_t tab_row4 = TABLE_SMEM ? (C / 4) : (table_stride_r / 4);

    bool table_ready = !(TABLE_SMEM && use_tma);
    for (uint64_t i0 = gwarp; i0 < n; i0 += (uint64_t)nwarps * VPW) {
        ValidEntry e[VPW];
        float4 old[VPW][CHUNKS];
        VoxelScalars sc[VPW];
        bool act[VPW];
#pragma unroll
        for (int v = 0; v < VPW; ++v) {
            const uint64_t i = i0 + (uint64_t)v * nwarps;
            act[v] = i < n;
            if (act[v]) e[v] = list[i];
        }
#pragma unroll
        for (int v = 0; v < VPW; ++v) {
            if (!act[v]) continue;
            const float4* row = reinterpret_cast<const float4*>(p.vol.clip_feat + (size_t)e[v].voxel * C);
#pragma unroll
            for (int j = 0; j < CHUNKS; ++j) old[v][j] = ld_stream_f4(row + j * 32 + lane);
            sc[v].w = p.vol.weight[e[v].voxel];
        }
        if (!table_ready) {
            mbar_wait(&bar, 0);
            table_ready = true;
        }
#pragma unroll
        for (int v = 0; v < VPW; ++v) {
            if (!act[v]) continue;
            // clip_seem_fusion.py:808-810
            sc[v].a = __frcp_rn(__int2float_rn(sc[v].w + 1));
            sc[v].b = __fmul_rn(__int2float_rn(sc[v].w), sc[v].a);
            Taps t;
            bilinear_setup(e[v].gx, e[v].gy, f.npx, f.npy, t);
            float4* row = reinterpret_cast<float4*>(p.vol.clip_feat + (size_t)e[v].voxel * C);
            const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < CHUNKS; ++j) {
                const int col = j * 32 + lane;
                const float4 t0 = t.idx[0] >= 0 ? tab4[t.idx[0] * tab_row4 + col] : zero;
                const float4 t1 = t.idx[1] >= 0 ? tab4[t.idx[1] * tab_row4 + col] : zero;
                const float4 t2 = t.idx[2] >= 0 ? tab4[t.idx[2] * tab_row4 + col] : zero;
                const float4 t3 = t.idx[3] >= 0 ? tab4[t.idx[3] * tab_row4 + col] : zero;
                st_stream_f4(row + col, blend4(mix4(t0, t1, t2, t3, t.w), old[v][j], sc[v].a, sc[v].b));
            }
            update_small_state(p, f, e[v], sc[v], lane);
        }
    }
    if (!table_ready) mbar_wait(&bar, 0);  // never leave a bulk copy in flight at exit
}

// Any feature_dim: VEC = 4 (C % 4 == 0, 16-byte aligned rows) or 1.  Table rows read from global.
template <int VEC>
__global__ void __launch_bounds__(kK3Threads) feature_accumulate_generic_kernel(const FusionParams p,
                                                                                const float* __restrict__ table,
                                                                                int64_t table_stride_r)
{
    const saf_frame& f = p.frames[p.frame_index];
    const int C = p.vol.feature_dim;
    const int lane = threadIdx.x & 31;
    const uint32_t n = p.hdr->n_valid[p.frame_index];
    const uint32_t nwarps = gridDim.x * (kK3Threads / 32);
    const uint32_t gwarp = blockIdx.x * (kK3Threads / 32) + (threadIdx.x >> 5);
    const ValidEntry* __restrict__ list = p.lists + (uint64_t)p.frame_index * p.list_cap;
    for (uint64_t i = gwarp; i < n; i += nwarps) {
        const ValidEntry e = list[i];
        VoxelScalars sc;
        sc.w = p.vol.weight[e.voxel];
        sc.a = __frcp_rn(__int2float_rn(sc.w + 1));
        sc.b = __fmul_rn(__int2float_rn(sc.w), sc.a);
        Taps t;
        bilinear_setup(e.gx, e.gy, f.npx, f.npy, t);
        float* row = p.vol.clip_feat + (size_t)e.voxel * C;
        if (VEC == 4) {
            const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int col = lane; col < C / 4; col += 32) {
                const float4 o = ld_stream_f4(reinterpret_cast<const float4*>(row) + col);
                float4 tv[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tv[k] = t.idx[k] >= 0 ? __ldg(reinterpret_cast<const float4*>(table + t.idx[k] * table_stride_r) + col)
                                          : zero;
                st_stream_f4(reinterpret_cast<float4*>(row) + col, blend4(mix4(tv[0], tv[1], tv[2], tv[3], t.w), o, sc.a, sc.b));
            }
        } else {
            for (int c = lane; c < C; c += 32) {
                float tv[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) tv[k] = t.idx[k] >= 0 ? __ldg(table + t.idx[k] * table_stride_r + c) : 0.0f;
                const float smp = bilinear_mix(tv[0], tv[1], tv[2], tv[3], t.w);
                row[c] = __fadd_rn(__fmul_rn(smp, sc.a), __fmul_rn(row[c], sc.b));
            }
        }
        update_small_state(p, f, e, sc, lane);
    }
}

// ---------------------------------------------------------------------------------------------
// label argmax (clip_seem_fusion.py:315-325)
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) label_argmax_kernel(const int32_t* __restrict__ labels, int64_t n, int n_classes,
                                                           long long* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x / 32);
    for (int64_t v = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5); v < n; v += nwarps) {
        const int32_t* row = labels + v * n_classes;
        int best_val = INT_MIN, best_idx = INT_MAX;
        bool any = false;
        for (int c = lane; c < n_classes; c += 32) {
            const int x = __ldg(row + c);
            any |= (x != 0);
            if (x > best_val) {  // strictly greater keeps the first maximum within a lane
                best_val = x;
                best_idx = c;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int ov = __shfl_xor_sync(0xffffffffu, best_val, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_idx, o);
            if (ov > best_val || (ov == best_val && oi < best_idx)) {
                best_val = ov;
                best_idx = oi;
            }
        }
        any = __any_sync(0xffffffffu, any);
        if (lane == 0) out[v] = any ? (long long)best_idx : -1ll;
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

int device_sm_count(int* sms, int* smem_optin)
{
    int dev = 0;
    SAF_CUDA_TRY(cudaGetDevice(&dev));
    int major = 0;
    SAF_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) return SAF_ERR_DEVICE;
    SAF_CUDA_TRY(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
    if (smem_optin) SAF_CUDA_TRY(cudaDeviceGetAttribute(smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    return 0;
}

static int build_params(const saf_grid_desc* grid, const saf_volume* vol, const saf_frame* frames, int32_t batch,
                        int32_t H, int32_t W, float trunc, int32_t rgb_mode, const saf_workspace* ws, FusionParams* p)
{
    if (!grid || !frames || !ws || !ws->base) return SAF_ERR_NULL;
    if (batch < 1 || batch > SAF_MAX_BATCH || batch > ws->max_batch) return SAF_ERR_BATCH;
    if (H <= 0 || W <= 0 || (int64_t)H * W >= (1ll << 31) / 3) return SAF_ERR_SHAPE;
    if (((uintptr_t)ws->base & 255u) != 0) return SAF_ERR_ALIGNMENT;
    WsLayout L;
    int rc = compute_layout(grid, ws->max_batch, ws->max_table_elems, &L);
    if (rc) return rc;
    if (ws->bytes < L.bytes) return SAF_ERR_WORKSPACE;
    memset(p, 0, sizeof(*p));
    p->grid = *grid;
    if (vol) p->vol = *vol;
    for (int b = 0; b < batch; ++b) {
        p->frames[b] = frames[b];
        if (!frames[b].depth) return SAF_ERR_NULL;
        if ((frames[b].pose_device == nullptr) != (frames[b].K_device == nullptr)) return SAF_ERR_NULL;
    }
    p->batch = batch;
    p->H = H;
    p->W = W;
    p->trunc = trunc;
    p->rgb_mode = rgb_mode;
    p->nb[0] = L.nb[0];
    p->nb[1] = L.nb[1];
    p->nb[2] = L.nb[2];
    p->nblocks_total = L.nblocks_total;
    p->nxs = (uint32_t)(grid->x_end - grid->x_begin);
    p->nslab = L.list_cap;
    p->list_cap = L.list_cap;
    p->max_table_elems = (uint64_t)ws->max_table_elems;
    unsigned char* base = (unsigned char*)ws->base;
    p->hdr = (WsHeader*)base;
    p->block_list = (uint32_t*)(base + L.off_blocks);
    p->lists = (ValidEntry*)(base + L.off_lists);
    p->tables = (float*)(base + L.off_tables);
    return 0;
}

static int check_feature_args(const saf_volume* vol, const saf_frame* frames, int32_t batch, int32_t rgb_mode,
                              const saf_workspace* ws, uint32_t* pack_mask)
{
    if (!vol || !vol->tsdf || !vol->tsdf_weight || !vol->weight || !vol->rgb || !vol->clip_feat) return SAF_ERR_NULL;
    if (vol->feature_dim <= 0) return SAF_ERR_SHAPE;
    if (vol->labels_one_hot && vol->n_classes <= 0) return SAF_ERR_SHAPE;
    if (rgb_mode != SAF_RGB_NEAREST && rgb_mode != SAF_RGB_BILINEAR) return SAF_ERR_UNSUPPORTED;
    *pack_mask = 0;
    for (int b = 0; b < batch; ++b) {
        const saf_frame& f = frames[b];
        if (!f.rgb || !f.table) return SAF_ERR_NULL;
        if (f.npy <= 0 || f.npx <= 0) return SAF_ERR_SHAPE;
        if (f.seg && (f.seg_dtype < SAF_SEG_U8 || f.seg_dtype > SAF_SEG_F32)) return SAF_ERR_DTYPE;
        const int64_t elems = (int64_t)f.npy * f.npx * vol->feature_dim;
        if (f.table_stride_c != 1) {
            if (elems > ws->max_table_elems) return SAF_ERR_WORKSPACE;
            *pack_mask |= 1u << b;
        }
    }
    return 0;
}

static int launch_k1(const FusionParams& p, cudaStream_t st)
{
    const uint32_t cull_ctas = (p.nblocks_total + kK1Threads - 1) / kK1Threads;
    const uint32_t pack_ctas = p.pack_mask ? 64u : 0u;
    frame_setup_kernel<<<cull_ctas + pack_ctas, kK1Threads, 0, st>>>(p, cull_ctas);
    return (int)cudaGetLastError();
}

static int launch_k2(const FusionParams& p, int sms, cudaStream_t st)
{
    const uint32_t grid = (uint32_t)min((uint64_t)p.nblocks_total, (uint64_t)sms * 16u);
    if (p.batch == 1)
        tsdf_update_kernel<1><<<grid, kK2Threads, 0, st>>>(p);
    else
        tsdf_update_kernel<0><<<grid, kK2Threads, 0, st>>>(p);
    return (int)cudaGetLastError();
}

template <int CHUNKS, int VPW, bool SMEM>
static int launch_k3_fixed(const FusionParams& p, const float* table, int64_t stride_r, int use_tma, size_t smem,
                           int sms, cudaStream_t st)
{
    auto kern = feature_accumulate_kernel<CHUNKS, VPW, SMEM>;
    if (smem > 48 * 1024) SAF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<sms, kK3Threads, smem, st>>>(p, table, stride_r, use_tma);
    return (int)cudaGetLastError();
}

static int launch_k3(FusionParams& p, int frame_index, int sms, int smem_optin, cudaStream_t st)
{
    p.frame_index = frame_index;
    const saf_frame& f = p.frames[frame_index];
    const int C = p.vol.feature_dim;
    const int R = f.npy * f.npx;
    const bool packed = (p.pack_mask >> frame_index) & 1u;
    const float* table = packed ? p.tables + (uint64_t)frame_index * p.max_table_elems : f.table;
    const int64_t stride_r = packed ? C : f.table_stride_r;
    const bool rows16 = (C % 4 == 0) && (stride_r % 4 == 0) && (((uintptr_t)table & 15u) == 0) &&
                        (((uintptr_t)p.vol.clip_feat & 15u) == 0);
    const size_t tab_bytes = (size_t)R * C * 4;
    const bool fits = tab_bytes + 1024 <= (size_t)smem_optin;
    if (rows16 && (C == 512 || C == 768 || C == 1024)) {
        const int use_tma = 1;
        if (fits) {
            switch (C) {
                case 512: return launch_k3_fixed<4, 2, true>(p, table, stride_r, use_tma, tab_bytes, sms, st);
                case 768: return launch_k3_fixed<6, 2, true>(p, table, stride_r, use_tma, tab_bytes, sms, st);
                default: return launch_k3_fixed<8, 2, true>(p, table, stride_r, use_tma, tab_bytes, sms, st);
            }
        }
        switch (C) {
            case 512: return launch_k3_fixed<4, 2, false>(p, table, stride_r, 0, 0, sms, st);
            case 768: return launch_k3_fixed<6, 2, false>(p, table, stride_r, 0, 0, sms, st);
            default: return launch_k3_fixed<8, 2, false>(p, table, stride_r, 0, 0, sms, st);
        }
    }
    if (rows16)
        feature_accumulate_generic_kernel<4><<<sms * 2, kK3Threads, 0, st>>>(p, table, stride_r);
    else
        feature_accumulate_generic_kernel<1><<<sms * 2, kK3Threads, 0, st>>>(p, table, stride_r);
    return (int)cudaGetLastError();
}

}  // namespace saf

using namespace saf;

extern "C" {

int saf_workspace_bytes(const saf_grid_desc* grid, int32_t max_batch, int64_t max_table_elems, uint64_t* bytes_out)
{
    if (!bytes_out) return SAF_ERR_NULL;
    WsLayout L;
    int rc = compute_layout(grid, max_batch, max_table_elems, &L);
    if (rc) return rc;
    *bytes_out = L.bytes;
    return 0;
}

int saf_workspace_init(const saf_workspace* ws, const saf_grid_desc* grid, void* stream)
{
    if (!ws || !ws->base) return SAF_ERR_NULL;
    if (((uintptr_t)ws->base & 255u) != 0) return SAF_ERR_ALIGNMENT;
    WsLayout L;
    int rc = compute_layout(grid, ws->max_batch, ws->max_table_elems, &L);
    if (rc) return rc;
    if (ws->bytes < L.bytes) return SAF_ERR_WORKSPACE;
    WsHeader h;
    memset(&h, 0, sizeof(h));
    h.magic = kWsMagic;
    h.bytes = L.bytes;
    h.list_cap = L.list_cap;
    h.max_table_elems = (uint64_t)ws->max_table_elems;
    h.off_blocks = L.off_blocks;
    h.off_lists = L.off_lists;
    h.off_tables = L.off_tables;
    h.nblocks_total = L.nblocks_total;
    h.max_batch = (uint32_t)ws->max_batch;
    h.nb[0] = L.nb[0];
    h.nb[1] = L.nb[1];
    h.nb[2] = L.nb[2];
    cudaStream_t st = (cudaStream_t)stream;
    SAF_CUDA_TRY(cudaMemsetAsync(ws->base, 0, 512, st));
    // header is tiny: a synchronous-with-stream copy from pageable memory is fine here
    SAF_CUDA_TRY(cudaMemcpyAsync(ws->base, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    SAF_CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

int saf_read_stats(const saf_workspace* ws, saf_stats* out, void* stream)
{
    if (!ws || !ws->base || !out) return SAF_ERR_NULL;
    WsHeader h;
    cudaStream_t st = (cudaStream_t)stream;
    SAF_CUDA_TRY(cudaMemcpyAsync(&h, ws->base, sizeof(h), cudaMemcpyDeviceToHost, st));
    SAF_CUDA_TRY(cudaStreamSynchronize(st));
    if (h.magic != kWsMagic) return SAF_ERR_WORKSPACE;
    memset(out, 0, sizeof(*out));
    out->total_frames = h.total_frames;
    out->total_valid = h.total_valid;
    out->total_tsdf_valid = h.total_tsdf_valid;
    out->total_blocks = h.total_blocks;
    out->last_blocks = h.last_blocks;
    for (int b = 0; b < SAF_MAX_BATCH; ++b) {
        out->last_valid[b] = h.n_valid[b];
        out->last_tsdf_valid[b] = h.n_tsdf_valid[b];
    }
    out->error_flags = h.error_flags;
    return 0;
}

int saf_frustum_cull(const saf_grid_desc* grid, const saf_frame* frames, int32_t batch, int32_t H, int32_t W,
                     float trunc, const saf_workspace* ws, void* stream)
{
    int sms = 0;
    int rc = device_sm_count(&sms, nullptr);
    if (rc) return rc;
    FusionParams p;
    rc = build_params(grid, nullptr, frames, batch, H, W, trunc, SAF_RGB_BILINEAR, ws, &p);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    // stand-alone use: do not rely on K2 having re-armed the counter
    SAF_CUDA_TRY(cudaMemsetAsync(&p.hdr->n_blocks, 0, sizeof(uint32_t), st));
    return launch_k1(p, st);
}

int saf_tsdf_update(const saf_grid_desc* grid, const saf_volume* vol, const saf_frame* frames, int32_t batch, int32_t H,
                    int32_t W, float trunc, const saf_workspace* ws, uint8_t* valid_out, uint8_t* tsdf_valid_out,
                    void* stream)
{
    int sms = 0;
    int rc = device_sm_count(&sms, nullptr);
    if (rc) return rc;
    if (!vol || !vol->tsdf || !vol->tsdf_weight) return SAF_ERR_NULL;
    if ((valid_out == nullptr) != (tsdf_valid_out == nullptr)) return SAF_ERR_NULL;
    if (!(trunc > 0.f)) return SAF_ERR_SHAPE;
    FusionParams p;
    rc = build_params(grid, vol, frames, batch, H, W, trunc, SAF_RGB_BILINEAR, ws, &p);
    if (rc) return rc;
    p.valid_out = valid_out;
    p.tsdf_valid_out = tsdf_valid_out;
    return launch_k2(p, sms, (cudaStream_t)stream);
}

int saf_feature_accumulate(const saf_grid_desc* grid, const saf_volume* vol, const saf_frame* frames, int32_t batch,
                           int32_t frame_index, int32_t H, int32_t W, int32_t rgb_mode, const saf_workspace* ws,
                           void* stream)
{
    int sms = 0, smem_optin = 0;
    int rc = device_sm_count(&sms, &smem_optin);
    if (rc) return rc;
    if (frame_index < 0 || frame_index >= batch) return SAF_ERR_BATCH;
    FusionParams p;
    rc = build_params(grid, vol, frames, batch, H, W, 1.0f, rgb_mode, ws, &p);
    if (rc) return rc;
    rc = check_feature_args(vol, frames, batch, rgb_mode, ws, &p.pack_mask);
    if (rc) return rc;
    return launch_k3(p, frame_index, sms, smem_optin, (cudaStream_t)stream);
}

int saf_integrate(const saf_grid_desc* grid, const saf_volume* vol, const saf_frame* frames, int32_t batch, int32_t H,
                  int32_t W, float trunc, int32_t rgb_mode, const saf_workspace* ws, void* stream)
{
    int sms = 0, smem_optin = 0;
    int rc = device_sm_count(&sms, &smem_optin);
    if (rc) return rc;
    if (!(trunc > 0.f)) return SAF_ERR_SHAPE;
    FusionParams p;
    rc = build_params(grid, vol, frames, batch, H, W, trunc, rgb_mode, ws, &p);
    if (rc) return rc;
    rc = check_feature_args(vol, frames, batch, rgb_mode, ws, &p.pack_mask);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = launch_k1(p, st))) return rc;
    if ((rc = launch_k2(p, sms, st))) return rc;
    for (int b = 0; b < batch; ++b)
        if ((rc = launch_k3(p, b, sms, smem_optin, st))) return rc;
    return 0;
}

int saf_integrate_sequence(const saf_grid_desc* grid, const saf_volume* vol, const saf_frame* frames, int32_t n_frames,
                           int32_t H, int32_t W, float trunc, int32_t rgb_mode, const saf_workspace* ws, void* stream)
{
    if (n_frames < 0) return SAF_ERR_BATCH;
    for (int32_t i = 0; i < n_frames; ++i) {
        int rc = saf_integrate(grid, vol, frames + i, 1, H, W, trunc, rgb_mode, ws, stream);
        if (rc) return rc;
    }
    return 0;
}

int saf_label_argmax(const int32_t* labels, int64_t n, int32_t n_classes, int64_t* out, void* stream)
{
    int sms = 0;
    int rc = device_sm_count(&sms, nullptr);
    if (rc) return rc;
    if (!labels || !out) return SAF_ERR_NULL;
    if (n < 0 || n_classes <= 0) return SAF_ERR_SHAPE;
    if (n == 0) return 0;
    const int64_t want = (n + 7) / 8;
    const int grid = (int)std::min<int64_t>(want, (int64_t)sms * 8);
    label_argmax_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(labels, n, n_classes, (long long*)out);
    return (int)cudaGetLastError();
}

}  // extern "C"
