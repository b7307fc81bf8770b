// saf_fusion.cu -- RGB-D integration on sm_100a.
//
// Replaces ClipSeemFusion.integrate (/root/reference/clip_seem_fusion.py:676-822) and
// ClipFusion.integrate (/root/reference/clipfusion.py:627-721), and the frame loop around them
// (clip_seem_fusion.py:305-313).  Kernels of one call (a frame, a reference batch, or a window of up to
// SAF_MAX_BATCH consecutive frames):
//   K0 depth_tiles_kernel      32x32-pixel depth maxima per frame (only while the adaptive depth cull is on)
//   K1 frame_setup_kernel      per 8^3 voxel block and frame: conservative frustum test (+ depth-reach test)
//                              -> ascending list of {block, per-frame mask}; repacks feature images to [R,C]
//   K2 tsdf_update_kernel      exact per-voxel projection / depth sample / masks / TSDF running average for
//                              the listed blocks and frames; emits each block's `valid` voxels in voxel
//                              order plus per-block counts and their prefix sums.  Window mode: the average
//                              is advanced frame by frame in registers and ONE union list is emitted
//   K3 feature_accumulate_*    single frame: one warp per listed voxel, the voxel's C-float feature row is
//                              pulled into a per-warp shared-memory ring by TMA bulk copies, blended with the
//                              bilinear sample of the frame's [R,C] table (TMA-staged in shared memory) and
//                              streamed back with 128-bit stores; rgb, label counter and weight one lane per
//                              voxel
//   K2T window_tile_setup_kernel  window mode: per 16-voxel tile of the union list the update records of every valid
//                              (voxel, frame) (bilinear weights, a, b, table rows) for K3W, and the window's rgb /
//                              weight / label-counter updates; one warp per tile over the whole GPU
//   K3W feature_accumulate_window_tile_kernel  window mode: each listed row is read once, every valid frame's
//                              update is applied in frame order in registers, and it is written once; a set of 8
//                              voxels shares the table-row loads of a frame (older variants: pair / per-voxel)
// The lists are deterministic and spatially ordered (ascending block, then voxel): warps work on runs of
// neighbouring voxels, which keeps their feature rows close together in DRAM, lets the per-voxel 4-byte
// accesses (weight, rgb, label counter) of the lanes share sectors and keeps the table rows in L1.
// All decisions that feed masks use explicitly rounded intrinsics in the reference's op order;
// this file is compiled with -fmad=false so nothing is contracted behind our back.
#include <limits.h>
#include <math.h>
#include <algorithm>
#include <mutex>
#include <unordered_map>
#include <vector>
#include <stdio.h>
#include <stdlib.h>
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
    int32_t sequential;        // window mode: the batch is a run of consecutive single-frame calls
    uint32_t pack_mask;        // frames whose feature image is repacked into the workspace
    uint32_t nb[3];
    uint32_t nblocks_total;
    uint32_t nxs;              // slab extent in x
    uint32_t nyl;              // slab extent in y (the grid's ny, or the sheared layout's local extent)
    uint64_t nslab;
    uint64_t list_cap;
    uint64_t max_table_elems;
    uint64_t table_slot_elems; // floats between consecutive frames' repacked tables
    uint32_t n_k1;             // number of K1 cull CTAs (segments of block_seg)
    uint32_t slot;             // which of the two per-call scratch slots this call uses
    WsHeader* hdr;
    uint32_t* cta_count;       // [n_k1]              visible blocks found by each K1 CTA
    float* tile_dmax;          // [batch][kMaxDepthTiles] largest depth in each (1 << tile_shift)^2 tile of the depth image
    int32_t tile_shift, ntx, nty;
    uint2* block_seg;          // [n_k1*256]          per-CTA ordered segments of {block id, per-frame mask}
    uint32_t* blk_count;       // [batch][nblocks_total]   valid voxels per visible block (by rank)
    uint32_t* blk_offset;      // [batch][nblocks_total+1] exclusive prefix of blk_count
    ValidEntry* lists;         // [batch][nblocks_total*512] rank r's entries start at r*512
    WinEntry* ulist;           // window mode: [nblocks_total*512] voxels valid in any frame, rank r's start at r*512
    float2* wcoords;           // window mode: [nblocks_total][batch][512] (gx, gy) per (block rank, frame, local voxel)
    float* tables;
    int32_t* w0_list;          // window mode: the union voxels' weights before the window (K2T -> K3W), indexed like ulist
    void* tile_meta;           // window mode: [tile_meta_cap] TileMeta<2> written by K2T for K3W's tiles (idle list regions)
    uint32_t tile_meta_cap;    // tiles of the union list K2T prepares; later tiles are prepared inside K3W
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

// Same taps and weights for a table stored with a one-cell zero border ([(H+2)*(W+2)] rows): dropped taps are
// redirected to the border, so all four indices are valid rows.
__device__ __forceinline__ void bilinear_setup_padded(float gx, float gy, int W, int H, Taps& t)
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
    const int cw = (ixw > -1 && ixw < W) ? ixw + 1 : 0, ce = (ixw + 1 > -1 && ixw + 1 < W) ? ixw + 2 : 0;
    const int rn = (iyn > -1 && iyn < H) ? iyn + 1 : 0, rs = (iyn + 1 > -1 && iyn + 1 < H) ? iyn + 2 : 0;
    const int pw = W + 2;
    t.idx[0] = rn * pw + cw;
    t.idx[1] = rn * pw + ce;
    t.idx[2] = rs * pw + cw;
    t.idx[3] = rs * pw + ce;
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

// Feature taps of one (voxel, frame): the reference's bilinear taps over the patch grid, or - segment-table mode,
// see SAF_TABLE_SEGMENTS - the single row of the voxel's nearest-sampled class id with weight 1.  Both feed the same
// four-tap mul/fma chain.  `flags`: the workspace's error word (a bad id samples zeros and is reported).
__device__ __forceinline__ int segment_row(const saf_frame& f, float gx, float gy, int W, int H, uint32_t* flags)
{
    const int px = nearest_index(gx, W), py = nearest_index(gy, H);
    const float lf = (px >= 0 && py >= 0) ? load_class_id(f.seg, f.seg_dtype, py * W + px) : 0.0f;
    const long long id = (long long)lf;
    if (id < 0 || id >= f.npx) {
        atomicOr(flags, SAF_FLAG_BAD_CLASS_ID);
        return -1;
    }
    return (int)id;
}
__device__ __forceinline__ void feature_taps(const saf_frame& f, float gx, float gy, int W, int H, uint32_t* flags, Taps& t)
{
    if (f.table_mode == SAF_TABLE_SEGMENTS) {
        t.idx[0] = segment_row(f, gx, gy, W, H, flags);
        t.idx[1] = t.idx[2] = t.idx[3] = -1;
        t.w[0] = 1.0f;
        t.w[1] = t.w[2] = t.w[3] = 0.0f;
    } else {
        bilinear_setup(gx, gy, f.npx, f.npy, t);
    }
}
// ... for the window kernels' repacked tables: zero-bordered patch grid, or [zero row, segment rows]
__device__ __forceinline__ void feature_taps_padded(const saf_frame& f, float gx, float gy, int W, int H, uint32_t* flags,
                                                    Taps& t)
{
    if (f.table_mode == SAF_TABLE_SEGMENTS) {
        t.idx[0] = segment_row(f, gx, gy, W, H, flags) + 1;
        t.idx[1] = t.idx[2] = t.idx[3] = 0;
        t.w[0] = 1.0f;
        t.w[1] = t.w[2] = t.w[3] = 0.0f;
    } else {
        bilinear_setup_padded(gx, gy, f.npx, f.npy, t);
    }
}

// Sensor-format inputs: what the reference's dataset classes hold before they convert
// (clipfusion.py:185-188, 245, 254, 355, 362): rgb = float(u8) / 255, depth = float(u16 millimetres) / 1000.
// Both quotients are produced without a division: q0 = x * fl(1/d), r = fma(-d, q0, x), q = fma(r, fl(1/d), q0)
// is the correctly rounded x / d for every u8 / u16 input (checked exhaustively: oracle/saf_oracle.c
// saf_oracle_check_sensor_conversions, tests/test_oracle_golden.py).
__device__ __forceinline__ float depth_from_mm(uint16_t v)
{
    const float x = __uint2float_rn((uint32_t)v), y = 1.0f / 1000.0f;
    const float q0 = __fmul_rn(x, y);
    return __fmaf_rn(__fmaf_rn(-1000.0f, q0, x), y, q0);
}
__device__ __forceinline__ float unorm_from_u8(uint8_t v)
{
    const float x = __uint2float_rn((uint32_t)v), y = 1.0f / 255.0f;
    const float q0 = __fmul_rn(x, y);
    return __fmaf_rn(__fmaf_rn(-255.0f, q0, x), y, q0);
}
__device__ __forceinline__ float load_depth(const saf_frame& f, size_t pix)
{
    if (f.depth_dtype == SAF_DEPTH_U16_MM) return depth_from_mm(__ldg(reinterpret_cast<const uint16_t*>(f.depth) + pix));
    return __ldg(reinterpret_cast<const float*>(f.depth) + pix);
}
__device__ __forceinline__ float load_rgb(const saf_frame& f, size_t elem)
{
    if (f.rgb_dtype == SAF_RGB_U8) return unorm_from_u8(__ldg(reinterpret_cast<const uint8_t*>(f.rgb) + elem));
    return __ldg(reinterpret_cast<const float*>(f.rgb) + elem);
}

// ---------------------------------------------------------------------------------------------
// K1: frame set-up (frustum cull of voxel blocks + feature-image repack)
// ---------------------------------------------------------------------------------------------

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

// bounding sphere (world centre, radius) of the voxel centres of block (bx,by,bz) of the slab
__device__ __forceinline__ void block_sphere(const FusionParams& p, uint32_t bx, uint32_t by, uint32_t bz, float& cx,
                                             float& cy, float& cz, float& r, float* half = nullptr)
{
    const int x0 = (int)bx * kBlockEdge, z0 = (int)bz * kBlockEdge;
    const int y0 = slab_global_y(p.grid, x0, (int)by * kBlockEdge);   // global y of the block's first row
    // (a local block column of the sheared layout that lies beyond the grid gets ey <= 0: see block_in_grid)
    const int ex = min(kBlockEdge, (int)p.nxs - x0), ey = max(1, min(kBlockEdge, p.grid.nvox[1] - y0)),
              ez = min(kBlockEdge, p.grid.nvox[2] - z0);
    const float vs = p.grid.voxel_size;
    const float hx = 0.5f * vs * (float)(ex - 1), hy = 0.5f * vs * (float)(ey - 1), hz = 0.5f * vs * (float)(ez - 1);
    cx = p.grid.origin[0] + vs * (float)slab_global_x(p.grid, x0) + hx;   // a block never straddles two stripes
    cy = p.grid.origin[1] + vs * (float)y0 + hy;
    cz = p.grid.origin[2] + vs * (float)z0 + hz;
    r = sqrtf(hx * hx + hy * hy + hz * hz) + 0.01f * vs;
    if (half) {
        half[0] = hx + 0.01f * vs;
        half[1] = hy + 0.01f * vs;
        half[2] = hz + 0.01f * vs;
    }
}

// sheared layout: the last local block column of an x-block may lie beyond the grid's y extent
__device__ __forceinline__ bool block_in_grid(const FusionParams& p, uint32_t bx, uint32_t by)
{
    return slab_global_y(p.grid, (int)bx * kBlockEdge, (int)by * kBlockEdge) < p.grid.nvox[1];
}

// smallest camera-space z = K[2,:] . R^T (x - t) over the block's box of voxel centres (exact minimum of a
// linear function over an axis-aligned box), minus rounding slack
__device__ __forceinline__ float block_z_min(const Geom& g, float cx, float cy, float cz, const float* half)
{
    const float d[3] = {cx - g.P[3], cy - g.P[7], cz - g.P[11]};
    float zc = 0.0f, ext = 0.0f, mag = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        // world-space gradient of z: n_i = sum_k R[i][k] K[2][k]
        const float n = g.P[4 * i] * g.K[6] + g.P[4 * i + 1] * g.K[7] + g.P[4 * i + 2] * g.K[8];
        zc += n * d[i];
        ext += fabsf(n) * half[i];
        mag += fabsf(n) * (fabsf(d[i]) + half[i]);
    }
    return zc - ext - 1e-3f * mag - 1e-6f;
}

// Largest depth any voxel of the block can sample: the maximum over the depth tiles its projected bounding
// sphere can touch (interval arithmetic on x/z, y/z), or the whole-image maximum when that is not cheap to
// bound (sphere reaches z <= 0, general intrinsics, footprint of more than 64 tiles).
__device__ __forceinline__ float block_depth_bound(const FusionParams& p, const Geom& g, float cx, float cy, float cz,
                                                   float r, const float* tiles_smem, const float* tiles_gmem,
                                                   float image_max)
{
    if (g.K[6] != 0.0f || g.K[7] != 0.0f || g.K[8] != 1.0f) return image_max;
    const float d0 = cx - g.P[3], d1 = cy - g.P[7], d2 = cz - g.P[11];
    float pc[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) pc[k] = g.P[k] * d0 + g.P[4 + k] * d1 + g.P[8 + k] * d2;
    const float rr = r * 1.01f + 1e-4f * (fabsf(pc[0]) + fabsf(pc[1]) + fabsf(pc[2]));
    const float z0 = pc[2] - rr, z1 = pc[2] + rr;
    if (!(z0 > 1e-4f)) return image_max;
    float lo[2], hi[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const float x0 = pc[a] - rr, x1 = pc[a] + rr;
        lo[a] = fminf(x0 / z0, x0 / z1);
        hi[a] = fmaxf(x1 / z0, x1 / z1);
    }
    // u = K00 ax + K01 ay + K02,  v = K10 ax + K11 ay + K12
    float uv_lo[2], uv_hi[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const float k0 = g.K[3 * a], k1 = g.K[3 * a + 1], k2 = g.K[3 * a + 2];
        uv_lo[a] = fminf(k0 * lo[0], k0 * hi[0]) + fminf(k1 * lo[1], k1 * hi[1]) + k2;
        uv_hi[a] = fmaxf(k0 * lo[0], k0 * hi[0]) + fmaxf(k1 * lo[1], k1 * hi[1]) + k2;
        const float pad = 1.5f + 1e-3f * fmaxf(fabsf(uv_lo[a]), fabsf(uv_hi[a]));  // nearest-pixel rounding + fp slack
        uv_lo[a] -= pad;
        uv_hi[a] += pad;
    }
    if (!(uv_lo[0] == uv_lo[0]) || !(uv_hi[0] == uv_hi[0]) || !(uv_lo[1] == uv_lo[1]) || !(uv_hi[1] == uv_hi[1]))
        return image_max;
    const int tx0 = max(0, (int)floorf(fmaxf(uv_lo[0], -1.0f))) >> p.tile_shift;
    const int ty0 = max(0, (int)floorf(fmaxf(uv_lo[1], -1.0f))) >> p.tile_shift;
    const int tx1 = min(p.ntx - 1, (int)floorf(fminf(uv_hi[0], (float)p.W)) >> p.tile_shift);
    const int ty1 = min(p.nty - 1, (int)floorf(fminf(uv_hi[1], (float)p.H)) >> p.tile_shift);
    if (uv_hi[0] < 0.0f || uv_hi[1] < 0.0f || tx0 > tx1 || ty0 > ty1) return 0.0f;  // only out-of-image pixels: depth 0
    if ((tx1 - tx0 + 1) * (ty1 - ty0 + 1) > 64) return image_max;
    float m = 0.0f;
    for (int ty = ty0; ty <= ty1; ++ty)
        for (int tx = tx0; tx <= tx1; ++tx)
            m = fmaxf(m, tiles_smem ? tiles_smem[ty * p.ntx + tx] : __ldcg(tiles_gmem + ty * p.ntx + tx));
    return m;
}

// exclusive prefix sum of n values (global, read through L2) by one CTA (blockDim.x a multiple of 32,
// at most 1024); returns the total to every thread.  scratch: 33 uint32 in shared memory.
__device__ uint32_t cta_exclusive_scan(const uint32_t* src, uint32_t* dst, uint32_t n, uint32_t* scratch)
{
    const uint32_t nt = blockDim.x, t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarp = nt >> 5;
    const uint32_t per = (n + nt - 1) / nt;
    const uint32_t lo = min(n, t * per), hi = min(n, lo + per);
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; ++i) sum += __ldcg(src + i);
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += up;
    }
    if (lane == 31) scratch[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t ws = lane < nwarp ? scratch[lane] : 0u;
        uint32_t wi = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= (uint32_t)o) wi += up;
        }
        if (lane < nwarp) scratch[lane] = wi - ws;  // exclusive warp base
        if (lane == 31) scratch[32] = wi;           // grand total
    }
    __syncthreads();
    uint32_t base = scratch[warp] + incl - sum;
    const uint32_t total = scratch[32];
    for (uint32_t i = lo; i < hi; ++i) {
        dst[i] = base;
        base += __ldcg(src + i);
    }
    __syncthreads();
    return total;
}

// K0: largest depth of every (1 << tile_shift)^2 tile of the frames' depth images, for K1's depth cull.  One warp
// per (frame, tile), each lane walks pixel rows.  Launched with every call but returns at once while the cull is
// switched off.
__global__ void __launch_bounds__(256) depth_tiles_kernel(const FusionParams p)
{
    if (!p.hdr->depth_cull) return;
    const int lane = threadIdx.x & 31;
    const int ts = 1 << p.tile_shift, ntiles = p.ntx * p.nty;
    const int item = (int)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (item >= p.batch * ntiles) return;
    const int b = item / ntiles, t = item - b * ntiles;
    const saf_frame& f = p.frames[b];
    const int x0 = (t % p.ntx) * ts, y0 = (t / p.ntx) * ts;
    const int x1 = min(p.W, x0 + ts), y1 = min(p.H, y0 + ts);
    float dm = 0.0f;
    for (int y = y0; y < y1; ++y)
        for (int x = x0 + lane; x < x1; x += 32) dm = fmaxf(dm, load_depth(f, (size_t)y * p.W + x));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dm = fmaxf(dm, __shfl_xor_sync(0xffffffffu, dm, o));
    if (lane == 0) p.tile_dmax[(size_t)b * kMaxDepthTiles + t] = dm;
}

__global__ void __launch_bounds__(kK1Threads) frame_setup_kernel(const FusionParams p, uint32_t cull_ctas)
{
    if (blockIdx.x >= cull_ctas) {
        // repack channel-major feature images into [R,C] rows (only frames in pack_mask)
        const uint32_t pack_ctas = gridDim.x - cull_ctas;
        const int C = p.vol.feature_dim;
        if (p.sequential) {
            // window mode: [(npy+2)*(npx+2), C] rows with a zero border, so that the feature kernel's four taps are
            // always in range (dropped taps land on a zero row: grid_sample's zeros padding).  The (frame, row)
            // pairs of the whole window are dealt over the pack CTAs, one 16-byte access per thread where the
            // layout allows - a loop over the frames with scalar copies was K1's long pole (16 dependent passes).
            int rp_max = 0;
            for (int b = 0; b < p.batch; ++b) {
                const saf_frame& f = p.frames[b];
                rp_max = max(rp_max, f.table_mode == SAF_TABLE_SEGMENTS ? f.npx + 1 : (f.npy + 2) * (f.npx + 2));
            }
            for (int idx = (int)(blockIdx.x - cull_ctas); idx < p.batch * rp_max; idx += (int)pack_ctas) {
                const int b = idx / rp_max, r = idx - b * rp_max;
                const saf_frame& f = p.frames[b];
                const bool segs = f.table_mode == SAF_TABLE_SEGMENTS;   // [zero row, one row per segment]
                const int pw = f.npx + 2;
                const int Rp = segs ? f.npx + 1 : (f.npy + 2) * pw;
                if (r >= Rp) continue;
                float* dst = p.tables + (uint64_t)b * p.table_slot_elems + (int64_t)r * C;
                const int py = segs ? 0 : r / pw - 1, px = segs ? r - 1 : r % pw - 1;
                const bool in = py >= 0 && py < f.npy && px >= 0 && px < f.npx;
                const float* src = f.table + ((int64_t)py * f.npx + px) * f.table_stride_r;
                const bool vec = f.table_stride_c == 1 && (C & 3) == 0 && (p.table_slot_elems & 3) == 0 &&
                                 ((reinterpret_cast<uintptr_t>(p.tables) | reinterpret_cast<uintptr_t>(f.table)) & 15u) == 0 &&
                                 (f.table_stride_r & 3) == 0;
                if (vec) {
                    for (int c = threadIdx.x; c < C / 4; c += kK1Threads)
                        reinterpret_cast<float4*>(dst)[c] =
                            in ? __ldg(reinterpret_cast<const float4*>(src) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                } else {
                    for (int c = threadIdx.x; c < C; c += kK1Threads)
                        dst[c] = in ? src[(int64_t)c * f.table_stride_c] : 0.0f;
                }
            }
            return;
        }
        for (int b = 0; b < p.batch; ++b) {
            const saf_frame& f = p.frames[b];
            float* dst = p.tables + (uint64_t)b * p.table_slot_elems;
            if (!((p.pack_mask >> b) & 1u)) continue;
            const int64_t R = (int64_t)f.npy * f.npx;
            for (int64_t e = (int64_t)(blockIdx.x - cull_ctas) * kK1Threads + threadIdx.x; e < R * C;
                 e += (int64_t)pack_ctas * kK1Threads) {
                const int64_t r = e / C, c = e - r * C;
                dst[e] = f.table[c * f.table_stride_c + r * f.table_stride_r];
            }
        }
        return;
    }
    __shared__ uint32_t s_warp[kK1Threads / 32];
    __shared__ float s_zfar[SAF_MAX_BATCH];     // whole-image depth bound per frame
    extern __shared__ float s_tiles[];          // [batch][ntiles] depth-tile maxima (K0), only while the cull is on
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t blk = blockIdx.x * kK1Threads + threadIdx.x;
    // Depth cull (switched on by K2's policy only while it pays): a voxel is tsdf_valid only if z < depth + trunc
    // (clip_seem_fusion.py:722-728), so a block whose nearest z lies behind (largest depth over the image tiles
    // its projection can touch) + trunc cannot be touched by that frame.  NaN depths never win fmaxf; pixels
    // outside the image sample depth 0.
    const bool depth_cull = p.hdr->depth_cull != 0;
    const int ntiles = p.ntx * p.nty;
    if (depth_cull) {
        if (threadIdx.x < SAF_MAX_BATCH) s_zfar[threadIdx.x] = 0.0f;
        __syncthreads();
        for (int b = 0; b < p.batch; ++b) {
            float dm = 0.0f;
            for (int t = threadIdx.x; t < ntiles; t += kK1Threads) {
                const float v = __ldcg(p.tile_dmax + (size_t)b * kMaxDepthTiles + t);
                s_tiles[b * ntiles + t] = v;
                dm = fmaxf(dm, v);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dm = fmaxf(dm, __shfl_xor_sync(0xffffffffu, dm, o));
            // non-negative floats order like their bit patterns
            if (lane == 0) atomicMax(reinterpret_cast<int*>(&s_zfar[b]), __float_as_int(dm));
        }
        __syncthreads();
    }
    uint32_t vis = 0;       // bit b: the block may be touched by frame b
    uint32_t frustum = 0;   // ... before the depth test
    if (blk < p.nblocks_total && block_in_grid(p, blk / (p.nb[2] * p.nb[1]), (blk / p.nb[2]) % p.nb[1])) {
        const uint32_t bz = blk % p.nb[2];
        const uint32_t by = (blk / p.nb[2]) % p.nb[1];
        const uint32_t bx = blk / (p.nb[2] * p.nb[1]);
        float cx, cy, cz, r, half[3];
        block_sphere(p, bx, by, bz, cx, cy, cz, r, half);
        const float fW = (float)p.W, fH = (float)p.H;
        for (int b = 0; b < p.batch; ++b) {
            Geom g;
            load_geom(p.frames[b], g);
            if (!block_maybe_visible(g, cx, cy, cz, r, fW, fH)) continue;
            frustum |= 1u << b;
            if (depth_cull) {
                const float dfar = block_depth_bound(p, g, cx, cy, cz, r, s_tiles + b * ntiles, nullptr, s_zfar[b]);
                if (block_z_min(g, cx, cy, cz, half) > (dfar + p.trunc) * 1.001f + 1e-5f) continue;  // NaN -> keep
            }
            vis |= 1u << b;
        }
    }
    if (depth_cull) {
        const uint32_t nf = (uint32_t)__popc(__ballot_sync(0xffffffffu, frustum != 0u));
        if (lane == 0 && nf) atomicAdd(&p.hdr->slot[p.slot].n_frustum_acc, nf);
    }
    // ordered compaction inside the CTA: this CTA's visible blocks, ascending, into its own segment
    const unsigned m = __ballot_sync(0xffffffffu, vis != 0u);
    if (lane == 0) s_warp[warp] = (uint32_t)__popc(m);
    __syncthreads();
    uint32_t base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kK1Threads / 32; ++w) {
        if (w < warp) base += s_warp[w];
        total += s_warp[w];
    }
    if (vis) p.block_seg[(size_t)blockIdx.x * kK1Threads + base + __popc(m & ((1u << lane) - 1u))] = make_uint2(blk, vis);
    if (threadIdx.x == 0) p.cta_count[blockIdx.x] = total;
}

// ---------------------------------------------------------------------------------------------
// K2: TSDF update over the visible blocks (clip_seem_fusion.py:698-744)
// ---------------------------------------------------------------------------------------------

constexpr int kK2Threads = 128;
constexpr int kK2Iter = kBlockVoxels / kK2Threads;  // voxels per thread per block

// SEQ = window mode: the batch is a run of consecutive single-frame integrate() calls.  Each voxel's TSDF running
// average is advanced frame by frame in registers (read once, written once per window) and the kernel emits ONE
// list of the voxels valid in at least one frame (WinEntry + per-frame coordinates) for the window feature kernel.
template <bool BATCH1, bool SEQ>
#ifndef SAF_K2_MINBLOCKS
#define SAF_K2_MINBLOCKS 10   // window-mode K2: 10 CTAs (40 warps) per SM at 48 registers; 8: -2 % on the step, 12: -1 %, 16: -5 %
#endif
__global__ void __launch_bounds__(kK2Threads, SEQ ? SAF_K2_MINBLOCKS : 1) tsdf_update_kernel(const FusionParams p)
{
    static_assert(!(BATCH1 && SEQ), "a one-frame window is the plain single-frame path");
    __shared__ uint32_t s_cnt[kK2Iter][kK2Threads / 32];
    __shared__ uint32_t s_fcnt[2][SAF_MAX_BATCH];   // SEQ: per-frame tsdf_valid / valid counts of this CTA
    __shared__ uint32_t s_scan[33];
    extern __shared__ uint32_t s_off[];  // [n_k1]
    __shared__ bool is_last;
    WsHeader* hdr = p.hdr;
    SlotCounters* sc = &hdr->slot[p.slot];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // K1 left one ordered segment of visible blocks per cull CTA: rank -> (segment, position)
    const uint32_t n_blocks = cta_exclusive_scan(p.cta_count, s_off, p.n_k1, s_scan);
    const bool depth_cull = hdr->depth_cull != 0;   // K1 applied the depth test to the block list (policy below)
    __shared__ uint32_t s_rank;                // list segment claimed for the block being processed
    const int B = BATCH1 ? 1 : p.batch;
    const float fW = (float)p.W, fH = (float)p.H;
    const int ny = p.grid.nvox[1], nz = p.grid.nvox[2], nyl = (int)p.nyl;
    uint32_t tv_count[BATCH1 ? 1 : SAF_MAX_BATCH];
#pragma unroll
    for (int b = 0; b < (BATCH1 ? 1 : SAF_MAX_BATCH); ++b) tv_count[b] = 0;
    Geom g;
    if (BATCH1) load_geom(p.frames[0], g);
    if (SEQ) {
        if (threadIdx.x < 2 * SAF_MAX_BATCH) s_fcnt[threadIdx.x / SAF_MAX_BATCH][threadIdx.x % SAF_MAX_BATCH] = 0;
        __syncthreads();
    }
    for (uint32_t bi = blockIdx.x; bi < n_blocks; bi += gridDim.x) {
        uint32_t lo = 0, hi = p.n_k1;  // s_off[lo] <= bi < s_off[hi] (with s_off[n_k1] = n_blocks)
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_off[mid] <= bi)
                lo = mid;
            else
                hi = mid;
        }
        const uint2 packed = p.block_seg[(size_t)lo * kK1Threads + (bi - s_off[lo])];
        const uint32_t blk = packed.x;
        const uint32_t fmask = packed.y;  // frames that can touch the block (K1: frustum and depth reach)
        const uint32_t bz = blk % p.nb[2];
        const uint32_t by = (blk / p.nb[2]) % p.nb[1];
        const uint32_t bx = blk / (p.nb[2] * p.nb[1]);
        // claim the next list segment (arrival order; the list order does not affect any result)
        if (threadIdx.x == 0) s_rank = atomicAdd(&sc->n_processed, 1u);
        float xw[kK2Iter], yw[kK2Iter], zw[kK2Iter], t_old[kK2Iter], bt[kK2Iter];
        uint32_t v[kK2Iter];
        int tw[kK2Iter], bw[kK2Iter];
        bool inside[kK2Iter];
        // voxel centres, and the voxel's TSDF state fetched up front (one latency instead of two)
#pragma unroll
        for (int j = 0; j < kK2Iter; ++j) {
            const int local = threadIdx.x + j * kK2Threads;
            const int lx = (int)bx * kBlockEdge + (local >> 6);  // slab-local x
            const int ly = (int)by * kBlockEdge + ((local >> 3) & 7);   // slab-local y
            const int iy = slab_global_y(p.grid, lx, ly);
            const int iz = (int)bz * kBlockEdge + (local & 7);
            inside[j] = lx < (int)p.nxs && iy < ny && iz < nz;
            v[j] = inside[j] ? (uint32_t)(((uint64_t)lx * nyl + ly) * nz + iz) : 0u;
            xw[j] = voxel_centre(slab_global_x(p.grid, lx), p.grid.voxel_size, p.grid.origin[0]);
            yw[j] = voxel_centre(iy, p.grid.voxel_size, p.grid.origin[1]);
            zw[j] = voxel_centre(iz, p.grid.voxel_size, p.grid.origin[2]);
            t_old[j] = inside[j] ? p.vol.tsdf[v[j]] : 0.0f;
            tw[j] = inside[j] ? p.vol.tsdf_weight[v[j]] : 0;
            bw[j] = 0;
            bt[j] = 0.0f;
        }
        if (SEQ) {
            __syncthreads();  // s_rank
            const uint32_t rank = s_rank;
            uint32_t m[kK2Iter];
            bool dirty[kK2Iter];
#pragma unroll
            for (int j = 0; j < kK2Iter; ++j) {
                m[j] = 0;
                dirty[j] = false;
            }
            for (int b = 0; b < B; ++b) {
                if (!((fmask >> b) & 1u)) continue;  // CTA-uniform
                const saf_frame& f = p.frames[b];
                load_geom(f, g);
                float2* cdst = p.wcoords + ((uint64_t)rank * B + b) * kBlockVoxels;
                uint32_t n_tv = 0, n_v = 0;
#pragma unroll
                for (int j = 0; j < kK2Iter; ++j) {
                    float gx, gy, z;
                    project(g.P, g.K, xw[j], yw[j], zw[j], fW, fH, gx, gy, z);
                    const int px = nearest_index(gx, p.W), py = nearest_index(gy, p.H);
                    const float d = (inside[j] && px >= 0 && py >= 0) ? load_depth(f, (size_t)py * p.W + px) : 0.0f;
                    const float sdf = __fdiv_rn(__fsub_rn(d, z), p.trunc);
                    const bool in_view = inside[j] && (fabsf(gx) <= 1.0f) && (fabsf(gy) <= 1.0f) && (z > 0.0f);
                    const bool valid = in_view && (fabsf(sdf) <= 1.0f);
                    const bool tv = in_view && (sdf > -1.0f);
                    if (tv) {
                        // one single-frame call of clip_seem_fusion.py:730-744: bw = 1, bt = clamp(sdf)
                        const int nw = tw[j] + 1;
                        const float fnw = __int2float_rn(nw);
                        t_old[j] = __fadd_rn(__fdiv_rn(fminf(fmaxf(sdf, -1.0f), 1.0f), fnw),
                                             __fmul_rn(t_old[j], __fdiv_rn(__int2float_rn(tw[j]), fnw)));
                        tw[j] = nw;
                        dirty[j] = true;
                    }
                    if (valid) {
                        m[j] |= 1u << b;
                        cdst[threadIdx.x + j * kK2Threads] = make_float2(gx, gy);
                    }
                    if (p.valid_out && inside[j]) {
                        if (valid) p.valid_out[(uint64_t)b * p.nslab + v[j]] = 1;
                        if (tv) p.tsdf_valid_out[(uint64_t)b * p.nslab + v[j]] = 1;
                    }
                    n_tv += (uint32_t)__popc(__ballot_sync(0xffffffffu, tv));
                    n_v += (uint32_t)__popc(__ballot_sync(0xffffffffu, valid));
                }
                if (lane == 0) {
                    if (n_tv) atomicAdd(&s_fcnt[0][b], n_tv);
                    if (n_v) atomicAdd(&s_fcnt[1][b], n_v);
                }
            }
            // the block's union list, in voxel order
            unsigned um[kK2Iter];
#pragma unroll
            for (int j = 0; j < kK2Iter; ++j) {
                um[j] = __ballot_sync(0xffffffffu, m[j] != 0u);
                if (lane == 0) s_cnt[j][warp] = (uint32_t)__popc(um[j]);
            }
            __syncthreads();
            uint32_t run = 0;
            WinEntry* seg = p.ulist + (uint64_t)rank * kBlockVoxels;
#pragma unroll
            for (int j = 0; j < kK2Iter; ++j) {
#pragma unroll
                for (int w = 0; w < kK2Threads / 32; ++w) {
                    if (w == warp && m[j] != 0u) {
                        WinEntry e;
                        e.voxel = v[j];
                        e.mask_local = m[j] | ((uint32_t)(threadIdx.x + j * kK2Threads) << 16);
                        seg[run + __popc(um[j] & ((1u << lane) - 1u))] = e;
                    }
                    run += s_cnt[j][w];
                }
            }
            if (threadIdx.x == 0) p.blk_count[rank] = run;
#pragma unroll
            for (int j = 0; j < kK2Iter; ++j) {
                if (dirty[j]) {
                    p.vol.tsdf[v[j]] = t_old[j];
                    p.vol.tsdf_weight[v[j]] = tw[j];
                }
            }
            __syncthreads();  // s_cnt / s_rank are reused by the next block
        }
        for (int b = 0; b < (SEQ ? 0 : B); ++b) {
            if (!BATCH1 && !((fmask >> b) & 1u)) {  // CTA-uniform: no voxel of the block can be in view of frame b
                __syncthreads();  // s_rank
                if (threadIdx.x == 0) p.blk_count[(uint64_t)b * p.nblocks_total + s_rank] = 0;
                continue;
            }
            const saf_frame& f = p.frames[b];
            if (!BATCH1) load_geom(f, g);
            float gx[kK2Iter], gy[kK2Iter], z[kK2Iter], d[kK2Iter];
#pragma unroll
            for (int j = 0; j < kK2Iter; ++j) {
                project(g.P, g.K, xw[j], yw[j], zw[j], fW, fH, gx[j], gy[j], z[j]);
                const int px = nearest_index(gx[j], p.W), py = nearest_index(gy[j], p.H);
                d[j] = (inside[j] && px >= 0 && py >= 0) ? load_depth(f, (size_t)py * p.W + px) : 0.0f;
            }
            bool valid[kK2Iter];
            unsigned vmask[kK2Iter];
#pragma unroll
            for (int j = 0; j < kK2Iter; ++j) {
                const float sdf = __fdiv_rn(__fsub_rn(d[j], z[j]), p.trunc);
                const bool in_view = inside[j] && (fabsf(gx[j]) <= 1.0f) && (fabsf(gy[j]) <= 1.0f) && (z[j] > 0.0f);
                valid[j] = in_view && (fabsf(sdf) <= 1.0f);
                const bool tv = in_view && (sdf > -1.0f);
                if (tv) {
                    bw[j] += 1;
                    bt[j] = __fadd_rn(bt[j], fminf(fmaxf(sdf, -1.0f), 1.0f));
                    tv_count[BATCH1 ? 0 : b] += 1;
                }
                if (p.valid_out && inside[j]) {
                    if (valid[j]) p.valid_out[(uint64_t)b * p.nslab + v[j]] = 1;
                    if (tv) p.tsdf_valid_out[(uint64_t)b * p.nslab + v[j]] = 1;
                }
                vmask[j] = __ballot_sync(0xffffffffu, valid[j]);
                if (lane == 0) s_cnt[j][warp] = (uint32_t)__popc(vmask[j]);
            }
            __syncthreads();
            // entries in voxel order: rank = (# valid with a smaller local index)
            uint32_t run = 0;
            const uint32_t rank = s_rank;  // written before the barrier above
            ValidEntry* seg = p.lists + (uint64_t)b * p.list_cap + (uint64_t)rank * kBlockVoxels;
#pragma unroll
            for (int j = 0; j < kK2Iter; ++j) {
#pragma unroll
                for (int w = 0; w < kK2Threads / 32; ++w) {
                    if (w == warp && valid[j]) {
                        ValidEntry e;
                        e.voxel = v[j];
                        e.gx = gx[j];
                        e.gy = gy[j];
                        e.pad = 0;
                        seg[run + __popc(vmask[j] & ((1u << lane) - 1u))] = e;
                    }
                    run += s_cnt[j][w];
                }
            }
            if (threadIdx.x == 0) p.blk_count[(uint64_t)b * p.nblocks_total + rank] = run;
            __syncthreads();
        }
#pragma unroll
        for (int j = 0; j < kK2Iter; ++j) {
            if (!SEQ && bw[j] > 0) {
                // clip_seem_fusion.py:736-744: three separately rounded fp32 ops
                const int nw = tw[j] + bw[j];
                const float fnw = __int2float_rn(nw);
                p.vol.tsdf[v[j]] =
                    __fadd_rn(__fdiv_rn(bt[j], fnw), __fmul_rn(t_old[j], __fdiv_rn(__int2float_rn(tw[j]), fnw)));
                p.vol.tsdf_weight[v[j]] = nw;
            }
        }
    }
    if (SEQ) {
        __syncthreads();
        if ((int)threadIdx.x < B) {
            if (s_fcnt[0][threadIdx.x]) atomicAdd(&sc->n_tsdf_valid[threadIdx.x], s_fcnt[0][threadIdx.x]);
            if (s_fcnt[1][threadIdx.x]) atomicAdd(&sc->acc_valid[threadIdx.x], s_fcnt[1][threadIdx.x]);
        }
    }
    // per-frame tsdf_valid totals: one atomic per warp
#pragma unroll
    for (int b = 0; b < (BATCH1 ? 1 : SAF_MAX_BATCH); ++b) {
        if (SEQ || b >= B) break;
        uint32_t c = tv_count[b];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0 && c) atomicAdd(&sc->n_tsdf_valid[b], c);
    }
    // the last CTA turns the per-block counts into list offsets and folds the call into the totals
    // (one fence by the ticket-taking thread after the CTA barrier publishes the whole CTA's writes)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = (atomicAdd(&sc->k2_done, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    unsigned long long sv = 0, stv = 0;
    const uint32_t n_proc = atomicAdd(&sc->n_processed, 0u);  // == n_blocks: every listed block claimed a segment
    if (SEQ) {
        // one union list for the window; the per-frame counts come from the atomics
        const uint32_t total = cta_exclusive_scan(p.blk_count, p.blk_offset, n_proc, s_scan);
        if (threadIdx.x == 0) {
            p.blk_offset[n_proc] = total;
            sc->n_union = total;
            hdr->total_union += total;
            for (int b = 0; b < B; ++b) {
                const uint32_t nv = atomicExch(&sc->acc_valid[b], 0u);
                const uint32_t t = atomicExch(&sc->n_tsdf_valid[b], 0u);
                sc->n_valid[b] = nv;
                sc->last_tsdf_valid[b] = t;
                sv += nv;
                stv += t;
            }
        }
    }
    for (int b = 0; b < (SEQ ? 0 : B); ++b) {
        uint32_t* off = p.blk_offset + (uint64_t)b * (p.nblocks_total + 1);
        const uint32_t total = cta_exclusive_scan(p.blk_count + (uint64_t)b * p.nblocks_total, off, n_proc, s_scan);
        if (threadIdx.x == 0) {
            off[n_proc] = total;
            sc->n_valid[b] = total;
            const uint32_t t = atomicExch(&sc->n_tsdf_valid[b], 0u);
            sc->last_tsdf_valid[b] = t;
            sv += total;
            stv += t;
        }
    }
    if (threadIdx.x == 0) {
        // depth-cull policy for the next calls: on while fewer than a quarter of the visited voxels were
        // tsdf_valid (many blocks behind the surfaces), off again (with a cool-down) once it removes < 1/8
        // (depth_cull_cooldown counts idle calls while on, remaining cool-down calls while off)
        sc->n_processed = 0;
        sc->last_processed = n_proc;
        const uint32_t n_frustum = depth_cull ? atomicExch(&sc->n_frustum_acc, 0u) : n_blocks;
        if (!depth_cull) {
            if (hdr->depth_cull_cooldown) hdr->depth_cull_cooldown -= 1;
            else if (n_blocks >= 64 && stv * 4ull < (unsigned long long)n_proc * kBlockVoxels * (unsigned long long)B) {
                hdr->depth_cull = 1;
                hdr->depth_cull_cooldown = 0;
            }
        } else if ((n_frustum - n_blocks) * 8u < n_frustum) {
            // ineffective on this call; give up only after 16 such calls in a row (cameras may alternate
            // between views where it helps and views where it does not)
            if (++hdr->depth_cull_cooldown >= 16) {
                hdr->depth_cull = 0;
                hdr->depth_cull_cooldown = 32;
            }
        } else {
            hdr->depth_cull_cooldown = 0;
        }
        hdr->total_valid += sv;
        hdr->total_tsdf_valid += stv;
        hdr->total_blocks += n_blocks;
        hdr->total_calls += 1ull;
        hdr->total_frames += (unsigned long long)B;
        hdr->last_slot = p.slot;
        sc->k3_next = 0;            // K3W's chunk dispenser
        sc->n_blocks = n_proc;      // list segments K3 searches
        sc->n_frustum_blocks = n_frustum;
        sc->k2_done = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// K0: which frames of a sequence can touch this slab at all?  (multi-GPU x-slabs: every rank is handed every
// frame, clip_seem_fusion.py:305-313 has no notion of slabs.)  One CTA per frame: depth-tile maxima into shared
// memory, then K1's frustum test and K2's depth-reach test for the slab's blocks until one passes.  A frame
// that fails for every block cannot produce a tsdf_valid voxel here, so the host drops it from the windows.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) frame_reach_kernel(const FusionParams p, const saf_frame* __restrict__ frames,
                                                          int32_t n_frames, uint32_t* __restrict__ reach)
{
    __shared__ float s_tile[kMaxDepthTiles];
    __shared__ float s_max[8];
    const int fi = blockIdx.x;
    if (fi >= n_frames) return;
    const saf_frame& f = frames[fi];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ts = 1 << p.tile_shift, ntiles = p.ntx * p.nty;
    float wmax = 0.0f;
    for (int t = warp; t < ntiles; t += 8) {
        const int x0 = (t % p.ntx) * ts, y0 = (t / p.ntx) * ts;
        const int x1 = min(p.W, x0 + ts);
        float dm = 0.0f;
        for (int y = y0 + lane; y < min(p.H, y0 + ts); y += 32)
            for (int x = x0; x < x1; ++x) dm = fmaxf(dm, load_depth(f, (size_t)y * p.W + x));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dm = fmaxf(dm, __shfl_xor_sync(0xffffffffu, dm, o));
        if (lane == 0) s_tile[t] = dm;
        wmax = fmaxf(wmax, dm);
    }
    if (lane == 0) s_max[warp] = wmax;
    __syncthreads();
    float image_max = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) image_max = fmaxf(image_max, s_max[w]);
    Geom g;
    load_geom(f, g);
    const float fW = (float)p.W, fH = (float)p.H;
    for (uint32_t b0 = 0; b0 < p.nblocks_total; b0 += 256) {
        const uint32_t blk = b0 + threadIdx.x;
        bool hit = false;
        if (blk < p.nblocks_total) {
            const uint32_t bz = blk % p.nb[2];
            const uint32_t by = (blk / p.nb[2]) % p.nb[1];
            const uint32_t bx = blk / (p.nb[2] * p.nb[1]);
            float cx, cy, cz, r, half[3];
            block_sphere(p, bx, by, bz, cx, cy, cz, r, half);
            if (block_maybe_visible(g, cx, cy, cz, r, fW, fH)) {
                const float dfar = block_depth_bound(p, g, cx, cy, cz, r, s_tile, nullptr, image_max);
                hit = !(block_z_min(g, cx, cy, cz, half) > (dfar + p.trunc) * 1.001f + 1e-5f);  // NaN -> keep
            }
        }
        if (__syncthreads_or(hit)) {
            if (threadIdx.x == 0) reach[fi] = 1u;
            return;
        }
    }
    if (threadIdx.x == 0) reach[fi] = 0u;
}

__global__ void add_skipped_frames_kernel(WsHeader* hdr, unsigned long long n) { hdr->total_frames += n; }

// ---------------------------------------------------------------------------------------------
// K3: per-voxel feature / rgb / label accumulation (clip_seem_fusion.py:751-822)
// ---------------------------------------------------------------------------------------------

#ifndef K3_MAXNREG
#define K3_MAXNREG 64
#endif
constexpr int kK3Warps = 16;
constexpr int kK3Threads = kK3Warps * 32;

// flat list position -> entry: position i lives in visible block r = upper_bound(offset, i) - 1
__device__ __forceinline__ ValidEntry fetch_entry(const ValidEntry* __restrict__ list, const uint32_t* __restrict__ off,
                                                  uint32_t n_blocks, uint32_t i)
{
    uint32_t lo = 0, hi = n_blocks;  // invariant: off[lo] <= i < off[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= i)
            lo = mid;
        else
            hi = mid;
    }
    return list[(uint64_t)lo * kBlockVoxels + (i - __ldg(off + lo))];
}

// rgb running average, label counter and weight of ONE voxel, done by one lane
// (clip_seem_fusion.py:786-798, 808-822; nearest rgb: clipfusion.py:701-706)
__device__ __forceinline__ void update_small_state(const FusionParams& p, const saf_frame& f, const ValidEntry& e, int w)
{
    const float a = __frcp_rn(__int2float_rn(w + 1));
    const float b = __fmul_rn(__int2float_rn(w), a);
    float smp[3];
    const int px = nearest_index(e.gx, p.W), py = nearest_index(e.gy, p.H);
    if (p.rgb_mode == SAF_RGB_NEAREST) {
#pragma unroll
        for (int c = 0; c < 3; ++c) smp[c] = (px >= 0 && py >= 0) ? load_rgb(f, ((size_t)py * p.W + px) * 3 + c) : 0.0f;
    } else {
        Taps t;
        bilinear_setup(e.gx, e.gy, p.W, p.H, t);
        float val[4][3];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int c = 0; c < 3; ++c) val[k][c] = t.idx[k] >= 0 ? load_rgb(f, (size_t)t.idx[k] * 3 + c) : 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) smp[c] = bilinear_mix(val[0][c], val[1][c], val[2][c], val[3][c], t.w);
    }
    float* dst = p.vol.rgb + (size_t)e.voxel * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) dst[c] = __fadd_rn(__fmul_rn(smp[c], a), __fmul_rn(dst[c], b));
    if (p.vol.labels_one_hot && f.seg) {
        const float lf = (px >= 0 && py >= 0) ? load_class_id(f.seg, f.seg_dtype, py * p.W + px) : 0.0f;
        const long long id = (long long)lf;
        if (id >= 0 && id < p.vol.n_classes)
            p.vol.labels_one_hot[(size_t)e.voxel * p.vol.n_classes + id] += 1;
        else
            atomicOr(&p.hdr->error_flags, SAF_FLAG_BAD_CLASS_ID);
    }
    p.vol.weight[e.voxel] = w + 1;
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

// Shared-memory plan of feature_accumulate_kernel:
//   [ table: R*C floats (TABLE_SMEM) ][ ring: kK3Warps * NST rows of C floats ][ mbarriers ]
// CHUNKS = C/128 float4 per lane; NST = ring stages per warp (rows in flight per warp).
template <int CHUNKS, int NST, bool TABLE_SMEM>
__global__ void __maxnreg__(K3_MAXNREG)
feature_accumulate_kernel(const FusionParams p, const float* __restrict__ table, int64_t table_stride_r, int table_tma)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int C = CHUNKS * 128;
    const saf_frame& f = p.frames[p.frame_index];
    const int R = f.npy * f.npx;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const SlotCounters* sc = &p.hdr->slot[p.slot];
    const uint32_t n = sc->n_valid[p.frame_index];
    if (n == 0) return;
    const uint32_t n_blocks = sc->n_blocks;

    float* tab = reinterpret_cast<float*>(smem_raw);
    float* ring = tab + (TABLE_SMEM ? (size_t)R * C : 0);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)kK3Warps * NST * C);  // [0] table, then [warp][stage]
    float* my_ring = ring + (size_t)warp * NST * C;
    uint64_t* my_bars = bars + 1 + warp * NST;

    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        for (int i = 0; i < kK3Warps * NST; ++i) mbar_init(&bars[1 + i], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (TABLE_SMEM) {
        if (table_tma) {
            if (warp == 0) {
                if (lane == 0) mbar_arrive_expect_tx(&bars[0], (uint32_t)R * (uint32_t)C * 4u);
                __syncwarp();
                for (int r = lane; r < R; r += 32)
                    tma_bulk_g2s(tab + (size_t)r * C, table + (size_t)r * table_stride_r, (uint32_t)C * 4u, &bars[0]);
            }
        } else {
            for (int64_t e = threadIdx.x; e < (int64_t)R * C; e += kK3Threads) {
                const int64_t r = e / C, c = e - r * C;
                tab[e] = table[r * table_stride_r + c];
            }
            __syncthreads();
        }
    }
    bool table_ready = !(TABLE_SMEM && table_tma);

    // each warp owns a contiguous run of the (spatially ordered) list: its voxels are neighbours, so the
    // rows it streams are close in memory and the lanes' small-state accesses share sectors
    const uint32_t nwarps = gridDim.x * kK3Warps;
    const uint32_t gwarp = blockIdx.x * kK3Warps + warp;
    const uint32_t per_warp = (n + nwarps - 1) / nwarps;
    const uint32_t first = min(n, gwarp * per_warp);
    const uint32_t k_total = min(n, first + per_warp) - first;  // entries this warp owns
    const ValidEntry* __restrict__ list = p.lists + (uint64_t)p.frame_index * p.list_cap;
    const uint32_t* __restrict__ off = p.blk_offset + (uint64_t)p.frame_index * (p.nblocks_total + 1);
    const float4* tab4 = reinterpret_cast<const float4*>(TABLE_SMEM ? tab : table);
    const int64_t tab_row4 = TABLE_SMEM ? (C / 4) : (table_stride_r / 4);
    uint32_t t_use = 0;  // rows this warp has pushed through the ring (stage = t % NST, parity = (t / NST) & 1)

    for (uint32_t kb = 0; kb < k_total; kb += 32) {
        const uint32_t cnt = min(32u, k_total - kb);
        // lane l owns entry kb + l of this batch: fetch it, then do its small state
        ValidEntry mine;
        mine.voxel = 0;
        mine.gx = mine.gy = 0.0f;
        int my_w = 0;
        if (lane < cnt) {
            mine = fetch_entry(list, off, n_blocks, first + kb + lane);
            my_w = p.vol.weight[mine.voxel];
        }
        // start the first rows of the batch before spending time on the small state
        const uint32_t pre = min((uint32_t)NST, cnt);
        for (uint32_t q = 0; q < pre; ++q) {
            const uint32_t vq = __shfl_sync(0xffffffffu, mine.voxel, q);
            if (lane == 0) {
                const uint32_t s = (t_use + q) % NST;
                mbar_arrive_expect_tx(&my_bars[s], (uint32_t)C * 4u);
                tma_bulk_g2s(my_ring + (size_t)s * C, p.vol.clip_feat + (size_t)vq * C, (uint32_t)C * 4u, &my_bars[s]);
            }
        }
        if (lane < cnt) update_small_state(p, f, mine, my_w);
        if (!table_ready) {
            mbar_wait(&bars[0], 0);
            table_ready = true;
        }
        for (uint32_t q = 0; q < cnt; ++q) {
            const uint32_t s = t_use % NST;
            const uint32_t parity = (t_use / NST) & 1u;
            const uint32_t vq = __shfl_sync(0xffffffffu, mine.voxel, q);
            const float gx = __shfl_sync(0xffffffffu, mine.gx, q);
            const float gy = __shfl_sync(0xffffffffu, mine.gy, q);
            const int w = __shfl_sync(0xffffffffu, my_w, q);
            // clip_seem_fusion.py:808-810
            const float a = __frcp_rn(__int2float_rn(w + 1));
            const float b = __fmul_rn(__int2float_rn(w), a);
            Taps t;
            feature_taps(f, gx, gy, p.W, p.H, &p.hdr->error_flags, t);
            float4* row = reinterpret_cast<float4*>(p.vol.clip_feat + (size_t)vq * C);
            const float4* old4 = reinterpret_cast<const float4*>(my_ring + (size_t)s * C);
            const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
            mbar_wait(&my_bars[s], parity);
#pragma unroll
            for (int j = 0; j < CHUNKS; ++j) {
                const int col = j * 32 + lane;
                const float4 t0 = t.idx[0] >= 0 ? tab4[t.idx[0] * tab_row4 + col] : zero;
                const float4 t1 = t.idx[1] >= 0 ? tab4[t.idx[1] * tab_row4 + col] : zero;
                const float4 t2 = t.idx[2] >= 0 ? tab4[t.idx[2] * tab_row4 + col] : zero;
                const float4 t3 = t.idx[3] >= 0 ? tab4[t.idx[3] * tab_row4 + col] : zero;
                st_stream_f4(row + col, blend4(mix4(t0, t1, t2, t3, t.w), old4[col], a, b));
            }
            __syncwarp();  // every lane has consumed stage s
            if (q + NST < cnt) {
                const uint32_t vn = __shfl_sync(0xffffffffu, mine.voxel, q + NST);
                if (lane == 0) {
                    mbar_arrive_expect_tx(&my_bars[s], (uint32_t)C * 4u);
                    tma_bulk_g2s(my_ring + (size_t)s * C, p.vol.clip_feat + (size_t)vn * C, (uint32_t)C * 4u, &my_bars[s]);
                }
            }
            ++t_use;
        }
    }
    if (!table_ready) mbar_wait(&bars[0], 0);  // never leave a bulk copy in flight at exit
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
    const SlotCounters* sc = &p.hdr->slot[p.slot];
    const uint32_t n = sc->n_valid[p.frame_index];
    const uint32_t n_blocks = sc->n_blocks;
    const uint32_t nwarps = gridDim.x * kK3Warps;
    const uint32_t gwarp = blockIdx.x * kK3Warps + (threadIdx.x >> 5);
    const ValidEntry* __restrict__ list = p.lists + (uint64_t)p.frame_index * p.list_cap;
    const uint32_t* __restrict__ off = p.blk_offset + (uint64_t)p.frame_index * (p.nblocks_total + 1);
    for (uint64_t i = gwarp; i < n; i += nwarps) {
        const ValidEntry e = fetch_entry(list, off, n_blocks, (uint32_t)i);
        const int w = p.vol.weight[e.voxel];
        const float a = __frcp_rn(__int2float_rn(w + 1));
        const float b = __fmul_rn(__int2float_rn(w), a);
        Taps t;
        feature_taps(f, e.gx, e.gy, p.W, p.H, &p.hdr->error_flags, t);
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
                st_stream_f4(reinterpret_cast<float4*>(row) + col, blend4(mix4(tv[0], tv[1], tv[2], tv[3], t.w), o, a, b));
            }
        } else {
            for (int c = lane; c < C; c += 32) {
                float tv[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) tv[k] = t.idx[k] >= 0 ? __ldg(table + t.idx[k] * table_stride_r + c) : 0.0f;
                const float smp = bilinear_mix(tv[0], tv[1], tv[2], tv[3], t.w);
                row[c] = __fadd_rn(__fmul_rn(smp, a), __fmul_rn(row[c], b));
            }
        }
        __syncwarp();
        if (lane == 0) update_small_state(p, f, e, w);
    }
}

// ---------------------------------------------------------------------------------------------
// K3W: window mode of K3 (saf_integrate_sequence).  One warp per voxel that is valid in at least one of
// the window's frames: the feature row is pulled in once (TMA ring), every valid frame's update is applied
// to it in registers in frame order - the exact arithmetic of that many single-frame calls - and it is
// written back once.  DRAM traffic per voxel update drops by the number of window frames that see the
// voxel (consecutive frames of a scan overlap almost completely); the per-frame table rows are read through
// L1 (the 16 warps of a CTA walk neighbouring voxels, so they share the same few rows of each frame).
// ---------------------------------------------------------------------------------------------

struct WindowTables {
    const float* ptr[SAF_MAX_BATCH];     // [R,C]-row view of each frame's feature image
    int64_t stride_r[SAF_MAX_BATCH];
};

__device__ __forceinline__ void sample_rgb(const FusionParams& p, const saf_frame& f, float gx, float gy, int px, int py,
                                           float (&smp)[3])
{
    if (p.rgb_mode == SAF_RGB_NEAREST) {
#pragma unroll
        for (int c = 0; c < 3; ++c) smp[c] = (px >= 0 && py >= 0) ? load_rgb(f, ((size_t)py * p.W + px) * 3 + c) : 0.0f;
    } else {
        Taps t;
        bilinear_setup(gx, gy, p.W, p.H, t);
        float val[4][3];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int c = 0; c < 3; ++c) val[k][c] = t.idx[k] >= 0 ? load_rgb(f, (size_t)t.idx[k] * 3 + c) : 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) smp[c] = bilinear_mix(val[0][c], val[1][c], val[2][c], val[3][c], t.w);
    }
}

// One CTA of kW3Warps warps per SM: the kernel is latency-, not HBM-bound (24 warps x 80 registers for the
// one-voxel-per-pass kernel, 16 x 128 for the pair kernel).
#ifndef K3W2_MAXNREG
#define K3W2_MAXNREG 128
#endif
#ifndef K3W2_WARPS
#define K3W2_WARPS 16
#endif
#ifndef K3W_CHUNK
#define K3W_CHUNK 8
#endif
constexpr int kW3Chunk = K3W_CHUNK;                  // union-list entries a warp claims at a time (dynamic balancing: the
                                             // number of updates per entry varies from 1 to the window length)

template <int CHUNKS, int kW3Warps>
__global__ void __maxnreg__(80)
feature_accumulate_window_kernel(const __grid_constant__ FusionParams p, const __grid_constant__ WindowTables wt)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int C = CHUNKS * 128;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SlotCounters* sc = &p.hdr->slot[p.slot];
    const uint32_t n = sc->n_union;
    if (n == 0) return;
    const uint32_t n_blocks = sc->n_blocks;
    const int B = p.batch;

    // [ one feature row per warp (TMA landing zone) ][ coords: warp x chunk x frame ][ one mbarrier per warp ]
    float* ring = reinterpret_cast<float*>(smem_raw);
    float2* coords = reinterpret_cast<float2*>(ring + (size_t)kW3Warps * C);
    uint64_t* bars = reinterpret_cast<uint64_t*>(coords + (size_t)kW3Warps * kW3Chunk * SAF_MAX_BATCH);
    float* my_row = ring + (size_t)warp * C;
    float2* my_coords = coords + (size_t)warp * kW3Chunk * SAF_MAX_BATCH;
    uint64_t* my_bar = bars + warp;
    __shared__ uint32_t s_ticket;
    __shared__ volatile uint32_t s_sbase[8], s_ssize[8], s_sgen[8];
    if (threadIdx.x == 0) {
        for (int i = 0; i < kW3Warps; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
        s_ticket = 0;
    }
    if (threadIdx.x < 8) s_sgen[threadIdx.x] = 0xffffffffu;
    __syncthreads();

    const uint32_t* __restrict__ off = p.blk_offset;
    uint32_t t_use = 0;  // rows this warp has received (mbarrier parity = t_use & 1)

    // Work dispenser.  Warps take tickets from a CTA counter; every kW3Warps tickets form one "super chunk" of
    // kW3Warps * kW3Chunk consecutive list entries that the warp holding its first ticket claims from the global
    // counter.  The CTA's warps therefore walk neighbouring voxels (same few table rows per frame -> L1 hits),
    // while the units stay small enough to balance the very uneven number of updates per entry.
    for (;;) {
        uint32_t base = 0, chunk = kW3Chunk;
        if (lane == 0) {
            const uint32_t t = atomicAdd(&s_ticket, 1u);
            const uint32_t k = t / kW3Warps, c = t % kW3Warps, slot = k & 7u;
            if (c == 0) {
                // towards the end of the list the units are halved, so that the CTAs finish closer together
                const uint32_t seen = *reinterpret_cast<volatile uint32_t*>(&sc->k3_next);
                const uint32_t csz = (seen < n && n - seen < gridDim.x * (uint32_t)(kW3Warps * kW3Chunk) * 2u)
                                         ? (uint32_t)kW3Chunk / 2u : (uint32_t)kW3Chunk;
                base = atomicAdd(&sc->k3_next, (uint32_t)kW3Warps * csz);
                s_sbase[slot] = base;
                s_ssize[slot] = csz;
                __threadfence_block();
                s_sgen[slot] = k;
                chunk = csz;
            } else {
                while (s_sgen[slot] != k) {
                }
                __threadfence_block();
                base = s_sbase[slot];
                chunk = s_ssize[slot];
            }
            base = min(base, n) + c * chunk;
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if (base >= n) break;
        const uint32_t cnt = min(chunk, n - base);
        uint32_t my_voxel = 0, my_mask = 0;
        int my_w = 0;
        if (lane < cnt) {
            const uint32_t i = base + lane;
            uint32_t lo = 0, hi = n_blocks;  // off[lo] <= i < off[hi]
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (__ldg(off + mid) <= i)
                    lo = mid;
                else
                    hi = mid;
            }
            const WinEntry e = p.ulist[(uint64_t)lo * kBlockVoxels + (i - __ldg(off + lo))];
            my_voxel = e.voxel;
            my_mask = e.mask_local & 0xffffu;
            my_w = p.vol.weight[my_voxel];
            const float2* src = p.wcoords + (uint64_t)lo * B * kBlockVoxels + (e.mask_local >> 16);
            for (uint32_t mm = my_mask; mm; mm &= mm - 1u) {
                const int b = __ffs(mm) - 1;
                my_coords[lane * SAF_MAX_BATCH + b] = src[(size_t)b * kBlockVoxels];
            }
        }
        {   // the landing zone is free (its last row went to registers): start the chunk's first row
            const uint32_t v0 = __shfl_sync(0xffffffffu, my_voxel, 0);
            if (lane == 0) {
                fence_proxy_async();
                mbar_arrive_expect_tx(my_bar, (uint32_t)C * 4u);
                tma_bulk_g2s(my_row, p.vol.clip_feat + (size_t)v0 * C, (uint32_t)C * 4u, my_bar);
            }
        }
        // rgb running average, label counters and weight of this lane's voxel, frame by frame
        // (clip_seem_fusion.py:786-798, 808-822)
        if (lane < cnt) {
            float* dst = p.vol.rgb + (size_t)my_voxel * 3;
            float acc[3] = {dst[0], dst[1], dst[2]};
            int w = my_w;
            for (uint32_t mm = my_mask; mm; mm &= mm - 1u) {
                const int b = __ffs(mm) - 1;
                const saf_frame& f = p.frames[b];
                const float2 g = my_coords[lane * SAF_MAX_BATCH + b];
                const float a = __frcp_rn(__int2float_rn(w + 1));
                const float bb = __fmul_rn(__int2float_rn(w), a);
                const int px = nearest_index(g.x, p.W), py = nearest_index(g.y, p.H);
                float smp[3];
                sample_rgb(p, f, g.x, g.y, px, py, smp);
#pragma unroll
                for (int c = 0; c < 3; ++c) acc[c] = __fadd_rn(__fmul_rn(smp[c], a), __fmul_rn(acc[c], bb));
                if (p.vol.labels_one_hot && f.seg) {
                    const float lf = (px >= 0 && py >= 0) ? load_class_id(f.seg, f.seg_dtype, py * p.W + px) : 0.0f;
                    const long long id = (long long)lf;
                    if (id >= 0 && id < p.vol.n_classes)
                        p.vol.labels_one_hot[(size_t)my_voxel * p.vol.n_classes + id] += 1;
                    else
                        atomicOr(&p.hdr->error_flags, SAF_FLAG_BAD_CLASS_ID);
                }
                ++w;
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) dst[c] = acc[c];
            p.vol.weight[my_voxel] = w;
        }
        __syncwarp();  // my_coords written by the lanes, read by the whole warp below
        for (uint32_t q = 0; q < cnt; ++q) {
            const uint32_t vq = __shfl_sync(0xffffffffu, my_voxel, q);
            const uint32_t mq = __shfl_sync(0xffffffffu, my_mask, q);
            int w = __shfl_sync(0xffffffffu, my_w, q);
            float4* row = reinterpret_cast<float4*>(p.vol.clip_feat + (size_t)vq * C);
            const float4* old4 = reinterpret_cast<const float4*>(my_row);
            mbar_wait(my_bar, t_use & 1u);
            ++t_use;
            float4 acc[CHUNKS];
#pragma unroll
            for (int j = 0; j < CHUNKS; ++j) acc[j] = old4[j * 32 + lane];
            __syncwarp();  // the row is in registers: the landing zone can take the next voxel's row already
            if (q + 1 < cnt) {
                const uint32_t vn = __shfl_sync(0xffffffffu, my_voxel, q + 1);
                if (lane == 0) {
                    fence_proxy_async();
                    mbar_arrive_expect_tx(my_bar, (uint32_t)C * 4u);
                    tma_bulk_g2s(my_row, p.vol.clip_feat + (size_t)vn * C, (uint32_t)C * 4u, my_bar);
                }
            }
            for (uint32_t mm = mq; mm; mm &= mm - 1u) {
                const int b = __ffs(mm) - 1;
                const float2 g = my_coords[q * SAF_MAX_BATCH + b];
                // clip_seem_fusion.py:808-810
                const float a = __frcp_rn(__int2float_rn(w + 1));
                const float bb = __fmul_rn(__int2float_rn(w), a);
                Taps t;
                feature_taps_padded(p.frames[b], g.x, g.y, p.W, p.H, &p.hdr->error_flags, t);
                const float4* tab4 = reinterpret_cast<const float4*>(wt.ptr[b]) + lane;   // zero-bordered [R',C] rows
                const float4* r0 = tab4 + t.idx[0] * (C / 4);
                const float4* r1 = tab4 + t.idx[1] * (C / 4);
                const float4* r2 = tab4 + t.idx[2] * (C / 4);
                const float4* r3 = tab4 + t.idx[3] * (C / 4);
#pragma unroll
                for (int j = 0; j < CHUNKS; ++j) {
                    const float4 t0 = __ldg(r0 + j * 32), t1 = __ldg(r1 + j * 32);
                    const float4 t2 = __ldg(r2 + j * 32), t3 = __ldg(r3 + j * 32);
                    acc[j] = blend4(mix4(t0, t1, t2, t3, t.w), acc[j], a, bb);
                }
                ++w;
            }
#pragma unroll
            for (int j = 0; j < CHUNKS; ++j) st_stream_f4(row + j * 32 + lane, acc[j]);
        }
        __syncwarp();  // the next chunk overwrites my_coords
    }
}

template <int CHUNKS, int kW3Warps>
__global__ void __maxnreg__(K3W2_MAXNREG)
feature_accumulate_window_pair_kernel(const __grid_constant__ FusionParams p, const __grid_constant__ WindowTables wt)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int C = CHUNKS * 128;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SlotCounters* sc = &p.hdr->slot[p.slot];
    const uint32_t n = sc->n_union;
    if (n == 0) return;
    const uint32_t n_blocks = sc->n_blocks;
    const int B = p.batch;

    // [ one feature row per warp (TMA landing zone) ][ coords: warp x chunk x frame ][ one mbarrier per warp ]
    float* ring = reinterpret_cast<float*>(smem_raw);
    float2* coords = reinterpret_cast<float2*>(ring + (size_t)kW3Warps * 2 * C);
    uint64_t* bars = reinterpret_cast<uint64_t*>(coords + (size_t)kW3Warps * kW3Chunk * SAF_MAX_BATCH);
    float* my_row = ring + (size_t)warp * 2 * C;   // two landing rows: the pair of voxels the warp updates together
    float2* my_coords = coords + (size_t)warp * kW3Chunk * SAF_MAX_BATCH;
    uint64_t* my_bar = bars + warp;
    __shared__ uint32_t s_ticket;
    __shared__ volatile uint32_t s_sbase[8], s_ssize[8], s_sgen[8];
    if (threadIdx.x == 0) {
        for (int i = 0; i < kW3Warps; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
        s_ticket = 0;
    }
    if (threadIdx.x < 8) s_sgen[threadIdx.x] = 0xffffffffu;
    __syncthreads();

    const uint32_t* __restrict__ off = p.blk_offset;
    uint32_t t_use = 0;  // rows this warp has received (mbarrier parity = t_use & 1)

    // Work dispenser.  Warps take tickets from a CTA counter; every kW3Warps tickets form one "super chunk" of
    // kW3Warps * kW3Chunk consecutive list entries that the warp holding its first ticket claims from the global
    // counter.  The CTA's warps therefore walk neighbouring voxels (same few table rows per frame -> L1 hits),
    // while the units stay small enough to balance the very uneven number of updates per entry.
    for (;;) {
        uint32_t base = 0, chunk = kW3Chunk;
        if (lane == 0) {
            const uint32_t t = atomicAdd(&s_ticket, 1u);
            const uint32_t k = t / kW3Warps, c = t % kW3Warps, slot = k & 7u;
            if (c == 0) {
                // towards the end of the list the units are halved, so that the CTAs finish closer together
                const uint32_t seen = *reinterpret_cast<volatile uint32_t*>(&sc->k3_next);
                const uint32_t csz = (seen < n && n - seen < gridDim.x * (uint32_t)(kW3Warps * kW3Chunk) * 2u)
                                         ? (uint32_t)kW3Chunk / 2u : (uint32_t)kW3Chunk;
                base = atomicAdd(&sc->k3_next, (uint32_t)kW3Warps * csz);
                s_sbase[slot] = base;
                s_ssize[slot] = csz;
                __threadfence_block();
                s_sgen[slot] = k;
                chunk = csz;
            } else {
                while (s_sgen[slot] != k) {
                }
                __threadfence_block();
                base = s_sbase[slot];
                chunk = s_ssize[slot];
            }
            base = min(base, n) + c * chunk;
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if (base >= n) break;
        const uint32_t cnt = min(chunk, n - base);
        uint32_t my_voxel = 0, my_mask = 0;
        int my_w = 0;
        if (lane < cnt) {
            const uint32_t i = base + lane;
            uint32_t lo = 0, hi = n_blocks;  // off[lo] <= i < off[hi]
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (__ldg(off + mid) <= i)
                    lo = mid;
                else
                    hi = mid;
            }
            const WinEntry e = p.ulist[(uint64_t)lo * kBlockVoxels + (i - __ldg(off + lo))];
            my_voxel = e.voxel;
            my_mask = e.mask_local & 0xffffu;
            my_w = p.vol.weight[my_voxel];
            const float2* src = p.wcoords + (uint64_t)lo * B * kBlockVoxels + (e.mask_local >> 16);
            for (uint32_t mm = my_mask; mm; mm &= mm - 1u) {
                const int b = __ffs(mm) - 1;
                my_coords[lane * SAF_MAX_BATCH + b] = src[(size_t)b * kBlockVoxels];
            }
        }
        {   // the landing zone is free (its last rows went to registers): start the chunk's first pair of rows
            const uint32_t v0 = __shfl_sync(0xffffffffu, my_voxel, 0);
            const uint32_t v1 = __shfl_sync(0xffffffffu, my_voxel, 1);
            if (lane == 0) {
                fence_proxy_async();
                mbar_arrive_expect_tx(my_bar, (cnt > 1 ? 2u : 1u) * (uint32_t)C * 4u);
                tma_bulk_g2s(my_row, p.vol.clip_feat + (size_t)v0 * C, (uint32_t)C * 4u, my_bar);
                if (cnt > 1) tma_bulk_g2s(my_row + C, p.vol.clip_feat + (size_t)v1 * C, (uint32_t)C * 4u, my_bar);
            }
        }
        // rgb running average, label counters and weight of this lane's voxel, frame by frame
        // (clip_seem_fusion.py:786-798, 808-822)
        if (lane < cnt) {
            float* dst = p.vol.rgb + (size_t)my_voxel * 3;
            float acc[3] = {dst[0], dst[1], dst[2]};
            int w = my_w;
            for (uint32_t mm = my_mask; mm; mm &= mm - 1u) {
                const int b = __ffs(mm) - 1;
                const saf_frame& f = p.frames[b];
                const float2 g = my_coords[lane * SAF_MAX_BATCH + b];
                const float a = __frcp_rn(__int2float_rn(w + 1));
                const float bb = __fmul_rn(__int2float_rn(w), a);
                const int px = nearest_index(g.x, p.W), py = nearest_index(g.y, p.H);
                float smp[3];
                sample_rgb(p, f, g.x, g.y, px, py, smp);
#pragma unroll
                for (int c = 0; c < 3; ++c) acc[c] = __fadd_rn(__fmul_rn(smp[c], a), __fmul_rn(acc[c], bb));
                if (p.vol.labels_one_hot && f.seg) {
                    const float lf = (px >= 0 && py >= 0) ? load_class_id(f.seg, f.seg_dtype, py * p.W + px) : 0.0f;
                    const long long id = (long long)lf;
                    if (id >= 0 && id < p.vol.n_classes)
                        p.vol.labels_one_hot[(size_t)my_voxel * p.vol.n_classes + id] += 1;
                    else
                        atomicOr(&p.hdr->error_flags, SAF_FLAG_BAD_CLASS_ID);
                }
                ++w;
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) dst[c] = acc[c];
            p.vol.weight[my_voxel] = w;
        }
        __syncwarp();  // my_coords written by the lanes, read by the whole warp below
        for (uint32_t q = 0; q < cnt; q += 2) {
            // two neighbouring voxels per pass: they nearly always use the same four table rows in a frame, so the
            // rows are loaded once for both, and the two running averages are independent dependency chains
            const bool two = q + 1 < cnt;
            const uint32_t v0 = __shfl_sync(0xffffffffu, my_voxel, q);
            const uint32_t v1 = __shfl_sync(0xffffffffu, my_voxel, two ? q + 1 : q);
            const uint32_t m0 = __shfl_sync(0xffffffffu, my_mask, q);
            const uint32_t m1 = two ? __shfl_sync(0xffffffffu, my_mask, q + 1) : 0u;
            int w0 = __shfl_sync(0xffffffffu, my_w, q);
            int w1 = __shfl_sync(0xffffffffu, my_w, two ? q + 1 : q);
            const float4* old4 = reinterpret_cast<const float4*>(my_row);
            mbar_wait(my_bar, t_use & 1u);
            ++t_use;
            float4 acc0[CHUNKS], acc1[CHUNKS];
#pragma unroll
            for (int j = 0; j < CHUNKS; ++j) {
                acc0[j] = old4[j * 32 + lane];
                acc1[j] = old4[(two ? C / 4 : 0) + j * 32 + lane];
            }
            __syncwarp();  // both rows are in registers: the landing zone can take the next pair already
            if (q + 2 < cnt) {
                const bool two_n = q + 3 < cnt;
                const uint32_t n0 = __shfl_sync(0xffffffffu, my_voxel, q + 2);
                const uint32_t n1 = __shfl_sync(0xffffffffu, my_voxel, two_n ? q + 3 : q + 2);
                if (lane == 0) {
                    fence_proxy_async();
                    mbar_arrive_expect_tx(my_bar, (two_n ? 2u : 1u) * (uint32_t)C * 4u);
                    tma_bulk_g2s(my_row, p.vol.clip_feat + (size_t)n0 * C, (uint32_t)C * 4u, my_bar);
                    if (two_n) tma_bulk_g2s(my_row + C, p.vol.clip_feat + (size_t)n1 * C, (uint32_t)C * 4u, my_bar);
                }
            }
            for (uint32_t mm = m0 | m1; mm; mm &= mm - 1u) {
                const int b = __ffs(mm) - 1;
                const bool in0 = (m0 >> b) & 1u, in1 = (m1 >> b) & 1u;
                const float4* tab4 = reinterpret_cast<const float4*>(wt.ptr[b]) + lane;   // zero-bordered [R',C] rows
                Taps t0, t1;
                float a0 = 0.f, b0 = 0.f, a1 = 0.f, b1 = 0.f;
                if (in0) {   // clip_seem_fusion.py:808-810
                    const float2 g = my_coords[q * SAF_MAX_BATCH + b];
                    feature_taps_padded(p.frames[b], g.x, g.y, p.W, p.H, &p.hdr->error_flags, t0);
                    a0 = __frcp_rn(__int2float_rn(w0 + 1));
                    b0 = __fmul_rn(__int2float_rn(w0), a0);
                    ++w0;
                }
                if (in1) {
                    const float2 g = my_coords[(q + 1) * SAF_MAX_BATCH + b];
                    feature_taps_padded(p.frames[b], g.x, g.y, p.W, p.H, &p.hdr->error_flags, t1);
                    a1 = __frcp_rn(__int2float_rn(w1 + 1));
                    b1 = __fmul_rn(__int2float_rn(w1), a1);
                    ++w1;
                }
                if (in0 && in1 && t0.idx[0] == t1.idx[0] && t0.idx[1] == t1.idx[1] && t0.idx[2] == t1.idx[2] &&
                    t0.idx[3] == t1.idx[3]) {
                    const float4* r0 = tab4 + t0.idx[0] * (C / 4);
                    const float4* r1 = tab4 + t0.idx[1] * (C / 4);
                    const float4* r2 = tab4 + t0.idx[2] * (C / 4);
                    const float4* r3 = tab4 + t0.idx[3] * (C / 4);
#pragma unroll
                    for (int j = 0; j < CHUNKS; ++j) {
                        const float4 x0 = __ldg(r0 + j * 32), x1 = __ldg(r1 + j * 32);
                        const float4 x2 = __ldg(r2 + j * 32), x3 = __ldg(r3 + j * 32);
                        acc0[j] = blend4(mix4(x0, x1, x2, x3, t0.w), acc0[j], a0, b0);
                        acc1[j] = blend4(mix4(x0, x1, x2, x3, t1.w), acc1[j], a1, b1);
                    }
                } else {
                    if (in0) {
                        const float4* r0 = tab4 + t0.idx[0] * (C / 4);
                        const float4* r1 = tab4 + t0.idx[1] * (C / 4);
                        const float4* r2 = tab4 + t0.idx[2] * (C / 4);
                        const float4* r3 = tab4 + t0.idx[3] * (C / 4);
#pragma unroll
                        for (int j = 0; j < CHUNKS; ++j) {
                            const float4 x0 = __ldg(r0 + j * 32), x1 = __ldg(r1 + j * 32);
                            const float4 x2 = __ldg(r2 + j * 32), x3 = __ldg(r3 + j * 32);
                            acc0[j] = blend4(mix4(x0, x1, x2, x3, t0.w), acc0[j], a0, b0);
                        }
                    }
                    if (in1) {
                        const float4* r0 = tab4 + t1.idx[0] * (C / 4);
                        const float4* r1 = tab4 + t1.idx[1] * (C / 4);
                        const float4* r2 = tab4 + t1.idx[2] * (C / 4);
                        const float4* r3 = tab4 + t1.idx[3] * (C / 4);
#pragma unroll
                        for (int j = 0; j < CHUNKS; ++j) {
                            const float4 x0 = __ldg(r0 + j * 32), x1 = __ldg(r1 + j * 32);
                            const float4 x2 = __ldg(r2 + j * 32), x3 = __ldg(r3 + j * 32);
                            acc1[j] = blend4(mix4(x0, x1, x2, x3, t1.w), acc1[j], a1, b1);
                        }
                    }
                }
            }
            float4* row0 = reinterpret_cast<float4*>(p.vol.clip_feat + (size_t)v0 * C);
#pragma unroll
            for (int j = 0; j < CHUNKS; ++j) st_stream_f4(row0 + j * 32 + lane, acc0[j]);
            if (two) {
                float4* row1 = reinterpret_cast<float4*>(p.vol.clip_feat + (size_t)v1 * C);
#pragma unroll
                for (int j = 0; j < CHUNKS; ++j) st_stream_f4(row1 + j * 32 + lane, acc1[j]);
            }
        }
        __syncwarp();  // the next chunk overwrites my_coords
    }
}

// ---------------------------------------------------------------------------------------------
// K3W, tile kernel (default for C = 512 / 768 / 1024).  The pair kernel above pulls four table rows through L1 for
// every (voxel, frame) update - 12 KB of on-chip traffic per 6 KB row read-modify-write, with a table working set
// (16 frames x 4-6 rows x C floats) that does not fit next to the landing zones, so a fifth of those loads miss.
// Here the roles are swapped: a tile of G = NSET * 8 neighbouring voxels keeps its accumulators in REGISTERS for
// the whole window and the frames are walked in order, so the four table rows of a frame are loaded once per
// (frame, set of 8 voxels) and reused by every voxel of the set that the frame sees.
//
//   warps [0, NBUF)              producers: claim G consecutive union-list entries and start the TMA bulk copies of
//                                their feature rows into a landing buffer and of the tile's update metadata (per
//                                valid (frame, voxel): bilinear weights, a = 1/(w+1), b = w*a, table rows - computed
//                                ONCE per update instead of redundantly by 32 lanes) into shared memory.  The
//                                metadata was prepared, and the voxels' small state (rgb average, label counters,
//                                weight) updated, by K2T (window_tile_setup_kernel, below) just before this kernel;
//                                tiles past K2T's capacity are prepared here by the same routine (fill_tile_meta)
//   warps [NBUF, NBUF+NSET*CHUNKS)  compute: warp (set, chunk) owns the 128-channel slice `chunk` of the set's 8
//                                voxels: accumulators from the landing buffer to registers (the buffer is handed back
//                                to the producer at once), then for every frame of the window in order the exact
//                                mul/fma chains as packed f32x2 operations, the frame's table rows software-pipelined
//                                one frame ahead in registers; rows are written back with 128-bit streaming stores
//
// mbarriers: full[2*NBUF] (producer arrival + TMA bytes), rows_free[NBUF], meta_free[2*NBUF] (one arrival per compute
// warp).  Producer p fills ring positions p, p+NBUF, ...; its landing buffer is p, its metadata alternates between
// slots p and p+NBUF, so metadata is never waited for.  A producer that finds the list exhausted publishes a tile
// with n_rows = 0.  Arithmetic per voxel = the single-frame calls in frame order (clip_seem_fusion.py:800-814).
// ---------------------------------------------------------------------------------------------

#ifndef SAF_TILE_NBUF
#define SAF_TILE_NBUF 3         // producer warps = landing buffers per CTA (C = 512 / 768)
#endif
constexpr int kTileSlots = 8;    // voxels per set (accumulator registers: 8 x float4 per thread)

struct __align__(16) TileUpdate {   // one (frame, voxel) feature update
    float w[4];                     // bilinear weights nw, ne, sw, se
    float a, b;                     // clip_seem_fusion.py:808-810
    uint32_t rows;                  // the four rows of the frame's zero-bordered table, one byte each
    float one;                      // 1.0f, a multiplicand ptxas cannot see through (mix_blend2)
};

template <int NSET>
struct __align__(128) TileMeta {
    uint32_t n_rows;                                // 0: the producer found the list exhausted
    uint32_t fmask[NSET];                           // frames in which any voxel of the set is valid
    uint32_t voxel[NSET * kTileSlots];
    uint32_t prim_rows[NSET][SAF_MAX_BATCH];        // rows of the set's first valid voxel in the frame
    uint32_t alt_rows[NSET][SAF_MAX_BATCH];         // rows of its first valid voxel that samples other rows (uniform = 0)
    uint8_t vmask[NSET][SAF_MAX_BATCH];             // voxels of the set valid in the frame
    uint8_t uniform[NSET][SAF_MAX_BATCH];           // 1: every valid voxel of the set uses prim_rows in the frame
    uint8_t pmask[NSET][SAF_MAX_BATCH];             // uniform = 0: valid voxels that use prim_rows,
    uint8_t amask[NSET][SAF_MAX_BATCH];             //              and those that use alt_rows (the rest: a third cell)
    TileUpdate upd[NSET][SAF_MAX_BATCH][kTileSlots];
};

struct RowRegs {   // this thread's 4-float column of the four table rows of one frame
    f32x2_t lo[4], hi[4];
};

__device__ __forceinline__ void load_rows(RowRegs& T, const float* __restrict__ table, uint32_t rows, int C, int col4)
{
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(table + (size_t)((rows >> (8 * k)) & 0xffu) * C) + col4);
        T.lo[k] = v.x;
        T.hi[k] = v.y;
    }
}

// bilinear_mix + blend4 on one packed pair: ((t0 w0 + t1 w1) + t2 w2) + t3 w3 as mul / fma / fma / fma, then
// smp a + old b as mul, mul, add - each lane of the pair rounded like the scalar code.  The add is
// fma(x, 1.0f, y) = x + y correctly rounded (signed zeros included) with the 1.0f read from the update record: ptxas
// (12.9) contracts a packed mul + add pair into one FFMA2 even with --fmad false and explicit .rn, which would drop
// one of the reference's three roundings, and it cannot see through a multiplicand that comes from memory.
__device__ __forceinline__ f32x2_t mix_blend2(const f32x2_t (&t)[4], const f32x2_t (&w)[4], f32x2_t a, f32x2_t b,
                                              f32x2_t one, f32x2_t old)
{
    const f32x2_t smp = fma2_rn(t[3], w[3], fma2_rn(t[2], w[2], fma2_rn(t[1], w[1], mul2_rn(t[0], w[0]))));
    return fma2_rn(mul2_rn(smp, a), one, mul2_rn(old, b));
}

// One lane's voxel of a tile of the window's union list (tile = G consecutive entries starting at `base`).
struct TileLane {
    uint32_t voxel, mask, local, rank;   // slab-local voxel, frames that see it, index in its block, block rank
    int w0;                              // the voxel's weight before the window
    float rgb0[3];
};

// SMALL: the caller is K2T, which owns the window's small state: weight and rgb are read from the volume and the
// weight before the window is left in w0_list for K3W's own producers (they run while K2T already works on the next
// window and must not look at the volume's weights).
template <bool SMALL>
__device__ __forceinline__ TileLane load_tile_lane(const FusionParams& p, uint32_t n_blocks, uint32_t base, uint32_t cnt,
                                                   int lane)
{
    const uint32_t* __restrict__ off = p.blk_offset;
    // block rank of the tile's first entry: 32-ary search by the whole warp (off[lo] <= base < off[hi])
    uint32_t lo = 0, hi = n_blocks;
    while (hi - lo > 1) {
        const uint32_t step = (hi - lo + 31u) / 32u;
        const uint32_t probe = lo + (uint32_t)lane * step;
        const bool le = probe < hi && __ldg(off + probe) <= base;
        const uint32_t k = (uint32_t)__popc(__ballot_sync(0xffffffffu, le));   // >= 1: lane 0 probes lo
        lo += (k - 1u) * step;
        hi = min(hi, lo + step);
    }
    TileLane L = {0u, 0u, 0u, 0u, 0, {0.f, 0.f, 0.f}};
    if (lane < cnt) {
        const uint32_t i = base + lane;
        uint32_t r = lo;
        while (r + 1u < n_blocks && __ldg(off + r + 1u) <= i) ++r;   // the tile spans a few blocks at most
        const uint64_t slot = (uint64_t)r * kBlockVoxels + (i - __ldg(off + r));
        const WinEntry e = p.ulist[slot];
        L.voxel = e.voxel;
        L.mask = e.mask_local & 0xffffu;
        L.local = e.mask_local >> 16;
        L.rank = r;
        if (SMALL) {
            L.w0 = p.vol.weight[L.voxel];
            p.w0_list[slot] = L.w0;
            const float* src3 = p.vol.rgb + (size_t)L.voxel * 3;
            L.rgb0[0] = src3[0];
            L.rgb0[1] = src3[1];
            L.rgb0[2] = src3[2];
        } else {
            L.w0 = p.w0_list[slot];
        }
    }
    return L;
}

// The tile's metadata for the compute warps (update records, masks, prefetch rows) into M - shared memory when K3W's
// producer prepares the tile itself, global memory when K2T prepares it ahead - and the tile's small state: rgb
// running average, label counters, weight.  my_smp: G x SAF_MAX_BATCH float4 of shared-memory scratch of this warp.
// The valid (voxel, frame) CELLS of the tile are spread over the 32 lanes and handled independently.
// META: write M (K2T skips it for tiles past its capacity).  SMALL: update the small state (K2T only).
template <int NSET, bool META, bool SMALL>
__device__ __forceinline__ void fill_tile_meta(const FusionParams& p, const TileLane& L, uint32_t cnt, TileMeta<NSET>* M,
                                               float4* my_smp, int lane)
{
    constexpr int G = NSET * kTileSlots;
    constexpr int kCellIters = G * SAF_MAX_BATCH / 32;
    const int B = p.batch;
    // per frame: which voxels of the tile are valid (lane b keeps frame b's ballot)
    uint32_t my_ballot = 0;
#pragma unroll
    for (int b = 0; b < SAF_MAX_BATCH; ++b) {
        const uint32_t bal = __ballot_sync(0xffffffffu, (L.mask >> b) & 1u);
        if (lane == b) my_ballot = bal;
    }
    if (META && lane < SAF_MAX_BATCH) {
#pragma unroll
        for (int s = 0; s < NSET; ++s) M->vmask[s][lane] = (uint8_t)((my_ballot >> (s * kTileSlots)) & 0xffu);
    }
#pragma unroll
    for (int s = 0; s < NSET; ++s) {
        const uint32_t fm = __ballot_sync(0xffffffffu, lane < SAF_MAX_BATCH &&
                                                           ((my_ballot >> (s * kTileSlots)) & 0xffu) != 0u);
        if (META && lane == 0) M->fmask[s] = fm;
    }
    if (META && lane == 0) M->n_rows = cnt;
    if (META && lane < cnt) M->voxel[lane] = L.voxel;
    // cells: c -> (frame c / G, voxel c % G).  Pass 1 requests every valid cell's image coordinates.
    float2 cg[kCellIters];
#pragma unroll
    for (int it = 0; it < kCellIters; ++it) {
        const int c = it * 32 + lane, b = c / G, v = c % G;
        const uint32_t mv = __shfl_sync(0xffffffffu, L.mask, v);
        const uint32_t rk = __shfl_sync(0xffffffffu, L.rank, v);
        const uint32_t lc = __shfl_sync(0xffffffffu, L.local, v);
        cg[it] = make_float2(0.f, 0.f);
        if ((mv >> b) & 1u) cg[it] = p.wcoords[((uint64_t)rk * B + b) * kBlockVoxels + lc];
    }
    // Pass 2: per valid cell the compute warps' update record (clip_seem_fusion.py:800-810), the rgb sample
    // and the label counter (:786-798, 820-822).
#pragma unroll
    for (int it = 0; it < kCellIters; ++it) {
        const int c = it * 32 + lane, b = c / G, v = c % G;
        const uint32_t mv = __shfl_sync(0xffffffffu, L.mask, v);
        const uint32_t vx = __shfl_sync(0xffffffffu, L.voxel, v);
        const int w0 = __shfl_sync(0xffffffffu, L.w0, v);
        const bool cell_valid = (mv >> b) & 1u;
        uint32_t cell_rows = 0;
        if (cell_valid) {
            const saf_frame& f = p.frames[b];
            const float2 g = cg[it];
            if (META) {
                const int w = w0 + __popc(mv & ((1u << b) - 1u));   // single-frame calls before this one
                const float a = __frcp_rn(__int2float_rn(w + 1));
                const float bb = __fmul_rn(__int2float_rn(w), a);
                Taps t;
                feature_taps_padded(f, g.x, g.y, p.W, p.H, &p.hdr->error_flags, t);
                const uint32_t rows = (uint32_t)t.idx[0] | ((uint32_t)t.idx[1] << 8) | ((uint32_t)t.idx[2] << 16) |
                                      ((uint32_t)t.idx[3] << 24);
                TileUpdate u;
                u.w[0] = t.w[0];
                u.w[1] = t.w[1];
                u.w[2] = t.w[2];
                u.w[3] = t.w[3];
                u.a = a;
                u.b = bb;
                u.rows = rows;
                u.one = 1.0f;
                M->upd[v / kTileSlots][b][v % kTileSlots] = u;
                cell_rows = rows;
            }
            const int px = nearest_index(g.x, p.W), py = nearest_index(g.y, p.H);
            float smp[3] = {0.f, 0.f, 0.f};
            if (SMALL) {
                sample_rgb(p, f, g.x, g.y, px, py, smp);
                my_smp[v * SAF_MAX_BATCH + b] = make_float4(smp[0], smp[1], smp[2], 0.f);
            }
            if (SMALL && p.vol.labels_one_hot && f.seg) {
                const float lf = (px >= 0 && py >= 0) ? load_class_id(f.seg, f.seg_dtype, py * p.W + px) : 0.0f;
                const long long id = (long long)lf;
                if (id >= 0 && id < p.vol.n_classes)
                    atomicAdd(p.vol.labels_one_hot + (size_t)vx * p.vol.n_classes + id, 1);   // RED: no round trip
                else
                    atomicOr(&p.hdr->error_flags, SAF_FLAG_BAD_CLASS_ID);
            }
        }
        // The 8 lanes of a group hold one set's voxels for one frame (G = 16: two frames per iteration).  The
        // group's first valid voxel provides the rows the compute warps prefetch for (set, frame); if every
        // valid voxel of the group uses those rows - nearly always - the frame takes the compare-free path.
        static_assert(G == 16, "cell groups assume 16 voxels per tile");
        const uint32_t vb = __ballot_sync(0xffffffffu, cell_valid);
        const int group = lane >> 3;
        const uint32_t gb = (vb >> (group * 8)) & 0xffu;
        const int first = group * 8 + (gb ? __ffs(gb) - 1 : 0);
        const uint32_t prim = __shfl_sync(0xffffffffu, cell_rows, first);
        const uint32_t mb = __ballot_sync(0xffffffffu, cell_valid && cell_rows != prim);
        // a set that straddles a table cell boundary (~30 % of the frames on cfg2) nearly always sees just two cells:
        // its voxels are grouped by cell so that the compute warps load each cell's rows once per frame
        const uint32_t mg = (mb >> (group * 8)) & 0xffu;
        const int first_alt = group * 8 + (mg ? __ffs(mg) - 1 : 0);
        const uint32_t alt = __shfl_sync(0xffffffffu, cell_rows, first_alt);
        const uint32_t ob = __ballot_sync(0xffffffffu, cell_valid && cell_rows != prim && cell_rows != alt);
        if (META && gb && lane == first) {
            const int st = v / kTileSlots;
            M->prim_rows[st][b] = prim;
            M->uniform[st][b] = mg == 0u ? 1 : 0;
            M->alt_rows[st][b] = alt;
            M->pmask[st][b] = (uint8_t)(gb & ~mg);
            M->amask[st][b] = (uint8_t)(mg & ~((ob >> (group * 8)) & 0xffu));
        }
    }
    __syncwarp();   // samples and update records of every cell are in shared memory
    if (SMALL && lane < cnt) {
        // the rgb running average walks the voxel's frames in order (clip_seem_fusion.py:808-813)
        float acc[3] = {L.rgb0[0], L.rgb0[1], L.rgb0[2]};
        int w = L.w0;
        for (uint32_t mm = L.mask; mm; mm &= mm - 1u, ++w) {
            const int b = __ffs(mm) - 1;
            const float4 sm = my_smp[lane * SAF_MAX_BATCH + b];
            const float a = __frcp_rn(__int2float_rn(w + 1)), bb = __fmul_rn(__int2float_rn(w), a);   // the records' a, b
            acc[0] = __fadd_rn(__fmul_rn(sm.x, a), __fmul_rn(acc[0], bb));
            acc[1] = __fadd_rn(__fmul_rn(sm.y, a), __fmul_rn(acc[1], bb));
            acc[2] = __fadd_rn(__fmul_rn(sm.z, a), __fmul_rn(acc[2], bb));
        }
        float* dst = p.vol.rgb + (size_t)L.voxel * 3;
        dst[0] = acc[0];
        dst[1] = acc[1];
        dst[2] = acc[2];
        p.vol.weight[L.voxel] = L.w0 + __popc(L.mask);
    }
    __syncwarp();   // every lane's metadata is written (and my_smp read) before the arrival publishes it
}

// Register budget.  The producers are one warpgroup (warps 0-3, the last one idle) that hands most of its registers to
// the compute warpgroups right after start-up (setmaxnreg): with 12 compute warps at C = 768 that is 152 registers per
// compute thread instead of 128, which pays for a third set of table-row registers - a frame's rows are requested
// TWO frames ahead, so their L2 round trip is covered by two frames of arithmetic instead of one.
constexpr int kTileProducerWarps = 4;
constexpr int kTileProducerRegs = 56;
template <int NCW>
struct TileRegs {
    static constexpr int kLaunch = 65536 / ((NCW + kTileProducerWarps) * 32) / 8 * 8;   // what __launch_bounds__ grants
    static constexpr int kPool = (NCW + kTileProducerWarps) * 32 * kLaunch;
    static constexpr int kRaw = (kPool - kTileProducerWarps * 32 * kTileProducerRegs) / (NCW * 32) / 8 * 8;
    static constexpr int kCompute = kRaw > 232 ? 232 : kRaw;
};

template <int CHUNKS, int NSET, int NBUF>
__global__ void __launch_bounds__((CHUNKS * NSET + kTileProducerWarps) * 32, 1)
feature_accumulate_window_tile_kernel(const __grid_constant__ FusionParams p, const __grid_constant__ WindowTables wt)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int C = CHUNKS * 128;
    constexpr int G = NSET * kTileSlots;
    constexpr int NCW = CHUNKS * NSET;
    static_assert(G <= 32, "one producer lane per voxel of the tile");
    static_assert(NBUF < kTileProducerWarps + 1 && NCW % 4 == 0, "warpgroups");
    using Meta = TileMeta<NSET>;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SlotCounters* sc = &p.hdr->slot[p.slot];
    const uint32_t n = sc->n_union;
    if (n == 0) return;

    float* rows_buf = reinterpret_cast<float*>(smem_raw);                                   // [NBUF][G][C]
    Meta* metas = reinterpret_cast<Meta*>(smem_raw + (size_t)NBUF * G * C * sizeof(float));  // [2*NBUF]
    uint64_t* bars = reinterpret_cast<uint64_t*>(metas + 2 * NBUF);
    uint64_t* full = bars;                    // [2*NBUF]
    uint64_t* meta_free = bars + 2 * NBUF;    // [2*NBUF]
    uint64_t* rows_free = bars + 4 * NBUF;    // [NBUF]
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2 * NBUF; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&meta_free[i], NCW);
        }
        for (int i = 0; i < NBUF; ++i) mbar_init(&rows_free[i], NCW);
        fence_mbar_init();
    }
    __syncthreads();

    if (warp < kTileProducerWarps) {
        // ------------------------------- producer -------------------------------
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kTileProducerRegs));
        if (warp >= NBUF) return;
        const uint32_t n_blocks = sc->n_blocks;
        float* my_rows = rows_buf + (size_t)warp * G * C;
        const Meta* prepared = reinterpret_cast<const Meta*>(p.tile_meta);
        for (uint32_t j = 0;; ++j) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&sc->k3_next, (uint32_t)G);
            base = __shfl_sync(0xffffffffu, base, 0);
            const uint32_t cnt = base < n ? min((uint32_t)G, n - base) : 0u;
            const uint32_t ms = (uint32_t)warp + (uint32_t)NBUF * (j & 1u);
            Meta* M = metas + ms;
            if (cnt == 0) {
                mbar_wait(&meta_free[ms], ((j >> 1) & 1u) ^ 1u);
                if (lane == 0) {
                    M->n_rows = 0;
                    mbar_arrive(&full[ms]);
                }
                break;
            }
            if (NSET == 2 && base / (uint32_t)G < p.tile_meta_cap) {
                // K2T prepared this tile (metadata in global memory, small state already updated): one bulk copy of
                // the metadata next to those of the feature rows
                const Meta* GM = prepared + base / (uint32_t)G;
                const uint32_t vox = lane < cnt ? __ldg(&GM->voxel[lane]) : 0u;
                mbar_wait(&rows_free[warp], (j & 1u) ^ 1u);
                mbar_wait(&meta_free[ms], ((j >> 1) & 1u) ^ 1u);
                if (lane == 0) {
                    mbar_expect_tx(&full[ms], cnt * (uint32_t)C * 4u + (uint32_t)sizeof(Meta));
                    tma_bulk_g2s(M, GM, (uint32_t)sizeof(Meta), &full[ms]);
                }
                __syncwarp();
                if (lane < cnt)
                    tma_bulk_g2s(my_rows + (size_t)lane * C, p.vol.clip_feat + (size_t)vox * C, (uint32_t)C * 4u, &full[ms]);
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[ms]);
                continue;
            }
            // a tile past K2T's capacity: its metadata is prepared here (the small state is always K2T's)
            const TileLane L = load_tile_lane<false>(p, n_blocks, base, cnt, lane);
            // the landing buffer was handed back when the compute warps took its rows to registers
            mbar_wait(&rows_free[warp], (j & 1u) ^ 1u);
            if (lane == 0) mbar_expect_tx(&full[ms], cnt * (uint32_t)C * 4u);
            __syncwarp();
            if (lane < cnt)
                tma_bulk_g2s(my_rows + (size_t)lane * C, p.vol.clip_feat + (size_t)L.voxel * C, (uint32_t)C * 4u, &full[ms]);
            // the metadata slot was last read two of this producer's tiles ago
            mbar_wait(&meta_free[ms], ((j >> 1) & 1u) ^ 1u);
            fill_tile_meta<NSET, true, false>(p, L, cnt, M, nullptr, lane);
            __syncwarp();   // every lane's metadata is written (and my_smp read) before the arrival publishes it
            if (lane == 0) mbar_arrive(&full[ms]);
        }
        return;
    }

    // ------------------------------- compute -------------------------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TileRegs<NCW>::kCompute));
    const int cw = warp - kTileProducerWarps;
    const int set = cw / CHUNKS, chunk = cw % CHUNKS;
    const int col4 = chunk * 32 + lane;   // this thread's float4 column of every row
    uint32_t done = 0;                    // producers that have published their last tile
#ifdef SAF_TILE_TIMING
    long long tm_wait = 0, tm_load = 0, tm_frames = 0, tm_store = 0, tm_c0 = clock64(), tm_c1;
    uint32_t tm_tiles = 0, tm_nframes = 0, tm_upd = 0, tm_nonuni = 0;
#define TM_MARK(acc) do { tm_c1 = clock64(); acc += tm_c1 - tm_c0; tm_c0 = tm_c1; } while (0)
#else
#define TM_MARK(acc) do { } while (0)
#endif
    for (uint32_t t = 0; done != (1u << NBUF) - 1u; ++t) {
        const uint32_t pi = t % NBUF, j = t / NBUF;
        if ((done >> pi) & 1u) continue;
        const uint32_t ms = pi + (uint32_t)NBUF * (j & 1u);
        mbar_wait(&full[ms], (j >> 1) & 1u);
        TM_MARK(tm_wait);
        const Meta* M = metas + ms;
        const uint32_t n_rows = M->n_rows;
        if (n_rows == 0) {
            done |= 1u << pi;
            continue;
        }
        // accumulators of the set's voxels: landing buffer -> registers, then the buffer goes back to the producer
        f32x2_t acc_lo[kTileSlots], acc_hi[kTileSlots];
        const ulonglong2* land = reinterpret_cast<const ulonglong2*>(rows_buf + (size_t)pi * G * C) + col4;
#pragma unroll
        for (int s = 0; s < kTileSlots; ++s) {
            if ((uint32_t)(set * kTileSlots + s) < n_rows) {
                const ulonglong2 v = land[(size_t)(set * kTileSlots + s) * (C / 4)];
                acc_lo[s] = v.x;
                acc_hi[s] = v.y;
            } else {
                acc_lo[s] = 0ull;
                acc_hi[s] = 0ull;
            }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&rows_free[pi]);
        TM_MARK(tm_load);

        const uint32_t fm = M->fmask[set];
#ifdef SAF_TILE_TIMING
        tm_tiles += 1;
        tm_nframes += __popc(fm);
        for (uint32_t q = fm; q; q &= q - 1u) {
            const int b = __ffs(q) - 1;
            tm_upd += __popc((uint32_t)M->vmask[set][b]);
            tm_nonuni += M->uniform[set][b] ? 0 : 1;
        }
#endif
        if (fm) {
            // one update: the record (weights, a, b, rows, 1.0f) is two 16-byte broadcast loads
            auto update = [&](int b, int s, const uint4& m0, const uint4& m1, const RowRegs& T) {
                const float w0 = __uint_as_float(m0.x), w1 = __uint_as_float(m0.y), w2 = __uint_as_float(m0.z),
                            w3 = __uint_as_float(m0.w), a = __uint_as_float(m1.x), bb = __uint_as_float(m1.y),
                            one = __uint_as_float(m1.w);
                const f32x2_t wp[4] = {pack2(w0, w0), pack2(w1, w1), pack2(w2, w2), pack2(w3, w3)};
                const f32x2_t ap = pack2(a, a), bp = pack2(bb, bb), op = pack2(one, one);
                acc_lo[s] = mix_blend2(T.lo, wp, ap, bp, op, acc_lo[s]);
                acc_hi[s] = mix_blend2(T.hi, wp, ap, bp, op, acc_hi[s]);
            };
            // one frame of the window: every voxel of the set that the frame sees, with the frame's rows in T
            auto frame_updates = [&](int b, const RowRegs& T) {
                const uint32_t vm = M->vmask[set][b];
                const uint4* rec = reinterpret_cast<const uint4*>(&M->upd[set][b][0]);
                if (M->uniform[set][b]) {
                    // every valid voxel of the set samples the prefetched rows: no per-voxel row check
#pragma unroll
                    for (int s = 0; s < kTileSlots; ++s)
                        if ((vm >> s) & 1u) update(b, s, rec[2 * s], rec[2 * s + 1], T);
                    return;
                }
                // the set straddles a table cell boundary: the second cell's rows are requested now and arrive while
                // the voxels of the first cell are updated (the order of a frame's voxels does not matter)
                const uint32_t pm = M->pmask[set][b], am = M->amask[set][b];
                RowRegs TA;
                uint32_t alt_rows = M->alt_rows[set][b];
                load_rows(TA, wt.ptr[b], alt_rows, C, col4);
#pragma unroll
                for (int s = 0; s < kTileSlots; ++s)
                    if ((pm >> s) & 1u) update(b, s, rec[2 * s], rec[2 * s + 1], T);
#pragma unroll
                for (int s = 0; s < kTileSlots; ++s)
                    if ((am >> s) & 1u) update(b, s, rec[2 * s], rec[2 * s + 1], TA);
                const uint32_t rest = vm & ~(pm | am);
                if (rest) {   // a third (fourth) cell: rare
#pragma unroll
                    for (int s = 0; s < kTileSlots; ++s) {
                        if ((rest >> s) & 1u) {
                            const uint4 m0 = rec[2 * s], m1 = rec[2 * s + 1];
                            if (m1.z != alt_rows) {
                                alt_rows = m1.z;
                                load_rows(TA, wt.ptr[b], alt_rows, C, col4);
                            }
                            update(b, s, m0, m1, TA);
                        }
                    }
                }
            };
            uint32_t rest = fm;
            auto next_frame = [&]() {
                const int b = rest ? __ffs(rest) - 1 : -1;
                rest &= rest - 1u;
                return b;
            };
            auto request = [&](int b, RowRegs& T) {
                if (b >= 0) load_rows(T, wt.ptr[b], M->prim_rows[set][b], C, col4);
            };
            if constexpr (TileRegs<NCW>::kCompute >= 144) {
                // a frame's rows are requested two frames before its arithmetic starts; three register sets take turns
                RowRegs T0, T1, T2;
                int b0 = next_frame(), b1 = next_frame(), b2;
                request(b0, T0);
                request(b1, T1);
                for (;;) {
                    b2 = next_frame();
                    request(b2, T2);
                    frame_updates(b0, T0);
                    if (b1 < 0) break;
                    b0 = next_frame();
                    request(b0, T0);
                    frame_updates(b1, T1);
                    if (b2 < 0) break;
                    b1 = next_frame();
                    request(b1, T1);
                    frame_updates(b2, T2);
                    if (b0 < 0) break;
                }
            } else {
                // C = 1024 (16 compute warps, 104 registers each): one frame ahead, two register sets
                RowRegs T0, T1;
                int b0 = next_frame(), b1;
                request(b0, T0);
                for (;;) {
                    b1 = next_frame();
                    request(b1, T1);
                    frame_updates(b0, T0);
                    if (b1 < 0) break;
                    b0 = next_frame();
                    request(b0, T0);
                    frame_updates(b1, T1);
                    if (b0 < 0) break;
                }
            }
            TM_MARK(tm_frames);
#pragma unroll
            for (int s = 0; s < kTileSlots; ++s) {
                if ((uint32_t)(set * kTileSlots + s) < n_rows) {
                    const uint32_t v = M->voxel[set * kTileSlots + s];
                    st_stream_b64x2(reinterpret_cast<ulonglong2*>(p.vol.clip_feat + (size_t)v * C) + col4, acc_lo[s], acc_hi[s]);
                }
            }
        }
        fence_proxy_async();   // prepared tiles: the slot is refilled by a bulk copy
        __syncwarp();
        if (lane == 0) mbar_arrive(&meta_free[ms]);
        TM_MARK(tm_store);
    }
#ifdef SAF_TILE_TIMING
    if (lane == 0 && (cw == 0 || cw == NCW - 1) && (blockIdx.x % 37) == 0)
        printf("[k3w timing] cta %d warp %d: tiles %u frames %u updates %u non-uniform frames %u | cycles wait %lld load %lld "
               "frames %lld store %lld | per update %.1f per frame %.1f\n", blockIdx.x, cw, tm_tiles, tm_nframes, tm_upd,
               tm_nonuni, tm_wait, tm_load, tm_frames, tm_store, (double)tm_frames / max(1u, tm_upd),
               (double)tm_frames / max(1u, tm_nframes));
#endif
#undef TM_MARK
}

// ---------------------------------------------------------------------------------------------
// K2T: prepares K3W's tiles ahead of it.  Inside K3W the three producer warps of a CTA do this work one tile at a
// time, each a chain of memory round trips (union entries, image coordinates, rgb taps, class ids) that 3 warps per
// SM cannot hide: measured on cfg2, K3W takes 461 us with its producers doing everything, 251 us with the compute
// warps idle, 323 us with the producers reduced to nothing.  Here the same per-tile routine (fill_tile_meta) runs
// one warp per tile over the whole GPU, tens of warps per SM, writes the tile's metadata to global memory and
// updates the small state; K3W's producers then only start two bulk copies per tile.
// ---------------------------------------------------------------------------------------------
constexpr int kK2TWarps = 8;

#ifndef SAF_K2T_MINBLOCKS
#define SAF_K2T_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(kK2TWarps * 32, SAF_K2T_MINBLOCKS) window_tile_setup_kernel(const __grid_constant__ FusionParams p)
{
    constexpr int NSET = 2, G = NSET * kTileSlots;
    __shared__ float4 s_smp[kK2TWarps][G * SAF_MAX_BATCH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const SlotCounters* sc = &p.hdr->slot[p.slot];
    const uint32_t n = sc->n_union, n_blocks = sc->n_blocks;
    const uint32_t n_tiles = (n + G - 1) / G;
    TileMeta<NSET>* metas = reinterpret_cast<TileMeta<NSET>*>(p.tile_meta);
    for (uint32_t tile = blockIdx.x * kK2TWarps + warp; tile < n_tiles; tile += gridDim.x * kK2TWarps) {
        const uint32_t base = tile * G, cnt = min((uint32_t)G, n - base);
        const TileLane L = load_tile_lane<true>(p, n_blocks, base, cnt, lane);
        if (tile < p.tile_meta_cap)
            fill_tile_meta<NSET, true, true>(p, L, cnt, metas + tile, s_smp[warp], lane);
        else   // no room for the metadata (K3W prepares it itself): the small state only
            fill_tile_meta<NSET, false, true>(p, L, cnt, metas, s_smp[warp], lane);
        __syncwarp();   // s_smp is reused by the warp's next tile
    }
}

// Any feature_dim (window mode): rows straight from global memory, VEC = 4 or 1.
template <int VEC>
__global__ void __launch_bounds__(kK3Threads) feature_accumulate_window_generic_kernel(
    const __grid_constant__ FusionParams p, const __grid_constant__ WindowTables wt)
{
    const int C = p.vol.feature_dim;
    const int lane = threadIdx.x & 31;
    const SlotCounters* sc = &p.hdr->slot[p.slot];
    const uint32_t n = sc->n_union;
    const uint32_t n_blocks = sc->n_blocks;
    const int B = p.batch;
    const uint32_t nwarps = gridDim.x * kK3Warps;
    const uint32_t gwarp = blockIdx.x * kK3Warps + (threadIdx.x >> 5);
    const uint32_t* __restrict__ off = p.blk_offset;
    for (uint64_t i = gwarp; i < n; i += nwarps) {
        uint32_t lo = 0, hi = n_blocks;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(off + mid) <= (uint32_t)i)
                lo = mid;
            else
                hi = mid;
        }
        const WinEntry e = p.ulist[(uint64_t)lo * kBlockVoxels + ((uint32_t)i - __ldg(off + lo))];
        const float2* src = p.wcoords + (uint64_t)lo * B * kBlockVoxels + (e.mask_local >> 16);
        const int w0 = p.vol.weight[e.voxel];
        float* row = p.vol.clip_feat + (size_t)e.voxel * C;
        int w = w0;
        for (uint32_t mm = e.mask_local & 0xffffu; mm; mm &= mm - 1u) {
            const int b = __ffs(mm) - 1;
            const float2 g = src[(size_t)b * kBlockVoxels];
            const float a = __frcp_rn(__int2float_rn(w + 1));
            const float bb = __fmul_rn(__int2float_rn(w), a);
            Taps t;
            feature_taps_padded(p.frames[b], g.x, g.y, p.W, p.H, &p.hdr->error_flags, t);
            const float* table = wt.ptr[b];   // zero-bordered [R',C] rows
            const int64_t sr = wt.stride_r[b];
            if (VEC == 4) {
                for (int col = lane; col < C / 4; col += 32) {
                    const float4 o = reinterpret_cast<const float4*>(row)[col];
                    float4 tv[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) tv[k] = __ldg(reinterpret_cast<const float4*>(table + t.idx[k] * sr) + col);
                    reinterpret_cast<float4*>(row)[col] = blend4(mix4(tv[0], tv[1], tv[2], tv[3], t.w), o, a, bb);
                }
            } else {
                for (int c = lane; c < C; c += 32) {
                    float tv[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) tv[k] = __ldg(table + t.idx[k] * sr + c);
                    const float smp = bilinear_mix(tv[0], tv[1], tv[2], tv[3], t.w);
                    row[c] = __fadd_rn(__fmul_rn(smp, a), __fmul_rn(row[c], bb));
                }
            }
            ++w;
        }
        __syncwarp();
        if (lane == 0) {
            // small state, frame by frame (weights advance exactly as above)
            w = w0;
            for (uint32_t mm = e.mask_local & 0xffffu; mm; mm &= mm - 1u) {
                const int b = __ffs(mm) - 1;
                const float2 g = src[(size_t)b * kBlockVoxels];
                ValidEntry ve;
                ve.voxel = e.voxel;
                ve.gx = g.x;
                ve.gy = g.y;
                ve.pad = 0;
                update_small_state(p, p.frames[b], ve, w);
                ++w;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// label argmax (clip_seem_fusion.py:315-325)
// ---------------------------------------------------------------------------------------------

// kArgmaxRows rows per warp iteration: all their loads are issued before the first reduction, which is what keeps
// enough bytes in flight for HBM (one 572-byte row per iteration reached 2.5 TB/s)
constexpr int kArgmaxRows = 6;

__global__ void __launch_bounds__(256) label_argmax_kernel(const int32_t* __restrict__ labels, int64_t n, int n_classes,
                                                           long long* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x / 32);
    const int64_t gwarp = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (n_classes > 160) {  // generic path: one row at a time
        for (int64_t v = gwarp; v < n; v += nwarps) {
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
        return;
    }
    // n_classes <= 160 (the reference's 143): five elements per lane and row, kArgmaxRows rows in flight
    for (int64_t v0 = gwarp * kArgmaxRows; v0 < n; v0 += nwarps * kArgmaxRows) {
        int x[kArgmaxRows][5];
#pragma unroll
        for (int r = 0; r < kArgmaxRows; ++r) {
            const int64_t v = v0 + r;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const int c = lane + 32 * k;
                x[r][k] = (v < n && c < n_classes) ? __ldg(labels + v * n_classes + c) : INT_MIN;
            }
        }
#pragma unroll
        for (int r = 0; r < kArgmaxRows; ++r) {
            const int64_t v = v0 + r;
            int best_val = INT_MIN, best_idx = INT_MAX;
            bool any = false;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const int c = lane + 32 * k;
                any |= (c < n_classes) && (x[r][k] != 0);
                if (x[r][k] > best_val) {
                    best_val = x[r][k];
                    best_idx = c;
                }
            }
            // warp argmax with the first maximum winning: redux.sync max of the values, then min of the indices at it
            const int warp_max = __reduce_max_sync(0xffffffffu, best_val);
            const int warp_idx = __reduce_min_sync(0xffffffffu, best_val == warp_max ? best_idx : INT_MAX);
            any = __any_sync(0xffffffffu, any);
            if (lane == 0 && v < n) out[v] = any ? (long long)warp_idx : -1ll;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

int device_sm_count(int* sms, int* smem_optin)
{
    // per-device attributes are constants: queried once per device
    struct Info { int ok, major, sms, smem; };
    static Info cache[64] = {};
    int dev = 0;
    SAF_CUDA_TRY(cudaGetDevice(&dev));
    Info local;
    Info& info = (dev >= 0 && dev < 64) ? cache[dev] : local;
    if (&info == &local || !info.ok) {
        Info q = {};
        SAF_CUDA_TRY(cudaDeviceGetAttribute(&q.major, cudaDevAttrComputeCapabilityMajor, dev));
        SAF_CUDA_TRY(cudaDeviceGetAttribute(&q.sms, cudaDevAttrMultiProcessorCount, dev));
        SAF_CUDA_TRY(cudaDeviceGetAttribute(&q.smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        q.ok = 1;
        info = q;
    }
    if (info.major != 10) return SAF_ERR_DEVICE;
    *sms = info.sms;
    if (smem_optin) *smem_optin = info.smem;
    return 0;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device, size) instead of once per launch
static int ensure_dynamic_smem_impl(const void* kern, size_t smem)
{
    static std::mutex mu;
    static std::unordered_map<const void*, size_t> granted[64];
    int dev = 0;
    SAF_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) {
        SAF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        return 0;
    }
    std::lock_guard<std::mutex> lock(mu);
    size_t& g = granted[dev][kern];
    if (g < smem) {
        SAF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        g = smem;
    }
    return 0;
}
template <typename K>
static int ensure_dynamic_smem(K kern, size_t smem)
{
    return ensure_dynamic_smem_impl(reinterpret_cast<const void*>(kern), smem);
}

static int build_params(const saf_grid_desc* grid, const saf_volume* vol, const saf_frame* frames, int32_t batch,
                        int32_t H, int32_t W, float trunc, int32_t rgb_mode, const saf_workspace* ws, FusionParams* p,
                        uint32_t slot = 0)
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
        if (frames[b].depth_dtype != SAF_DEPTH_F32 && frames[b].depth_dtype != SAF_DEPTH_U16_MM) return SAF_ERR_DTYPE;
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
    p->n_k1 = L.n_k1;
    p->nxs = (uint32_t)slab_planes(*grid);
    p->nyl = (uint32_t)slab_ny_local(*grid);
    p->nslab = (uint64_t)p->nxs * (uint64_t)p->nyl * (uint64_t)grid->nvox[2];
    p->list_cap = L.list_cap;
    p->max_table_elems = (uint64_t)ws->max_table_elems;
    p->table_slot_elems = L.table_slot_elems;
    unsigned char* base = (unsigned char*)ws->base;
    p->hdr = (WsHeader*)base;
    p->slot = slot;
    unsigned char* sb = base + L.slot0 + (uint64_t)slot * L.slot_stride;
    p->cta_count = (uint32_t*)(sb + L.off_cta_count);
    p->tile_dmax = (float*)(sb + L.off_cta_dmax);
    p->tile_shift = 5;
    for (;;) {
        p->ntx = (W + (1 << p->tile_shift) - 1) >> p->tile_shift;
        p->nty = (H + (1 << p->tile_shift) - 1) >> p->tile_shift;
        if ((uint32_t)p->ntx * (uint32_t)p->nty <= kMaxDepthTiles) break;
        p->tile_shift += 1;
    }
    p->block_seg = (uint2*)(sb + L.off_block_seg);
    p->blk_count = (uint32_t*)(sb + L.off_blk_count);
    p->blk_offset = (uint32_t*)(sb + L.off_blk_offset);
    p->lists = (ValidEntry*)(sb + L.off_lists);
    // window mode reuses the per-frame list regions: region 0 holds the union list, regions 1.. the coordinates
    p->ulist = (WinEntry*)p->lists;
    p->wcoords = (float2*)(p->lists + L.list_cap);
    p->w0_list = (int32_t*)(p->ulist + L.list_cap);   // second half of list region 0
    p->tables = (float*)(sb + L.off_tables);
    // a window workspace's list regions past the coordinates are idle in window mode: K2T's tile metadata
    {
        const uint64_t used = (uint64_t)L.list_cap * (sizeof(ValidEntry) + (uint64_t)batch * sizeof(float2));
        const uint64_t total = (uint64_t)ws->max_batch * L.list_cap * sizeof(ValidEntry);
        const uint64_t start = align_up(used, 128);
        p->tile_meta = (unsigned char*)p->lists + start;
        static const bool setup_on = !(getenv("SAF_TILE_SETUP") && atoi(getenv("SAF_TILE_SETUP")) == 0);
        p->tile_meta_cap = (setup_on && total > start)
                               ? (uint32_t)std::min<uint64_t>((total - start) / sizeof(TileMeta<2>), 0x7fffffffu) : 0u;
    }
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
        if (f.rgb_dtype != SAF_RGB_F32 && f.rgb_dtype != SAF_RGB_U8) return SAF_ERR_DTYPE;
        if (f.table_mode != SAF_TABLE_PATCH_GRID && f.table_mode != SAF_TABLE_SEGMENTS) return SAF_ERR_UNSUPPORTED;
        if (f.table_mode == SAF_TABLE_SEGMENTS && (!f.seg || f.npy != 1)) return SAF_ERR_NULL;
        const int64_t elems = (int64_t)f.npy * f.npx * vol->feature_dim;
        if (f.table_stride_c != 1) {
            if (elems > ws->max_table_elems) return SAF_ERR_WORKSPACE;
            *pack_mask |= 1u << b;
        }
    }
    return 0;
}

// window mode repacks every frame's feature image into its table slot of the workspace
static int check_window_tables(const saf_volume* vol, const saf_frame* frames, int32_t batch, const saf_workspace* ws)
{
    for (int b = 0; b < batch; ++b)
        if ((int64_t)frames[b].npy * frames[b].npx * vol->feature_dim > ws->max_table_elems) return SAF_ERR_WORKSPACE;
    return 0;
}

static int launch_k1(const FusionParams& p, cudaStream_t st)
{
    const uint32_t pack_ctas = p.sequential ? 256u : (p.pack_mask ? 64u : 0u);
    const int ntiles = p.ntx * p.nty;
    depth_tiles_kernel<<<(p.batch * ntiles + 7) / 8, 256, 0, st>>>(p);
    SAF_CHECK_LAUNCH("depth_tiles_kernel (K0)", st);
    const size_t k1_smem = (size_t)p.batch * ntiles * sizeof(float);
    if (k1_smem > 48u * 1024u)
        { int rc_ = ensure_dynamic_smem(frame_setup_kernel, k1_smem); if (rc_) return rc_; }
    frame_setup_kernel<<<p.n_k1 + pack_ctas, kK1Threads, k1_smem, st>>>(p, p.n_k1);
    SAF_CHECK_LAUNCH("frame_setup_kernel (K1)", st);
    return 0;
}

static int launch_k2(const FusionParams& p, int sms, cudaStream_t st)
{
    const uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)p.nblocks_total, (uint64_t)sms * (p.sequential ? 32u : 8u));
    if (p.batch == 1)
        tsdf_update_kernel<true, false><<<grid, kK2Threads, (size_t)p.n_k1 * 4, st>>>(p);
    else if (p.sequential)
        tsdf_update_kernel<false, true><<<grid, kK2Threads, (size_t)p.n_k1 * 4, st>>>(p);
    else
        tsdf_update_kernel<false, false><<<grid, kK2Threads, (size_t)p.n_k1 * 4, st>>>(p);
    SAF_CHECK_LAUNCH("tsdf_update_kernel (K2)", st);
    return 0;
}

template <int CHUNKS, int NST, bool SMEM>
static int launch_k3_fixed(const FusionParams& p, const float* table, int64_t stride_r, int table_tma, size_t smem,
                           int sms, cudaStream_t st)
{
    auto kern = feature_accumulate_kernel<CHUNKS, NST, SMEM>;
    { int rc_ = ensure_dynamic_smem(kern, smem); if (rc_) return rc_; }
    kern<<<sms, kK3Threads, smem, st>>>(p, table, stride_r, table_tma);
    SAF_CHECK_LAUNCH("feature_accumulate_kernel (K3)", st);
    return 0;
}

template <int CHUNKS>
static int launch_k3_chunks(const FusionParams& p, const float* table, int64_t stride_r, int R, int sms,
                            int smem_optin, cudaStream_t st)
{
    constexpr size_t row = (size_t)CHUNKS * 128 * 4;
    const size_t tab = (size_t)R * row;
    const size_t budget = (size_t)smem_optin - 256;
    auto need = [&](int nst, bool smem_tab) { return (smem_tab ? tab : 0) + (size_t)kK3Warps * nst * row + 8 * (1 + kK3Warps * nst); };
    // prefer the table in shared memory with as many ring stages as fit (at most 3)
    if (need(3, true) <= budget) return launch_k3_fixed<CHUNKS, 3, true>(p, table, stride_r, 1, need(3, true), sms, st);
    if (need(2, true) <= budget) return launch_k3_fixed<CHUNKS, 2, true>(p, table, stride_r, 1, need(2, true), sms, st);
    if (need(1, true) <= budget) return launch_k3_fixed<CHUNKS, 1, true>(p, table, stride_r, 1, need(1, true), sms, st);
    // table too large: read its rows through L1/L2 instead
    if (need(3, false) <= budget) return launch_k3_fixed<CHUNKS, 3, false>(p, table, stride_r, 0, need(3, false), sms, st);
    return launch_k3_fixed<CHUNKS, 2, false>(p, table, stride_r, 0, need(2, false), sms, st);
}

static int launch_k3(FusionParams& p, int frame_index, int sms, int smem_optin, cudaStream_t st)
{
    p.frame_index = frame_index;
    const saf_frame& f = p.frames[frame_index];
    const int C = p.vol.feature_dim;
    const int R = f.npy * f.npx;
    const bool packed = (p.pack_mask >> frame_index) & 1u;
    const float* table = packed ? p.tables + (uint64_t)frame_index * p.table_slot_elems : f.table;
    const int64_t stride_r = packed ? C : f.table_stride_r;
    const bool rows16 = (C % 4 == 0) && (stride_r % 4 == 0) && (((uintptr_t)table & 15u) == 0) &&
                        (((uintptr_t)p.vol.clip_feat & 15u) == 0);
    if (rows16) {
        switch (C) {
            case 512: return launch_k3_chunks<4>(p, table, stride_r, R, sms, smem_optin, st);
            case 768: return launch_k3_chunks<6>(p, table, stride_r, R, sms, smem_optin, st);
            case 1024: return launch_k3_chunks<8>(p, table, stride_r, R, sms, smem_optin, st);
            default: break;
        }
        feature_accumulate_generic_kernel<4><<<sms * 2, kK3Threads, 0, st>>>(p, table, stride_r);
    } else {
        feature_accumulate_generic_kernel<1><<<sms * 2, kK3Threads, 0, st>>>(p, table, stride_r);
    }
    SAF_CHECK_LAUNCH("feature_accumulate_generic_kernel (K3)", st);
    return 0;
}

template <int CHUNKS, int kW3Warps>
static int launch_k3w_pair(const FusionParams& p, const WindowTables& wt, int sms, cudaStream_t st)
{
    constexpr int kW3Threads = kW3Warps * 32;
    constexpr size_t row = (size_t)CHUNKS * 128 * 4;
    const size_t smem = (size_t)kW3Warps * 2 * row + (size_t)kW3Warps * kW3Chunk * SAF_MAX_BATCH * sizeof(float2) +
                        8 * (size_t)kW3Warps;
    auto kern = feature_accumulate_window_pair_kernel<CHUNKS, kW3Warps>;
    { int rc_ = ensure_dynamic_smem(kern, smem); if (rc_) return rc_; }
    kern<<<sms, kW3Threads, smem, st>>>(p, wt);
    SAF_CHECK_LAUNCH("feature_accumulate_window_pair_kernel (K3W)", st);
    return 0;
}

template <int CHUNKS, int kW3Warps>
static int launch_k3w_fixed(const FusionParams& p, const WindowTables& wt, int sms, cudaStream_t st)
{
    constexpr int kW3Threads = kW3Warps * 32;
    constexpr size_t row = (size_t)CHUNKS * 128 * 4;
    const size_t smem = (size_t)kW3Warps * row + (size_t)kW3Warps * kW3Chunk * SAF_MAX_BATCH * sizeof(float2) +
                        8 * (size_t)kW3Warps;
    auto kern = feature_accumulate_window_kernel<CHUNKS, kW3Warps>;
    { int rc_ = ensure_dynamic_smem(kern, smem); if (rc_) return rc_; }
    kern<<<sms, kW3Threads, smem, st>>>(p, wt);
    SAF_CHECK_LAUNCH("feature_accumulate_window_kernel (K3W)", st);
    return 0;
}

template <int CHUNKS, int NSET, int NBUF>
static int launch_k3w_tile(const FusionParams& p, const WindowTables& wt, int sms, cudaStream_t st)
{
    constexpr int kThreads = (CHUNKS * NSET + kTileProducerWarps) * 32;
    constexpr size_t smem = (size_t)NBUF * NSET * kTileSlots * CHUNKS * 128 * sizeof(float) +
                            2 * (size_t)NBUF * sizeof(TileMeta<NSET>) + 5 * (size_t)NBUF * sizeof(uint64_t);
    static_assert(smem <= 227 * 1024, "tile kernel shared memory");
    auto kern = feature_accumulate_window_tile_kernel<CHUNKS, NSET, NBUF>;
    { int rc_ = ensure_dynamic_smem(kern, smem); if (rc_) return rc_; }
    // A shard of a grid (multi-GPU) has small windows: their cost is the chain of dependent launches, not
    // throughput.  K3W then leaves some SMs free (it holds one CTA per SM and fills every SM it is given), so that
    // K0 / K1 / K2 of the NEXT window - already queued on the side stream - run beside it instead of after it.
    // Whole grids keep all SMs: there both kernels are throughput-bound and sharing would only slow K3W.
    static const int reserve_env = getenv("SAF_K3W_RESERVE_SMS") ? atoi(getenv("SAF_K3W_RESERVE_SMS")) : -1;
    // share of the grid this volume holds.  Measured on the cfg3 grid (one GPU running one rank's shard, K2T on its
    // own stream): a half grid is fastest with ~8 of 148 SMs left free (there K3W is most of the window), quarter and
    // eighth shards with ~54 (K0 -> K1 -> K2 -> K2T of the next window take as long as K3W of this one, and the two
    // chains run side by side: 1.75 -> 1.62 ms per 100 frames at 1/4, 1.30 -> 1.08 at 1/8 against 8 free SMs)
    const double share = (double)p.nslab / ((double)p.grid.nvox[0] * p.grid.nvox[1] * p.grid.nvox[2]);
    const int reserve = reserve_env >= 0 ? reserve_env : (share >= 0.75 ? 0 : (share >= 0.4 ? (int)(sms * 0.055) : (int)(sms * 0.365)));
    const int grid = std::max(1, sms - std::min(reserve, sms - 1));
    kern<<<grid, kThreads, smem, st>>>(p, wt);
    SAF_CHECK_LAUNCH("feature_accumulate_window_tile_kernel (K3W)", st);
    return 0;
}

// K3W variant: 2 = tile kernel (default), 1 = pair kernel, 0 = one voxel per warp pass.  SAF_K3W_VARIANT in the
// environment selects one of the older kernels for A/B timing.
static int k3w_variant()
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SAF_K3W_VARIANT");
        v = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 2;
    }
    return v;
}

// does this window take the tile kernel (and therefore K2T)?
static bool k3w_uses_tiles(const FusionParams& p)
{
    const int C = p.vol.feature_dim;
    if (k3w_variant() != 2 || (C != 512 && C != 768 && C != 1024)) return false;
    if ((((uintptr_t)p.vol.clip_feat & 15u) != 0) || (p.table_slot_elems % 4 != 0)) return false;
    // the tile kernel addresses table rows with one byte each
    for (int b = 0; b < p.batch; ++b)
        if ((p.frames[b].table_mode == SAF_TABLE_SEGMENTS ? p.frames[b].npx + 1
                                                          : (p.frames[b].npy + 2) * (p.frames[b].npx + 2)) > 256)
            return false;
    return true;
}

// K2T: after K2 of the same window, on K2's stream - it touches rgb / weight / label counters, which K3W never
// does, so it may run while K3W is still busy with the previous window
static int launch_k2t(const FusionParams& p, int sms, cudaStream_t st)
{
    if (!k3w_uses_tiles(p)) return 0;   // the other window kernels update the small state themselves
    window_tile_setup_kernel<<<sms * SAF_K2T_MINBLOCKS, kK2TWarps * 32, 0, st>>>(p);
    SAF_CHECK_LAUNCH("window_tile_setup_kernel (K2T)", st);
    return 0;
}

static int launch_k3w(const FusionParams& p, int sms, cudaStream_t st)
{
    const int C = p.vol.feature_dim;
    WindowTables wt;
    // K1 repacked every frame's feature image into the workspace (zero-bordered [R',C] rows)
    const bool rows16 = (C % 4 == 0) && (((uintptr_t)p.vol.clip_feat & 15u) == 0) && (p.table_slot_elems % 4 == 0);
    for (int b = 0; b < SAF_MAX_BATCH; ++b) {
        wt.ptr[b] = b < p.batch ? p.tables + (uint64_t)b * p.table_slot_elems : nullptr;
        wt.stride_r[b] = C;
    }
    if (k3w_uses_tiles(p)) {
        switch (C) {
            case 512: return launch_k3w_tile<4, 2, SAF_TILE_NBUF>(p, wt, sms, st);
            case 768: return launch_k3w_tile<6, 2, SAF_TILE_NBUF>(p, wt, sms, st);
            default: return launch_k3w_tile<8, 2, 2>(p, wt, sms, st);
        }
    }
    if (rows16) {
        if (k3w_variant() >= 1) {
            switch (C) {
                case 512: return launch_k3w_pair<4, K3W2_WARPS>(p, wt, sms, st);
                case 768: return launch_k3w_pair<6, K3W2_WARPS>(p, wt, sms, st);
                default: break;   // C = 1024: two rows of accumulators do not fit the register budget
            }
        }
        switch (C) {
            case 512: return launch_k3w_fixed<4, 24>(p, wt, sms, st);
            case 768: return launch_k3w_fixed<6, 24>(p, wt, sms, st);
            case 1024: return launch_k3w_fixed<8, 24>(p, wt, sms, st);
            default: break;
        }
        feature_accumulate_window_generic_kernel<4><<<sms * 2, kK3Threads, 0, st>>>(p, wt);
    } else {
        feature_accumulate_window_generic_kernel<1><<<sms * 2, kK3Threads, 0, st>>>(p, wt);
    }
    SAF_CHECK_LAUNCH("feature_accumulate_window_generic_kernel (K3W)", st);
    return 0;
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
    h.nblocks_total = L.nblocks_total;
    h.max_batch = (uint32_t)ws->max_batch;
    h.nb[0] = L.nb[0];
    h.nb[1] = L.nb[1];
    h.nb[2] = L.nb[2];
    cudaStream_t st = (cudaStream_t)stream;
    SAF_CUDA_TRY(cudaMemsetAsync(ws->base, 0, kWsHeaderBytes, st));
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
    out->total_calls = h.total_calls;
    out->total_union = h.total_union;
    out->last_union = h.slot[h.last_slot & 1u].n_union;
    const SlotCounters& sc = h.slot[h.last_slot & 1u];
    out->last_blocks = sc.n_frustum_blocks;
    for (int b = 0; b < SAF_MAX_BATCH; ++b) {
        out->last_valid[b] = sc.n_valid[b];
        out->last_tsdf_valid[b] = sc.last_tsdf_valid[b];
    }
    out->error_flags = h.error_flags;
    out->last_processed = sc.last_processed;
    out->depth_cull_on = h.depth_cull;
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
    return launch_k1(p, (cudaStream_t)stream);
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

// Window-mode counterparts of saf_tsdf_update / saf_feature_accumulate: the batch is a run of consecutive
// single-frame calls (what saf_integrate_sequence issues per window), after saf_frustum_cull on the same batch.
int saf_tsdf_update_window(const saf_grid_desc* grid, const saf_volume* vol, const saf_frame* frames, int32_t batch,
                           int32_t H, int32_t W, float trunc, const saf_workspace* ws, void* stream)
{
    int sms = 0;
    int rc = device_sm_count(&sms, nullptr);
    if (rc) return rc;
    if (!vol || !vol->tsdf || !vol->tsdf_weight) return SAF_ERR_NULL;
    if (!(trunc > 0.f)) return SAF_ERR_SHAPE;
    if (batch < 2) return SAF_ERR_BATCH;
    FusionParams p;
    rc = build_params(grid, vol, frames, batch, H, W, trunc, SAF_RGB_BILINEAR, ws, &p);
    if (rc) return rc;
    p.sequential = 1;
    return launch_k2(p, sms, (cudaStream_t)stream);
}

int saf_feature_accumulate_window(const saf_grid_desc* grid, const saf_volume* vol, const saf_frame* frames,
                                  int32_t batch, int32_t H, int32_t W, int32_t rgb_mode, const saf_workspace* ws,
                                  void* stream)
{
    return saf_feature_accumulate_window_stages(grid, vol, frames, batch, H, W, rgb_mode, ws,
                                                SAF_STAGE_TILE_SETUP | SAF_STAGE_ACCUMULATE, stream);
}

int saf_feature_accumulate_window_stages(const saf_grid_desc* grid, const saf_volume* vol, const saf_frame* frames,
                                         int32_t batch, int32_t H, int32_t W, int32_t rgb_mode, const saf_workspace* ws,
                                         int32_t stages, void* stream)
{
    if (stages < 1 || stages > (SAF_STAGE_TILE_SETUP | SAF_STAGE_ACCUMULATE)) return SAF_ERR_UNSUPPORTED;
    int sms = 0;
    int rc = device_sm_count(&sms, nullptr);
    if (rc) return rc;
    if (batch < 2) return SAF_ERR_BATCH;
    FusionParams p;
    rc = build_params(grid, vol, frames, batch, H, W, 1.0f, rgb_mode, ws, &p);
    if (rc) return rc;
    rc = check_feature_args(vol, frames, batch, rgb_mode, ws, &p.pack_mask);
    if (rc) return rc;
    if ((rc = check_window_tables(vol, frames, batch, ws))) return rc;
    p.sequential = 1;
    if (stages & SAF_STAGE_TILE_SETUP) {
        // saf_frustum_cull does not know the call is a window: repack the feature images here (K1's pack CTAs only)
        frame_setup_kernel<<<256, kK1Threads, 0, (cudaStream_t)stream>>>(p, 0u);
        SAF_CHECK_LAUNCH("frame_setup_kernel (table repack)", (cudaStream_t)stream);
        if ((rc = launch_k2t(p, sms, (cudaStream_t)stream))) return rc;
    }
    return (stages & SAF_STAGE_ACCUMULATE) ? launch_k3w(p, sms, (cudaStream_t)stream) : 0;
}

// K1 + K2 of one integrate() call on `st_geo`, then its K3 launches on `st_feat`.
static int integrate_call(const saf_grid_desc* grid, const saf_volume* vol, const saf_frame* frames, int32_t batch,
                          int32_t H, int32_t W, float trunc, int32_t rgb_mode, const saf_workspace* ws, uint32_t slot,
                          int sms, int smem_optin, cudaStream_t st_geo, cudaStream_t st_feat, cudaEvent_t geo_done,
                          cudaEvent_t feat_done, bool sequential = false, cudaStream_t st_tile = nullptr,
                          cudaEvent_t tile_done = nullptr)
{
    FusionParams p;
    int rc = build_params(grid, vol, frames, batch, H, W, trunc, rgb_mode, ws, &p, slot);
    if (rc) return rc;
    rc = check_feature_args(vol, frames, batch, rgb_mode, ws, &p.pack_mask);
    if (rc) return rc;
    p.sequential = (sequential && batch > 1) ? 1 : 0;
    if (p.sequential && (rc = check_window_tables(vol, frames, batch, ws))) return rc;
    // (the window's union list and coordinates live in list regions 0 and 1 .. ceil(batch / 2) <= batch - 1)
    if ((rc = launch_k1(p, st_geo))) return rc;
    if ((rc = launch_k2(p, sms, st_geo))) return rc;
    if (geo_done) SAF_CUDA_TRY(cudaEventRecord(geo_done, st_geo));
    if (p.sequential && st_tile && tile_done && geo_done) {
        // K2T on its own stream: behind K2 of this window, beside K0 / K1 / K2 of the next one (st_geo) and beside
        // K3W of the previous one (st_feat), which never touches the small state K2T updates
        SAF_CUDA_TRY(cudaStreamWaitEvent(st_tile, geo_done, 0));
        if ((rc = launch_k2t(p, sms, st_tile))) return rc;
        SAF_CUDA_TRY(cudaEventRecord(tile_done, st_tile));
        SAF_CUDA_TRY(cudaStreamWaitEvent(st_feat, tile_done, 0));
    } else {
        if (geo_done) SAF_CUDA_TRY(cudaStreamWaitEvent(st_feat, geo_done, 0));
        if (p.sequential && (rc = launch_k2t(p, sms, st_feat))) return rc;
    }
    if (p.sequential) {
        if ((rc = launch_k3w(p, sms, st_feat))) return rc;
    } else {
        for (int b = 0; b < batch; ++b)
            if ((rc = launch_k3(p, b, sms, smem_optin, st_feat))) return rc;
    }
    if (feat_done) SAF_CUDA_TRY(cudaEventRecord(feat_done, st_feat));
    return 0;
}

int saf_integrate(const saf_grid_desc* grid, const saf_volume* vol, const saf_frame* frames, int32_t batch, int32_t H,
                  int32_t W, float trunc, int32_t rgb_mode, const saf_workspace* ws, void* stream)
{
    int sms = 0, smem_optin = 0;
    int rc = device_sm_count(&sms, &smem_optin);
    if (rc) return rc;
    if (!(trunc > 0.f)) return SAF_ERR_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    return integrate_call(grid, vol, frames, batch, H, W, trunc, rgb_mode, ws, 0, sms, smem_optin, st, st, nullptr, nullptr);
}

// Host-side reach test for contiguous sub-slabs: can the camera frustum of a frame (pose and intrinsics only, no
// depth) contain any voxel centre of the slab?  The five half-spaces of block_maybe_visible, taken to world space,
// against the slab's axis-aligned box of voxel centres (padded by a voxel).  Conservative (NaN -> keep); costs
// nothing on the device and needs no synchronisation, so it runs before any frame is handed to the kernels -
// a caller can use saf_frame_reaches_slab to decide BEFORE copying a frame to the device at all.
static bool frame_may_reach_slab_host(const saf_grid_desc& g, const float* pose, const float* K, int H, int W)
{
    const double vs = g.voxel_size;
    const double lo[3] = {g.origin[0] + vs * (g.x_begin - 1), g.origin[1] - vs, g.origin[2] - vs};
    const double hi[3] = {g.origin[0] + vs * g.x_end, g.origin[1] + vs * g.nvox[1], g.origin[2] + vs * g.nvox[2]};
    const double t[3] = {pose[3], pose[7], pose[11]};
    double n[5][3];
    for (int k = 0; k < 3; ++k) {
        n[0][k] = K[6 + k];
        n[1][k] = K[k] + 0.5 * K[6 + k];
        n[2][k] = (W - 0.5) * K[6 + k] - K[k];
        n[3][k] = K[3 + k] + 0.5 * K[6 + k];
        n[4][k] = (H - 0.5) * K[6 + k] - K[3 + k];
    }
    for (int i = 0; i < 5; ++i) {
        double best = 0.0, scale = 0.0, nlen = 0.0;
        for (int a = 0; a < 3; ++a) {
            // world-space normal m = R n_i, R = pose[0:3, 0:3] (camera -> world)
            const double m = pose[4 * a] * n[i][0] + pose[4 * a + 1] * n[i][1] + pose[4 * a + 2] * n[i][2];
            const double v0 = m * (lo[a] - t[a]), v1 = m * (hi[a] - t[a]);
            best += v0 > v1 ? v0 : v1;
            scale += fabs(m) * (fabs(lo[a] - t[a]) + fabs(hi[a] - t[a]));
            nlen += m * m;
        }
        if (best + 1e-3 * scale + 1e-6 * sqrt(nlen) < 0.0) return false;   // NaN compares false -> kept
    }
    return true;
}

// Side stream and events of saf_integrate_sequence's two-slot overlap: created once per (host thread, device) and
// kept (creating and destroying a stream and five events per call cost more than a window's kernels).
struct SeqStreams {
    cudaStream_t side = nullptr, tile = nullptr;
    cudaEvent_t fork = nullptr, geo_done[2] = {nullptr, nullptr}, feat_done[2] = {nullptr, nullptr}, join = nullptr;
    cudaEvent_t tile_done[2] = {nullptr, nullptr}, tile_join = nullptr;
    bool ok = false;
};
static int seq_streams(SeqStreams** out)
{
    static thread_local SeqStreams cache[64];
    int dev = 0;
    SAF_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return SAF_ERR_DEVICE;
    SeqStreams& s = cache[dev];
    if (!s.ok) {
        SAF_CUDA_TRY(cudaStreamCreateWithFlags(&s.side, cudaStreamNonBlocking));
        SAF_CUDA_TRY(cudaStreamCreateWithFlags(&s.tile, cudaStreamNonBlocking));
        SAF_CUDA_TRY(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
        SAF_CUDA_TRY(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
        SAF_CUDA_TRY(cudaEventCreateWithFlags(&s.tile_join, cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) {
            SAF_CUDA_TRY(cudaEventCreateWithFlags(&s.geo_done[i], cudaEventDisableTiming));
            SAF_CUDA_TRY(cudaEventCreateWithFlags(&s.feat_done[i], cudaEventDisableTiming));
            SAF_CUDA_TRY(cudaEventCreateWithFlags(&s.tile_done[i], cudaEventDisableTiming));
        }
        s.ok = true;
    }
    *out = &s;
    return 0;
}

int saf_frame_reaches_slab(const saf_grid_desc* grid, const float* pose, const float* K, int32_t H, int32_t W)
{
    if (!grid || !pose || !K) return SAF_ERR_NULL;
    if (!slab_desc_ok(*grid) || H <= 0 || W <= 0) return SAF_ERR_GRID;
    if (grid->x_span != 0 || grid->y_ranks > 1) return 1;   // stripes / block columns span the grid
    return frame_may_reach_slab_host(*grid, pose, K, H, W) ? 1 : 0;
}

// The reference's frame loop.  Frames are independent except through the volume, and K1/K2 touch only the
// TSDF state while K3 touches only weight / rgb / features / labels, so K1+K2 of window i+1 run on a side
// stream underneath K3W of window i (the two scratch slots of the workspace alternate).
int saf_integrate_sequence(const saf_grid_desc* grid, const saf_volume* vol, const saf_frame* frames, int32_t n_frames,
                           int32_t H, int32_t W, float trunc, int32_t rgb_mode, const saf_workspace* ws, void* stream)
{
    if (n_frames < 0) return SAF_ERR_BATCH;
    int sms = 0, smem_optin = 0;
    int rc = device_sm_count(&sms, &smem_optin);
    if (rc) return rc;
    if (!(trunc > 0.f)) return SAF_ERR_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    // Window mode: up to SAF_MAX_BATCH consecutive frames share one K0 / K1 / K2 / K3W launch quartet (per-voxel
    // updates are applied frame by frame inside the kernels, so the result is that of the single-frame calls).  The
    // window is the workspace's max_batch; a max_batch = 1 workspace runs frame by frame.
    const int32_t window = ws ? std::max<int32_t>(1, std::min<int32_t>(SAF_MAX_BATCH, ws->max_batch)) : 1;
    // Contiguous sub-slab volumes (multi-GPU): drop the frames that cannot touch this slab before forming windows.
    // First on the host from the pose alone (no device work, no synchronisation); then, for long sequences, the
    // depth-aware pre-pass kernel (walls between the camera and the slab), which costs one small kernel, a copy
    // of the frame descriptors into the workspace and one stream synchronisation.  SAF_REACH_PREPASS=0 in the
    // environment keeps only the host test.
    std::vector<saf_frame> kept;
    if (grid && frames && ws && ws->base && grid->x_span == 0 && (grid->x_begin > 0 || grid->x_end < grid->nvox[0])) {
        FusionParams p;
        rc = build_params(grid, vol, frames, 1, H, W, trunc, rgb_mode, ws, &p, 0);
        if (rc) return rc;
        kept.reserve((size_t)n_frames);
        for (int32_t i = 0; i < n_frames; ++i)
            if (frames[i].pose_device || frame_may_reach_slab_host(*grid, frames[i].pose, frames[i].K, H, W))
                kept.push_back(frames[i]);
        static const bool prepass = !(getenv("SAF_REACH_PREPASS") && getenv("SAF_REACH_PREPASS")[0] == '0');
        const int32_t n_host = (int32_t)kept.size();
        if (prepass && n_host >= 16) {
            // scratch: the (idle) list region of slot 0 holds the frame descriptors and the flags - no allocation
            const size_t fbytes = sizeof(saf_frame) * (size_t)n_host, rbytes = sizeof(uint32_t) * (size_t)n_host;
            if (fbytes + rbytes <= (size_t)p.list_cap * sizeof(ValidEntry)) {   // else: keep every frame
                unsigned char* dbuf = reinterpret_cast<unsigned char*>(p.lists);
                std::vector<uint32_t> reach((size_t)n_host, 1u);
                SAF_CUDA_TRY(cudaMemcpyAsync(dbuf, kept.data(), fbytes, cudaMemcpyHostToDevice, st));
                frame_reach_kernel<<<n_host, 256, 0, st>>>(p, (const saf_frame*)dbuf, n_host, (uint32_t*)(dbuf + fbytes));
                SAF_CHECK_LAUNCH("frame_reach_kernel", st);
                SAF_CUDA_TRY(cudaMemcpyAsync(reach.data(), dbuf + fbytes, rbytes, cudaMemcpyDeviceToHost, st));
                SAF_CUDA_TRY(cudaStreamSynchronize(st));
                size_t o = 0;
                for (int32_t i = 0; i < n_host; ++i)
                    if (reach[(size_t)i]) kept[o++] = kept[(size_t)i];
                kept.resize(o);
            }
        }
        const unsigned long long skipped = (unsigned long long)n_frames - kept.size();
        if (getenv("SAF_DEBUG_REACH"))
            fprintf(stderr, "[saf] reach: %d frames, %d after the pose test, %zu kept\n", n_frames, n_host, kept.size());
        if (skipped) {
            add_skipped_frames_kernel<<<1, 1, 0, st>>>(p.hdr, skipped);
            SAF_CHECK_LAUNCH("add_skipped_frames_kernel", st);
        }
        frames = kept.data();
        n_frames = (int32_t)kept.size();
        if (n_frames == 0) return 0;
    }
    // windows of (nearly) equal length: 100 frames are 7 windows of 14-15 frames, not 6 of 16 and one of 4 whose
    // rows would be read and written for a quarter of the updates
    const int32_t n_calls = (n_frames + window - 1) / window;
    const int32_t w_base = n_frames / n_calls, w_rem = n_frames % n_calls;
    auto call_begin = [&](int32_t c) { return c * w_base + std::min(c, w_rem); };
    if (n_calls < 4) {
        for (int32_t c = 0; c < n_calls; ++c) {
            const int32_t i0 = call_begin(c), nb = call_begin(c + 1) - i0;
            rc = integrate_call(grid, vol, frames + i0, nb, H, W, trunc, rgb_mode, ws, 0, sms, smem_optin, st, st, nullptr,
                                nullptr, true);
            if (rc) return rc;
        }
        return 0;
    }
    SeqStreams* ss = nullptr;
    if ((rc = seq_streams(&ss))) return rc;
    // the side stream starts after everything already queued on the caller's stream
    SAF_CUDA_TRY(cudaEventRecord(ss->fork, st));
    SAF_CUDA_TRY(cudaStreamWaitEvent(ss->side, ss->fork, 0));
    for (int32_t c = 0; c < n_calls && rc == 0; ++c) {
        const uint32_t slot = (uint32_t)(c & 1);
        const int32_t i0 = call_begin(c), nb = call_begin(c + 1) - i0;
        // slot reuse: the feature kernel of call c-2 must have finished reading this slot's lists
        if (c >= 2) rc = (int)cudaStreamWaitEvent(ss->side, ss->feat_done[slot], 0);
        if (rc == 0)
            rc = integrate_call(grid, vol, frames + i0, nb, H, W, trunc, rgb_mode, ws, slot, sms, smem_optin, ss->side, st,
                                ss->geo_done[slot], ss->feat_done[slot], true, ss->tile, ss->tile_done[slot]);
    }
    // Whatever happened, the caller's stream must not run ahead of the kernels already queued on the side stream
    // (after an error the caller drops the frame tensors and the workspace; the last K2 is otherwise awaited
    // through geo_done).
    if (cudaEventRecord(ss->join, ss->side) == cudaSuccess) cudaStreamWaitEvent(st, ss->join, 0);
    if (cudaEventRecord(ss->tile_join, ss->tile) == cudaSuccess) cudaStreamWaitEvent(st, ss->tile_join, 0);
    return rc;
}

int saf_label_argmax(const int32_t* labels, int64_t n, int32_t n_classes, int64_t* out, void* stream)
{
    int sms = 0;
    int rc = device_sm_count(&sms, nullptr);
    if (rc) return rc;
    if (!labels || !out) return SAF_ERR_NULL;
    if (n < 0 || n_classes <= 0) return SAF_ERR_SHAPE;
    if (n == 0) return 0;
    const int64_t want = (n + 8 * kArgmaxRows - 1) / (8 * kArgmaxRows);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sms * 8));
    label_argmax_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(labels, n, n_classes, (long long*)out);
    SAF_CHECK_LAUNCH("label_argmax_kernel", (cudaStream_t)stream);
    return 0;
}

}  // extern "C"
