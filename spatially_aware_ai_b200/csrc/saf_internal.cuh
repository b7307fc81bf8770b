// saf_internal.cuh -- workspace layout and small device helpers shared by the kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "saf_b200.h"

#define SAF_CUDA_TRY(expr)                         \
    do {                                           \
        cudaError_t _e = (expr);                   \
        if (_e != cudaSuccess) return (int)_e;     \
    } while (0)

namespace saf {

constexpr uint64_t kWsMagic = 0x5341465f42323030ull;  // "SAF_B200"
constexpr int kBlockEdge = SAF_BLOCK_EDGE;
constexpr int kBlockVoxels = kBlockEdge * kBlockEdge * kBlockEdge;

// Per-call counters.  There are two slots so that K1/K2 of frame i+1 (side stream) can run while K3 of
// frame i (main stream) is still consuming frame i's lists (saf_integrate_sequence).  They are produced
// by the kernels themselves (no host memsets): K2's last CTA publishes the per-block list offsets,
// n_valid, and folds the call into the running totals.
struct SlotCounters {
    uint32_t n_blocks;                      // list segments (blocks K2 processed) of the call, written by K2's last CTA
    uint32_t k2_done;                       // CTA completion ticket of K2
    uint32_t reserved_;
    uint32_t k3_next;                       // window mode: next unclaimed entry of the union list (K3W's dispenser)
    uint32_t n_valid[SAF_MAX_BATCH];        // length of each frame's valid list
    uint32_t n_tsdf_valid[SAF_MAX_BATCH];   // accumulators (atomics), folded and zeroed by K2's last CTA
    uint32_t last_tsdf_valid[SAF_MAX_BATCH];
    uint32_t n_processed;                   // list segments claimed by K2's CTAs (atomic, zeroed like above)
    uint32_t n_frustum_acc;                 // blocks inside a frustum before K1's depth test (atomic, zeroed like above)
    uint32_t last_processed;
    uint32_t n_frustum_blocks;              // blocks K1 listed for the call
    uint32_t n_union;                       // window mode: voxels valid in at least one frame of the window
    uint32_t acc_valid[SAF_MAX_BATCH];      // window mode: per-frame valid counts (atomics, folded like n_tsdf_valid)
};

// Device-resident workspace header (kWsHeaderBytes).
struct WsHeader {
    SlotCounters slot[2];
    uint32_t error_flags;
    uint32_t last_slot;                     // slot of the most recent call (for saf_read_stats)
    // Depth-aware block culling costs two extra L2 round trips on the K1 -> K2 critical path, so it is only
    // switched on (by K2's last CTA, for the following calls) while it pays: see tsdf_update_kernel.
    uint32_t depth_cull;
    uint32_t depth_cull_cooldown;
    unsigned long long total_frames;
    unsigned long long total_valid;
    unsigned long long total_tsdf_valid;
    unsigned long long total_blocks;
    unsigned long long total_calls;         // integrate() calls / windows (one K0+K1+K2 launch trio each)
    unsigned long long total_union;         // window mode: sum of the windows' union-list lengths
    // immutable after saf_workspace_init
    uint64_t magic;
    uint64_t bytes;
    uint64_t list_cap;                      // entries per per-frame valid list (= nblocks_total * 512)
    uint64_t max_table_elems;
    uint32_t nblocks_total;
    uint32_t max_batch;
    uint32_t nb[3];                         // blocks per axis of the slab
    uint32_t pad1_[1];
};
constexpr uint64_t kWsHeaderBytes = 1024;
static_assert(sizeof(WsHeader) <= kWsHeaderBytes, "workspace header grew past its slot");

// Window mode (saf_integrate_sequence): one voxel that is `valid` in at least one frame of the window.
// The (gx, gy) of frame b live in a separate array at [(rank * B + b) * 512 + local].
struct __align__(8) WinEntry {
    uint32_t voxel;        // slab-local flat index
    uint32_t mask_local;   // bits 0..15: frames of the window in which the voxel is valid; bits 16..24: index in its block
};

// One `valid` voxel of one frame: slab-local flat index and the normalised image coordinates
// the reference reuses for all three samplers (clip_seem_fusion.py:752).
struct __align__(16) ValidEntry {
    uint32_t voxel;
    float gx, gy;
    uint32_t pad;
};

constexpr int kK1Threads = 256;             // blocks tested per K1 CTA (one per thread)
constexpr uint32_t kMaxDepthTiles = 1024;   // depth image tiles (>= 32x32 px) whose maxima K1 publishes for K2's depth cull
constexpr uint32_t kMaxK1Ctas = 2048;       // K1's last CTA scans this many per-CTA counts in shared memory

// Workspace layout (all offsets 256-byte aligned):
//   header | slot 0 | slot 1        with, per slot,
//   cta_count[n_k1] | tile_dmax[max_batch][kMaxDepthTiles] | block_seg[n_k1*256] (uint2) | blk_count[max_batch][nblocks_total]
//   | blk_offset[max_batch][nblocks_total+1] | lists[max_batch][nblocks_total*512] | tables[max_batch][table_slot_elems]
struct WsLayout {
    uint64_t bytes;
    uint64_t list_cap;
    uint64_t slot0, slot_stride;            // byte offset of slot 0 and distance to slot 1
    uint64_t table_slot_elems;  // floats reserved per frame in tables[]
    uint64_t off_cta_count, off_cta_dmax, off_block_seg, off_blk_count, off_blk_offset, off_lists, off_tables;  // within a slot
    uint32_t nblocks_total;
    uint32_t n_k1;
    uint32_t nb[3];
};

inline uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

// x-planes held by a slab (contiguous, or block-cyclic stripes: see saf_grid_desc)
__host__ __device__ inline int slab_planes(const saf_grid_desc& g)
{
    const int len = g.x_end - g.x_begin;
    if (g.x_span <= 0) return len;
    const int full = len / g.x_stride, rem = len % g.x_stride;
    return full * g.x_span + (rem < g.x_span ? rem : g.x_span);
}
// global x index of slab-local plane lx
__host__ __device__ inline int slab_global_x(const saf_grid_desc& g, int lx)
{
    if (g.x_span <= 0) return g.x_begin + lx;
    return g.x_begin + (lx / g.x_span) * g.x_stride + (lx % g.x_span);
}
// sheared block-column layout (saf_grid_desc.y_ranks): local y extent and the global y of a local (x, y)
__host__ __device__ inline int slab_ny_local(const saf_grid_desc& g)
{
    if (g.y_ranks <= 1) return g.nvox[1];
    const int nby = (g.nvox[1] + SAF_BLOCK_EDGE - 1) / SAF_BLOCK_EDGE;
    return SAF_BLOCK_EDGE * ((nby + g.y_ranks - 1) / g.y_ranks);
}
__host__ __device__ inline int slab_global_y(const saf_grid_desc& g, int lx, int ly)
{
    if (g.y_ranks <= 1) return ly;
    const int n = g.y_ranks;
    const int shift = ((g.y_rank - lx / SAF_BLOCK_EDGE) % n + n) % n;
    return ((ly / SAF_BLOCK_EDGE) * n + shift) * SAF_BLOCK_EDGE + ly % SAF_BLOCK_EDGE;
}
inline bool slab_desc_ok(const saf_grid_desc& g)
{
    if (g.x_begin < 0 || g.x_end > g.nvox[0] || g.x_begin >= g.x_end) return false;
    if (g.y_ranks > 1)
        return g.x_begin == 0 && g.x_end == g.nvox[0] && g.x_span == 0 && g.y_rank >= 0 && g.y_rank < g.y_ranks;
    if (g.y_ranks < 0) return false;
    if (g.x_span == 0) return true;
    return g.x_span > 0 && g.x_span % SAF_BLOCK_EDGE == 0 && g.x_stride >= g.x_span && g.x_stride % SAF_BLOCK_EDGE == 0;
}

inline int compute_layout(const saf_grid_desc* g, int32_t max_batch, int64_t max_table_elems, WsLayout* L)
{
    if (!g || !L) return SAF_ERR_NULL;
    if (max_batch < 1 || max_batch > SAF_MAX_BATCH) return SAF_ERR_BATCH;
    if (g->nvox[0] <= 0 || g->nvox[1] <= 0 || g->nvox[2] <= 0 || !slab_desc_ok(*g) || !(g->voxel_size > 0.f))
        return SAF_ERR_GRID;
    if (max_table_elems < 0) return SAF_ERR_SHAPE;
    const uint64_t nxs = (uint64_t)slab_planes(*g);
    const uint64_t nyl = (uint64_t)slab_ny_local(*g);
    const uint64_t n = nxs * nyl * (uint64_t)g->nvox[2];
    if (n >= (1ull << 31)) return SAF_ERR_GRID;  // voxel indices are 32-bit
    L->nb[0] = (uint32_t)((nxs + kBlockEdge - 1) / kBlockEdge);
    L->nb[1] = (uint32_t)((nyl + kBlockEdge - 1) / kBlockEdge);
    L->nb[2] = (uint32_t)((g->nvox[2] + kBlockEdge - 1) / kBlockEdge);
    const uint64_t nblocks = (uint64_t)L->nb[0] * L->nb[1] * L->nb[2];
    L->n_k1 = (uint32_t)((nblocks + kK1Threads - 1) / kK1Threads);
    if (L->n_k1 > kMaxK1Ctas) return SAF_ERR_GRID;  // > 268 M voxels in one slab: shard it
    L->nblocks_total = (uint32_t)nblocks;
    L->list_cap = nblocks * kBlockVoxels;
    L->off_cta_count = 0;
    L->off_cta_dmax = align_up(L->off_cta_count + 4ull * L->n_k1, 256);
    L->off_block_seg = align_up(L->off_cta_dmax + 4ull * max_batch * kMaxDepthTiles, 256);
    L->off_blk_count = align_up(L->off_block_seg + 8ull * L->n_k1 * kK1Threads, 256);   // uint2 {block, frame mask}
    L->off_blk_offset = align_up(L->off_blk_count + 4ull * max_batch * nblocks, 256);
    L->off_lists = align_up(L->off_blk_offset + 4ull * max_batch * (nblocks + 1), 256);
    L->off_tables = align_up(L->off_lists + (uint64_t)max_batch * L->list_cap * sizeof(ValidEntry), 256);
    // window mode repacks every frame's feature image with a zero border: (npy+2)(npx+2) <= 9 npy npx rows
    L->table_slot_elems = (uint64_t)max_table_elems * (max_batch > 1 ? 9ull : 1ull);
    L->slot_stride = align_up(L->off_tables + (uint64_t)max_batch * L->table_slot_elems * 4ull, 256);
    L->slot0 = kWsHeaderBytes;
    L->bytes = L->slot0 + 2 * L->slot_stride;
    return 0;
}

// ---- device helpers ------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// streaming 128-bit accesses for the once-per-frame feature rows
__device__ __forceinline__ float4 ld_stream_f4(const float4* p)
{
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream_f4(float4* p, const float4& v)
{
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// ---- packed fp32 pairs (sm_100 {mul,fma}.rn.f32x2: two IEEE-rounded fp32 operations per instruction, the same
// roundings as the scalar intrinsics, half the issue slots; SASS FMUL2 / FFMA2).  No add2: ptxas fuses a packed
// mul + add pair into FFMA2 regardless of --fmad false, see mix_blend2 in saf_fusion.cu ----------------
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float lo, float hi)
{
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2_t v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t mul2_rn(f32x2_t a, f32x2_t b)
{
    f32x2_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t fma2_rn(f32x2_t a, f32x2_t b, f32x2_t c)
{
    f32x2_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ void st_stream_b64x2(void* p, f32x2_t lo, f32x2_t hi)
{
    asm volatile("st.global.L1::no_allocate.v2.b64 [%0], {%1,%2};" ::"l"(p), "l"(lo), "l"(hi) : "memory");
}
__device__ __forceinline__ void st_stream_b64(void* p, f32x2_t v)
{
    asm volatile("st.global.L1::no_allocate.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- exact fp32 scoring helpers shared by the fp32 query kernel and the top-k rescoring kernel, so that
// both produce bit-identical scores -------------------------------------------------------------------
__device__ __forceinline__ float row_scale(float norm2, int norm_mode)
{
    if (norm_mode == SAF_NORM_NONE) return 1.0f;
    const float nrm = sqrtf(norm2);
    if (norm_mode == SAF_NORM_CLAMP_MIN) return 1.0f / fmaxf(nrm, 0.1f);
    return nrm > 0.0f ? 1.0f / nrm : 0.0f;  // f/|f| then nan_to_num: zero rows stay zero
}
__device__ __forceinline__ float dot4_acc(const float4& f, const float4& x, float acc)
{
    return fmaf(f.x, x.x, fmaf(f.y, x.y, fmaf(f.z, x.z, fmaf(f.w, x.w, acc))));
}
__device__ __forceinline__ float sq4_acc(const float4& f, float acc)
{
    return acc + (f.x * f.x + f.y * f.y + f.z * f.z + f.w * f.w);
}

int device_sm_count(int* sms, int* smem_optin);

// SAF_DEBUG_SYNC=1 in the environment: synchronise after every launch and name the kernel that
// faulted on stderr (compute-sanitizer is not available on every pool).
int debug_check_launch(const char* what, cudaStream_t st);
#define SAF_CHECK_LAUNCH(what, st)                          \
    do {                                                    \
        int _rc = saf::debug_check_launch(what, st);        \
        if (_rc) return _rc;                                \
    } while (0)

}  // namespace saf
