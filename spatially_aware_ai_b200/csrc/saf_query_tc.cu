// saf_query_tc.cu -- the query GEMM S = F X^T on the 5th-gen tensor cores (tcgen05), sm_100a.
//
// Replaces the matmul of Clip.run_query / Clip.clip_feature_surgery
// (/root/reference/clipfusion.py:902, 909, 913-932) for large T.  F[M,C] is the fp32 feature
// matrix as it sits in HBM (the voxel grid's clip_feat or mesh-vertex features); it is fed to
// the tensor cores UNCONVERTED with kind::tf32 (the MMA reads the fp32 words and uses their top
// 19 bits), so the kernel moves every feature byte exactly once: TMA -> swizzled shared memory
// -> tcgen05.mma -> TMEM.  Row norms (the callers' F/|F|, clip_seem_fusion.py:507-511) are
// accumulated from the same shared-memory tiles by the epilogue warps while the MMAs run.
//
// One CTA per 128-row tile of F, all T (<= 256 per pass) texts at once:
//   warp 0      TMA producer: A tile [128 x 32 fp32] + B tile [N x 32 fp32] per k-block, 4 stages
//   warp 1      TMEM allocation, single-thread tcgen05.mma issue (4 x K=8 per k-block), commits
//   warps 2-5   per-row sum of squares from the A tiles; epilogue tcgen05.ld -> scale -> store
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "saf_internal.cuh"

namespace saf {

namespace tc {

constexpr int BM = 128;          // rows of F per CTA (UMMA M)
constexpr int BK = 32;           // fp32 per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 8;        // tf32 MMA K
constexpr int MAX_STAGES = 4;   // k-block ring depth (2 for the wide fused top-k kernel: two CTAs per SM)
constexpr int THREADS = 192;
constexpr uint32_t A_BYTES = BM * BK * 4;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128-byte swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart
// (cute::UMMA::SmemDescriptor: LBO = 1, SBO = 64 (x16 B), version = 1, layout = SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc(const void* smem_tile)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem_tile) >> 4) & 0x3FFFu);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)64 << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// cute::UMMA::InstrDescriptor for kind::tf32: D = F32, A = B = TF32, both K-major, M = 128, N = n.
__device__ __forceinline__ uint32_t umma_idesc_tf32(uint32_t n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// A row that may still belong to a text's top-k: tensor-core score and its error radius.
struct __align__(16) Candidate {
    float score;   // tf32 score (scaled)
    float eps;     // |exact - score| <= eps
    uint32_t row;  // row of F
    uint32_t pad;
};

// Filter state of the fused top-k (device arrays indexed by text).
struct FilterArgs {
    const float* thr;      // [T] a row is kept iff score + eps >= thr[t]
    const float* xnorm;    // [T] |X_t|
    Candidate* buckets;    // [T][cap]
    uint32_t* counts;      // [T]
    uint32_t* flags;       // bit 0: a bucket overflowed
    uint32_t cap;
};

// FILTER = false: out[m, t0 + t] = scale(m) * sum_c F[m,c] X[t0+t,c] for the CTA's 128 rows, t < t_valid.
// FILTER = true : nothing is written to `out`; instead every (row, text) whose score may still reach
//                 the text's current top-k threshold is appended to that text's candidate bucket.
template <bool FILTER>
__global__ void __launch_bounds__(THREADS, 1)
query_gemm_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int64_t M,
                       int C, int n_pad, int t_valid, int t0, uint32_t tmem_cols, int norm_mode, float* __restrict__ out,
                       int64_t ldo, int64_t tile0, const FilterArgs fa, const int STAGES)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t stage_bytes = A_BYTES + (uint32_t)n_pad * BK * 4;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_k = (C + BK - 1) / BK;
    const int64_t m0 = ((int64_t)blockIdx.x + tile0) * BM;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1 + 4);  // MMA commit + one arrive per norm warp
        }
        mbar_init(tmem_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---- TMA producer -------------------------------------------------------------------
        if (lane == 0) {
            for (int k = 0; k < num_k; ++k) {
                const int s = k % STAGES;
                const uint32_t ph = (uint32_t)(k / STAGES) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                unsigned char* a_dst = smem + (size_t)s * stage_bytes;
                mbar_arrive_expect_tx(&full[s], stage_bytes);
                tma_load_2d(a_dst, &map_a, k * BK, (int)m0, &full[s]);
                tma_load_2d(a_dst + A_BYTES, &map_b, k * BK, t0, &full[s]);
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer ---------------------------------------------------------------------
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32((uint32_t)n_pad);
            for (int k = 0; k < num_k; ++k) {
                const int s = k % STAGES;
                const uint32_t ph = (uint32_t)(k / STAGES) & 1u;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const unsigned char* a_src = smem + (size_t)s * stage_bytes;
                const uint64_t adesc = umma_desc(a_src), bdesc = umma_desc(a_src + A_BYTES);
#pragma unroll
                for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                    // advance 32 bytes inside the 128-byte swizzle row: +2 in the 16-byte address field
                    umma_tf32(tmem_base, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc,
                              (uint32_t)((k | kk) != 0));
                }
                umma_commit(&empty[s]);  // frees the stage once these MMAs have read it
            }
            umma_commit(tmem_full);      // accumulator complete
        }
    } else {
        // ---- row norms during the main loop, then the epilogue --------------------------------
        const int row = (warp & 3) * 32 + lane;  // TMEM lane group of this warp = warp % 4
        float norm2 = 0.0f;
        for (int k = 0; k < num_k; ++k) {
            const int s = k % STAGES;
            const uint32_t ph = (uint32_t)(k / STAGES) & 1u;
            mbar_wait(&full[s], ph);
            if (FILTER || norm_mode != SAF_NORM_NONE) {
                const float4* a4 = reinterpret_cast<const float4*>(smem + (size_t)s * stage_bytes) + row * 8;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    // the swizzle only permutes the eight 16-byte chunks inside the row: rotate the
                    // starting chunk by the row so that a quarter-warp touches all 32 banks
                    const float4 v = a4[(j + row) & 7];
                    norm2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, norm2))));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        const float scale = row_scale(norm2, norm_mode);
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        if (FILTER) {
            // tf32 keeps 10 explicit mantissa bits of each operand (truncation): every product is off by
            // less than 2^-9 relative, so |exact - tf32| <= 2^-9 |f| |x| (Cauchy-Schwarz) plus fp32 slack.
            const float eps_row = 0.001953125f * sqrtf(norm2) * scale;
            const int64_t m = m0 + row;
            // An all-zero row (an unobserved voxel: most of a fused grid) scores exactly 0 against every text.  It
            // is never a candidate here - millions of rows tied at 0 would flood the buckets - and only noted:
            // the final kernel asks for the exact path in the rare case that a text's k-th best score is <= 0.
            const bool zero_row = m < M && norm2 == 0.0f;
            if (__any_sync(0xffffffffu, zero_row) && lane == 0) fa.flags[1] = 1u;
            const uint32_t lane_base_f = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
            for (int c0 = 0; c0 < n_pad; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(lane_base_f + (uint32_t)c0, r);
                if (m < M && !zero_row) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int t = t0 + c0 + c;
                        if (c0 + c < t_valid) {
                            const float sc = __uint_as_float(r[c]) * scale;
                            const float e = eps_row * __ldg(fa.xnorm + t) + 1e-6f;
                            if (sc + e >= __ldg(fa.thr + t)) {
                                const uint32_t pos = atomicAdd(fa.counts + t, 1u);
                                if (pos < fa.cap) {
                                    Candidate cd;
                                    cd.score = sc;
                                    cd.eps = e;
                                    cd.row = (uint32_t)m;
                                    cd.pad = 0;
                                    fa.buckets[(size_t)t * fa.cap + pos] = cd;
                                } else {
                                    atomicOr(fa.flags, 1u);
                                }
                            }
                        }
                    }
                }
            }
        } else {
        // All MMAs have retired, so the stage buffers are free: each warp parks its 32 score rows
        // there (row pitch n_pad + 4 floats keeps the 128-bit shared stores conflict free) and
        // then streams them out row by row with fully coalesced global stores.
        const int pitch = n_pad + 4;
        float* park = reinterpret_cast<float*>(smem) + (size_t)(warp & 3) * 32 * pitch;
        const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        for (int c0 = 0; c0 < n_pad; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(lane_base + (uint32_t)c0, r);
            float4* dst = reinterpret_cast<float4*>(park + (size_t)lane * pitch + c0);
#pragma unroll
            for (int c = 0; c < 32; c += 4)
                if (c0 + c < n_pad)
                    dst[c >> 2] = make_float4(__uint_as_float(r[c]) * scale, __uint_as_float(r[c + 1]) * scale,
                                              __uint_as_float(r[c + 2]) * scale, __uint_as_float(r[c + 3]) * scale);
        }
        __syncwarp();
        const int64_t row0 = m0 + (warp & 3) * 32;
        const bool vec_ok = ((ldo | (int64_t)t0) & 3) == 0 && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
        for (int rr = 0; rr < 32; ++rr) {
            const int64_t m = row0 + rr;
            if (m >= M) break;
            const float* src = park + (size_t)rr * pitch;
            float* dst = out + m * ldo + t0;
            if (vec_ok) {
                for (int c = lane * 4; c < t_valid; c += 128) {
                    const float4 v = *reinterpret_cast<const float4*>(src + c);
                    if (c + 3 < t_valid) {
                        *reinterpret_cast<float4*>(dst + c) = v;
                    } else {
                        dst[c] = v.x;
                        if (c + 1 < t_valid) dst[c + 1] = v.y;
                        if (c + 2 < t_valid) dst[c + 2] = v.z;
                    }
                }
            } else {
                for (int c = lane; c < t_valid; c += 32) dst[c] = src[c];
            }
        }
        }  // !FILTER
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}


// ---------------------------------------------------------------------------------------------
// Persistent form of the kernel above (the default): one CTA per SM walks the row tiles blockIdx.x, blockIdx.x +
// gridDim.x, ... with TWO TMEM accumulators, so nothing drains between tiles - the TMA warp keeps the k-block ring
// full across tile boundaries, the MMA warp starts tile i+1 in the other accumulator while the epilogue warps read
// tile i out of TMEM, and barrier set-up / TMEM allocation are paid once per SM instead of once per 128 rows.
//   warp 0      TMA producer          warp 1      TMEM allocation + single-thread tcgen05.mma issue
//   warps 2-5   row norms from the A tiles (handed to the epilogue through shared memory)
//   warps 6-9   epilogue: tcgen05.ld -> scale -> (scores: 128-column halves parked in shared memory, coalesced
//               stores | top-k: candidate filter), TMEM lane group = warp % 4
// Measured with the one-tile-per-CTA kernel (12 M rows x 256 texts): neither HBM (47 %), L2 (35 %) nor the tensor
// pipe (36 %) was busy - a tile cost ~20 us of which the overlapped main loop is ~9.
// ---------------------------------------------------------------------------------------------
constexpr int P_EPI_WARPS = 8;                   // two per TMEM lane group, each takes half of the columns
constexpr int P_THREADS = (6 + P_EPI_WARPS) * 32;
constexpr int P_PARK_COLS = 64;                  // columns of a tile parked at a time by a scores-epilogue warp
constexpr int P_PARK_PITCH = P_PARK_COLS + 4;    // floats; keeps the 128-bit shared stores conflict free

template <bool FILTER>
__global__ void __launch_bounds__(P_THREADS, 1)
query_gemm_tf32_persistent_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                  int64_t M, int C, int n_pad, int t_valid, int t0, int norm_mode, float* __restrict__ out,
                                  int64_t ldo, int64_t tile0, int64_t n_tiles, const FilterArgs fa, const int STAGES)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t stage_bytes = A_BYTES + (uint32_t)n_pad * BK * 4;
    unsigned char* after_ring = smem + (size_t)STAGES * stage_bytes;
    // scores: [P_EPI_WARPS][32][P_PARK_PITCH] park buffers; top-k: [256] {threshold, |x|} per text of this pass
    float* park_all = reinterpret_cast<float*>(after_ring);
    float2* sthr = reinterpret_cast<float2*>(after_ring);
    float* snorm = reinterpret_cast<float*>(after_ring + (FILTER ? 256 * 8 : P_EPI_WARPS * 32 * P_PARK_PITCH * 4));   // [2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(snorm + 2 * BM);
    uint64_t* full = bars;                      // [STAGES]
    uint64_t* empty = full + STAGES;            // [STAGES]
    uint64_t* tmem_full = empty + STAGES;       // [2]
    uint64_t* tmem_empty = tmem_full + 2;       // [2]
    uint64_t* norm_full = tmem_empty + 2;       // [2]
    uint64_t* norm_empty = norm_full + 2;       // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(norm_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_k = (C + BK - 1) / BK;
    const int64_t my_tiles = (int64_t)blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1 + 4);  // MMA commit + one arrive per norm warp
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], P_EPI_WARPS);
            mbar_init(&norm_full[b], 4);
            mbar_init(&norm_empty[b], P_EPI_WARPS);
        }
        fence_mbar_init();
    }
    if (FILTER) {
        // the thresholds are constants of a launch (they rise between waves): one copy per CTA instead of two global
        // loads per score; a padding column can never be a candidate
        for (int t = threadIdx.x; t < 256; t += P_THREADS)
            sthr[t] = t < t_valid ? make_float2(__ldg(fa.thr + t0 + t), __ldg(fa.xnorm + t0 + t)) : make_float2(INFINITY, 0.f);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---- TMA producer: one k-block stream across all of this CTA's tiles --------------------------
        if (lane == 0) {
            uint32_t g = 0;
            for (int64_t i = 0; i < my_tiles; ++i) {
                const int64_t m0 = (tile0 + blockIdx.x + i * gridDim.x) * BM;
                for (int k = 0; k < num_k; ++k, ++g) {
                    const uint32_t s = g % (uint32_t)STAGES, ph = (g / (uint32_t)STAGES) & 1u;
                    mbar_wait(&empty[s], ph ^ 1u);
                    unsigned char* a_dst = smem + (size_t)s * stage_bytes;
                    mbar_arrive_expect_tx(&full[s], stage_bytes);
                    tma_load_2d(a_dst, &map_a, k * BK, (int)m0, &full[s]);
                    tma_load_2d(a_dst + A_BYTES, &map_b, k * BK, t0, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer: tile i accumulates in TMEM columns [(i & 1) * 256, +n_pad) --------------------
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32((uint32_t)n_pad);
            uint32_t g = 0;
            for (int64_t i = 0; i < my_tiles; ++i) {
                const uint32_t buf = (uint32_t)i & 1u, use = (uint32_t)(i >> 1);
                mbar_wait(&tmem_empty[buf], (use & 1u) ^ 1u);   // the epilogue has read this accumulator's previous tile
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * 256u;
                for (int k = 0; k < num_k; ++k, ++g) {
                    const uint32_t s = g % (uint32_t)STAGES, ph = (g / (uint32_t)STAGES) & 1u;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const unsigned char* a_src = smem + (size_t)s * stage_bytes;
                    const uint64_t adesc = umma_desc(a_src), bdesc = umma_desc(a_src + A_BYTES);
#pragma unroll
                    for (int kk = 0; kk < BK / UMMA_K; ++kk)
                        umma_tf32(tmem_d, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc,
                                  (uint32_t)((k | kk) != 0));
                    umma_commit(&empty[s]);   // frees the stage once these MMAs have read it
                }
                umma_commit(&tmem_full[buf]);   // accumulator complete
            }
        }
    } else if (warp < 6) {
        // ---- row norms, one tile ahead of the epilogue ---------------------------------------------
        const int row = (warp - 2) * 32 + lane;
        uint32_t g = 0;
        for (int64_t i = 0; i < my_tiles; ++i) {
            const uint32_t buf = (uint32_t)i & 1u, use = (uint32_t)(i >> 1);
            float norm2 = 0.0f;
            for (int k = 0; k < num_k; ++k, ++g) {
                const uint32_t s = g % (uint32_t)STAGES, ph = (g / (uint32_t)STAGES) & 1u;
                mbar_wait(&full[s], ph);
                if (FILTER || norm_mode != SAF_NORM_NONE) {
                    const float4* a4 = reinterpret_cast<const float4*>(smem + (size_t)s * stage_bytes) + row * 8;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        // the swizzle only permutes the eight 16-byte chunks inside the row: rotate the starting
                        // chunk by the row so that a quarter-warp touches all 32 banks.  (Same summation order as
                        // the one-tile kernel: the norms, and with them the scores, are bit-identical.)
                        const float4 v = a4[(j + row) & 7];
                        norm2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, norm2))));
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
            }
            mbar_wait(&norm_empty[buf], (use & 1u) ^ 1u);
            snorm[buf * BM + row] = norm2;
            __syncwarp();
            if (lane == 0) mbar_arrive(&norm_full[buf]);
        }
    } else {
        // ---- epilogue ------------------------------------------------------------------------------
        const int q = warp & 3;                    // TMEM lane group of this warp
        const int half = (warp - 6) >> 2;          // which half of the columns
        const int n_half = ((n_pad / 2 + 31) / 32) * 32;
        const int col_begin = half * n_half, col_end = min(n_pad, col_begin + n_half);
        const int row = q * 32 + lane;
        float* park = park_all + (size_t)(warp - 6) * 32 * P_PARK_PITCH;
        for (int64_t i = 0; i < my_tiles; ++i) {
            const uint32_t buf = (uint32_t)i & 1u, use = (uint32_t)(i >> 1);
            const int64_t m0 = (tile0 + blockIdx.x + i * gridDim.x) * BM;
            mbar_wait(&norm_full[buf], use & 1u);
            const float norm2 = snorm[buf * BM + row];
            __syncwarp();
            if (lane == 0) mbar_arrive(&norm_empty[buf]);
            const float scale = row_scale(norm2, norm_mode);
            mbar_wait(&tmem_full[buf], use & 1u);
            tc_fence_after();
            const uint32_t lane_base = tmem_base + buf * 256u + ((uint32_t)(q * 32) << 16);
            bool released = false;
            auto release = [&]() {   // this warp's part of the accumulator is in registers: hand it back to the MMA warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[buf]);
                released = true;
            };
            if (FILTER) {
                // (see query_gemm_tf32_kernel for the error radius and the treatment of all-zero rows)
                const float eps_row = 0.001953125f * sqrtf(norm2) * scale;
                const int64_t m = m0 + row;
                const bool zero_row = m < M && norm2 == 0.0f;
                if (half == 0 && __any_sync(0xffffffffu, zero_row) && lane == 0) fa.flags[1] = 1u;
                for (int c0 = col_begin; c0 < col_end; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(lane_base + (uint32_t)c0, r);
                    if (c0 + 32 >= col_end) release();
                    const bool live = m < M && !zero_row;
                    if (__any_sync(0xffffffffu, live)) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            const float2 tx = sthr[c0 + c];   // broadcast
                            const float sc = __uint_as_float(r[c]) * scale;
                            const float e = eps_row * tx.y + 1e-6f;
                            const bool hit = live && sc + e >= tx.x;
                            // one counter update per warp and text: while the thresholds are still -inf (first wave)
                            // every row is a candidate and 32 same-address atomics per column would serialise
                            const uint32_t bal = __ballot_sync(0xffffffffu, hit);
                            if (bal) {
                                const int t = t0 + c0 + c;
                                const int leader = __ffs(bal) - 1;
                                uint32_t base = 0;
                                if (lane == leader) base = atomicAdd(fa.counts + t, (uint32_t)__popc(bal));
                                base = __shfl_sync(0xffffffffu, base, leader);
                                if (hit) {
                                    const uint32_t pos = base + (uint32_t)__popc(bal & ((1u << lane) - 1u));
                                    if (pos < fa.cap) {
                                        Candidate cd;
                                        cd.score = sc;
                                        cd.eps = e;
                                        cd.row = (uint32_t)m;
                                        cd.pad = 0;
                                        fa.buckets[(size_t)t * fa.cap + pos] = cd;
                                    } else {
                                        atomicOr(fa.flags, 1u);
                                    }
                                }
                            }
                        }
                    }
                }
            } else {
                const int64_t row0 = m0 + q * 32;
                const bool vec_ok = ((ldo | (int64_t)t0) & 3) == 0 && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
                for (int h0 = col_begin; h0 < col_end; h0 += P_PARK_COLS) {
                    const int hcols = min(P_PARK_COLS, col_end - h0);
                    for (int c0 = 0; c0 < hcols; c0 += 32) {
                        uint32_t r[32];
                        tmem_ld32(lane_base + (uint32_t)(h0 + c0), r);
                        float4* dst = reinterpret_cast<float4*>(park + (size_t)lane * P_PARK_PITCH + c0);
#pragma unroll
                        for (int c = 0; c < 32; c += 4)
                            if (c0 + c < hcols)
                                dst[c >> 2] = make_float4(__uint_as_float(r[c]) * scale, __uint_as_float(r[c + 1]) * scale,
                                                          __uint_as_float(r[c + 2]) * scale, __uint_as_float(r[c + 3]) * scale);
                    }
                    if (h0 + P_PARK_COLS >= col_end) release();
                    __syncwarp();
                    const int hvalid = min(hcols, t_valid - h0);   // columns of this part that exist in `out`
                    if (hvalid > 0) {
                        if (vec_ok) {
                            // two rows per pass: 16 lanes x 16 bytes cover the 64 parked columns of a row
                            const int c = (lane & 15) * 4;
                            for (int rr = lane >> 4; rr < 32; rr += 2) {
                                const int64_t m = row0 + rr;
                                if (m >= M || c >= hvalid) continue;
                                const float4 v = *reinterpret_cast<const float4*>(park + (size_t)rr * P_PARK_PITCH + c);
                                float* dst = out + m * ldo + t0 + h0;
                                if (c + 3 < hvalid) {
                                    *reinterpret_cast<float4*>(dst + c) = v;
                                } else {
                                    dst[c] = v.x;
                                    if (c + 1 < hvalid) dst[c + 1] = v.y;
                                    if (c + 2 < hvalid) dst[c + 2] = v.z;
                                }
                            }
                        } else {
                            for (int rr = 0; rr < 32; ++rr) {
                                const int64_t m = row0 + rr;
                                if (m >= M) break;
                                for (int c = lane; c < hvalid; c += 32)
                                    out[m * ldo + t0 + h0 + c] = park[(size_t)rr * P_PARK_PITCH + c];
                            }
                        }
                    }
                    __syncwarp();   // the park buffer is reused by the next part / tile
                }
            }
            if (!released) release();   // a warp whose half of the columns is empty (few texts)
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode_fn(EncodeTiledFn* fn)
{
    static EncodeTiledFn cached = nullptr;
    if (!cached) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        SAF_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !p) return SAF_ERR_DEVICE;
        cached = (EncodeTiledFn)p;
    }
    *fn = cached;
    return 0;
}

// 2-D fp32 row-major matrix [rows, cols] with row pitch ld (elements); box = [box_rows, 32 cols], 128B swizzle.
static int make_map(EncodeTiledFn enc, CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols, uint64_t ld,
                    uint32_t box_rows)
{
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 4};
    cuuint32_t box[2] = {BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : SAF_ERR_SHAPE;
}



// Launches the GEMM over rows [row_begin, row_end) (row_begin a multiple of 128) for all T texts.
static int launch_gemm(bool filter, const float* feats, int64_t M, int32_t C, int64_t ldf, const float* text, int32_t T,
                       int32_t norm_mode, float* out, int64_t row_begin, int64_t row_end, const FilterArgs& fa,
                       cudaStream_t st)
{
    if ((((uintptr_t)feats | (uintptr_t)text) & 15u) != 0 || (ldf % 4) != 0 || (C % 4) != 0) return SAF_ERR_ALIGNMENT;
    if (M >= (1ll << 31)) return SAF_ERR_SHAPE;
    EncodeTiledFn enc;
    int rc = get_encode_fn(&enc);
    if (rc) return rc;
    CUtensorMap map_a;
    rc = make_map(enc, &map_a, feats, (uint64_t)M, (uint64_t)C, (uint64_t)ldf, BM);
    if (rc) return rc;
    const int64_t tile0 = row_begin / BM;
    const int64_t tiles = (row_end - row_begin + BM - 1) / BM;
    if (tiles <= 0) return 0;
    for (int t0 = 0; t0 < T; t0 += 256) {
        const int t_valid = std::min(256, T - t0);
        const int n_pad = std::max(16, (t_valid + 15) & ~15);
        uint32_t tmem_cols = 32;
        while (tmem_cols < (uint32_t)((n_pad + 31) & ~31)) tmem_cols <<= 1;
        CUtensorMap map_b;
        rc = make_map(enc, &map_b, text, (uint64_t)T, (uint64_t)C, (uint64_t)C, (uint32_t)n_pad);
        if (rc) return rc;
        const size_t stage_bytes = A_BYTES + (size_t)n_pad * BK * 4;
        // Persistent kernel (default): one CTA per SM over the tiles, two TMEM accumulators.  SAF_QUERY_PERSISTENT=0
        // selects the one-tile-per-CTA kernel.
        static const bool persistent = !(getenv("SAF_QUERY_PERSISTENT") && atoi(getenv("SAF_QUERY_PERSISTENT")) == 0);
        if (persistent) {
            int sms = 0;
            if ((rc = device_sm_count(&sms, nullptr))) return rc;
            const size_t extra = (filter ? (size_t)256 * 8 : (size_t)P_EPI_WARPS * 32 * P_PARK_PITCH * 4) + 2 * BM * 4 + 8 * 8 + 16;
            int STAGES = MAX_STAGES;
            while (STAGES > 2 && (size_t)STAGES * stage_bytes + 2 * STAGES * 8 + extra > 227 * 1024) --STAGES;
            const size_t smem = (size_t)STAGES * stage_bytes + 2 * STAGES * 8 + extra;
            const unsigned grid = (unsigned)std::min<int64_t>(tiles, sms);
            if (filter) {
                SAF_CUDA_TRY(cudaFuncSetAttribute(query_gemm_tf32_persistent_kernel<true>,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                query_gemm_tf32_persistent_kernel<true><<<grid, P_THREADS, smem, st>>>(
                    map_a, map_b, M, C, n_pad, t_valid, t0, norm_mode, out, (int64_t)T, tile0, tiles, fa, STAGES);
            } else {
                SAF_CUDA_TRY(cudaFuncSetAttribute(query_gemm_tf32_persistent_kernel<false>,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                query_gemm_tf32_persistent_kernel<false><<<grid, P_THREADS, smem, st>>>(
                    map_a, map_b, M, C, n_pad, t_valid, t0, norm_mode, out, (int64_t)T, tile0, tiles, fa, STAGES);
            }
            SAF_CHECK_LAUNCH("query_gemm_tf32_persistent_kernel", st);
            continue;
        }
        // The fused top-k kernel with more than 128 texts runs TWO CTAs per SM on a 2-deep ring each (2 x 96 KB of
        // shared memory, 2 x 256 TMEM columns): one CTA's filter epilogue overlaps the other's main loop.  The
        // scores kernel parks its output tile in the ring buffers and keeps 4 stages.  SAF_QUERY_STAGES overrides.
        static const int stages_env = getenv("SAF_QUERY_STAGES") ? atoi(getenv("SAF_QUERY_STAGES")) : 0;
        int STAGES = (filter && n_pad > 128) ? 2 : MAX_STAGES;
        if (filter && stages_env >= 2 && stages_env <= MAX_STAGES) STAGES = stages_env;
        const size_t smem = (size_t)STAGES * stage_bytes + (2 * STAGES + 1) * 8 + 16;
        if (filter) {
            SAF_CUDA_TRY(cudaFuncSetAttribute(query_gemm_tf32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem));
            query_gemm_tf32_kernel<true><<<(unsigned)tiles, THREADS, smem, st>>>(map_a, map_b, M, C, n_pad, t_valid, t0,
                                                                                tmem_cols, norm_mode, out, (int64_t)T,
                                                                                tile0, fa, STAGES);
        } else {
            SAF_CUDA_TRY(cudaFuncSetAttribute(query_gemm_tf32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem));
            query_gemm_tf32_kernel<false><<<(unsigned)tiles, THREADS, smem, st>>>(map_a, map_b, M, C, n_pad, t_valid, t0,
                                                                                 tmem_cols, norm_mode, out, (int64_t)T,
                                                                                 tile0, fa, STAGES);
        }
        SAF_CHECK_LAUNCH("query_gemm_tf32_kernel", st);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// fused top-k on the tensor cores: no score matrix, exact result
// ---------------------------------------------------------------------------------------------
//
// Rows are processed in waves (8 K rows, then x8 each).  The GEMM epilogue keeps only (row, text)
// pairs whose score interval [s - eps, s + eps] can still reach the text's threshold L_t = k-th
// largest lower bound seen so far; after each wave topk_tc_threshold_kernel raises L_t and compacts
// the bucket.  A rejected row had s + eps < L_t, and k rows with exact score >= L_t exist, so no
// rejected row can be in the exact top-k.  topk_tc_final_kernel rescoring the survivors in fp32
// (same arithmetic as the fp32 query kernel) therefore yields the exact top-k.

// Fused grids are smooth: thousands of neighbouring voxels can score within the tf32 error radius of a text's k-th
// best row.  The buckets are therefore generous, the waves stop growing at kWaveMax rows (the thresholds tighten after
// every wave, so a late wave admits few rows) and the final ranking takes a whole bucket.
constexpr uint32_t kBucketCap = 16384;
constexpr int kFinalMax = 16384;
constexpr int64_t kWaveMax = 1 << 20;

__device__ __forceinline__ uint32_t float_key(float f)
{
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void __launch_bounds__(128) topk_tc_init_kernel(const float* __restrict__ X, int C, float* xnorm, float* thr,
                                                           uint32_t* counts_a, uint32_t* counts_b, uint32_t* flags)
{
    __shared__ float red[4];
    const int t = blockIdx.x;
    float acc = 0.0f;
    for (int c = threadIdx.x; c < C; c += 128) acc = fmaf(X[(size_t)t * C + c], X[(size_t)t * C + c], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        xnorm[t] = sqrtf(red[0] + red[1] + red[2] + red[3]) * 1.0001f;
        thr[t] = -INFINITY;
        counts_a[t] = 0;
        counts_b[t] = 0;
        if (t == 0) {
            flags[0] = 0;
            flags[1] = 0;   // set by the filter when it skipped an all-zero row
        }
    }
}

// one CTA per text: L = k-th largest lower bound in `src`; survivors (upper bound >= L) -> `dst`
__global__ void __launch_bounds__(256) topk_tc_threshold_kernel(const Candidate* __restrict__ src_all, uint32_t* src_counts,
                                                                Candidate* __restrict__ dst_all, uint32_t* dst_counts,
                                                                float* thr, int k, uint32_t cap)
{
    extern __shared__ uint32_t keys[];  // [cap] orderable keys of the lower bounds
    __shared__ uint32_t s_cnt, s_out;
    const int t = blockIdx.x;
    const Candidate* src = src_all + (size_t)t * cap;
    Candidate* dst = dst_all + (size_t)t * cap;
    const uint32_t n = min(src_counts[t], cap);
    for (uint32_t i = threadIdx.x; i < n; i += 256) keys[i] = float_key(src[i].score - src[i].eps);
    if (threadIdx.x == 0) s_out = 0;
    __syncthreads();
    float L = -INFINITY;
    if (n >= (uint32_t)k) {
        uint32_t prefix = 0;
        for (int bit = 31; bit >= 0; --bit) {
            const uint32_t trial = prefix | (1u << bit);
            if (threadIdx.x == 0) s_cnt = 0;
            __syncthreads();
            uint32_t c = 0;
            for (uint32_t i = threadIdx.x; i < n; i += 256) c += keys[i] >= trial ? 1u : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
            __syncthreads();
            if (s_cnt >= (uint32_t)k) prefix = trial;
            __syncthreads();
        }
        L = key_float(prefix);
    }
    for (uint32_t i = threadIdx.x; i < n; i += 256) {
        const Candidate c = src[i];
        if (c.score + c.eps >= L) dst[atomicAdd(&s_out, 1u)] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        dst_counts[t] = s_out;
        src_counts[t] = 0;
        thr[t] = L;
    }
}

// one CTA per text: exact fp32 score of every survivor, then rank -> top-k (descending, ties to lower row)
__global__ void __launch_bounds__(256) topk_tc_final_kernel(const Candidate* __restrict__ buckets,
                                                            const uint32_t* __restrict__ counts, uint32_t cap,
                                                            const float* __restrict__ F, int C, int64_t ldf,
                                                            const float* __restrict__ X, int norm_mode, int k,
                                                            int64_t index_base, float* __restrict__ out_s,
                                                            long long* __restrict__ out_i, uint32_t* flags)
{
    extern __shared__ __align__(16) unsigned char final_smem[];
    float* ex_s = reinterpret_cast<float*>(final_smem);                 // [kFinalMax]
    uint32_t* ex_r = reinterpret_cast<uint32_t*>(ex_s + kFinalMax);     // [kFinalMax]
    const int t = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Candidate* b = buckets + (size_t)t * cap;
    uint32_t n = min(counts[t], cap);
    if (n > (uint32_t)kFinalMax) {
        if (threadIdx.x == 0) atomicOr(flags, 2u);
        n = kFinalMax;
    }
    for (int q = threadIdx.x; q < k; q += 256) {
        out_s[(size_t)t * k + q] = -INFINITY;
        out_i[(size_t)t * k + q] = -1;
    }
    const float4* x4 = reinterpret_cast<const float4*>(X + (size_t)t * C);
    for (uint32_t e = warp; e < n; e += 8) {
        const uint32_t row = b[e].row;
        const float4* f4 = reinterpret_cast<const float4*>(F + (size_t)row * ldf);
        float acc = 0.0f, norm2 = 0.0f;
        for (int c4 = lane; c4 < C / 4; c4 += 32) {
            const float4 f = __ldg(f4 + c4);
            norm2 = sq4_acc(f, norm2);
            acc = dot4_acc(f, __ldg(x4 + c4), acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            norm2 += __shfl_xor_sync(0xffffffffu, norm2, o);
            acc += __shfl_xor_sync(0xffffffffu, acc, o);
        }
        if (lane == 0) {
            ex_s[e] = acc * row_scale(norm2, norm_mode);
            ex_r[e] = row;
        }
    }
    __syncthreads();
    for (uint32_t a = threadIdx.x; a < n; a += 256) {
        const float sa = ex_s[a];
        const uint32_t ra = ex_r[a];
        int rank = 0;
        for (uint32_t j = 0; j < n; ++j) {
            const float sj = ex_s[j];
            rank += (sj > sa || (sj == sa && ex_r[j] < ra)) ? 1 : 0;
        }
        if (rank < k) {
            out_s[(size_t)t * k + rank] = sa;
            out_i[(size_t)t * k + rank] = (long long)ra + index_base;
        }
        // zero rows were left out of the buckets: they belong in the answer iff the k-th best score is not positive
        if (rank == k - 1 && !(sa > 0.0f) && flags[1]) atomicOr(flags, 1u);
    }
    if (threadIdx.x == 0 && n < (uint32_t)k && flags[1]) atomicOr(flags, 1u);
}

}  // namespace tc

uint64_t query_topk_tc_workspace_bytes(int32_t T)
{
    return 2ull * (uint64_t)T * tc::kBucketCap * sizeof(tc::Candidate) + 4ull * (uint64_t)T * 4 + 256;   // + flags[2]
}

// returns 0 on success, 1 when the caller must fall back to the exact chunked path (bucket overflow:
// e.g. millions of exactly tied rows), <0 / CUDA error otherwise.  Synchronises the stream.
int query_topk_tc(const float* feats, int64_t M, int32_t C, int64_t ldf, const float* text, int32_t T, int32_t norm_mode,
                  int32_t k, int64_t index_base, float* out_scores, int64_t* out_index, void* ws, cudaStream_t st)
{
    using namespace tc;
    unsigned char* base = (unsigned char*)ws;
    Candidate* bucket[2] = {(Candidate*)base, (Candidate*)(base + (uint64_t)T * kBucketCap * sizeof(Candidate))};
    uint32_t* u = (uint32_t*)(base + 2ull * (uint64_t)T * kBucketCap * sizeof(Candidate));
    uint32_t* counts[2] = {u, u + T};
    float* thr = (float*)(u + 2 * T);
    float* xnorm = (float*)(u + 3 * T);
    uint32_t* flags = u + 4 * T;
    topk_tc_init_kernel<<<T, 128, 0, st>>>(text, C, xnorm, thr, counts[0], counts[1], flags);
    SAF_CHECK_LAUNCH("topk_tc_init_kernel", st);
    SAF_CUDA_TRY(cudaFuncSetAttribute(topk_tc_threshold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(kBucketCap * 4)));
    int cur = 0;
    int64_t r0 = 0;
    int64_t wave = (int64_t)kBucketCap / BM * BM;
    while (r0 < M) {
        const int64_t r1 = std::min(M, r0 + wave);
        FilterArgs fa;
        fa.thr = thr;
        fa.xnorm = xnorm;
        fa.buckets = bucket[cur];
        fa.counts = counts[cur];
        fa.flags = flags;
        fa.cap = kBucketCap;
        int rc = launch_gemm(true, feats, M, C, ldf, text, T, norm_mode, nullptr, r0, r1, fa, st);
        if (rc) return rc;
        topk_tc_threshold_kernel<<<T, 256, kBucketCap * 4, st>>>(bucket[cur], counts[cur], bucket[cur ^ 1], counts[cur ^ 1],
                                                                 thr, k, kBucketCap);
        SAF_CHECK_LAUNCH("topk_tc_threshold_kernel", st);
        cur ^= 1;
        r0 = r1;
        wave = std::min<int64_t>(wave * 8, kWaveMax);
    }
    SAF_CUDA_TRY(cudaFuncSetAttribute(topk_tc_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFinalMax * 8));
    topk_tc_final_kernel<<<T, 256, kFinalMax * 8, st>>>(bucket[cur], counts[cur], kBucketCap, feats, C, ldf, text, norm_mode, k,
                                            index_base, out_scores, (long long*)out_index, flags);
    SAF_CHECK_LAUNCH("topk_tc_final_kernel", st);
    uint32_t h_flags = 0;
    SAF_CUDA_TRY(cudaMemcpyAsync(&h_flags, flags, 4, cudaMemcpyDeviceToHost, st));
    SAF_CUDA_TRY(cudaStreamSynchronize(st));
    return h_flags ? 1 : 0;
}

int query_scores_tc(const float* feats, int64_t M, int32_t C, int64_t ldf, const float* text, int32_t T,
                    int32_t norm_mode, int32_t precision, float* out, cudaStream_t st)
{
    if (precision != 1) return SAF_ERR_UNSUPPORTED;
    tc::FilterArgs fa;
    memset(&fa, 0, sizeof(fa));
    return tc::launch_gemm(false, feats, M, C, ldf, text, T, norm_mode, out, 0, M, fa, st);
}

}  // namespace saf
