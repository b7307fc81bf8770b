// saf_topk.cu -- top-k feature rows per text embedding without keeping the [M,T] score matrix.
//
// The reference ranks with torch.argsort over the whole score matrix
// (/root/reference/eval_scannet_segmentation.py:546-561) or thresholds it
// (query_mesh.py:59-73); at M = 24 M voxels x T = 256 texts that matrix is 24.6 GB, so scores are
// produced in row chunks and reduced on the fly:
//   topk_scan_kernel   each CTA owns 8 texts x one row split and keeps their k best (score, row)
//                      in shared memory, one warp per text; a score is queued for insertion only
//                      if it beats the current k-th best, which after warm-up is rare
//                      (~k ln(M/k) times per text)
//   topk_merge_kernel  one CTA per text ranks the splits' candidates -> final sorted top-k
// Order: descending score, ties to the lower row index (a total order, so the result does not
// depend on the insertion order).
#include <math.h>

#include <algorithm>

#include "saf_internal.cuh"

namespace saf {

uint64_t query_topk_tc_workspace_bytes(int32_t T);
int query_topk_tc(const float* feats, int64_t M, int32_t C, int64_t ldf, const float* text, int32_t T, int32_t norm_mode,
                  int32_t k, int64_t index_base, float* out_scores, int64_t* out_index, void* ws, cudaStream_t st);

constexpr int kScanThreads = 256;
constexpr int kTextsPerCta = 8;

__device__ __forceinline__ bool beats(float s, long long i, float s2, long long i2)
{
    return s > s2 || (s == s2 && i < i2);
}

struct Partial {  // [splits][T][k]
    float* score;
    long long* index;
};

// warp-wide argmin under `beats` of a packed list (the element every other one beats)
__device__ __forceinline__ void find_worst(const float* ls, const long long* li, int k, int lane, float& ws,
                                           long long& wi, int& wp)
{
    ws = INFINITY;
    wi = -1;
    wp = 0;
    for (int q = lane; q < k; q += 32) {
        const float qs = ls[q];
        const long long qi = li[q];
        if (beats(ws, wi, qs, qi)) {
            ws = qs;
            wi = qi;
            wp = q;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float os = __shfl_xor_sync(0xffffffffu, ws, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, wi, o);
        const int op = __shfl_xor_sync(0xffffffffu, wp, o);
        if (beats(ws, wi, os, oi)) {
            ws = os;
            wi = oi;
            wp = op;
        }
    }
}

// S: [rows, T] scores of rows [row0, row0 + rows) (global row = index_base + row0 + r).
// 256 threads = 8 warps; the CTA owns 8 texts and warp w is the only writer of text w's list, so
// no locks: per block of 256 rows every thread pushes its candidates (score beats the current
// k-th best) into the text's queue, then warp w drains queue w.
__global__ void __launch_bounds__(kScanThreads) topk_scan_kernel(const float* __restrict__ S, int64_t rows, int T,
                                                                 int64_t row_id0, int k, Partial part, int first_chunk)
{
    extern __shared__ __align__(16) unsigned char smem[];
    long long* best_i = reinterpret_cast<long long*>(smem);                       // [8][k]
    float* best_s = reinterpret_cast<float*>(best_i + (size_t)kTextsPerCta * k);  // [8][k]
    __shared__ long long q_i[kTextsPerCta][kScanThreads];
    __shared__ float q_s[kTextsPerCta][kScanThreads];
    __shared__ int q_n[kTextsPerCta];
    __shared__ float thr_s[kTextsPerCta];
    __shared__ long long thr_i[kTextsPerCta];

    const int t0 = blockIdx.x * kTextsPerCta;
    const int split = blockIdx.y, splits = gridDim.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // load this (split, text) partial list from the previous chunk; lists are packed: valid entries
    // (index >= 0) occupy [0, count)
    for (int e = threadIdx.x; e < kTextsPerCta * k; e += kScanThreads) {
        const int j = e / k, q = e - j * k;
        const int t = t0 + j;
        float sv = -INFINITY;
        long long iv = -1;
        if (!first_chunk && t < T) {
            sv = part.score[((size_t)split * T + t) * k + q];
            iv = part.index[((size_t)split * T + t) * k + q];
        }
        best_s[e] = sv;
        best_i[e] = iv;
    }
    if (threadIdx.x < kTextsPerCta) q_n[threadIdx.x] = 0;
    __syncthreads();
    // warp-private state of text `warp`
    float* ls = best_s + (size_t)warp * k;
    long long* li = best_i + (size_t)warp * k;
    int count = 0;
    for (int q = lane; q < k; q += 32) count += (li[q] >= 0) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
    float ws;
    long long wi;
    int wp;
    find_worst(ls, li, k, lane, ws, wi, wp);
    if (lane == 0) {
        thr_s[warp] = (count == k) ? ws : -INFINITY;
        thr_i[warp] = (count == k) ? wi : (long long)0x7fffffffffffffffll;
    }
    __syncthreads();

    const int64_t per_split = (rows + splits - 1) / splits;
    const int64_t r_begin = (int64_t)split * per_split;
    const int64_t r_end = min(rows, r_begin + per_split);
    for (int64_t rb = r_begin; rb < r_end; rb += kScanThreads) {
        const int64_t r = rb + threadIdx.x;
        const bool in = r < r_end;
        const long long gid = row_id0 + r;
#pragma unroll
        for (int j = 0; j < kTextsPerCta; ++j) {
            const float sc = (in && t0 + j < T) ? S[r * T + t0 + j] : -INFINITY;
            if (in && (t0 + j < T) && !(sc != sc) && beats(sc, gid, thr_s[j], thr_i[j])) {
                const int pos = atomicAdd(&q_n[j], 1);
                q_s[j][pos] = sc;
                q_i[j][pos] = gid;
            }
        }
        __syncthreads();
        const int nq = q_n[warp];
        for (int e = 0; e < nq; ++e) {
            const float s = q_s[warp][e];
            const long long i = q_i[warp][e];
            if (count < k) {
                if (lane == 0) {
                    ls[count] = s;
                    li[count] = i;
                }
                ++count;
                __syncwarp();
                if (count == k) find_worst(ls, li, k, lane, ws, wi, wp);
            } else if (beats(s, i, ws, wi)) {
                if (lane == 0) {
                    ls[wp] = s;
                    li[wp] = i;
                }
                __syncwarp();
                find_worst(ls, li, k, lane, ws, wi, wp);
            }
        }
        __syncwarp();
        if (lane == 0) {
            q_n[warp] = 0;
            if (count == k) {
                thr_s[warp] = ws;
                thr_i[warp] = wi;
            }
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < kTextsPerCta * k; e += kScanThreads) {
        const int j = e / k, q = e - j * k;
        const int t = t0 + j;
        if (t < T) {
            part.score[((size_t)split * T + t) * k + q] = best_s[e];
            part.index[((size_t)split * T + t) * k + q] = best_i[e];
        }
    }
}

// one CTA per text: rank every candidate among the splits' lists; ranks < k are the answer
__global__ void __launch_bounds__(256) topk_merge_kernel(Partial part, int splits, int T, int k, int64_t index_base,
                                                         float* __restrict__ out_s, long long* __restrict__ out_i)
{
    const int t = blockIdx.x;
    const int n = splits * k;
    for (int q = threadIdx.x; q < k; q += blockDim.x) {
        out_s[(size_t)t * k + q] = -INFINITY;
        out_i[(size_t)t * k + q] = -1;
    }
    __syncthreads();
    for (int a = threadIdx.x; a < n; a += blockDim.x) {
        const int sa = a / k, qa = a - sa * k;
        const float s = part.score[((size_t)sa * T + t) * k + qa];
        const long long i = part.index[((size_t)sa * T + t) * k + qa];
        if (i < 0) continue;
        int rank = 0;
        for (int b = 0; b < n; ++b) {
            const int sb = b / k, qb = b - sb * k;
            const long long i2 = part.index[((size_t)sb * T + t) * k + qb];
            if (i2 < 0) continue;
            rank += beats(part.score[((size_t)sb * T + t) * k + qb], i2, s, i) ? 1 : 0;
        }
        if (rank < k) {
            out_s[(size_t)t * k + rank] = s;
            out_i[(size_t)t * k + rank] = i + index_base;
        }
    }
}

constexpr int kTopkSplits = 32;
// A chunk's [rows, T] score block is written by the scoring kernel and read straight back by the
// scan kernel: keep it around 48 MB so that it lives in the 126 MB L2 and never travels to HBM.
constexpr int64_t kTopkChunkBytes = 48ll << 20;

struct TopkLayout {
    uint64_t off_scores, off_part_s, off_part_i, bytes;
    int64_t chunk_rows;
};

static int topk_layout(int64_t M, int T, int k, TopkLayout* L)
{
    if (M < 0 || T <= 0 || k <= 0 || k > 2048) return SAF_ERR_SHAPE;
    int64_t chunk = kTopkChunkBytes / ((int64_t)T * 4);
    chunk = std::max<int64_t>(4096, chunk) / 128 * 128;
    L->chunk_rows = M < chunk ? (M > 0 ? M : 1) : chunk;
    L->off_scores = 0;
    L->off_part_s = align_up((uint64_t)L->chunk_rows * T * 4, 256);
    L->off_part_i = align_up(L->off_part_s + (uint64_t)kTopkSplits * T * k * 4, 256);
    L->bytes = align_up(L->off_part_i + (uint64_t)kTopkSplits * T * k * 8, 256);
    L->bytes = std::max<uint64_t>(L->bytes, align_up(query_topk_tc_workspace_bytes(T), 256));
    return 0;
}

}  // namespace saf

using namespace saf;

extern "C" {

int saf_query_topk_workspace_bytes(int64_t M, int32_t T, int32_t k, uint64_t* bytes_out)
{
    if (!bytes_out) return SAF_ERR_NULL;
    TopkLayout L;
    int rc = topk_layout(M, T, k, &L);
    if (rc) return rc;
    *bytes_out = L.bytes;
    return 0;
}

int saf_query_topk(const float* feats, int64_t M, int32_t C, int64_t ldf, const float* text, int32_t T,
                   int32_t norm_mode, int32_t score_mode, const float* surgery_w, int32_t precision, int32_t k,
                   int64_t index_base, float* out_scores, int64_t* out_index, void* ws, uint64_t ws_bytes, void* stream)
{
    int sms = 0, smem_optin = 0;
    int rc = device_sm_count(&sms, &smem_optin);
    if (rc) return rc;
    if (!feats || !text || !out_scores || !out_index || !ws) return SAF_ERR_NULL;
    if (((uintptr_t)ws & 255u) != 0) return SAF_ERR_ALIGNMENT;
    TopkLayout L;
    rc = topk_layout(M, T, k, &L);
    if (rc) return rc;
    if (ws_bytes < L.bytes) return SAF_ERR_WORKSPACE;
    const size_t smem = (size_t)kTextsPerCta * k * 12;
    if (smem + 32 * 1024 > (size_t)smem_optin) return SAF_ERR_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == 1 && M > 0) {
        // tensor-core path: exact top-k from tf32 scores + fp32 rescoring; plain dot / cosine only
        if (score_mode != SAF_SCORE_DOT) return SAF_ERR_UNSUPPORTED;
        if ((C % 4) != 0 || (ldf % 4) != 0 || (((uintptr_t)feats | (uintptr_t)text) & 15u) != 0)
            return SAF_ERR_ALIGNMENT;
        rc = query_topk_tc(feats, M, C, ldf, text, T, norm_mode, k, index_base, out_scores, out_index, ws, st);
        if (rc != 1) return rc;
        precision = 0;  // bucket overflow (mass ties): redo exactly with the fp32 chunked path
    }
    unsigned char* base = (unsigned char*)ws;
    float* scores = (float*)(base + L.off_scores);
    Partial part;
    part.score = (float*)(base + L.off_part_s);
    part.index = (long long*)(base + L.off_part_i);
    SAF_CUDA_TRY(cudaFuncSetAttribute(topk_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const dim3 grid((T + kTextsPerCta - 1) / kTextsPerCta, kTopkSplits);
    int first = 1;
    for (int64_t r0 = 0; r0 < M || first; r0 += L.chunk_rows) {
        const int64_t rows = M - r0 < L.chunk_rows ? M - r0 : L.chunk_rows;
        if (rows > 0) {
            rc = saf_query_scores(feats + r0 * ldf, rows, C, ldf, text, T, norm_mode, score_mode, surgery_w, precision,
                                  scores, stream);
            if (rc) return rc;
        }
        SAF_CHECK_LAUNCH("query_scores (top-k chunk)", st);
        topk_scan_kernel<<<grid, kScanThreads, smem, st>>>(scores, rows > 0 ? rows : 0, T, r0, k, part, first);
        SAF_CHECK_LAUNCH("topk_scan_kernel", st);
        first = 0;
    }
    topk_merge_kernel<<<T, 256, 0, st>>>(part, kTopkSplits, T, k, index_base, out_scores, (long long*)out_index);
    SAF_CHECK_LAUNCH("topk_merge_kernel", st);
    return 0;
}

}  // extern "C"
