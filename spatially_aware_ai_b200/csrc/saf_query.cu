// saf_query.cu -- text-query scoring of fused features.
//
// Replaces Clip.run_query (/root/reference/clipfusion.py:899-904), Clip.clip_feature_surgery
// (clipfusion.py:906-934) and the callers' row normalisation (clip_seem_fusion.py:507-511,
// hypersim_eval.py:50-51).  S = F X^T with F[M,C] fp32 rows (the voxel grid or mesh-vertex
// features) and X[T,C] unit text embeddings, followed by a row epilogue.
//
// precision 0: fp32 CUDA-core kernel (this file): one warp per feature row, the text block staged
//              in shared memory, butterfly reduction that leaves 32 scores in 32 lanes.
// precision 1: tcgen05 tensor-core kernel (saf_query_tc.cu).
#include <limits.h>
#include <math.h>
#include <algorithm>
#include <string.h>

#include "saf_internal.cuh"

namespace saf {

int query_scores_tc(const float* feats, int64_t M, int32_t C, int64_t ldf, const float* text, int32_t T,
                    int32_t norm_mode, int32_t precision, float* out, cudaStream_t st);

constexpr int kQThreads = 256;

// out[m, t0 + t] = scale(m) * <F[m,:], X[t0+t,:]>   for t in [0, Tt)
template <int VEC>
__global__ void __launch_bounds__(kQThreads) query_scores_fp32_kernel(const float* __restrict__ F, int64_t M, int C,
                                                                       int64_t ldf, const float* __restrict__ X, int t0,
                                                                       int Tt, int norm_mode, float* __restrict__ out,
                                                                       int64_t ldo)
{
    extern __shared__ __align__(16) float xs[];  // [Tt_pad][C], zero padded to a multiple of 32 texts
    const int Tt_pad = (Tt + 31) & ~31;
    for (int64_t e = threadIdx.x; e < (int64_t)Tt_pad * C; e += kQThreads) {
        const int t = (int)(e / C);
        xs[e] = t < Tt ? X[(int64_t)(t0 + t) * C + (e - (int64_t)t * C)] : 0.0f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * (kQThreads / 32);
    for (int64_t m = (int64_t)blockIdx.x * (kQThreads / 32) + (threadIdx.x >> 5); m < M; m += nwarps) {
        const float* row = F + m * ldf;
        float norm2 = 0.0f;
        for (int tb = 0; tb < Tt_pad; tb += 32) {
            float acc[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = 0.0f;
            if (VEC == 4) {
                for (int c4 = lane; c4 < C / 4; c4 += 32) {
                    const float4 f = __ldg(reinterpret_cast<const float4*>(row) + c4);
                    if (tb == 0) norm2 = sq4_acc(f, norm2);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float4 x = reinterpret_cast<const float4*>(xs + (size_t)(tb + i) * C)[c4];
                        acc[i] = dot4_acc(f, x, acc[i]);
                    }
                }
            } else {
                for (int c = lane; c < C; c += 32) {
                    const float f = __ldg(row + c);
                    if (tb == 0) norm2 += f * f;
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = fmaf(f, xs[(size_t)(tb + i) * C + c], acc[i]);
                }
            }
            if (tb == 0) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) norm2 += __shfl_xor_sync(0xffffffffu, norm2, o);
            }
            // butterfly transpose-reduce: 31 shuffles; lane l ends up with the full sum of text tb+l
#pragma unroll
            for (int o = 16, len = 32; o >= 1; o >>= 1, len >>= 1) {
                const bool hi = (lane & o) != 0;
#pragma unroll
                for (int i = 0; i < len / 2; ++i) {
                    const float a = acc[i], b = acc[i + len / 2];
                    acc[i] = (hi ? b : a) + __shfl_xor_sync(0xffffffffu, hi ? a : b, o);
                }
            }
            const int t = tb + lane;
            if (t < Tt) out[m * ldo + t0 + t] = acc[0] * row_scale(norm2, norm_mode);
        }
    }
}

// Row epilogues on S[M,T] in place.
//   SOFTMAX100: softmax(100*s) over t                              (clipfusion.py:902-903)
//   SURGERY:    w_t s_t - (1/T) sum_s w_s s_s                      (clipfusion.py:918-932)
__global__ void __launch_bounds__(256) query_epilogue_kernel(float* __restrict__ S, int64_t M, int T, int mode,
                                                             const float* __restrict__ w)
{
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x / 32);
    for (int64_t m = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5); m < M; m += nwarps) {
        float* row = S + m * T;
        if (mode == SAF_SCORE_SOFTMAX100) {
            float mx = -INFINITY;
            for (int t = lane; t < T; t += 32) mx = fmaxf(mx, 100.0f * row[t]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float sum = 0.0f;
            for (int t = lane; t < T; t += 32) sum += expf(100.0f * row[t] - mx);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            for (int t = lane; t < T; t += 32) row[t] = expf(100.0f * row[t] - mx) / sum;
        } else if (mode == SAF_SCORE_SURGERY) {
            float sum = 0.0f;
            for (int t = lane; t < T; t += 32) sum += w[t] * row[t];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float mean = sum / (float)T;
            for (int t = lane; t < T; t += 32) row[t] = w[t] * row[t] - mean;
        }
    }
}

static int query_scores_fp32(const float* F, int64_t M, int C, int64_t ldf, const float* X, int T, int norm_mode,
                             float* out, int sms, int smem_optin, cudaStream_t st)
{
    const bool vec4 = (C % 4 == 0) && (ldf % 4 == 0) && (((uintptr_t)F & 15u) == 0);
    // texts per pass: whole multiples of 32 that fit the shared-memory budget
    int per_pass = (int)(((size_t)smem_optin - 1024) / ((size_t)C * 4)) & ~31;
    if (per_pass < 32) return SAF_ERR_UNSUPPORTED;  // feature_dim too large for one 32-text block
    const int grid = (int)std::min<int64_t>((M + 7) / 8, (int64_t)sms * 2);
    for (int t0 = 0; t0 < T; t0 += per_pass) {
        const int Tt = min(per_pass, T - t0);
        const size_t smem = (size_t)((Tt + 31) & ~31) * C * 4;
        if (vec4) {
            SAF_CUDA_TRY(cudaFuncSetAttribute(query_scores_fp32_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem));
            query_scores_fp32_kernel<4><<<grid, kQThreads, smem, st>>>(F, M, C, ldf, X, t0, Tt, norm_mode, out, T);
        } else {
            SAF_CUDA_TRY(cudaFuncSetAttribute(query_scores_fp32_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem));
            query_scores_fp32_kernel<1><<<grid, kQThreads, smem, st>>>(F, M, C, ldf, X, t0, Tt, norm_mode, out, T);
        }
        SAF_CUDA_TRY(cudaGetLastError());
    }
    return 0;
}

// ---- per-row consumers of the score block: the reference's segment() (eval_scannet_segmentation.py:546-561,
// argsort of softmax(100 cos) over the texts, of which its callers use the first 1 / 5 columns) and the
// hypersim presence test (hypersim_eval.py:80-89, max over rows of softmax(100 [bg.., target])[-1]).
// Both run on a chunk of rows whose [rows, T] score block stays in L2, so the [M,T] matrix never exists. ----

constexpr int kRowChunk = 1 << 16;    // rows per chunk: 64 MB of scores at T = 256
constexpr int kMaxRowTexts = 2048;    // texts a warp keeps in registers (64 per lane)

// labels[m, j] = index of the j-th largest score of row m (ties to the lower text index: the order of a stable
// descending sort; softmax(100 s) is monotonic in s, so this is the reference's argsort wherever the
// probabilities differ).  One warp per row, k selection rounds.
template <int NV>   // score slots per lane: T <= 32 * NV
__global__ void __launch_bounds__(256) row_topk_kernel(const float* __restrict__ S, int64_t rows, int T, int k,
                                                       long long* __restrict__ labels, float* __restrict__ probs)
{
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x / 32);
    for (int64_t m = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5); m < rows; m += nwarps) {
        const float* row = S + m * T;
        float v[NV];
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int t = i * 32 + lane;
            v[i] = t < T ? row[t] : -INFINITY;
            mx = fmaxf(mx, v[i]);
        }
        float sum = 0.0f;
        if (probs) {   // softmax(100 s) denominators (clipfusion.py:902-903 / eval_scannet_segmentation.py:557)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
#pragma unroll
            for (int i = 0; i < NV; ++i)
                if (i * 32 + lane < T) sum += expf(100.0f * v[i] - 100.0f * mx);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        }
        for (int j = 0; j < k; ++j) {
            float best = -INFINITY;
            int best_t = INT_MAX;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int t = i * 32 + lane;
                // NaN scores (0/0 rows are filtered by the norm modes) never win; -inf marks taken entries
                if (t < T && (v[i] > best || (v[i] == best && t < best_t && v[i] != -INFINITY))) {
                    best = v[i];
                    best_t = t;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int ot = __shfl_xor_sync(0xffffffffu, best_t, o);
                if (ob > best || (ob == best && ot < best_t)) {
                    best = ob;
                    best_t = ot;
                }
            }
            if (best_t == INT_MAX) {   // fewer than k finite scores left: remaining texts in index order
                best_t = -1;
            }
            if (lane == 0) {
                labels[m * k + j] = best_t;
                if (probs) probs[m * k + j] = best_t >= 0 ? expf(100.0f * best - 100.0f * mx) / sum : 0.0f;
            }
#pragma unroll
            for (int i = 0; i < NV; ++i)
                if (i * 32 + lane == best_t) v[i] = -INFINITY;
        }
    }
}

// out[i] = max over rows of softmax(100 [s_bg0 .. s_bg(nb-1), s_target_i])[-1]; scores row = [bg.., targets..].
// Probabilities are positive, so the float maximum is an integer atomicMax on the bit patterns.
__global__ void __launch_bounds__(256) text_presence_kernel(const float* __restrict__ S, int64_t rows, int nb, int L,
                                                            float* __restrict__ out)
{
    const int T = nb + L;
    for (int i = blockIdx.y; i < L; i += gridDim.y) {
        float best = 0.0f;
        for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < rows; m += (int64_t)gridDim.x * blockDim.x) {
            const float* row = S + m * T;
            const float xt = 100.0f * row[nb + i];
            float mx = xt;
            for (int j = 0; j < nb; ++j) mx = fmaxf(mx, 100.0f * row[j]);
            float sum = 0.0f;
            for (int j = 0; j < nb; ++j) sum += expf(100.0f * row[j] - mx);
            const float et = expf(xt - mx);
            sum += et;
            best = fmaxf(best, et / sum);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
        if ((threadIdx.x & 31) == 0 && best > 0.0f) atomicMax(reinterpret_cast<int*>(out + i), __float_as_int(best));
    }
}

}  // namespace saf

using namespace saf;

extern "C" {

int saf_query_scores(const float* feats, int64_t M, int32_t C, int64_t ldf, const float* text, int32_t T,
                     int32_t norm_mode, int32_t score_mode, const float* surgery_w, int32_t precision, float* out,
                     void* stream)
{
    int sms = 0, smem_optin = 0;
    int rc = device_sm_count(&sms, &smem_optin);
    if (rc) return rc;
    if (!feats || !text || !out) return SAF_ERR_NULL;
    if (M < 0 || C <= 0 || T <= 0 || ldf < C) return SAF_ERR_SHAPE;
    if (norm_mode < SAF_NORM_NONE || norm_mode > SAF_NORM_CLAMP_MIN) return SAF_ERR_UNSUPPORTED;
    if (score_mode < SAF_SCORE_DOT || score_mode > SAF_SCORE_SURGERY) return SAF_ERR_UNSUPPORTED;
    if (score_mode == SAF_SCORE_SURGERY && !surgery_w) return SAF_ERR_NULL;
    if (M == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == 0)
        rc = query_scores_fp32(feats, M, C, ldf, text, T, norm_mode, out, sms, smem_optin, st);
    else if (precision == 1)
        rc = query_scores_tc(feats, M, C, ldf, text, T, norm_mode, precision, out, st);
    else
        rc = SAF_ERR_UNSUPPORTED;
    if (rc) return rc;
    if (score_mode != SAF_SCORE_DOT) {
        const int grid = (int)std::min<int64_t>((M + 7) / 8, (int64_t)sms * 8);
        query_epilogue_kernel<<<grid, 256, 0, st>>>(out, M, T, score_mode, surgery_w);
        SAF_CUDA_TRY(cudaGetLastError());
    }
    return 0;
}

int saf_query_rows_workspace_bytes(int32_t T, uint64_t* bytes_out)
{
    if (!bytes_out) return SAF_ERR_NULL;
    if (T <= 0) return SAF_ERR_SHAPE;
    *bytes_out = (uint64_t)kRowChunk * (uint64_t)T * sizeof(float);
    return 0;
}

int saf_query_row_labels(const float* feats, int64_t M, int32_t C, int64_t ldf, const float* text, int32_t T,
                         int32_t norm_mode, int32_t precision, int32_t k, int64_t* out_labels, float* out_probs,
                         void* ws, uint64_t ws_bytes, void* stream)
{
    int sms = 0;
    int rc = device_sm_count(&sms, nullptr);
    if (rc) return rc;
    if (!feats || !text || !out_labels || !ws) return SAF_ERR_NULL;
    if (M < 0 || C <= 0 || T <= 0 || ldf < C || k <= 0 || k > T) return SAF_ERR_SHAPE;
    if (T > kMaxRowTexts) return SAF_ERR_UNSUPPORTED;
    if (ws_bytes < (uint64_t)kRowChunk * (uint64_t)T * sizeof(float) || ((uintptr_t)ws & 255u)) return SAF_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    for (int64_t m0 = 0; m0 < M; m0 += kRowChunk) {
        const int64_t rows = std::min<int64_t>(kRowChunk, M - m0);
        rc = saf_query_scores(feats + m0 * ldf, rows, C, ldf, text, T, norm_mode, SAF_SCORE_DOT, nullptr, precision,
                              (float*)ws, stream);
        if (rc) return rc;
        const int grid = (int)std::min<int64_t>((rows + 7) / 8, (int64_t)sms * 8);
        long long* lab = (long long*)out_labels + m0 * k;
        float* pr = out_probs ? out_probs + m0 * k : nullptr;
        if (T <= 256)
            row_topk_kernel<8><<<grid, 256, 0, st>>>((const float*)ws, rows, T, k, lab, pr);
        else
            row_topk_kernel<kMaxRowTexts / 32><<<grid, 256, 0, st>>>((const float*)ws, rows, T, k, lab, pr);
        SAF_CUDA_TRY(cudaGetLastError());
    }
    return 0;
}

int saf_query_text_presence(const float* feats, int64_t M, int32_t C, int64_t ldf, const float* text, int32_t n_background,
                            int32_t n_targets, int32_t norm_mode, int32_t precision, float* out, void* ws,
                            uint64_t ws_bytes, void* stream)
{
    int sms = 0;
    int rc = device_sm_count(&sms, nullptr);
    if (rc) return rc;
    if (!feats || !text || !out || !ws) return SAF_ERR_NULL;
    const int T = n_background + n_targets;
    if (M < 0 || C <= 0 || n_background < 0 || n_targets <= 0 || ldf < C) return SAF_ERR_SHAPE;
    if (ws_bytes < (uint64_t)kRowChunk * (uint64_t)T * sizeof(float) || ((uintptr_t)ws & 255u)) return SAF_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    SAF_CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)n_targets, st));
    for (int64_t m0 = 0; m0 < M; m0 += kRowChunk) {
        const int64_t rows = std::min<int64_t>(kRowChunk, M - m0);
        rc = saf_query_scores(feats + m0 * ldf, rows, C, ldf, text, T, norm_mode, SAF_SCORE_DOT, nullptr, precision,
                              (float*)ws, stream);
        if (rc) return rc;
        dim3 grid((unsigned)std::min<int64_t>((rows + 255) / 256, 64), (unsigned)std::min(n_targets, sms * 4));
        text_presence_kernel<<<grid, 256, 0, st>>>((const float*)ws, rows, n_background, n_targets, out);
        SAF_CUDA_TRY(cudaGetLastError());
    }
    return 0;
}

}  // extern "C"
