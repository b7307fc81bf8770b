// saf_query_tc.cu -- tcgen05 (5th-gen tensor core) kernels for the query GEMM.  Placeholder until
// the TMA + TMEM pipeline lands: reports "unsupported" so callers fail loudly instead of silently
// taking another path.
#include "saf_internal.cuh"

namespace saf {

int query_scores_tc(const float*, int64_t, int32_t, int64_t, const float*, int32_t, int32_t, int32_t, float*,
                    cudaStream_t)
{
    return SAF_ERR_UNSUPPORTED;
}

}  // namespace saf
