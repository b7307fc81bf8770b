// saf_abi.cu -- version and error strings of the C ABI.
#include <cuda_runtime.h>

#include <stdio.h>
#include <stdlib.h>

#include "saf_internal.cuh"

extern "C" {

int saf_abi_version(void) { return SAF_ABI_VERSION; }

const char* saf_error_string(int code)
{
    if (code == 0) return "ok";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    switch (code) {
        case SAF_ERR_NULL: return "null pointer argument";
        case SAF_ERR_BATCH: return "batch size outside [1, SAF_MAX_BATCH] or larger than the workspace was sized for";
        case SAF_ERR_GRID: return "invalid voxel grid / slab description";
        case SAF_ERR_WORKSPACE: return "workspace too small or not initialised for this grid, batch or feature image";
        case SAF_ERR_SHAPE: return "invalid shape or size argument";
        case SAF_ERR_DTYPE: return "unsupported element type";
        case SAF_ERR_ALIGNMENT: return "pointer not aligned as required";
        case SAF_ERR_UNSUPPORTED: return "unsupported mode";
        case SAF_ERR_DEVICE: return "current CUDA device is not sm_100 (B200); this library has no other code path";
        default: return "unknown error";
    }
}
}

namespace saf {

int debug_check_launch(const char* what, cudaStream_t st)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        fprintf(stderr, "[saf] launch of %s failed: %s\n", what, cudaGetErrorString(e));
        return (int)e;
    }
    static int debug = -1;
    if (debug < 0) {
        const char* v = getenv("SAF_DEBUG_SYNC");
        debug = (v && v[0] == '1') ? 1 : 0;
    }
    if (debug) {
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            fprintf(stderr, "[saf] kernel %s faulted: %s\n", what, cudaGetErrorString(e));
            return (int)e;
        }
    }
    return 0;
}

}  // namespace saf
