// saf_abi.cu -- version and error strings of the C ABI.
#include <cuda_runtime.h>

#include "saf_b200.h"

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
