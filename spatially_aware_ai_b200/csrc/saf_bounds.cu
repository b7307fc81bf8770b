// saf_bounds.cu -- scene-bounds pre-pass: backproject_pcd (/root/reference/clipfusion.py:510-572).
//
// The reference builds the full H*W ray table of every frame (get_pix_vecs, clipfusion.py:496-507) and then keeps
// 7x7 samples of it; here one thread computes one sample: ray = K^-1 (u, v, 1), camera point = ray * depth,
// world point = R * p + t, valid = depth not NaN, > 0 and < max_depth.  The percentiles that turn the points into
// the grid origin / size (clip_seem_fusion.py:278-288) stay on the host (49 points per frame).
#include <cuda_runtime.h>

#include "saf_internal.cuh"

namespace saf {
namespace {

__global__ void __launch_bounds__(128) backproject_samples_kernel(const float* __restrict__ depth, const float* __restrict__ poses,
                                                                  const float* __restrict__ kinv, const int32_t* __restrict__ us,
                                                                  const int32_t* __restrict__ vs, int n_frames, int H, int W, int nu,
                                                                  int nv, float max_depth, float* __restrict__ xyz,
                                                                  uint8_t* __restrict__ valid)
{
    const int per = nu * nv;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_frames * per) return;
    const int f = (int)(i / per), s = (int)(i - (int64_t)f * per);
    // meshgrid(u, v, indexing="xy") flattened: sample s = (row j over v, column i over u)
    const int u = us[s % nu], v = vs[s / nu];
    const float d = __ldg(depth + ((size_t)f * H + v) * W + u);
    const float* Ki = kinv + (size_t)f * 9;
    const float* P = poses + (size_t)f * 16;
    const float fu = (float)u, fv = (float)v;
    float ray[3], pc[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) ray[k] = Ki[3 * k] * fu + Ki[3 * k + 1] * fv + Ki[3 * k + 2];
#pragma unroll
    for (int k = 0; k < 3; ++k) pc[k] = ray[k] * d;
#pragma unroll
    for (int k = 0; k < 3; ++k)
        xyz[i * 3 + k] = (P[4 * k] * pc[0] + P[4 * k + 1] * pc[1] + P[4 * k + 2] * pc[2]) + P[4 * k + 3];
    valid[i] = (d == d) && (d > 0.0f) && (d < max_depth);
}

}  // namespace
}  // namespace saf

using namespace saf;

extern "C" int saf_backproject_samples(const float* depth, const float* poses, const float* k_inverse, const int32_t* us,
                                       const int32_t* vs, int32_t n_frames, int32_t height, int32_t width, int32_t nu,
                                       int32_t nv, float max_depth, float* xyz_out, uint8_t* valid_out, void* stream)
{
    int sms = 0;
    int rc = device_sm_count(&sms, nullptr);
    if (rc) return rc;
    if (n_frames < 0 || height <= 0 || width <= 0 || nu <= 0 || nv <= 0) return SAF_ERR_SHAPE;
    if (n_frames == 0) return 0;
    if (!depth || !poses || !k_inverse || !us || !vs || !xyz_out || !valid_out) return SAF_ERR_NULL;
    const int64_t n = (int64_t)n_frames * nu * nv;
    backproject_samples_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        depth, poses, k_inverse, us, vs, n_frames, height, width, nu, nv, max_depth, xyz_out, valid_out);
    SAF_CHECK_LAUNCH("backproject_samples_kernel", (cudaStream_t)stream);
    return 0;
}
