"""Scene-bounds pre-pass: ``backproject_pcd`` (/root/reference/clipfusion.py:510-572) and the percentile rule that
turns its sparse point cloud into the voxel grid's origin and size (/root/reference/clip_seem_fusion.py:268-288)."""
import ctypes

import numpy as np
import torch

from . import _lib


def sample_pixels(imwidth, imheight, uv_size=7):
    """clipfusion.py:517-519: round(linspace(0, size - 1, 7)) pixel columns / rows."""
    u = torch.round(torch.linspace(0, imwidth - 1, uv_size)).long()
    v = torch.round(torch.linspace(0, imheight - 1, uv_size)).long()
    return u, v


def backproject_samples(depth_imgs, poses, K, max_depth=float("inf"), uv_size=7):
    """depth_imgs [F,H,W], poses [F,4,4], K [F,3,3] (CUDA tensors) -> (xyz [n,3] world points of the valid samples,
    in frame order then sample order like the reference's concatenation, valid mask [F, uv_size**2])."""
    if not depth_imgs.is_cuda:
        raise RuntimeError("backproject_samples runs on a CUDA (sm_100) device only; there is no CPU path")
    dev = depth_imgs.device
    F, H, W = depth_imgs.shape
    depth = depth_imgs.to(torch.float32).contiguous()
    P = poses.to(device=dev, dtype=torch.float32).contiguous()
    Kinv = torch.linalg.inv(K.to(device=dev, dtype=torch.float32)).contiguous()     # clipfusion.py:503: K.inverse()
    u, v = sample_pixels(W, H, uv_size)
    us, vs = u.to(device=dev, dtype=torch.int32), v.to(device=dev, dtype=torch.int32)
    xyz = torch.empty((F, uv_size * uv_size, 3), dtype=torch.float32, device=dev)
    valid = torch.empty((F, uv_size * uv_size), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().saf_backproject_samples(depth.data_ptr(), P.data_ptr(), Kinv.data_ptr(), us.data_ptr(),
                                                   vs.data_ptr(), F, H, W, uv_size, uv_size, float(max_depth),
                                                   xyz.data_ptr(), valid.data_ptr(),
                                                   torch.cuda.current_stream(dev).cuda_stream), "saf_backproject_samples")
    valid = valid.bool()
    return xyz[valid], valid


def scene_bounds(xyz, voxel_size, trunc_vox):
    """clip_seem_fusion.py:276-288: 1st / 99th percentile of the points -/+ the truncation distance.
    Returns (origin float32 tensor [3], nvox int32 tensor [3], trunc_m)."""
    pts = xyz.detach().cpu().numpy()
    trunc_m = trunc_vox * voxel_size
    minbound = torch.tensor(np.percentile(pts, 1, axis=0)).float() - trunc_m
    maxbound = torch.tensor(np.percentile(pts, 99, axis=0)).float() + trunc_m
    nvox = ((maxbound - minbound) / voxel_size).round().int()
    return minbound, nvox, trunc_m
