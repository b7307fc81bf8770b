"""CPU oracle of the scene-bounds pre-pass (backproject_pcd, /root/reference/clipfusion.py:510-572).
TEST INFRASTRUCTURE ONLY.  Parity status: pinned - tests/golden/bounds.npz holds the output of the unmodified
reference backproject_pcd + the bounds rule of clip_seem_fusion.py:276-288 on synthetic frames."""
import numpy as np


def backproject_samples(depth, poses, K, max_depth=np.inf, uv_size=7):
    F, H, W = depth.shape
    u = np.rint(np.linspace(0, W - 1, uv_size)).astype(np.int64)
    v = np.rint(np.linspace(0, H - 1, uv_size)).astype(np.int64)
    uu, vv = np.meshgrid(u, v, indexing="xy")
    uv = np.stack((uu, vv), -1).reshape(-1, 2)
    pts, masks = [], []
    for f in range(F):
        Kinv = np.linalg.inv(K[f].astype(np.float64))
        rays = (Kinv @ np.stack((uv[:, 0], uv[:, 1], np.ones(len(uv))), 0)).T
        d = depth[f, uv[:, 1], uv[:, 0]].astype(np.float64)
        with np.errstate(invalid="ignore"):
            ok = ~np.isnan(d) & (d > 0) & (d < max_depth)
        world = (poses[f, :3, :3].astype(np.float64) @ (rays * d[:, None]).T).T + poses[f, :3, 3]
        pts.append(world[ok])
        masks.append(ok)
    return np.concatenate(pts).astype(np.float32), np.stack(masks)


def scene_bounds(xyz, voxel_size, trunc_vox):
    trunc_m = trunc_vox * voxel_size
    lo = np.percentile(xyz, 1, axis=0).astype(np.float32) - np.float32(trunc_m)
    hi = np.percentile(xyz, 99, axis=0).astype(np.float32) + np.float32(trunc_m)
    return lo, np.rint((hi - lo) / np.float32(voxel_size)).astype(np.int32)
