"""Seeded synthetic RGB-D scenes in the shape the reference's datasets deliver.

Frames are quantised like sensor data: ``rgb = uint8 / 255`` and ``depth = uint16 millimetres / 1000``
(clipfusion.py:185-188), and both representations are returned (``rgb`` / ``rgb_u8``, ``depth`` / ``depth_mm``).

The fusion path consumes, per frame, exactly what the reference's dataset classes return
(/root/reference/clipfusion.py:86-494): ``rgb[H,W,3]`` float32 in [0,1], ``depth[H,W]`` float32
metres (0 = missing), ``pose[4,4]`` camera->world with right-down-forward axes
(clipfusion.py:308-312) and ``K[3,3]``; plus the two per-frame producer outputs the fusion
call pulls in itself: the tiled-patch CLIP feature image ``[C,npy,npx]``
(clipfusion.py:808-839) and the kMaX class-id map ``[H,W]`` (handy_utils.py:103-161).
There are no datasets or checkpoints on the build/bench boxes, so the scene is the inside of
an axis-aligned box room with analytic z-depth, and the producer outputs are random
(SURVEY.md section 8d).  numpy only: identical bytes on every machine for a given seed.
"""
from dataclasses import dataclass, field
import math

import numpy as np


@dataclass
class SceneConfig:
    extent: tuple = (4.0, 4.0, 3.0)   # room size in metres
    voxel_size: float = 0.04
    trunc_vox: int = 2
    height: int = 192
    width: int = 256
    feature_dim: int = 768
    patch_size: int = 128
    patch_stride: int = 64
    frames: int = 60
    missing_fraction: float = 0.02
    n_seg_classes: int = 134          # ids 0..133 (133 = null), handy_utils.py:103-161
    seg_block: int = 16
    seed: int = 0
    name: str = "custom"
    laps: int = 1

    @property
    def trunc(self):
        # clip_seem_fusion.py:278 : trunc_m = trunc_vox * voxel_size (python floats)
        return self.trunc_vox * self.voxel_size

    @property
    def npatches(self):
        # clipfusion.py:792-796
        assert (self.height - self.patch_size) % self.patch_stride == 0
        assert (self.width - self.patch_size) % self.patch_stride == 0
        npx = int(np.round(1 + (self.width - self.patch_size) / self.patch_stride))
        npy = int(np.round(1 + (self.height - self.patch_size) / self.patch_stride))
        return npy, npx

    def grid(self):
        """origin float32[3], nvox int64[3]: the room padded by trunc like clip_seem_fusion.py:278-288."""
        trunc = self.trunc
        minb = np.array([-trunc] * 3, dtype=np.float64)
        maxb = np.array(self.extent, dtype=np.float64) + trunc
        nvox = np.round((maxb - minb) / self.voxel_size).astype(np.int64)
        return minb.astype(np.float32), nvox

    @property
    def n_voxels(self):
        return int(np.prod(self.grid()[1]))


# BASELINE.json configs.
def baseline_config(which, **overrides):
    presets = {
        # cfg-1: 3D Scanner App-format scene, CPU-runnable
        "cfg1": dict(extent=(4.0, 4.0, 3.0), voxel_size=0.04, height=192, width=256, patch_size=128,
                     patch_stride=64, frames=60, name="cfg1"),
        # cfg-2: ScanNet-scale room, single B200
        "cfg2": dict(extent=(6.0, 6.0, 3.0), voxel_size=0.02, height=480, width=640, patch_size=160,
                     patch_stride=80, frames=1000, laps=4, name="cfg2"),
        # cfg-3: large multi-room scan, slab-sharded
        "cfg3": dict(extent=(8.0, 8.0, 3.0), voxel_size=0.02, height=480, width=640, patch_size=160,
                     patch_stride=80, frames=5000, laps=16, name="cfg3"),
        # cfg-5: Scene Manager v00 -> v01 incremental re-fusion: the cfg-1 extent, 60 + 500 frames at 640x480
        # (SURVEY.md 8d); swept over feature_dim and voxel_size by the caller
        "cfg5": dict(extent=(4.0, 4.0, 3.0), voxel_size=0.02, height=480, width=640, patch_size=160,
                     patch_stride=80, frames=560, laps=4, name="cfg5"),
        # small case for unit tests
        "tiny": dict(extent=(1.6, 1.4, 1.2), voxel_size=0.08, height=48, width=64, patch_size=32,
                     patch_stride=16, frames=6, feature_dim=16, seg_block=8, name="tiny"),
    }
    kw = dict(presets[which])
    kw.update(overrides)
    return SceneConfig(**kw)


def intrinsics(cfg):
    f = 0.9 * cfg.width
    return np.array([[f, 0.0, cfg.width / 2 - 0.5], [0.0, f, cfg.height / 2 - 0.5], [0.0, 0.0, 1.0]],
                    dtype=np.float32)


def camera_pose(cfg, i):
    """Orbit about the room centre looking outward and slightly down; camera axes right-down-forward."""
    ext = np.asarray(cfg.extent, dtype=np.float64)
    centre = ext / 2
    frames_per_lap = max(1, cfg.frames // max(1, cfg.laps))
    theta = 2 * math.pi * (i % frames_per_lap) / frames_per_lap
    lap = (i // frames_per_lap) / max(1, cfg.laps - 1) if cfg.laps > 1 else 0.0
    r_max = 0.35 * min(ext[0], ext[1])
    radius = 0.3 + (r_max - 0.3) * min(1.0, lap)
    pos = centre + np.array([radius * math.cos(theta), radius * math.sin(theta), 0.0])
    fwd = np.array([math.cos(theta), math.sin(theta), -0.2])
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    pose = np.eye(4, dtype=np.float64)
    pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3] = right, down, fwd, pos
    return pose.astype(np.float32)


def render_depth(cfg, pose, K):
    """Analytic z-depth of the box room [0,extent] seen from inside."""
    H, W = cfg.height, cfg.width
    u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    K = K.astype(np.float64)
    dirs = np.stack([(u - K[0, 2]) / K[0, 0], (v - K[1, 2]) / K[1, 1], np.ones_like(u)], axis=-1)
    R = pose[:3, :3].astype(np.float64)
    pos = pose[:3, 3].astype(np.float64)
    dw = dirs @ R.T
    ext = np.asarray(cfg.extent, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        t_hi = (ext - pos) / dw
        t_lo = (0.0 - pos) / dw
    t = np.where(dw > 0, t_hi, np.where(dw < 0, t_lo, np.inf))
    return t.min(axis=-1).astype(np.float32)


def make_frame(cfg, i, table_layout="chw"):
    """Frame i of the scene.  Returns a dict of numpy arrays.

    table_layout "chw": contiguous [C,npy,npx] (what a fake producer's randn returns);
    "hwc": the same values stored [npy,npx,C] and returned as the permuted [C,npy,npx] view,
    which is the memory layout the reference's own Clip.img_inference_tiled produces
    (clipfusion.py:835-839 .view(B,npy,npx,C).permute(0,3,1,2)).
    """
    rng = np.random.default_rng([cfg.seed, i])
    K = intrinsics(cfg)
    pose = camera_pose(cfg, i)
    depth_mm, depth = quantize_depth(render_depth(cfg, pose, K))
    if cfg.missing_fraction > 0:
        missing = rng.random(depth.shape, dtype=np.float32) < cfg.missing_fraction
        depth[missing] = 0.0
        depth_mm[missing] = 0
    rgb_u8 = rng.integers(0, 256, size=(cfg.height, cfg.width, 3), dtype=np.uint8)
    rgb = rgb_u8.astype(np.float32) / np.float32(255)
    sb = cfg.seg_block
    coarse = rng.integers(0, cfg.n_seg_classes, size=(-(-cfg.height // sb), -(-cfg.width // sb)), dtype=np.int64)
    seg = np.kron(coarse, np.ones((sb, sb), dtype=np.int64))[: cfg.height, : cfg.width].astype(np.uint8)
    npy, npx = cfg.npatches
    table = rng.standard_normal((cfg.feature_dim, npy, npx), dtype=np.float32)
    if table_layout == "hwc":
        table = np.ascontiguousarray(table.transpose(1, 2, 0)).transpose(2, 0, 1)
    return dict(depth=depth, rgb=rgb, seg=np.ascontiguousarray(seg), table=table, pose=pose, K=K, index=i,
                depth_mm=depth_mm, rgb_u8=rgb_u8)


def quantize_depth(depth_m):
    """Metres -> (uint16 millimetres, float32 metres) the way a depth sensor file delivers them: the fp32 depth
    is float32(mm) / 1000, exactly what the reference's dataset classes compute (clipfusion.py:187-188), so the
    sensor-format and the fp32 inputs of a synthetic frame carry identical values."""
    mm = np.clip(np.rint(np.asarray(depth_m, np.float64) * 1000.0), 0, 65535).astype(np.uint16)
    return mm, mm.astype(np.float32) / np.float32(1000)


def perturbed_pose(cfg, i, rng):
    """A tilted / rolled variant of the orbit pose, for parity cases with oblique views."""
    pose = camera_pose(cfg, i).astype(np.float64)
    ang = rng.uniform(-0.4, 0.4, size=3)
    cx, sx, cy, sy, cz, sz = math.cos(ang[0]), math.sin(ang[0]), math.cos(ang[1]), math.sin(ang[1]), \
        math.cos(ang[2]), math.sin(ang[2])
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    pose[:3, :3] = pose[:3, :3] @ (rz @ ry @ rx)
    pose[:3, 3] += rng.uniform(-0.2, 0.2, size=3)
    return pose.astype(np.float32)


class FakeClip:
    """Duck-typed stand-in for the reference's Clip image producer (clipfusion.py:808-839): returns a
    pre-generated tiled-patch feature image.  DNN inference is upstream of the fusion path; tests and benches
    inject this where the reference injects a real open_clip model."""

    def __init__(self, feature_dim):
        self.feature_dim = feature_dim
        self.next_table = None

    def img_inference_tiled(self, rgb_imgs, patch_size, patch_stride):
        return self.next_table


class FakeSeg:
    """Duck-typed kMaX stand-in (handy_utils.py:103-161): returns pre-generated class maps."""

    def __init__(self):
        self.queue = []

    def run_on_image(self, img):
        return self.queue.pop(0)
