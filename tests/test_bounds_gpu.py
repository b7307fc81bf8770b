"""GPU: scene-bounds pre-pass (saf_backproject_samples) against the reference's backproject_pcd golden.
Floating point: 1e-5 absolute on the world points and the origin; the voxel counts must be identical."""
import numpy as np
import pytest
import torch

from spatially_aware_ai_b200 import bounds
from tests import helpers as Hh

pytestmark = pytest.mark.gpu


def test_backproject_and_bounds_match_reference_golden():
    g = Hh.load_golden("bounds")
    xyz, valid = bounds.backproject_samples(torch.from_numpy(g["depth"]).cuda(), torch.from_numpy(g["pose"]).cuda(),
                                            torch.from_numpy(g["K"]).cuda(), max_depth=float(g["max_depth"]))
    assert xyz.shape == g["xyz"].shape and valid.shape == (len(g["depth"]), 49)
    assert np.abs(xyz.cpu().numpy() - g["xyz"]).max() <= 1e-5
    origin, nvox, trunc_m = bounds.scene_bounds(xyz, float(g["voxel_size"]), int(g["trunc_vox"]))
    assert np.abs(origin.numpy() - g["origin"]).max() <= 1e-5
    assert np.array_equal(nvox.numpy(), g["nvox"]) and nvox.dtype == torch.int32
    assert abs(trunc_m - 0.15) < 1e-12


def test_backproject_all_invalid_and_empty():
    depth = torch.zeros((2, 10, 12), device="cuda")
    depth[1] = float("nan")
    pose = torch.eye(4, device="cuda").repeat(2, 1, 1)
    K = torch.tensor([[10.0, 0, 6], [0, 10, 5], [0, 0, 1]], device="cuda").repeat(2, 1, 1)
    xyz, valid = bounds.backproject_samples(depth, pose, K)
    assert xyz.shape == (0, 3) and not valid.any()
    xyz, valid = bounds.backproject_samples(depth[:0], pose[:0], K[:0])
    assert xyz.shape == (0, 3) and valid.shape == (0, 49)
