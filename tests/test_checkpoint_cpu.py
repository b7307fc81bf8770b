"""CPU: grid files / full-state checkpoint round trip (host I/O only; no kernels involved)."""
import os

import numpy as np
import pytest
import torch

import spatially_aware_ai_b200 as saf
from tests.helpers import FakeClip, FakeSeg


def _volume(nvox=(6, 5, 4), C=8, cls="seem", **kw):
    if cls == "seem":
        return saf.ClipSeemFusion(torch.tensor([0.1, 0.2, 0.3]), 0.05, torch.tensor(nvox), 0.1, False, 0, 0, FakeClip(C),
                                  FakeSeg(), **kw)
    return saf.ClipFusion(torch.tensor([0.1, 0.2, 0.3]), 0.05, torch.tensor(nvox), 0.1, False, FakeClip(C), None, 0, 0, **kw)


def _randomise(vol, seed):
    g = torch.Generator().manual_seed(seed)
    vol.tsdf.copy_(torch.rand(vol.tsdf.shape, generator=g) * 2 - 1)
    vol.rgb.copy_(torch.rand(vol.rgb.shape, generator=g))
    vol.clip_feat.copy_(torch.randn(vol.clip_feat.shape, generator=g))
    vol.weight.copy_(torch.randint(0, 9, vol.weight.shape, generator=g, dtype=torch.int32))
    vol.tsdf_weight.copy_(torch.randint(0, 9, vol.tsdf_weight.shape, generator=g, dtype=torch.int32))
    if hasattr(vol, "labels_one_hot"):
        vol.labels_one_hot.copy_(torch.randint(0, 3, vol.labels_one_hot.shape, generator=g, dtype=torch.int32))


def test_round_trip_and_reference_file_formats(tmp_path, monkeypatch):
    from spatially_aware_ai_b200 import checkpoint
    monkeypatch.setattr(checkpoint, "_CHUNK_BYTES", 1000)      # force many chunks
    vol = _volume()
    _randomise(vol, 1)
    meta = saf.save_state(vol, str(tmp_path))
    assert set(meta["files"]) == set(checkpoint._STATE_FILES)
    # the two files the reference itself writes / reloads (clip_seem_fusion.py:563-571, 204-221)
    rgb = np.load(os.path.join(tmp_path, "voxel_rgb.npy"))
    feats = np.load(os.path.join(tmp_path, "voxel_clip_feats.npy"))
    assert rgb.shape == (6, 5, 4, 3) and feats.shape == (6, 5, 4, 8) and rgb.dtype == np.float32
    assert np.array_equal(rgb, vol.rgb.view(6, 5, 4, -1).numpy())
    assert np.array_equal(feats, vol.clip_feat.view(6, 5, 4, -1).numpy())
    other = _volume()
    saf.load_state(other, str(tmp_path))
    for name in ("tsdf", "rgb", "clip_feat", "weight", "tsdf_weight", "labels_one_hot"):
        assert torch.equal(getattr(other, name), getattr(vol, name)), name


def test_slab_and_class_mismatch(tmp_path):
    vol = _volume(x_begin=2, x_end=5)
    _randomise(vol, 2)
    saf.save_state(vol, str(tmp_path))
    assert np.load(os.path.join(tmp_path, "voxel_tsdf.npy")).shape == (3, 5, 4)
    with pytest.raises(ValueError, match="x_begin"):
        saf.load_state(_volume(), str(tmp_path))
    with pytest.raises(ValueError, match="feature_dim"):
        saf.load_state(_volume(C=4, x_begin=2, x_end=5), str(tmp_path))
    plain = _volume(cls="fusion", x_begin=2, x_end=5)        # no label histogram: the file is skipped
    saf.load_state(plain, str(tmp_path))
    assert torch.equal(plain.clip_feat, vol.clip_feat) and torch.equal(plain.tsdf_weight, vol.tsdf_weight)


def test_mesh_ply_and_scene_knowledge_round_trip(tmp_path):
    from spatially_aware_ai_b200 import checkpoint
    rng = np.random.default_rng(4)
    verts = rng.standard_normal((50, 3)).astype(np.float32)
    faces = rng.integers(0, 50, (80, 3)).astype(np.int64)
    colors = rng.random((50, 3)).astype(np.float32)
    path = str(tmp_path / "mesh_rgb.ply")
    checkpoint.save_mesh_ply(path, verts, faces, torch.from_numpy(colors))
    head = open(path, "rb").read(400).decode("ascii", "ignore")
    assert head.startswith("ply\nformat binary_little_endian 1.0") and "element vertex 50" in head and "element face 80" in head
    v, f, c = checkpoint.load_mesh_ply(path)
    assert np.array_equal(v, verts) and np.array_equal(f, faces)
    assert c.shape == (50, 4) and (c[:, 3] == 255).all()
    assert np.array_equal(c[:, :3], np.rint(colors * 255).astype(np.uint8))
    # how the reference reloads the colours (clip_seem_fusion.py:228-231): uint8 / 255
    assert np.abs(c[:, :3] / 255.0 - colors).max() <= 0.5 / 255 + 1e-7
    checkpoint.save_mesh_ply(path, verts, faces)                       # no colours
    v, f, c = checkpoint.load_mesh_ply(path)
    assert np.array_equal(v, verts) and np.array_equal(f, faces) and c is None
    checkpoint.save_mesh_ply(path, verts[:0], faces[:0], colors[:0])   # empty mesh
    v, f, c = checkpoint.load_mesh_ply(path)
    assert v.shape == (0, 3) and f.shape == (0, 3)
    import json
    sk = {"unique_objects": {"chair:1": {"class_id": 56, "voxels": [(1, 2, 3)], "color": np.int64(7)}}, "scan_version": "v01"}
    checkpoint.save_scene_knowledge(str(tmp_path / "scene_knowledge.json"), sk)
    back = json.load(open(tmp_path / "scene_knowledge.json"))
    assert back["unique_objects"]["chair:1"]["voxels"] == [[1, 2, 3]] and back["scan_version"] == "v01"
