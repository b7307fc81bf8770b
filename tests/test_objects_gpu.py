"""GPU: device connected-component labelling (saf_label_components) against the reference's flood_fill_3d
golden and the CPU oracle.  Integer work: bit-exact."""
import numpy as np
import pytest
import torch

import spatially_aware_ai_b200 as saf
from oracle import components as CC
from tests import helpers as Hh

pytestmark = pytest.mark.gpu


def test_label_objects_matches_reference_golden():
    g = Hh.load_golden("objects")
    ids, n = saf.label_objects(torch.from_numpy(g["class_grid"]).cuda())
    assert ids.dtype == torch.int32 and n == len(g["obj_ids"])
    assert np.array_equal(ids.cpu().numpy(), g["voxel_obj_ids"])
    names = [str(s) for s in g["class_names"]]
    know = saf.build_scene_knowledge(ids, g["class_grid"], names, [[0, 0, 0]] * len(names))
    uo = know["unique_objects"]
    assert list(uo.keys()) == [str(s) for s in g["obj_ids"]]
    assert [uo[k]["class_id"] for k in uo] == g["obj_class_id"].tolist()
    assert [uo[k]["object_index"] for k in uo] == g["obj_index"].tolist()
    assert [len(uo[k]["voxels"]) for k in uo] == g["obj_size"].tolist()
    assert know["object_counts"] == dict(zip([str(s) for s in g["count_keys"]], g["count_vals"].tolist()))


@pytest.mark.parametrize("shape,seed", [((40, 37, 29), 1), ((7, 130, 5), 2), ((1, 1, 9), 3), ((64, 64, 64), 4)])
def test_label_objects_matches_oracle(shape, seed):
    """Percolating noise (long winding components, many merges), thin grids, and every class boundary."""
    rng = np.random.default_rng(seed)
    grid = rng.choice(np.array([-1, 133, 0, 1, 2, 142]), size=shape, p=[0.3, 0.1, 0.2, 0.2, 0.1, 0.1]).astype(np.int64)
    ids, n = saf.label_objects(torch.from_numpy(grid).cuda())
    ref, n_ref = CC.label_objects(grid)
    assert n == n_ref
    assert np.array_equal(ids.cpu().numpy(), ref)


def test_label_objects_on_fused_volume():
    """End to end: fuse a few frames, argmax the histogram, label objects; objects are single-class and >= 3 voxels."""
    from spatially_aware_ai_b200 import synth
    cfg = synth.SceneConfig(extent=(2.2, 2.0, 1.6), voxel_size=0.05, height=96, width=128, patch_size=64,
                            patch_stride=32, feature_dim=8, frames=4, seed=9)
    origin, nvox = cfg.grid()
    g = dict(cls="ClipSeemFusion", feature_dim=8, origin=origin, nvox=nvox, voxel_size=cfg.voxel_size, trunc=cfg.trunc)
    vol, clip, seg = Hh.make_gpu_volume(g)
    for i in range(cfg.frames):
        fr = synth.make_frame(cfg, i * 11)
        clip.next_table = torch.from_numpy(fr["table"]).cuda()[None]
        seg.queue = [torch.from_numpy(fr["seg"]).cuda()]
        vol.integrate(torch.from_numpy(fr["depth"]).cuda()[None], torch.from_numpy(fr["rgb"]).cuda()[None],
                      torch.from_numpy(fr["pose"]).cuda()[None], torch.from_numpy(fr["K"]).cuda()[None])
    grid = vol.label_argmax().view(*[int(v) for v in nvox])
    ids, n = saf.label_objects(grid)
    ref, n_ref = CC.label_objects(grid.cpu().numpy())
    assert n == n_ref > 0 and np.array_equal(ids.cpu().numpy(), ref)
    sizes = np.bincount(-ids.cpu().numpy().reshape(-1)[ids.cpu().numpy().reshape(-1) < -1] - 2)
    assert sizes.min() >= 3
