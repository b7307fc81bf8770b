"""CPU tests: the oracle's restatements of the query drivers, the sensor-format conversions and the tiled-patch
front end against golden vectors produced by the unmodified reference (tests/golden/make_golden.py), plus host
logic of the slab layouts."""
import hashlib

import numpy as np
import torch

from oracle import oracle as O
from tests import helpers as Hh


def _g(name):
    return Hh.load_golden(name)


def test_text_query_driver_matches_reference():
    """InSituManager.clip_text_query (clip_seem_fusion.py:507-533): normalise + nan_to_num, surgery, relevance."""
    g = _g("query_drivers")
    F = O.normalize_rows(g["feats"], "nan_to_num")
    sim = O.clip_feature_surgery(F[None], g["text"])[0]
    rel = O.relevance_minmax(sim[:, int(g["text_query_column"])])
    assert np.abs(rel - g["text_query_relevance"]).max() <= 2e-5
    assert np.abs(rel * np.float32(0.5) - g["text_query_alpha"]).max() <= 1e-5


def test_segment_matches_reference():
    """segment() (eval_scannet_segmentation.py:546-561): identical label order wherever the scores are apart."""
    g = _g("query_drivers")
    lab = O.segment_labels(g["segment_feats"], g["text"])
    ref = g["segment_labels"]
    assert np.array_equal(lab[:, :5], ref[:, :5])
    assert (lab == ref).mean() > 0.999


def test_query_mesh_expressions():
    g = _g("query_drivers")
    F = g["segment_feats"] / np.linalg.norm(g["segment_feats"], axis=-1, keepdims=True)
    rel = O.run_query(F, g["text"][:5])[:, -1]
    assert np.abs(O.relevance_half(rel) - g["query_mesh_half"]).max() <= 2e-5
    sim = O.minmax_per_text(O.clip_feature_surgery(F[None], g["text"])[0])
    assert np.abs(sim - g["query_mesh_minmax"]).max() <= 2e-5
    for n in range(sim.shape[1]):
        out = O.relevance_outliers(g["query_mesh_minmax"][:, n])
        assert np.array_equal(out != 0, g["query_mesh_outliers"][:, n] != 0)
        assert np.abs(out - g["query_mesh_outliers"][:, n]).max() <= 1e-6


def test_hypersim_presence():
    g = _g("query_drivers")
    pres = O.presence_scores(g["feats"], g["text"][:4], g["text"][4:])
    assert np.abs(pres - g["hypersim_presence"]).max() <= 2e-5
    thresholds = np.linspace(0, 1, 101, dtype=np.float32)
    agree = (pres[:, None] > thresholds[None]) == g["hypersim_preds"]
    assert agree.mean() > 0.99      # a presence value within rounding of a threshold may flip one column


def test_sensor_conversions_match_reference_dataset():
    """ScanNetDataset.__getitem__ (clipfusion.py:243-256): every uint16 / uint8 input value."""
    g = _g("sensor")
    assert np.array_equal(O.depth_from_mm(np.arange(65536, dtype=np.uint16)), g["depth_lut"])
    assert np.array_equal(O.rgb_from_u8(np.arange(256, dtype=np.uint8)), g["rgb_lut"])
    # the kernels' division-free evaluation of the same quotients, exhaustively
    assert O.check_sensor_conversions() == 0


def test_synth_frames_are_sensor_quantised():
    from spatially_aware_ai_b200 import synth
    cfg = synth.baseline_config("tiny")
    fr = synth.make_frame(cfg, 2)
    assert np.array_equal(fr["depth"], O.depth_from_mm(fr["depth_mm"]))
    assert np.array_equal(fr["rgb"], O.rgb_from_u8(fr["rgb_u8"]))
    assert (fr["depth_mm"] == 0).any() and fr["depth"].max() > 0.5


def test_tiled_front_end_matches_reference():
    """Clip.get_patches / img_inference_tiled (clipfusion.py:789-839) with the golden's stand-in encoder."""
    import spatially_aware_ai_b200 as saf
    g = _g("tiled")
    proj = torch.from_numpy(g["proj"])

    class _Backend:
        feature_dim = proj.shape[1]
        tokenizer = staticmethod(lambda s: s)

        @staticmethod
        def encode_image(x):
            return torch.nn.functional.adaptive_avg_pool2d(x, 4).reshape(len(x), -1) @ proj

    clip = saf.Clip(backend=_Backend())
    rgb = torch.from_numpy(g["rgb"])
    ps, st = int(g["patch_size"]), int(g["patch_stride"])
    patches = clip.get_patches(rgb, ps, st)
    assert list(patches.shape) == g["patches_shape"].tolist()
    assert hashlib.sha256(np.ascontiguousarray(patches.numpy()).tobytes()).hexdigest() == str(g["patches_sha256"])
    assert np.array_equal(patches[1, 0, 2].numpy(), g["patch_1_2"])
    feat = clip.img_inference_tiled(rgb, ps, st)
    assert np.array_equal(feat.numpy(), g["feat_img"])
    assert list(feat.stride()) == g["feat_img_strides"].tolist()    # [B,npy,npx,C] memory, permuted view


def test_default_prompt_ensemble_is_the_reference_default():
    from spatially_aware_ai_b200.prompt_templates import DEFAULT_PROMPT_TEMPLATES
    assert len(DEFAULT_PROMPT_TEMPLATES) == 85 and DEFAULT_PROMPT_TEMPLATES[0] == "a bad photo of a {}."
    assert DEFAULT_PROMPT_TEMPLATES[-1] == "this is one {} in the scene."


def test_cyclic_slab_layout_host_side():
    import spatially_aware_ai_b200 as saf
    from spatially_aware_ai_b200 import slab, synth
    nx, ny, nz = 44, 6, 5
    seen = []
    for r in range(3):
        kw = slab.cyclic_slab(nx, 3, r)
        vol = saf.ClipSeemFusion(torch.zeros(3), 0.05, torch.tensor([nx, ny, nz]), 0.1, False, 0, 0, synth.FakeClip(4),
                                 synth.FakeSeg(), **kw)
        xs = vol.global_x_planes()
        assert vol.tsdf.shape[0] == len(xs) * ny * nz
        assert all((x // 8) % 3 == r for x in xs)
        seen += xs
        rows = torch.tensor([0, ny * nz * 9 + 7, -1])
        glob = slab.local_to_global_rows(vol, rows)
        assert glob[0].item() == xs[0] * ny * nz and glob[1].item() == xs[9] * ny * nz + 7 and glob[2].item() == -1
        assert torch.allclose(vol.xyz_world[:, 0].reshape(len(xs), -1)[:, 0], torch.tensor(xs) * 0.05)
    assert sorted(seen) == list(range(nx))


def test_host_reach_test_is_conservative_and_useful():
    """saf_frame_reaches_slab (pose only, host): never 0 for a frame that updates a voxel of the slab (oracle), and
    0 for a good share of the frames that look away from it."""
    import ctypes
    from spatially_aware_ai_b200 import _lib, synth
    cfg = synth.SceneConfig(extent=(4.0, 2.0, 1.6), voxel_size=0.1, height=48, width=64, patch_size=32, patch_stride=16,
                            feature_dim=4, frames=40, seed=3)
    origin, nvox = cfg.grid()
    lib = _lib.load()
    rng = np.random.default_rng(0)
    dropped = 0
    for xb, xe in ((0, 8), (16, 24), (int(nvox[0]) - 8, int(nvox[0]))):
        g = _lib.GridDesc()
        g.origin[:] = origin.tolist()
        g.voxel_size = cfg.voxel_size
        g.nvox[:] = [int(v) for v in nvox]
        g.x_begin, g.x_end = xb, xe
        for i in range(cfg.frames):
            fr = synth.make_frame(cfg, i)
            if i % 3 == 1:
                fr["pose"] = synth.perturbed_pose(cfg, i, rng)
            orc = O.OracleVolume(origin, cfg.voxel_size, nvox, cfg.trunc, 4, x_begin=xb, x_end=xe, num_threads=0)
            cnt = orc.integrate(fr["depth"][None] * 0 + 50.0, fr["rgb"][None], fr["pose"][None], fr["K"][None],
                                fr["table"][None], fr["seg"][None], want_masks=False)   # far surface: all in-view voxels
            pose = np.ascontiguousarray(fr["pose"], np.float32).reshape(-1)
            K = np.ascontiguousarray(fr["K"], np.float32).reshape(-1)
            r = lib.saf_frame_reaches_slab(ctypes.byref(g), pose.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                           K.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), cfg.height, cfg.width)
            assert r in (0, 1)
            if cnt[0, 1] > 0:
                assert r == 1, (xb, i)
            dropped += (r == 0)
    assert dropped >= 10
    g.x_span, g.x_stride = 8, 16
    assert lib.saf_frame_reaches_slab(ctypes.byref(g), pose.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                      K.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), cfg.height, cfg.width) == 1


def test_sheared_block_column_layout_host_side():
    """Column (bx, by) on rank (bx + by) % n: the ranks' rows cover every voxel exactly once, a single x-plane and a
    single y-plane are both spread over all ranks, padding rows are marked."""
    import spatially_aware_ai_b200 as saf
    from spatially_aware_ai_b200 import slab, synth
    nx, ny, nz, world = 44, 30, 5, 3
    seen = torch.zeros(nx * ny * nz, dtype=torch.int64)
    for r in range(world):
        vol = saf.ClipSeemFusion(torch.zeros(3), 0.05, torch.tensor([nx, ny, nz]), 0.1, False, 0, 0, synth.FakeClip(4),
                                 synth.FakeSeg(), **slab.sheared_slab(world, r))
        rows = vol.global_rows()
        assert vol.ny_local == 8 * 2 and vol.tsdf.shape[0] == nx * vol.ny_local * nz == rows.numel()
        real = rows[rows >= 0]
        seen[real] += 1
        x, y = real // (ny * nz), real // nz % ny
        assert bool(((x // 8 + y // 8) % world == r).all())
        assert (x == 17).sum() > 0 and (y == 3).sum() > 0           # every rank holds part of any x- or y-plane
        local = torch.tensor([0, 5 * vol.ny_local * nz + 7, -1])
        out = slab.local_to_global_rows(vol, local)
        assert out[2].item() == -1 and out[0].item() == rows[0].item() and out[1].item() == rows[5 * vol.ny_local * nz + 7].item()
        xyz = vol.xyz_world
        ok = rows >= 0
        assert torch.allclose(xyz[ok][:, 1], (rows[ok] // nz % ny) * 0.05)
    assert bool((seen == 1).all())
