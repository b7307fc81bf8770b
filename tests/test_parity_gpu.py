"""GPU: the CUDA path (through the drop-in classes and the C ABI) against the reference's golden
vectors and the CPU oracle.

Tolerances (BASELINE.json north_star): voxel indices / valid / tsdf_valid / label histogram /
weights bit-exact; tsdf <= 1e-5 abs; features cosine >= 0.9999 per touched voxel (rgb <= 1e-5).
For single-frame calls the kernels reproduce the reference's fp32 roundings, so those cases are
additionally required to be bit-identical to the oracle.
"""
import ctypes

import numpy as np
import pytest
import torch

from oracle import oracle as O
from spatially_aware_ai_b200 import _lib, synth
from tests import helpers as Hh

pytestmark = pytest.mark.gpu

TSDF_TOL = 1e-5
RGB_TOL = 1e-5
COS_MIN = 0.9999


def _np(t):
    return t.detach().cpu().numpy()


def _check_against(vol, ref_tsdf, ref_w, ref_tw, ref_rgb, ref_feat, ref_labels, exact):
    assert np.array_equal(_np(vol.weight), ref_w)
    assert np.array_equal(_np(vol.tsdf_weight), ref_tw)
    if ref_labels is not None:
        assert np.array_equal(_np(vol.labels_one_hot), ref_labels)
    assert np.abs(_np(vol.tsdf) - ref_tsdf).max() <= TSDF_TOL
    assert np.abs(_np(vol.rgb) - ref_rgb).max() <= RGB_TOL
    touched = ref_w > 0
    assert Hh.cosine_rows(_np(vol.clip_feat)[touched], ref_feat[touched]).min() >= COS_MIN
    assert not _np(vol.clip_feat)[~touched].any()
    if exact:
        assert np.array_equal(_np(vol.tsdf), ref_tsdf)
        assert np.array_equal(_np(vol.rgb), ref_rgb)
        assert np.array_equal(_np(vol.clip_feat), ref_feat)


@pytest.mark.parametrize("name", Hh.FUSION_GOLDENS)
def test_fusion_matches_reference_golden(name):
    g = Hh.load_golden(name)
    vol, counts = Hh.replay_gpu(g)
    assert np.array_equal(counts, g["counts"])
    labels = Hh.golden_labels(g) if g["cls"] == "ClipSeemFusion" else None
    _check_against(vol, g["tsdf"], g["weight"], g["tsdf_weight"], g["rgb_state"], g["clip_feat"], labels,
                   exact=(g["batch"] == 1))


@pytest.mark.parametrize("name", ["seem_a", "seem_edge", "fusion_b2"])
def test_fusion_matches_oracle_bitwise(name):
    g = Hh.load_golden(name)
    vol, counts = Hh.replay_gpu(g)
    orc, ocounts = Hh.replay_oracle(g)
    assert np.array_equal(counts, ocounts)
    _check_against(vol, orc.tsdf, orc.weight, orc.tsdf_weight, orc.rgb, orc.clip_feat, orc.labels_one_hot, exact=True)


def test_masks_bit_exact_through_c_abi():
    """valid / tsdf_valid masks of K2 against the oracle's, voxel by voxel."""
    g = Hh.load_golden("seem_edge")
    lib = _lib.load()
    dev = "cuda"
    vol, clip, seg = Hh.make_gpu_volume(g, dev)
    orc = O.OracleVolume(g["origin"], g["voxel_size"], g["nvox"], g["trunc"], g["feature_dim"])
    n = vol.tsdf.numel()
    H, W = g["depth"].shape[1:]
    stream = torch.cuda.current_stream().cuda_stream
    for i in range(len(g["depth"])):
        table = torch.from_numpy(g["table"][i:i + 1]).to(dev)
        frames, keep, (B, _, _, telems) = vol._make_frames(
            torch.from_numpy(g["depth"][i:i + 1]).to(dev), torch.from_numpy(g["rgb"][i:i + 1]).to(dev),
            torch.from_numpy(g["pose"][i:i + 1]), torch.from_numpy(g["K"][i:i + 1]), table,
            [torch.from_numpy(g["seg"][i].astype(np.int64)).to(dev)])
        ws = vol._workspace(1, telems)
        valid = torch.zeros(n, dtype=torch.uint8, device=dev)
        tvalid = torch.zeros(n, dtype=torch.uint8, device=dev)
        gd, vd = vol._grid_desc(), vol._volume_desc()
        _lib.check(lib.saf_frustum_cull(ctypes.byref(gd), frames, 1, H, W, g["trunc"], ctypes.byref(ws), stream), "cull")
        _lib.check(lib.saf_tsdf_update(ctypes.byref(gd), ctypes.byref(vd), frames, 1, H, W, g["trunc"], ctypes.byref(ws),
                                       valid.data_ptr(), tvalid.data_ptr(), stream), "tsdf")
        _lib.check(lib.saf_feature_accumulate(ctypes.byref(gd), ctypes.byref(vd), frames, 1, 0, H, W,
                                              _lib.SAF_RGB_BILINEAR, ctypes.byref(ws), stream), "feat")
        orc.integrate(g["depth"][i:i + 1], g["rgb"][i:i + 1], g["pose"][i:i + 1], g["K"][i:i + 1], g["table"][i:i + 1],
                      g["seg"][i:i + 1])
        assert np.array_equal(_np(valid).astype(bool), orc.last_valid[0]), "valid mask, frame %d" % i
        assert np.array_equal(_np(tvalid).astype(bool), orc.last_tsdf_valid[0]), "tsdf_valid mask, frame %d" % i
        st = vol.stats()
        # the conservative block cull never hides a voxel the oracle touches, and does cull something
        assert st["last_valid"][0] == orc.last_counts[0, 0] and st["last_tsdf_valid"][0] == orc.last_counts[0, 1]
    assert np.array_equal(_np(vol.tsdf), orc.tsdf)
    assert np.array_equal(_np(vol.labels_one_hot), orc.labels_one_hot)


def test_slabs_concatenate_to_full_grid():
    g = Hh.load_golden("seem_a")
    nx = int(g["nvox"][0])
    cuts = [0, 13, 30, nx]
    parts = [Hh.replay_gpu(g, x_begin=a, x_end=b)[0] for a, b in zip(cuts[:-1], cuts[1:])]
    cat = lambda name: np.concatenate([_np(getattr(p, name)) for p in parts])  # noqa: E731
    assert np.array_equal(cat("tsdf"), g["tsdf"])
    assert np.array_equal(cat("weight"), g["weight"])
    assert np.array_equal(cat("tsdf_weight"), g["tsdf_weight"])
    assert np.array_equal(cat("clip_feat"), g["clip_feat"])
    assert np.array_equal(cat("rgb"), g["rgb_state"])
    assert np.array_equal(cat("labels_one_hot"), Hh.golden_labels(g))


@pytest.mark.parametrize("seg_dtype", [torch.uint8, torch.int16, torch.int32, torch.float32])
def test_class_map_dtypes_and_host_poses(seg_dtype):
    g = Hh.load_golden("seem_a")
    vol, _ = Hh.replay_gpu(g, seg_dtype=seg_dtype, pose_on_device=False, upto=4)
    ref, _ = Hh.replay_oracle(g, upto=4)
    assert np.array_equal(_np(vol.labels_one_hot), ref.labels_one_hot)
    assert np.array_equal(_np(vol.tsdf), ref.tsdf)
    assert np.array_equal(_np(vol.clip_feat), ref.clip_feat)


def test_bad_class_id_is_reported():
    g = Hh.load_golden("seem_a")
    vol, clip, seg = Hh.make_gpu_volume(g)
    clip.next_table = torch.from_numpy(g["table"][:1]).cuda()
    seg.queue = [torch.full(g["seg"][0].shape, 200, dtype=torch.int64, device="cuda")]
    vol.integrate(torch.from_numpy(g["depth"][:1]).cuda(), torch.from_numpy(g["rgb"][:1]).cuda(),
                  torch.from_numpy(g["pose"][:1]).cuda(), torch.from_numpy(g["K"][:1]).cuda())
    with pytest.raises(RuntimeError, match="num_classes"):
        vol.check_errors()


def test_label_argmax_matches_reference_rule():
    g = Hh.load_golden("seem_a")
    vol, _ = Hh.replay_gpu(g)
    lab = Hh.golden_labels(g)
    ref = np.where(lab.any(axis=1), lab.argmax(axis=1), -1)
    assert np.array_equal(_np(vol.label_argmax()), ref)


@pytest.mark.parametrize("C", [512, 768, 1024, 20, 6, 7])
def test_feature_dims(C):
    """The specialised (C = 512/768/1024, TMA-staged table) and generic feature kernels against the oracle."""
    cfg = synth.SceneConfig(extent=(1.6, 1.4, 1.2), voxel_size=0.05, height=48, width=64, patch_size=32,
                            patch_stride=16, feature_dim=C, frames=3, seed=21 + C)
    origin, nvox = cfg.grid()
    g = dict(cls="ClipSeemFusion", feature_dim=C, origin=origin, nvox=nvox, voxel_size=cfg.voxel_size, trunc=cfg.trunc)
    vol, clip, seg = Hh.make_gpu_volume(g)
    orc = O.OracleVolume(origin, cfg.voxel_size, nvox, cfg.trunc, C, num_threads=4)
    for i in range(cfg.frames):
        fr = synth.make_frame(cfg, i, table_layout="hwc" if i != 1 else "chw")
        t = torch.from_numpy(np.ascontiguousarray(fr["table"].transpose(1, 2, 0))).cuda().permute(2, 0, 1)[None] \
            if i != 1 else torch.from_numpy(fr["table"]).cuda()[None]
        clip.next_table = t
        seg.queue = [torch.from_numpy(fr["seg"]).cuda()]
        vol.integrate(torch.from_numpy(fr["depth"]).cuda()[None], torch.from_numpy(fr["rgb"]).cuda()[None],
                      torch.from_numpy(fr["pose"]).cuda()[None], torch.from_numpy(fr["K"]).cuda()[None])
        orc.integrate(fr["depth"][None], fr["rgb"][None], fr["pose"][None], fr["K"][None], fr["table"][None],
                      fr["seg"][None], want_masks=False)
    _check_against(vol, orc.tsdf, orc.weight, orc.tsdf_weight, orc.rgb, orc.clip_feat, orc.labels_one_hot, exact=True)


def test_cfg1_full_parity_run():
    """BASELINE config 1 (60 frames, 256x192, 4 cm voxels over 4x4x3 m, 768-d) end to end vs the oracle."""
    cfg = synth.baseline_config("cfg1")
    origin, nvox = cfg.grid()
    g = dict(cls="ClipSeemFusion", feature_dim=cfg.feature_dim, origin=origin, nvox=nvox,
             voxel_size=cfg.voxel_size, trunc=cfg.trunc)
    vol, clip, seg = Hh.make_gpu_volume(g)
    orc = O.OracleVolume(origin, cfg.voxel_size, nvox, cfg.trunc, cfg.feature_dim, num_threads=0)
    n_valid = 0
    for i in range(cfg.frames):
        fr = synth.make_frame(cfg, i, table_layout="hwc")
        clip.next_table = torch.from_numpy(np.ascontiguousarray(fr["table"].transpose(1, 2, 0))).cuda() \
            .permute(2, 0, 1)[None]
        seg.queue = [torch.from_numpy(fr["seg"]).cuda()]
        vol.integrate(torch.from_numpy(fr["depth"]).cuda()[None], torch.from_numpy(fr["rgb"]).cuda()[None],
                      torch.from_numpy(fr["pose"]).cuda()[None], torch.from_numpy(fr["K"]).cuda()[None])
        n_valid += int(orc.integrate(fr["depth"][None], fr["rgb"][None], fr["pose"][None], fr["K"][None],
                                     fr["table"][None], fr["seg"][None], want_masks=False)[0, 0])
    st = vol.stats()
    assert st["total_frames"] == cfg.frames and st["total_valid"] == n_valid
    _check_against(vol, orc.tsdf, orc.weight, orc.tsdf_weight, orc.rgb, orc.clip_feat, orc.labels_one_hot, exact=True)


@pytest.mark.parametrize("mode,n_frames,step", [("frame", 6, 37), ("window", 20, 3)])
def test_size_independent_properties_at_cfg2_scale(mode, n_frames, step):
    """At a ScanNet-scale grid (2 cm, 640x480) the oracle is too slow for a full comparison; check
    invariants instead: counter sums equal the kernels' own valid counts, every histogram row sums
    to the voxel's weight, untouched voxels stay zero, and a sampled x-slab matches the oracle.
    "frame": one integrate() per frame; "window": one integrate_sequence() call (two windows of 10)."""
    cfg = synth.baseline_config("cfg2", feature_dim=512, frames=n_frames, extent=(6.0, 6.0, 3.0))
    origin, nvox = cfg.grid()
    g = dict(cls="ClipSeemFusion", feature_dim=cfg.feature_dim, origin=origin, nvox=nvox,
             voxel_size=cfg.voxel_size, trunc=cfg.trunc)
    vol, clip, seg = Hh.make_gpu_volume(g)
    xs0, xs1 = 150, 162
    orc = O.OracleVolume(origin, cfg.voxel_size, nvox, cfg.trunc, cfg.feature_dim, x_begin=xs0, x_end=xs1, num_threads=0)
    frames = [synth.make_frame(cfg, i * step, table_layout="hwc") for i in range(cfg.frames)]
    for fr in frames:
        orc.integrate(fr["depth"][None], fr["rgb"][None], fr["pose"][None], fr["K"][None], fr["table"][None],
                      fr["seg"][None], want_masks=False)
    if mode == "frame":
        for fr in frames:
            clip.next_table = torch.from_numpy(np.ascontiguousarray(fr["table"].transpose(1, 2, 0))).cuda() \
                .permute(2, 0, 1)[None]
            seg.queue = [torch.from_numpy(fr["seg"]).cuda()]
            vol.integrate(torch.from_numpy(fr["depth"]).cuda()[None], torch.from_numpy(fr["rgb"]).cuda()[None],
                          torch.from_numpy(fr["pose"]).cuda()[None], torch.from_numpy(fr["K"]).cuda()[None])
    else:
        clip.next_table = torch.stack([torch.from_numpy(np.ascontiguousarray(f["table"].transpose(1, 2, 0)))
                                       for f in frames]).cuda().permute(0, 3, 1, 2)
        seg.queue = [torch.from_numpy(f["seg"]).cuda() for f in frames]
        vol.integrate_sequence(torch.stack([torch.from_numpy(f["depth"]) for f in frames]).cuda(),
                               torch.stack([torch.from_numpy(f["rgb"]) for f in frames]).cuda(),
                               torch.stack([torch.from_numpy(f["pose"]) for f in frames]),
                               torch.stack([torch.from_numpy(f["K"]) for f in frames]))
    st = vol.stats()
    assert st["total_frames"] == cfg.frames
    assert int(vol.weight.sum(dtype=torch.int64)) == st["total_valid"] > 0
    assert int(vol.tsdf_weight.sum(dtype=torch.int64)) == st["total_tsdf_valid"] >= st["total_valid"]
    assert torch.equal(vol.labels_one_hot.sum(dim=1, dtype=torch.int32), vol.weight)
    untouched = vol.weight == 0
    row_abs = torch.zeros_like(vol.tsdf)
    for c0 in range(0, cfg.feature_dim, 64):      # chunked: avoid a second 29 GB temporary
        row_abs += vol.clip_feat[:, c0:c0 + 64].abs().sum(dim=1)
    assert not row_abs[untouched].any() and not vol.rgb.abs().sum(dim=1)[untouched].any()
    assert (row_abs[~untouched] > 0).all()
    assert (vol.tsdf.abs() <= 1).all()
    ny, nz = int(nvox[1]), int(nvox[2])
    sl = slice(xs0 * ny * nz, xs1 * ny * nz)
    assert np.array_equal(_np(vol.weight[sl]), orc.weight)
    assert np.array_equal(_np(vol.tsdf_weight[sl]), orc.tsdf_weight)
    assert np.array_equal(_np(vol.tsdf[sl]), orc.tsdf)
    assert np.array_equal(_np(vol.clip_feat[sl]), orc.clip_feat)
    assert np.array_equal(_np(vol.labels_one_hot[sl]), orc.labels_one_hot)


# ---- query -------------------------------------------------------------------------------------

def test_query_scores_match_reference_golden():
    import spatially_aware_ai_b200 as saf
    g = Hh.load_golden("query")
    F, X = torch.from_numpy(g["F"]).cuda(), torch.from_numpy(g["X"]).cuda()
    Fraw = torch.from_numpy(g["F_raw"]).cuda()

    class Backend:
        feature_dim = X.shape[1]

        def tokenizer(self, labels):
            return labels

        def encode_text(self, tokens):
            return X

    clip = saf.Clip(backend=Backend())
    rel = clip.run_query(F, ["x"] * X.shape[0])
    assert np.allclose(_np(rel), g["relevance"], rtol=1e-4, atol=1e-6)
    sim = saf.Clip.clip_feature_surgery(F[None], X)
    assert np.allclose(_np(sim), g["surgery"], atol=2e-6)
    sim_red = saf.Clip.clip_feature_surgery(F[None], X, redundant_feats=torch.from_numpy(g["redundant"]).cuda())
    assert np.allclose(_np(sim_red), g["surgery_red"], atol=2e-6)
    F0 = F.clone()
    F0[0] = 0
    assert np.allclose(_np(saf.Clip.clip_feature_surgery(F0[None], X)), g["surgery_row0_zero"], atol=2e-6)
    # normalisation fused into the kernel == the callers' explicit normalisation
    fused = saf.query_scores(Fraw, X, norm="nan_to_num", mode="dot")
    assert np.allclose(_np(fused), g["F"] @ g["X"].T, atol=2e-6)
    # top-k identical to ranking the reference's own scores
    k = 5
    w = saf.surgery_weights(F[0], X)
    ts, ti = saf.query_topk(F, X, k, mode="surgery", surgery_w=w)
    assert np.array_equal(_np(ti), O.topk_indices(g["surgery"][0], k))


@pytest.mark.parametrize("M,C,T,k", [(5000, 768, 33, 7), (70000, 512, 8, 100), (3, 64, 2, 1), (1000, 20, 5, 1000)])
def test_query_topk_matches_oracle(M, C, T, k):
    import spatially_aware_ai_b200 as saf
    rng = np.random.default_rng(M + C)
    F = rng.standard_normal((M, C)).astype(np.float32)
    F[rng.integers(0, M, size=max(1, M // 50))] = 0
    if M > 10:
        F[7] = F[3]          # exact tie between two rows
    X = rng.standard_normal((T, C)).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    Fd, Xd = torch.from_numpy(F).cuda(), torch.from_numpy(X).cuda()
    scores = saf.query_scores(Fd, Xd, norm="nan_to_num", mode="dot")
    ref = O.normalize_rows(F) @ X.T
    assert np.allclose(_np(scores), ref, atol=3e-6)
    kk = min(k, M)
    ts, ti = saf.query_topk(Fd, Xd, k, norm="nan_to_num", mode="dot")
    # compare against ranking the kernel's own score matrix (exactly the same numbers)
    want = O.topk_indices(_np(scores), kk)
    assert np.array_equal(_np(ti)[:, :kk], want)
    assert (_np(ti)[:, kk:] == -1).all()
    got_scores = np.take_along_axis(_np(scores).T, want, axis=1)
    assert np.array_equal(_np(ts)[:, :kk], got_scores)


@pytest.mark.parametrize("M,C,T", [(1000, 768, 33), (300, 64, 256), (4097, 512, 300), (129, 100, 5), (70000, 768, 64)])
def test_query_tensor_core_scores(M, C, T):
    """tcgen05 kind::tf32 GEMM (precision='tf32') against the fp32 kernel and the numpy oracle.
    tf32 keeps 10 mantissa bits of each operand: |error| <= 2^-9 * |f| * |x| per score."""
    import spatially_aware_ai_b200 as saf
    rng = np.random.default_rng(M * 7 + T)
    F = rng.standard_normal((M, C)).astype(np.float32)
    F[::17] = 0
    X = rng.standard_normal((T, C)).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    Fd, Xd = torch.from_numpy(F).cuda(), torch.from_numpy(X).cuda()
    ref = O.normalize_rows(F) @ X.T
    got = _np(saf.query_scores(Fd, Xd, norm="nan_to_num", mode="dot", precision="tf32"))
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 2.0 ** -9 + 1e-5
    assert not got[::17].any()
    # un-normalised dot product and the softmax epilogue on top of the tensor-core scores
    raw = _np(saf.query_scores(Fd, Xd, norm=None, mode="dot", precision="tf32"))
    bound = 2.0 ** -9 * np.linalg.norm(F, axis=1, keepdims=True) + 1e-4
    assert (np.abs(raw - F @ X.T) <= bound).all()
    sm = _np(saf.query_scores(Fd, Xd, norm="nan_to_num", mode="softmax100", precision="tf32"))
    assert np.allclose(sm.sum(axis=1), 1.0, atol=1e-4)


@pytest.mark.parametrize("M,C,T,k", [(50000, 768, 40, 10), (300000, 512, 256, 100), (1000, 64, 3, 5),
                                     (20000, 768, 8, 2000)])
def test_query_topk_tensor_core_is_exact(M, C, T, k):
    """Fused tensor-core top-k (tf32 GEMM + candidate filter + fp32 rescoring) returns exactly the
    ranking of the fp32 score matrix: same indices, same score bits."""
    import spatially_aware_ai_b200 as saf
    rng = np.random.default_rng(M + T)
    F = rng.standard_normal((M, C)).astype(np.float32)
    F[rng.integers(0, M, size=M // 20)] = 0
    F[11] = F[5]                                  # an exact tie
    X = rng.standard_normal((T, C)).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    Fd, Xd = torch.from_numpy(F).cuda(), torch.from_numpy(X).cuda()
    scores = _np(saf.query_scores(Fd, Xd, norm="nan_to_num", mode="dot", precision="fp32"))
    ts, ti = saf.query_topk(Fd, Xd, k, norm="nan_to_num", mode="dot", precision="tf32", index_base=7)
    kk = min(k, M)
    want = O.topk_indices(scores, kk)
    assert np.array_equal(_np(ti)[:, :kk], want + 7)
    assert np.array_equal(_np(ts)[:, :kk], np.take_along_axis(scores.T, want, axis=1))


def test_query_topk_tensor_core_falls_back_on_mass_ties():
    """All-zero features: every row ties at score 0, the candidate buckets overflow and the call must
    fall back to the exact chunked path (lowest indices win)."""
    import spatially_aware_ai_b200 as saf
    M, C, T, k = 40000, 64, 4, 6
    Fd = torch.zeros((M, C), device="cuda")
    Xd = torch.randn((T, C), device="cuda")
    ts, ti = saf.query_topk(Fd, Xd, k, norm="nan_to_num", mode="dot", precision="tf32")
    assert np.array_equal(_np(ti), np.tile(np.arange(k), (T, 1)))
    assert not _np(ts).any()


def test_depth_aware_block_culling_keeps_results():
    """Near occluders (depth clipped to 0.6 m in a 2 m room) leave most frustum blocks behind the surface:
    K2's adaptive depth cull switches on after the first frames and must not change any result."""
    cfg = synth.SceneConfig(extent=(2.2, 2.0, 1.6), voxel_size=0.05, height=48, width=64, patch_size=32,
                            patch_stride=16, feature_dim=8, frames=8, seed=31)
    origin, nvox = cfg.grid()
    g = dict(cls="ClipSeemFusion", feature_dim=8, origin=origin, nvox=nvox, voxel_size=cfg.voxel_size, trunc=cfg.trunc)
    vol, clip, seg = Hh.make_gpu_volume(g)
    orc = O.OracleVolume(origin, cfg.voxel_size, nvox, cfg.trunc, 8, num_threads=4)
    blocks = []
    for i in range(cfg.frames):
        fr = synth.make_frame(cfg, i)
        fr["depth"] = np.minimum(fr["depth"], np.float32(0.6))
        if i == 5:
            fr["depth"][:4] = np.nan          # NaN depths must neither crash the cull nor be used as the maximum
        clip.next_table = torch.from_numpy(fr["table"]).cuda()[None]
        seg.queue = [torch.from_numpy(fr["seg"]).cuda()]
        vol.integrate(torch.from_numpy(fr["depth"]).cuda()[None], torch.from_numpy(fr["rgb"]).cuda()[None],
                      torch.from_numpy(fr["pose"]).cuda()[None], torch.from_numpy(fr["K"]).cuda()[None])
        orc.integrate(fr["depth"][None], fr["rgb"][None], fr["pose"][None], fr["K"][None], fr["table"][None],
                      fr["seg"][None], want_masks=False)
        st = vol.stats()
        blocks.append(st["last_blocks"])
        assert st["last_valid"][0] == orc.last_counts[0, 0] and st["last_tsdf_valid"][0] == orc.last_counts[0, 1]
    _check_against(vol, orc.tsdf, orc.weight, orc.tsdf_weight, orc.rgb, orc.clip_feat, orc.labels_one_hot, exact=True)


# ---- window mode: saf_integrate_sequence fuses up to 16 consecutive frames per K0/K1/K2/K2T/K3W launch set ----

@pytest.mark.parametrize("name", ["seem_a", "fusion_a", "seem_edge"])
def test_sequence_window_matches_reference_golden(name):
    """One integrate_sequence call == the reference's frame loop of single-frame integrate() calls: bit-exact
    against the golden state the unmodified reference produced frame by frame."""
    g = Hh.load_golden(name)
    vol = Hh.replay_gpu_sequence(g)
    labels = Hh.golden_labels(g) if g["cls"] == "ClipSeemFusion" else None
    _check_against(vol, g["tsdf"], g["weight"], g["tsdf_weight"], g["rgb_state"], g["clip_feat"], labels, exact=True)
    st = vol.stats()
    assert st["total_frames"] == len(g["counts"])
    assert st["total_valid"] == int(g["counts"][:, 0].sum()) and st["total_tsdf_valid"] == int(g["counts"][:, 1].sum())


@pytest.mark.parametrize("n_frames,C,step", [(19, 768, 1), (43, 512, 3), (9, 20, 7)])
def test_sequence_window_matches_oracle(n_frames, C, step):
    """Longer sequences (several windows, the overlapped two-slot path for >= 4 windows, a ragged last window),
    the TMA-ring kernel (C = 768 / 512) and the generic one (C = 20), against the oracle run frame by frame."""
    cfg = synth.SceneConfig(extent=(2.2, 2.0, 1.6), voxel_size=0.05, height=96, width=128, patch_size=64,
                            patch_stride=32, feature_dim=C, frames=60, seed=21)
    origin, nvox = cfg.grid()
    g = dict(cls="ClipSeemFusion", feature_dim=C, origin=origin, nvox=nvox, voxel_size=cfg.voxel_size, trunc=cfg.trunc)
    vol, clip, seg = Hh.make_gpu_volume(g)
    orc = O.OracleVolume(origin, cfg.voxel_size, nvox, cfg.trunc, C, num_threads=0)
    frames = [synth.make_frame(cfg, (i * step) % cfg.frames) for i in range(n_frames)]
    total_valid = 0
    for fr in frames:
        cnt = orc.integrate(fr["depth"][None], fr["rgb"][None], fr["pose"][None], fr["K"][None], fr["table"][None],
                            fr["seg"][None], want_masks=False)
        total_valid += int(cnt[0, 0])
    clip.next_table = torch.stack([torch.from_numpy(f["table"]) for f in frames]).cuda()
    seg.queue = [torch.from_numpy(f["seg"]).cuda() for f in frames]
    vol.integrate_sequence(torch.stack([torch.from_numpy(f["depth"]) for f in frames]).cuda(),
                           torch.stack([torch.from_numpy(f["rgb"]) for f in frames]).cuda(),
                           torch.stack([torch.from_numpy(f["pose"]) for f in frames]).cuda(),
                           torch.stack([torch.from_numpy(f["K"]) for f in frames]).cuda())
    _check_against(vol, orc.tsdf, orc.weight, orc.tsdf_weight, orc.rgb, orc.clip_feat, orc.labels_one_hot, exact=True)
    st = vol.stats()
    assert st["total_frames"] == n_frames and st["total_valid"] == total_valid


def test_sequence_then_single_frames_interleave():
    """Window calls and plain integrate() calls share the workspace and the counters."""
    g = Hh.load_golden("seem_a")
    vol, clip, seg = Hh.make_gpu_volume(g)
    n = len(g["counts"])

    def feed(lo, hi, sequence):
        clip.next_table = torch.stack([torch.from_numpy(np.ascontiguousarray(Hh.golden_table(g, i)))
                                       for i in range(lo, hi)]).cuda()
        seg.queue = [torch.from_numpy(g["seg"][i].astype(np.int64)).cuda() for i in range(lo, hi)]
        args = [torch.from_numpy(g[k][lo:hi]).cuda() for k in ("depth", "rgb", "pose", "K")]
        (vol.integrate_sequence if sequence else vol.integrate)(*args)

    feed(0, 1, False)
    feed(1, 6, True)
    for i in range(6, n):
        feed(i, i + 1, False)
    _check_against(vol, g["tsdf"], g["weight"], g["tsdf_weight"], g["rgb_state"], g["clip_feat"], Hh.golden_labels(g),
                   exact=True)
    assert vol.stats()["total_valid"] == int(g["counts"][:, 0].sum())


def test_sequence_on_slab_drops_unreachable_frames():
    """Multi-GPU layout: two rooms side by side along x, this volume holds only room 0's slab and is handed every
    frame (cameras alternate between the rooms).  Frames that cannot touch the slab are dropped by the reach
    pre-pass; the result must equal the oracle fed every frame."""
    cfg = synth.SceneConfig(extent=(2.0, 2.0, 1.6), voxel_size=0.05, height=48, width=64, patch_size=32,
                            patch_stride=16, feature_dim=8, frames=40, seed=17)
    origin, nvox_room = cfg.grid()
    slab_nx = int(nvox_room[0]) + 4
    nvox = nvox_room.copy()
    nvox[0] = slab_nx * 2
    g = dict(cls="ClipSeemFusion", feature_dim=8, origin=origin, nvox=nvox, voxel_size=cfg.voxel_size, trunc=cfg.trunc)
    for own in (0, 1):
        xb, xe = own * slab_nx, (own + 1) * slab_nx
        vol, clip, seg = Hh.make_gpu_volume(g, x_begin=xb, x_end=xe)
        orc = O.OracleVolume(origin, cfg.voxel_size, nvox, cfg.trunc, 8, x_begin=xb, x_end=xe, num_threads=0)
        frames = []
        for i in range(24):
            fr = synth.make_frame(cfg, i)
            fr["pose"] = fr["pose"].copy()
            fr["pose"][0, 3] += (i % 2) * slab_nx * cfg.voxel_size
            frames.append(fr)
            orc.integrate(fr["depth"][None], fr["rgb"][None], fr["pose"][None], fr["K"][None], fr["table"][None],
                          fr["seg"][None], want_masks=False)
        clip.next_table = torch.stack([torch.from_numpy(f["table"]) for f in frames]).cuda()
        seg.queue = [torch.from_numpy(f["seg"]).cuda() for f in frames]
        vol.integrate_sequence(torch.stack([torch.from_numpy(f["depth"]) for f in frames]).cuda(),
                               torch.stack([torch.from_numpy(f["rgb"]) for f in frames]).cuda(),
                               torch.stack([torch.from_numpy(f["pose"]) for f in frames]),
                               torch.stack([torch.from_numpy(f["K"]) for f in frames]))
        _check_against(vol, orc.tsdf, orc.weight, orc.tsdf_weight, orc.rgb, orc.clip_feat, orc.labels_one_hot, exact=True)
        st = vol.stats()
        assert st["total_frames"] == 24 and st["total_valid"] == int(orc.weight.sum()) > 0


def test_checkpoint_resume_equals_uninterrupted_fusion(tmp_path):
    """Scene Manager v00 -> v01 (BASELINE config 5): integrate, save, load into a fresh volume, integrate more
    frames == integrating everything into one volume, bit for bit."""
    import spatially_aware_ai_b200 as saf
    g = Hh.load_golden("seem_a")
    n = len(g["counts"])

    def feed(vol, clip, seg, lo, hi):
        for i in range(lo, hi):
            clip.next_table = torch.from_numpy(np.ascontiguousarray(Hh.golden_table(g, i))).cuda()[None]
            seg.queue = [torch.from_numpy(g["seg"][i].astype(np.int64)).cuda()]
            vol.integrate(*[torch.from_numpy(g[k][i:i + 1]).cuda() for k in ("depth", "rgb", "pose", "K")])

    first, clip, seg = Hh.make_gpu_volume(g)
    feed(first, clip, seg, 0, 5)
    saf.save_state(first, str(tmp_path))
    resumed, clip2, seg2 = Hh.make_gpu_volume(g)
    saf.load_state(resumed, str(tmp_path))
    feed(resumed, clip2, seg2, 5, n)
    _check_against(resumed, g["tsdf"], g["weight"], g["tsdf_weight"], g["rgb_state"], g["clip_feat"],
                   Hh.golden_labels(g), exact=True)
