"""GPU: round-2 additions - the window tile kernel at every compiled width, sensor-format inputs, block-cyclic
slabs, the query drivers and the per-row / per-text scorers - against the oracle and the reference's goldens."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from spatially_aware_ai_b200 import slab, synth
from tests import helpers as Hh
from tests.test_parity_gpu import _check_against, _np

pytestmark = pytest.mark.gpu


def _scene(C, seed=21, frames=60, h=96, w=128, patch=64, stride=32, extent=(2.2, 2.0, 1.6)):
    cfg = synth.SceneConfig(extent=extent, voxel_size=0.05, height=h, width=w, patch_size=patch, patch_stride=stride,
                            feature_dim=C, frames=frames, seed=seed)
    origin, nvox = cfg.grid()
    g = dict(cls="ClipSeemFusion", feature_dim=C, origin=origin, nvox=nvox, voxel_size=cfg.voxel_size, trunc=cfg.trunc)
    return cfg, origin, nvox, g


def _oracle_run(cfg, origin, nvox, C, frames, **kw):
    orc = O.OracleVolume(origin, cfg.voxel_size, nvox, cfg.trunc, C, num_threads=0, **kw)
    for fr in frames:
        orc.integrate(fr["depth"][None], fr["rgb"][None], fr["pose"][None], fr["K"][None], fr["table"][None],
                      fr["seg"][None], want_masks=False)
    return orc


def _sequence(vol, clip, seg, frames, sensor=False, pass_producers=False):
    tables = torch.stack([torch.from_numpy(f["table"]) for f in frames]).cuda()
    segs = [torch.from_numpy(f["seg"]).cuda() for f in frames]
    if sensor:
        depth = torch.stack([torch.from_numpy(f["depth_mm"]) for f in frames]).cuda()
        rgb = torch.stack([torch.from_numpy(f["rgb_u8"]) for f in frames]).cuda()
    else:
        depth = torch.stack([torch.from_numpy(f["depth"]) for f in frames]).cuda()
        rgb = torch.stack([torch.from_numpy(f["rgb"]) for f in frames]).cuda()
    poses = torch.stack([torch.from_numpy(f["pose"]) for f in frames])
    Ks = torch.stack([torch.from_numpy(f["K"]) for f in frames])
    if pass_producers:
        vol.integrate_sequence(depth, rgb, poses, Ks, clip_feat_img=tables, seg_maps=torch.stack(segs))
    else:
        clip.next_table, seg.queue = tables, segs
        vol.integrate_sequence(depth, rgb, poses, Ks)


@pytest.mark.parametrize("C,n_frames,step", [(1024, 21, 1), (768, 37, 2), (512, 18, 5), (6, 11, 3), (132, 9, 4)])
def test_window_kernels_every_width(C, n_frames, step):
    """Window mode for the tile kernel (C = 512 / 768 / 1024, incl. ragged last tiles and windows), the generic
    16-byte kernel (C = 132) and the scalar one (C = 6: rows are not 16-byte multiples), bit-exact vs the oracle."""
    cfg, origin, nvox, g = _scene(C)
    vol, clip, seg = Hh.make_gpu_volume(g)
    frames = [synth.make_frame(cfg, (i * step) % cfg.frames) for i in range(n_frames)]
    orc = _oracle_run(cfg, origin, nvox, C, frames)
    _sequence(vol, clip, seg, frames)
    _check_against(vol, orc.tsdf, orc.weight, orc.tsdf_weight, orc.rgb, orc.clip_feat, orc.labels_one_hot, exact=True)
    st = vol.stats()
    assert st["total_valid"] == int(orc.weight.sum()) and 0 < st["total_union"] <= st["total_valid"]


@pytest.mark.parametrize("variant", ["0", "1"])
def test_older_window_kernels_still_exact(variant, monkeypatch):
    """SAF_K3W_VARIANT selects the one-voxel / pair kernels (kept for A/B timing).  The variable is read once per
    process, so this runs in a subprocess."""
    import os
    import subprocess
    import sys
    code = ("import os, sys; sys.path.insert(0, %r)\n"
            "import tests.test_round2_gpu as T\nT.test_window_kernels_every_width(768, 19, 1)\nprint('ok')\n"
            % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ, SAF_K3W_VARIANT=variant)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


@pytest.mark.parametrize("cls", ["ClipSeemFusion", "ClipFusion"])
def test_sensor_format_inputs_equal_fp32_inputs(cls):
    """uint16-mm depth + uint8 rgb through integrate() and integrate_sequence() == the fp32 tensors the reference's
    dataset classes would have produced from them (clipfusion.py:185-188), bit for bit, for both rgb samplers."""
    C = 512
    cfg, origin, nvox, g = _scene(C, seed=8)
    g["cls"] = cls
    frames = [synth.make_frame(cfg, i) for i in range(20)]
    a, clip_a, seg_a = Hh.make_gpu_volume(g)
    b, clip_b, seg_b = Hh.make_gpu_volume(g)
    # first three frames one by one, the rest as a sequence
    for vol, clip, seg, sensor in ((a, clip_a, seg_a, False), (b, clip_b, seg_b, True)):
        for fr in frames[:3]:
            clip.next_table = torch.from_numpy(fr["table"]).cuda()[None]
            seg.queue = [torch.from_numpy(fr["seg"]).cuda()]
            d = torch.from_numpy(fr["depth_mm"] if sensor else fr["depth"]).cuda()[None]
            r = torch.from_numpy(fr["rgb_u8"] if sensor else fr["rgb"]).cuda()[None]
            vol.integrate(d, r, torch.from_numpy(fr["pose"])[None], torch.from_numpy(fr["K"])[None])
        _sequence(vol, clip, seg, frames[3:], sensor=sensor, pass_producers=sensor)
    for name in ("tsdf", "weight", "tsdf_weight", "rgb", "clip_feat") + (("labels_one_hot",) if cls == "ClipSeemFusion" else ()):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    assert int(a.weight.sum()) > 0
    orc = O.OracleVolume(origin, cfg.voxel_size, nvox, cfg.trunc, C, with_labels=(cls == "ClipSeemFusion"), num_threads=0)
    for fr in frames:
        orc.integrate(O.depth_from_mm(fr["depth_mm"])[None], O.rgb_from_u8(fr["rgb_u8"])[None], fr["pose"][None],
                      fr["K"][None], fr["table"][None], fr["seg"][None] if cls == "ClipSeemFusion" else None,
                      want_masks=False)
    _check_against(b, orc.tsdf, orc.weight, orc.tsdf_weight, orc.rgb, orc.clip_feat, orc.labels_one_hot, exact=True)


@pytest.mark.parametrize("world", [2, 3])
def test_block_cyclic_slabs_concatenate_to_the_full_grid(world):
    """Every rank's block-cyclic volume (emulated in one process) holds exactly the oracle's rows of its stripes,
    through single-frame calls and window mode; the stripes cover the grid once."""
    C = 768
    cfg, origin, nvox, g = _scene(C, seed=5, extent=(2.6, 1.8, 1.5))
    frames = [synth.make_frame(cfg, i * 3 % cfg.frames) for i in range(22)]
    orc = _oracle_run(cfg, origin, nvox, C, frames)
    nx, plane = int(nvox[0]), int(nvox[1]) * int(nvox[2])
    covered = np.zeros(nx, int)
    for r in range(world):
        kw = slab.cyclic_slab(nx, world, r)
        vol, clip, seg = Hh.make_gpu_volume(g, **kw)
        fr = frames[0]
        clip.next_table = torch.from_numpy(fr["table"]).cuda()[None]
        seg.queue = [torch.from_numpy(fr["seg"]).cuda()]
        vol.integrate(torch.from_numpy(fr["depth"]).cuda()[None], torch.from_numpy(fr["rgb"]).cuda()[None],
                      torch.from_numpy(fr["pose"])[None], torch.from_numpy(fr["K"])[None])
        _sequence(vol, clip, seg, frames[1:])
        xs = np.array(vol.global_x_planes())
        covered[xs] += 1
        rows = (xs[:, None] * plane + np.arange(plane)[None]).reshape(-1)
        _check_against(vol, orc.tsdf[rows], orc.weight[rows], orc.tsdf_weight[rows], orc.rgb[rows], orc.clip_feat[rows],
                       orc.labels_one_hot[rows], exact=True)
        # query rows map back to global voxel ids
        X = torch.nn.functional.normalize(torch.randn(3, C, generator=torch.Generator().manual_seed(1)), dim=-1).cuda()
        import spatially_aware_ai_b200 as saf
        ts, ti = saf.query_topk(vol.clip_feat, X, 4, norm="nan_to_num", mode="dot")
        gi = _np(slab.local_to_global_rows(vol, ti))
        # neighbouring voxels can score within an ulp of each other: rank the kernel's own fp32 scores
        scores = _np(saf.query_scores(vol.clip_feat, X, norm="nan_to_num", mode="dot"))
        assert np.abs(scores - O.normalize_rows(orc.clip_feat[rows]) @ _np(X).T).max() <= 1e-5
        assert np.array_equal(gi, rows[O.topk_indices(scores, 4)])
    assert (covered == 1).all()


def test_text_query_driver_on_device():
    """clip_text_query (clip_seem_fusion.py:507-533) through the kernels: nan_to_num normalisation fused into the
    scorer, surgery epilogue, relevance post-processing - against the unmodified driver's output."""
    import spatially_aware_ai_b200 as saf
    g = Hh.load_golden("query_drivers")
    F, X = torch.from_numpy(g["feats"]).cuda(), torch.from_numpy(g["text"]).cuda()
    Fn = torch.nan_to_num(F / F.norm(dim=-1, keepdim=True))
    sim = saf.Clip.clip_feature_surgery(Fn[None], X)
    rel = saf.relevance_minmax(sim[0, :, int(g["text_query_column"])])
    assert np.abs(_np(rel) - g["text_query_relevance"]).max() <= 2e-5
    # the same with the normalisation inside the kernel (row 0's weights from the normalised row 0)
    w = saf.surgery_weights(Fn[0], X)
    sim2 = saf.query_scores(F, X, norm="nan_to_num", mode="surgery", surgery_w=w)
    assert np.abs(_np(saf.relevance_minmax(sim2[:, int(g["text_query_column"])])) - g["text_query_relevance"]).max() <= 2e-5


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_segment_labels_match_reference(precision):
    """segment() (eval_scannet_segmentation.py:546-561): clamp_min normalisation in-kernel, labels per row."""
    import spatially_aware_ai_b200 as saf
    g = Hh.load_golden("query_drivers")
    F, X = torch.from_numpy(g["segment_feats"]).cuda(), torch.from_numpy(g["text"]).cuda()
    ref = g["segment_labels"]
    full = _np(saf.segment_labels(F, X, precision=precision))
    assert full.shape == ref.shape
    if precision == "fp32":
        assert np.array_equal(full[:, :5], ref[:, :5]) and (full == ref).mean() > 0.999
    else:
        assert (full[:, 0] == ref[:, 0]).mean() > 0.99
    top5, probs = saf.segment_labels(F, X, k=5, precision=precision, return_probs=True)
    assert np.array_equal(_np(top5), full[:, :5])
    rel = O._softmax(np.float32(100) * (O.normalize_rows(g["segment_feats"], "clamp_min") @ g["text"].T), -1)
    want = np.take_along_axis(rel, ref[:, :5], axis=1)
    assert np.abs(_np(probs) - want).max() <= (2e-5 if precision == "fp32" else 5e-2)
    # many rows: several 64 Ki-row chunks, ragged tail
    M = 150_001
    big = torch.randn(M, 24, generator=torch.Generator().manual_seed(3)).cuda()
    lab = _np(saf.segment_labels(big, X, k=3, precision="fp32"))
    want = O.segment_labels(_np(big), g["text"])[:, :3]
    assert (lab == want).mean() > 0.9999


def test_presence_scores_match_reference():
    """hypersim_eval.py:80-89: max relevance per target label in one pass."""
    import spatially_aware_ai_b200 as saf
    g = Hh.load_golden("query_drivers")
    F, X = torch.from_numpy(g["feats"]).cuda(), torch.from_numpy(g["text"]).cuda()
    pres = _np(saf.presence_scores(F, X[:4], X[4:]))
    assert np.abs(pres - g["hypersim_presence"]).max() <= 2e-5


def test_query_mesh_post_processing_on_device():
    import spatially_aware_ai_b200 as saf
    g = Hh.load_golden("query_drivers")
    sim = torch.from_numpy(g["query_mesh_minmax"]).cuda()
    for n in range(sim.shape[1]):
        out = _np(saf.relevance_outliers(sim[:, n]))
        assert np.array_equal(out, g["query_mesh_outliers"][:, n])
    F = torch.from_numpy(g["segment_feats"]).cuda()
    F = F / F.norm(dim=-1, keepdim=True)
    X = torch.from_numpy(g["text"]).cuda()
    s = saf.Clip.clip_feature_surgery(F[None], X)
    assert np.abs(_np(saf.minmax_per_text(s)[0]) - g["query_mesh_minmax"]).max() <= 2e-5
    rel = saf.query_scores(F, X[:5], mode="softmax100")[:, -1]
    assert np.abs(_np(saf.relevance_half(rel)) - g["query_mesh_half"]).max() <= 2e-5


def test_volume_on_a_non_current_device_is_refused_or_correct():
    """Every ctypes call runs with the tensors' device current (a volume on cuda:1 while cuda:0 is current)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    g = Hh.load_golden("seem_a")
    torch.cuda.set_device(0)
    vol, _ = Hh.replay_gpu(g, device="cuda:1")
    _check_against(vol, g["tsdf"], g["weight"], g["tsdf_weight"], g["rgb_state"], g["clip_feat"], Hh.golden_labels(g),
                   exact=True)


@pytest.mark.parametrize("C", [768, 20])
def test_segment_table_mode(C):
    """north_star's segment -> CLIP table (SAF_TABLE_SEGMENTS): a voxel's sample is the row of its nearest-sampled
    class id.  Single-frame kernels and window mode (tile kernel at C = 768, generic at C = 20) against the
    oracle's definition of the mode, bit for bit; a bad class id is reported."""
    cfg, origin, nvox, g = _scene(C, seed=13)
    n_seg = 134
    rng = np.random.default_rng(5)
    frames = [synth.make_frame(cfg, i) for i in range(19)]
    tables = [rng.standard_normal((n_seg, C), dtype=np.float32) for _ in frames]
    orc = O.OracleVolume(origin, cfg.voxel_size, nvox, cfg.trunc, C, num_threads=0)
    for fr, tab in zip(frames, tables):
        orc.integrate(fr["depth"][None], fr["rgb"][None], fr["pose"][None], fr["K"][None],
                      np.ascontiguousarray(tab.T)[None, :, None, :], fr["seg"][None], want_masks=False, table_mode=1)
    vol, clip, seg = Hh.make_gpu_volume(g)
    vol.feature_source = "segment_table"
    dev_tables = torch.from_numpy(np.stack(tables)).cuda()

    class _SegClip:
        def segment_features(self, rgb_chw, seg_maps):
            return self.next

    vol.clip = _SegClip()
    for i in range(3):          # the reference's call pattern: one frame per integrate()
        fr = frames[i]
        vol.clip.next = dev_tables[i:i + 1]
        seg.queue = [torch.from_numpy(fr["seg"]).cuda()]
        vol.integrate(torch.from_numpy(fr["depth"]).cuda()[None], torch.from_numpy(fr["rgb"]).cuda()[None],
                      torch.from_numpy(fr["pose"])[None], torch.from_numpy(fr["K"])[None])
    rest = frames[3:]
    vol.integrate_sequence(torch.stack([torch.from_numpy(f["depth"]) for f in rest]).cuda(),
                           torch.stack([torch.from_numpy(f["rgb"]) for f in rest]).cuda(),
                           torch.stack([torch.from_numpy(f["pose"]) for f in rest]),
                           torch.stack([torch.from_numpy(f["K"]) for f in rest]),
                           clip_feat_img=dev_tables[3:], seg_maps=torch.stack([torch.from_numpy(f["seg"]) for f in rest]).cuda())
    _check_against(vol, orc.tsdf, orc.weight, orc.tsdf_weight, orc.rgb, orc.clip_feat, orc.labels_one_hot, exact=True)
    assert int(orc.weight.sum()) > 0
    # a table with fewer rows than the class map's ids: zeros are sampled and the sticky flag is raised
    vol.integrate_sequence(torch.from_numpy(frames[0]["depth"]).cuda()[None].repeat(2, 1, 1),
                           torch.from_numpy(frames[0]["rgb"]).cuda()[None].repeat(2, 1, 1, 1),
                           torch.from_numpy(frames[0]["pose"])[None].repeat(2, 1, 1),
                           torch.from_numpy(frames[0]["K"])[None].repeat(2, 1, 1),
                           clip_feat_img=dev_tables[:2, :40], seg_maps=torch.from_numpy(frames[0]["seg"]).cuda()[None].repeat(2, 1, 1))
    with pytest.raises(RuntimeError):
        vol.check_errors()


@pytest.mark.parametrize("world", [2, 3])
def test_sheared_block_columns_concatenate_to_the_full_grid(world):
    """Sheared block-column layout (column (bx, by) on rank (bx + by) % n): every rank's volume holds exactly the
    oracle's rows of its columns, padding rows stay zero, query rows map back to global voxel ids."""
    import spatially_aware_ai_b200 as saf
    C = 768
    cfg, origin, nvox, g = _scene(C, seed=6, extent=(2.6, 1.9, 1.5))
    frames = [synth.make_frame(cfg, i * 3 % cfg.frames) for i in range(21)]
    orc = _oracle_run(cfg, origin, nvox, C, frames)
    covered = np.zeros(orc.n, int)
    for r in range(world):
        vol, clip, seg = Hh.make_gpu_volume(g, **slab.sheared_slab(world, r))
        fr = frames[0]
        clip.next_table = torch.from_numpy(fr["table"]).cuda()[None]
        seg.queue = [torch.from_numpy(fr["seg"]).cuda()]
        vol.integrate(torch.from_numpy(fr["depth"]).cuda()[None], torch.from_numpy(fr["rgb"]).cuda()[None],
                      torch.from_numpy(fr["pose"])[None], torch.from_numpy(fr["K"])[None])
        _sequence(vol, clip, seg, frames[1:])
        rows = vol.global_rows().numpy()
        real = rows >= 0
        covered[rows[real]] += 1
        sel = np.where(real, rows, 0)
        pad0 = lambda a: np.where(real.reshape((-1,) + (1,) * (a.ndim - 1)), a[sel], 0).astype(a.dtype)   # noqa: E731
        _check_against(vol, pad0(orc.tsdf), pad0(orc.weight), pad0(orc.tsdf_weight), pad0(orc.rgb), pad0(orc.clip_feat),
                       pad0(orc.labels_one_hot), exact=True)
        X = torch.nn.functional.normalize(torch.randn(3, C, generator=torch.Generator().manual_seed(2)), dim=-1).cuda()
        ts, ti = saf.query_topk(vol.clip_feat, X, 4, norm="nan_to_num", mode="dot")
        scores = _np(saf.query_scores(vol.clip_feat, X, norm="nan_to_num", mode="dot"))
        assert np.array_equal(_np(slab.local_to_global_rows(vol, ti)), rows[O.topk_indices(scores, 4)])
    assert (covered == 1).all()


def test_fused_topk_on_a_mostly_unobserved_grid():
    """A fused grid is mostly all-zero rows (unobserved voxels).  The tensor-core top-k must neither rank them above
    positive scores nor drown in them (they are skipped by the filter, and asked for only when needed)."""
    import spatially_aware_ai_b200 as saf
    gen = torch.Generator().manual_seed(9)
    M, C, T, k = 300_000, 768, 40, 100
    F = torch.zeros(M, C)
    obs = torch.randperm(M, generator=gen)[:20_000]
    obs = obs[obs > 150_000]                       # the first half of the rows is entirely unobserved
    F[obs] = torch.randn(len(obs), C, generator=gen)
    X = torch.nn.functional.normalize(torch.randn(T, C, generator=gen), dim=-1)
    Fd, Xd = F.cuda(), X.cuda()
    ts, ti = saf.query_topk(Fd, Xd, k, norm="nan_to_num", mode="dot", precision="tf32")
    scores = _np(saf.query_scores(Fd, Xd, norm="nan_to_num", mode="dot"))
    assert np.array_equal(_np(ti), O.topk_indices(scores, k))
    assert float(ts.min()) > 0
    # fewer positive rows than k: zero rows (lowest indices first) complete the list, through the exact path
    F2 = torch.zeros(5000, C)
    F2[4000:4030] = torch.randn(30, C, generator=gen)
    ts2, ti2 = saf.query_topk(F2.cuda(), Xd, k, norm="nan_to_num", mode="dot", precision="tf32")
    s2 = _np(saf.query_scores(F2.cuda(), Xd, norm="nan_to_num", mode="dot"))
    assert np.array_equal(_np(ti2), O.topk_indices(s2, k))


# ---- BASELINE configs 3 / 4 at full size ------------------------------------------------------------------

def _free_gb():
    import gc
    gc.collect()
    torch.cuda.empty_cache()      # the previous tests' volumes sit in torch's caching allocator
    free, _ = torch.cuda.mem_get_info()
    return free / 2 ** 30


def test_sampled_slab_at_cfg3_scale_window16():
    """BASELINE config 3 (8x8x3 m at 2 cm: 25.1 M voxels x 768-d) is far too large for the oracle; 32 frames of its
    orbit are fused in two 16-frame windows (K0/K1/K2/K2T/K3W) and a 12-plane x-slab of the result must equal the
    oracle bit for bit, next to the size-independent invariants."""
    if _free_gb() < 130:
        pytest.skip("needs ~110 GB of device memory")
    cfg = synth.baseline_config("cfg3", frames=32)
    origin, nvox = cfg.grid()
    g = dict(cls="ClipSeemFusion", feature_dim=cfg.feature_dim, origin=origin, nvox=nvox,
             voxel_size=cfg.voxel_size, trunc=cfg.trunc)
    vol, clip, seg = Hh.make_gpu_volume(g)
    frames = [synth.make_frame(cfg, i * 3, table_layout="hwc") for i in range(cfg.frames)]
    clip.next_table = torch.stack([torch.from_numpy(np.ascontiguousarray(f["table"].transpose(1, 2, 0)))
                                   for f in frames]).cuda().permute(0, 3, 1, 2)
    seg.queue = [torch.from_numpy(f["seg"]).cuda() for f in frames]
    vol.integrate_sequence(torch.stack([torch.from_numpy(f["depth"]) for f in frames]).cuda(),
                           torch.stack([torch.from_numpy(f["rgb"]) for f in frames]).cuda(),
                           torch.stack([torch.from_numpy(f["pose"]) for f in frames]),
                           torch.stack([torch.from_numpy(f["K"]) for f in frames]))
    # the 12 planes these frames update most (the oracle only computes that slab)
    nx, ny, nz = (int(v) for v in nvox)
    per_plane = vol.weight.view(nx, ny * nz).sum(dim=1, dtype=torch.int64)
    win = torch.nn.functional.avg_pool1d(per_plane.double()[None, None], 12, 1)[0, 0]
    xs0 = int(win.argmax())
    xs1 = xs0 + 12
    orc = O.OracleVolume(origin, cfg.voxel_size, nvox, cfg.trunc, cfg.feature_dim, x_begin=xs0, x_end=xs1, num_threads=0)
    for fr in frames:
        orc.integrate(fr["depth"][None], fr["rgb"][None], fr["pose"][None], fr["K"][None], fr["table"][None],
                      fr["seg"][None], want_masks=False)
    st = vol.stats()
    assert st["total_frames"] == cfg.frames and st["total_calls"] == 2
    assert int(vol.weight.sum(dtype=torch.int64)) == st["total_valid"] > 0
    assert int(vol.tsdf_weight.sum(dtype=torch.int64)) == st["total_tsdf_valid"] >= st["total_valid"]
    assert torch.equal(vol.labels_one_hot.sum(dim=1, dtype=torch.int32), vol.weight)
    sl = slice(xs0 * ny * nz, xs1 * ny * nz)
    assert orc.weight.sum() > 10000
    assert np.array_equal(_np(vol.weight[sl]), orc.weight)
    assert np.array_equal(_np(vol.tsdf_weight[sl]), orc.tsdf_weight)
    assert np.array_equal(_np(vol.tsdf[sl]), orc.tsdf)
    assert np.array_equal(_np(vol.rgb[sl]), orc.rgb)
    assert np.array_equal(_np(vol.clip_feat[sl]), orc.clip_feat)
    assert np.array_equal(_np(vol.labels_one_hot[sl]), orc.labels_one_hot)


def test_fused_topk_at_cfg4_scale_equals_chunked_fp32_ranking():
    """BASELINE config 4: 256 texts against 24 M feature rows x 768-d.  The fused tensor-core top-100 (persistent tf32
    GEMM + candidate filter + fp32 rescoring, no [M,T] matrix) must return exactly the ranking of the fp32 scores,
    here ranked chunk by chunk (1 M rows at a time) and merged."""
    import spatially_aware_ai_b200 as saf
    if _free_gb() < 100:
        pytest.skip("needs ~80 GB of device memory")
    M, C, T, k, chunk = 24_000_000, 768, 256, 100, 1_000_000
    gen = torch.Generator(device="cuda").manual_seed(24)
    F = torch.empty((M, C), device="cuda")
    for r0 in range(0, M, chunk):                      # a fused grid: most rows unobserved (zero), the rest smooth-ish
        blk = torch.randn((chunk, C), device="cuda", generator=gen)
        keep = torch.rand((chunk, 1), device="cuda", generator=gen) < 0.3
        F[r0:r0 + chunk] = blk * keep
    X = torch.nn.functional.normalize(torch.randn((T, C), device="cuda", generator=gen), dim=-1)
    ts, ti = saf.query_topk(F, X, k, norm="nan_to_num", mode="dot", precision="tf32")
    best_s = torch.full((T, 0), 0.0, device="cuda")
    best_i = torch.zeros((T, 0), dtype=torch.int64, device="cuda")
    for r0 in range(0, M, chunk):
        sc = saf.query_scores(F[r0:r0 + chunk], X, norm="nan_to_num", mode="dot", precision="fp32").T.contiguous()   # [T, chunk]
        cs, ci = torch.topk(sc, k, dim=1)
        best_s = torch.cat([best_s, cs], dim=1)
        best_i = torch.cat([best_i, ci + r0], dim=1)
        # keep the k best so far; ties broken by the lower row index, as the library does
        order = torch.argsort(best_i, dim=1, stable=True)
        best_s, best_i = torch.gather(best_s, 1, order), torch.gather(best_i, 1, order)
        order = torch.argsort(best_s, dim=1, descending=True, stable=True)[:, :k]
        best_s, best_i = torch.gather(best_s, 1, order), torch.gather(best_i, 1, order)
    assert torch.equal(ts, best_s)
    assert torch.equal(ti.to(torch.int64), best_i)


def test_window_path_without_prepared_tiles():
    """SAF_TILE_SETUP=0: K2T keeps the small state but leaves every tile's metadata to K3W's own producers (the
    path taken by tiles past K2T's capacity).  The library reads the switch once per process, hence a subprocess."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, SAF_TILE_SETUP="0")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-m", "pytest", "tests/test_parity_gpu.py", "tests/test_round2_gpu.py", "-q", "-x", "-m", "gpu",
                          "-k", "sequence_window_matches_oracle or window_kernels_every_width or segment_table_mode",
                          "-p", "no:cacheprovider"], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-1000:]
    assert " passed" in res.stdout
