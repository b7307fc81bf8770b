"""GPU: extract_mesh (device marching cubes + vertex sampling, through the C ABI) against the reference's
golden output and the CPU oracle (oracle/mc.py).

Vertex positions, face indices and the nearest-sampled attributes are required bit-exact; the trilinear
samples reproduce torch-CPU's roundings and are required bit-exact too (tolerance 0), which is tighter than
the north-star bound (1e-5 abs / cosine >= 0.9999)."""
import numpy as np
import pytest
import torch

from oracle import mc
from tests import helpers as Hh

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy()


def _load_state(vol, g):
    vol.tsdf.copy_(torch.from_numpy(g["tsdf"]))
    vol.weight.copy_(torch.from_numpy(g["weight"]))
    vol.rgb.copy_(torch.from_numpy(g["rgb_state"]))
    vol.clip_feat.copy_(torch.from_numpy(g["clip_feat"]))


def test_extract_mesh_seem_matches_reference_golden():
    m, g = Hh.load_golden("mesh"), Hh.load_golden("seem_a")
    vol, _, _ = Hh.make_gpu_volume(g)
    _load_state(vol, g)
    vol.voxel_obj_idx = torch.from_numpy(m["seem_obj"]).cuda()
    vol.objects_segmentation_color = torch.from_numpy(m["seem_seg_color"]).cuda()
    verts, faces, colors, feats, obj, seg = vol.extract_mesh()
    assert isinstance(verts, np.ndarray) and verts.dtype == np.float32 and faces.dtype == np.int64
    assert colors.is_cuda and feats.is_cuda and obj.shape == (len(verts), 1)
    assert np.array_equal(verts, m["seem_verts"])
    assert np.array_equal(faces, m["seem_faces"])
    assert np.array_equal(_np(colors), m["seem_colors"])
    assert np.array_equal(_np(feats), m["seem_feats"])
    assert np.array_equal(_np(obj), m["seem_vertex_obj"])
    assert np.array_equal(_np(seg), m["seem_vertex_seg"])


def test_extract_mesh_fusion_matches_reference_golden():
    m, g = Hh.load_golden("mesh"), Hh.load_golden("fusion_a")
    vol, _, _ = Hh.make_gpu_volume(g)
    _load_state(vol, g)
    verts, faces, colors, feats = vol.extract_mesh()
    assert np.array_equal(verts, m["fusion_verts"]) and np.array_equal(faces, m["fusion_faces"])
    assert np.array_equal(_np(colors), m["fusion_colors"]) and np.array_equal(_np(feats), m["fusion_feats"])


def _random_volume(nvox, C, seed, x_begin=0, x_end=None):
    rng = np.random.default_rng(seed)
    n = int(np.prod(nvox))
    g = dict(cls="ClipFusion", feature_dim=C, origin=np.array([-1.0, 0.5, 2.0], np.float32),
             nvox=np.asarray(nvox, np.int32), voxel_size=0.03, trunc=0.09)
    state = dict(tsdf=rng.uniform(-1, 1, n).astype(np.float32), weight=rng.integers(0, 4, n).astype(np.int32),
                 rgb_state=rng.uniform(-0.2, 1.2, (n, 3)).astype(np.float32),
                 clip_feat=rng.standard_normal((n, C)).astype(np.float32))
    state["tsdf"][rng.random(n) < 0.02] = 0.0      # exact zeros: degenerate triangles are kept
    return g, state


@pytest.mark.parametrize("nvox,C", [((23, 19, 17), 12), ((9, 40, 33), 7), ((2, 2, 2), 4), ((40, 1, 30), 4)])
def test_marching_cubes_random_volume_matches_oracle(nvox, C):
    """Noise exercises every one of the 256 cases, the ambiguous faces, NaN corners and the grid boundary."""
    g, state = _random_volume(nvox, C, seed=sum(nvox))
    vol, _, _ = Hh.make_gpu_volume(g)
    _load_state(vol, state)
    verts, faces, colors, feats = vol.extract_mesh()
    rv, rf, rc, rfeat, _, _ = mc.extract_mesh(state["tsdf"], state["weight"], state["rgb_state"], state["clip_feat"],
                                               g["nvox"], g["voxel_size"], g["origin"])
    assert np.array_equal(verts, rv) and np.array_equal(faces, rf)
    assert np.array_equal(_np(colors), rc) and np.array_equal(_np(feats), rfeat)


def test_extract_mesh_all_unobserved_is_empty():
    g, state = _random_volume((8, 8, 8), 4, seed=1)
    state["weight"][:] = 0
    vol, _, _ = Hh.make_gpu_volume(g)
    _load_state(vol, state)
    verts, faces, colors, feats = vol.extract_mesh()
    assert verts.shape == (0, 3) and faces.shape == (0, 3) and colors.shape == (0, 3) and feats.shape == (0, 4)


def test_extract_mesh_on_x_slab():
    """A slab volume meshes its own cells with global coordinates (taps outside the slab are dropped)."""
    nvox, C, xb, xe = (20, 14, 12), 8, 6, 15
    g, state = _random_volume(nvox, C, seed=5)
    ny_nz = nvox[1] * nvox[2]
    sl = slice(xb * ny_nz, xe * ny_nz)
    vol, _, _ = Hh.make_gpu_volume(g, x_begin=xb, x_end=xe)
    _load_state(vol, {k: v[sl] for k, v in state.items()})
    verts, faces, colors, feats = vol.extract_mesh()
    sub = np.array([xe - xb, nvox[1], nvox[2]], np.int32)
    raw_v, raw_f = mc.filter_mesh(*mc.marching_cubes_raw(mc.masked_tsdf(state["tsdf"][sl], state["weight"][sl], sub),
                                                         x_offset=xb))
    world = (raw_v * np.float32(g["voxel_size"]) + g["origin"]).astype(np.float32)
    assert np.array_equal(faces, raw_f) and np.array_equal(verts, world)
    # sampling: the full-grid oracle, restricted to taps inside the slab == oracle on a zero-padded copy
    padded = np.zeros_like(state["clip_feat"])
    padded[sl] = state["clip_feat"][sl]
    assert np.array_equal(_np(feats), mc.sample_trilinear(padded, raw_v, g["nvox"]))


def test_mesh_of_fused_scene_is_closed_band():
    """End to end at a modest size: fuse a synthetic room, extract the mesh, check it against the oracle applied
    to the same device state and that every vertex lies within one voxel of an observed surface voxel."""
    from spatially_aware_ai_b200 import synth
    cfg = synth.SceneConfig(extent=(2.2, 2.0, 1.6), voxel_size=0.05, height=96, width=128, patch_size=64,
                            patch_stride=32, feature_dim=16, frames=6, seed=8)
    origin, nvox = cfg.grid()
    g = dict(cls="ClipSeemFusion", feature_dim=16, origin=origin, nvox=nvox, voxel_size=cfg.voxel_size, trunc=cfg.trunc)
    vol, clip, seg = Hh.make_gpu_volume(g)
    for i in range(cfg.frames):
        fr = synth.make_frame(cfg, i * 9)
        clip.next_table = torch.from_numpy(fr["table"]).cuda()[None]
        seg.queue = [torch.from_numpy(fr["seg"]).cuda()]
        vol.integrate(torch.from_numpy(fr["depth"]).cuda()[None], torch.from_numpy(fr["rgb"]).cuda()[None],
                      torch.from_numpy(fr["pose"]).cuda()[None], torch.from_numpy(fr["K"]).cuda()[None])
    vol.voxel_obj_idx = vol.label_argmax().view(*[int(v) for v in nvox])
    vol.objects_segmentation_color = vol.rgb.clone()
    verts, faces, colors, feats, obj, segc = vol.extract_mesh()
    assert len(verts) > 300 and len(faces) > 300
    ref = mc.extract_mesh(_np(vol.tsdf), _np(vol.weight), _np(vol.rgb), _np(vol.clip_feat), nvox, cfg.voxel_size,
                          origin, _np(vol.voxel_obj_idx), _np(vol.objects_segmentation_color))
    for got, want in zip((verts, faces, _np(colors), _np(feats), _np(obj), _np(segc)), ref):
        assert np.array_equal(got, want)
    idx = np.rint((verts - origin) / cfg.voxel_size).astype(int)
    w = _np(vol.weight).reshape([int(v) for v in nvox])
    assert (w[idx[:, 0], idx[:, 1], idx[:, 2]] > 0).mean() > 0.99


def test_slab_meshes_with_halo_weld_to_the_full_grid_mesh():
    """Multi-GPU mesh extraction, emulated in one process: three slab volumes, each given its successor's first
    plane (what slab.exchange_halo delivers over NCCL), mesh their own cells plus the cells across the cut;
    welded, the result is the mesh of the whole grid: identical vertex positions and faces, attributes of the
    vertices identical except on the cut planes (there a tap of weight ~1e-7 lies in the predecessor's slab:
    1e-5 tolerance)."""
    from spatially_aware_ai_b200 import slab
    nvox, C = (22, 13, 11), 8
    g, state = _random_volume(nvox, C, seed=77)
    full, _, _ = Hh.make_gpu_volume(g)
    _load_state(full, state)
    fv, ff, fc, ffeat = full.extract_mesh()
    ny_nz = nvox[1] * nvox[2]
    cuts = [0, 7, 15, nvox[0]]
    vols = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        v, _, _ = Hh.make_gpu_volume(g, x_begin=a, x_end=b)
        _load_state(v, {k: s[a * ny_nz:b * ny_nz] for k, s in state.items()})
        vols.append(v)
    parts, planes = [], []
    for r, v in enumerate(vols):
        halo = slab.first_plane(vols[r + 1]) if r + 1 < len(vols) else None
        verts, faces, cols, feats, ids = v.extract_mesh(halo=halo, return_edge_ids=True)
        # against the oracle on the slab plus its halo plane (global coordinates)
        a, b = cuts[r], cuts[r + 1]
        hi = min(b + 1, nvox[0])
        ov, of, oids = mc.filter_mesh(*mc.marching_cubes_raw(
            mc.masked_tsdf(state["tsdf"][a * ny_nz:hi * ny_nz], state["weight"][a * ny_nz:hi * ny_nz], (hi - a, nvox[1], nvox[2])),
            x_offset=a, return_ids=True))
        world = (ov * np.float32(g["voxel_size"]) + g["origin"]).astype(np.float32)
        assert np.array_equal(faces, of) and np.array_equal(verts, world) and np.array_equal(ids, oids)
        parts.append((verts, faces, ids, (_np(cols), _np(feats))))
        if b < nvox[0]:
            planes.append(np.float32(np.float32(b) * np.float32(g["voxel_size"])) + g["origin"][0])
    wv, wf, (wc, wfeat) = slab.weld_slab_meshes(parts)
    assert len(wv) == len(fv) and len(wf) == len(ff)
    # both meshes list their vertices in ascending edge-id order, slab by slab: after welding the orders coincide
    assert np.array_equal(wv, fv) and np.array_equal(wf, ff)
    dc = np.abs(wc - _np(fc)).max(axis=1)
    df = np.abs(wfeat - _np(ffeat)).max(axis=1)
    on_cut = np.isin(wv[:, 0], np.array(planes, np.float32))
    assert on_cut.any() and not dc[~on_cut].any() and not df[~on_cut].any()
    assert dc.max() <= 1e-5 and df.max() <= 1e-5
