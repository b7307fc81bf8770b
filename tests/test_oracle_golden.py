"""CPU: the oracle (oracle/saf_oracle.c, oracle/oracle.py) against the reference's golden vectors."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import helpers as Hh


@pytest.mark.parametrize("name", Hh.FUSION_GOLDENS)
def test_fusion_oracle_matches_reference(name):
    g = Hh.load_golden(name)
    vol, counts = Hh.replay_oracle(g)
    # integer / mask state: bit-exact
    assert np.array_equal(vol.weight, g["weight"])
    assert np.array_equal(vol.tsdf_weight, g["tsdf_weight"])
    assert np.array_equal(counts, g["counts"])
    if g["cls"] == "ClipSeemFusion":
        assert np.array_equal(vol.labels_one_hot, Hh.golden_labels(g))
    if g["batch"] == 1:
        # float state: the restatement reproduces torch-CPU's roundings exactly in this regime
        assert np.array_equal(vol.tsdf, g["tsdf"])
        assert np.array_equal(vol.rgb, g["rgb_state"])
        assert np.array_equal(vol.clip_feat, g["clip_feat"])
    else:
        # B > 1: MKL's batched sgemm runs each batch item on one thread with a different fp32
        # summation order (see make_golden.py), so z differs in the last bits -> north_star tolerances.
        assert np.abs(vol.tsdf - g["tsdf"]).max() <= 1e-5
        assert np.abs(vol.rgb - g["rgb_state"]).max() <= 1e-5
        assert Hh.cosine_rows(vol.clip_feat, g["clip_feat"]).min() >= 0.9999


@pytest.mark.parametrize("name", ["seem_a", "fusion_b2"])
def test_fusion_oracle_mid_state(name):
    g = Hh.load_golden(name)
    vol, _ = Hh.replay_oracle(g, upto=len(g["counts"]) // 2)
    assert np.abs(vol.tsdf - g["mid_tsdf"]).max() <= (0 if g["batch"] == 1 else 1e-5)
    assert np.array_equal(vol.weight, g["mid_weight"])
    assert np.array_equal(vol.tsdf_weight, g["mid_tsdf_weight"])


def test_fusion_oracle_slabs_concatenate():
    """x-slabs with global indices reproduce the full grid exactly (SURVEY 7.3 slab bit-exactness)."""
    g = Hh.load_golden("seem_a")
    nx = int(g["nvox"][0])
    cuts = [0, 13, 30, nx]
    parts = [Hh.replay_oracle(g, x_begin=a, x_end=b)[0] for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.array_equal(np.concatenate([p.tsdf for p in parts]), g["tsdf"])
    assert np.array_equal(np.concatenate([p.weight for p in parts]), g["weight"])
    assert np.array_equal(np.concatenate([p.clip_feat for p in parts]), g["clip_feat"])
    assert np.array_equal(np.concatenate([p.labels_one_hot for p in parts]), Hh.golden_labels(g))


def test_oracle_rejects_bad_class_id():
    g = Hh.load_golden("seem_a")
    vol = O.OracleVolume(g["origin"], g["voxel_size"], g["nvox"], g["trunc"], g["feature_dim"])
    seg = np.full_like(g["seg"][:1], 200)  # >= 143: torch one_hot raises in the reference
    with pytest.raises(RuntimeError):
        vol.integrate(g["depth"][:1], g["rgb"][:1], g["pose"][:1], g["K"][:1], g["table"][:1], seg)


def test_query_oracle_matches_reference():
    g = Hh.load_golden("query")
    F, X = g["F"], g["X"]
    assert np.allclose(O.normalize_rows(g["F_raw"]), F, atol=1e-6)
    assert np.allclose(O.run_query(F, X), g["relevance"], rtol=1e-4, atol=1e-6)
    assert np.allclose(O.clip_feature_surgery(F[None], X), g["surgery"], atol=2e-6)
    assert np.allclose(O.clip_feature_surgery_literal(F[None], X), g["surgery"], atol=2e-6)
    assert np.allclose(O.clip_feature_surgery(F[None], X, g["redundant"]), g["surgery_red"], atol=2e-6)
    F0 = F.copy()
    F0[0] = 0
    assert np.allclose(O.clip_feature_surgery(F0[None], X), g["surgery_row0_zero"], atol=2e-6)
    # top-k sets from oracle scores and from reference scores agree
    k = 5
    assert np.array_equal(O.topk_indices(O.clip_feature_surgery(F[None], X)[0], k),
                          O.topk_indices(g["surgery"][0], k))


def test_extract_mesh_by_object_oracle():
    g = Hh.load_golden("query")
    v, f, c = O.extract_mesh_by_object(g["mesh_verts"], g["mesh_faces"], g["mesh_colors"], g["mesh_vidx"], 1)
    assert np.array_equal(v, g["obj1_verts"])
    assert np.array_equal(f, g["obj1_faces"])
    assert np.array_equal(c, g["obj1_colors"])


def test_label_argmax_oracle():
    g = Hh.load_golden("seem_a")
    vol, _ = Hh.replay_oracle(g)
    lab = Hh.golden_labels(g)
    ref = np.where(lab.any(axis=1), lab.argmax(axis=1), -1)   # clip_seem_fusion.py:315-325
    assert np.array_equal(vol.label_argmax(), ref)


# ---- extract_mesh (oracle/mc.py) -------------------------------------------------------------------

def test_mesh_oracle_matches_reference():
    """The reference's extract_mesh (run with oracle marching cubes in place of the absent skimage call)
    against the numpy restatement of everything around that call: bit-exact."""
    from oracle import mc
    m = Hh.load_golden("mesh")
    g = Hh.load_golden("seem_a")
    out = mc.extract_mesh(g["tsdf"], g["weight"], g["rgb_state"], g["clip_feat"], g["nvox"], g["voxel_size"],
                          g["origin"], m["seem_obj"], m["seem_seg_color"])
    for got, key in zip(out, ("seem_verts", "seem_faces", "seem_colors", "seem_feats", "seem_vertex_obj",
                              "seem_vertex_seg")):
        assert np.array_equal(got, m[key]), key
    g = Hh.load_golden("fusion_a")
    out = mc.extract_mesh(g["tsdf"], g["weight"], g["rgb_state"], g["clip_feat"], g["nvox"], g["voxel_size"],
                          g["origin"])
    for got, key in zip(out[:4], ("fusion_verts", "fusion_faces", "fusion_colors", "fusion_feats")):
        assert np.array_equal(got, m[key]), key


def _edge_census(faces):
    from collections import Counter
    c = Counter()
    for a, b, d in faces:
        for e in ((a, b), (b, d), (d, a)):
            c[e] += 1
    return c


def test_marching_cubes_oracle_is_watertight_and_oriented():
    """Properties of the from-scratch case table: closed surface of a sphere is a 2-manifold with outward
    normals and vertices on the iso-surface; random noise (every ambiguous case) has no interior cracks."""
    from oracle import mc
    n = 20
    grid = np.mgrid[0:n, 0:n, 0:n].astype(np.float32)
    centre, radius = 9.3, 6.2
    vol = (np.sqrt(((grid - centre) ** 2).sum(0)) - radius).astype(np.float32)
    verts, faces = mc.marching_cubes_raw(vol)
    c = _edge_census(faces)
    assert max(c.values()) == 1 and all(c[(b, a)] == 1 for (a, b) in c)
    p = verts[faces]
    normals = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    assert ((normals * (p.mean(1) - centre)).sum(1) > 0).all()          # towards positive values
    assert np.abs(np.linalg.norm(verts - centre, axis=1) - radius).max() < 0.03
    assert len(verts) - len(c) // 2 + len(faces) == 2                    # Euler characteristic of a sphere
    rng = np.random.default_rng(3)
    vol = rng.standard_normal((10, 10, 10)).astype(np.float32)
    verts, faces = mc.marching_cubes_raw(vol)
    c = _edge_census(faces)
    assert max(c.values()) == 1
    open_edges = [(a, b) for (a, b) in c if c[(b, a)] != 1]
    on_boundary = lambda i: bool(((verts[i] <= 0) | (verts[i] >= 9)).any())
    assert all(on_boundary(a) and on_boundary(b) for a, b in open_edges)
    assert set(np.unique(mc.N_TRIS)) <= set(range(6)) and mc.N_TRIS[0] == mc.N_TRIS[255] == 0


def test_marching_cubes_oracle_nan_handling():
    """Unobserved voxels are NaN: faces touching an edge with a NaN end are dropped by the reference's filter."""
    from oracle import mc
    vol = np.full((6, 6, 6), np.nan, np.float32)
    vol[1:5, 1:5, 2] = -0.5
    vol[1:5, 1:5, 3] = 0.5
    verts, faces = mc.filter_mesh(*mc.marching_cubes_raw(vol))
    assert len(faces) == 3 * 3 * 2 and len(verts) == 16 and not np.isnan(verts).any()
    assert np.allclose(verts[:, 2], 2.5)


# ---- object labelling (oracle/components.py) --------------------------------------------------------

def test_objects_oracle_matches_reference():
    """The reference's flood_fill_3d (pure Python, run unmodified for the golden) against the scipy restatement."""
    from oracle import components as CC
    g = Hh.load_golden("objects")
    ids, n = CC.label_objects(g["class_grid"])
    assert n == len(g["obj_ids"]) == 99
    assert np.array_equal(ids, g["voxel_obj_ids"])


# ---- scene bounds (oracle/bounds.py) -------------------------------------------------------------------

def test_bounds_oracle_matches_reference():
    """backproject_pcd + percentile bounds of the unmodified reference against the numpy restatement.
    Floating point (K^-1 by LU in torch vs numpy): 1e-5 absolute on the points, identical grid size."""
    from oracle import bounds as B
    g = Hh.load_golden("bounds")
    xyz, _ = B.backproject_samples(g["depth"], g["pose"], g["K"], float(g["max_depth"]))
    assert xyz.shape == g["xyz"].shape and np.abs(xyz - g["xyz"]).max() <= 1e-5
    origin, nvox = B.scene_bounds(xyz, float(g["voxel_size"]), int(g["trunc_vox"]))
    assert np.abs(origin - g["origin"]).max() <= 1e-5 and np.array_equal(nvox, g["nvox"])
