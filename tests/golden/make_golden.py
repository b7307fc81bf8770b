"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on torch-CPU.

Run in the build container only (the reference checkout does not travel):
    python tests/golden/make_golden.py
Each file holds the inputs and the reference's outputs for one scenario; tests compare the
oracle (tests/test_oracle_golden.py, CPU) and the CUDA path (tests/test_parity_gpu.py) to them.

Regime note.  The reference delegates `R^T @ (xyz - t)` to torch.bmm, i.e. MKL sgemm with K=3.
MKL picks different kernels (different fp32 summation orders) depending on thread count and
N: with >= 2 threads and N >= ~50 000 voxels it is the left-to-right fma chain (also what a GPU
computes); single-threaded or tiny N it is (a0 x0 + a2 x2) + a1 x1 without fma.  Real scenes
(0.2 M - 24 M voxels, multi-core) are in the first regime, so every golden grid has >= 60 000
voxels and this script asserts the regime it ran in.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import _ref_loader  # noqa: E402
from spatially_aware_ai_b200 import synth  # noqa: E402

clipfusion, clip_seem_fusion, handy_utils = _ref_loader.load_reference()


class FakeClip(torch.nn.Module):
    """Duck-typed stand-in for clipfusion.Clip: returns the pre-generated feature image."""

    def __init__(self, feature_dim):
        super().__init__()
        self.feature_dim = feature_dim
        self.next_table = None

    def img_inference_tiled(self, rgb_imgs, patch_size, patch_stride):
        return self.next_table


class FakeSeg:
    def __init__(self):
        self.queue = []

    def run_on_image(self, img):
        return self.queue.pop(0)


def assert_fma_regime(n_voxels):
    torch.manual_seed(1)
    A = torch.randn(1, 3, 3)
    X = (torch.randn(n_voxels, 3)[None] - torch.randn(1, 1, 3)).transpose(1, 2)
    Y = (A.transpose(1, 2) @ X)[0].numpy()
    At, Xn = A.transpose(1, 2)[0].numpy().astype(np.float64), X[0].numpy().astype(np.float64)
    acc = (At[:, 0:1] * Xn[0]).astype(np.float32).astype(np.float64)
    acc = (At[:, 1:2] * Xn[1] + acc).astype(np.float32).astype(np.float64)
    acc = (At[:, 2:3] * Xn[2] + acc).astype(np.float32)
    mism = int((acc != Y).sum())
    assert mism <= 2, "torch.bmm is not in the fma-chain regime here (%d mismatches)" % mism


SCENES = {
    # name: (SceneConfig kwargs, class, batch, n_calls, special)
    "seem_a": dict(cfg=dict(extent=(2.2, 2.0, 1.6), voxel_size=0.05, height=48, width=64, patch_size=32,
                            patch_stride=16, feature_dim=8, frames=8, missing_fraction=0.05, seg_block=8,
                            seed=11, name="seem_a"),
                   cls="ClipSeemFusion", batch=1, calls=8),
    "fusion_a": dict(cfg=dict(extent=(2.2, 2.0, 1.6), voxel_size=0.05, height=48, width=64, patch_size=32,
                              patch_stride=16, feature_dim=8, frames=6, missing_fraction=0.05, seg_block=8,
                              seed=12, name="fusion_a"),
                     cls="ClipFusion", batch=1, calls=6),
    "fusion_b2": dict(cfg=dict(extent=(2.2, 2.0, 1.6), voxel_size=0.05, height=48, width=64, patch_size=32,
                               patch_stride=16, feature_dim=8, frames=6, missing_fraction=0.05, seg_block=8,
                               seed=13, name="fusion_b2"),
                      cls="ClipFusion", batch=2, calls=3),
    # odd image size (not a power of two), denser patch grid, trunc 3 voxels, plus degenerate frames
    "seem_edge": dict(cfg=dict(extent=(2.4, 1.8, 1.5), voxel_size=0.05, trunc_vox=3, height=44, width=60,
                               patch_size=20, patch_stride=8, feature_dim=12, frames=7, missing_fraction=0.1,
                               seg_block=4, seed=14, name="seem_edge"),
                      cls="ClipSeemFusion", batch=1, calls=7, edge=True),
}


def scene_frames(spec):
    cfg = synth.SceneConfig(**spec["cfg"])
    rng = np.random.default_rng(cfg.seed + 1000)
    frames = []
    for i in range(cfg.frames):
        fr = synth.make_frame(cfg, i, table_layout="hwc" if i % 3 == 2 else "chw")
        if i % 2 == 1:
            fr["pose"] = synth.perturbed_pose(cfg, i, rng)
            fr["depth"] = synth.render_depth(cfg, fr["pose"], fr["K"])
            fr["depth"][rng.random(fr["depth"].shape) < cfg.missing_fraction] = 0.0
        if spec.get("edge"):
            if i == 2:      # all depth missing: sdf = -z/trunc, only voxels within trunc of the camera
                fr["depth"][:] = 0.0
            elif i == 3:    # camera far outside the grid looking away: nothing valid
                fr["pose"][:3, 3] += np.array([30.0, 0.0, 0.0], np.float32)
            elif i == 4:    # surface far beyond the grid: tsdf_valid everywhere in view, valid nowhere
                fr["depth"][:] = 50.0
            elif i == 5:    # anisotropic / off-centre intrinsics
                fr["K"] = np.array([[70.0, 0.0, 20.0], [0.0, 40.0, 30.0], [0.0, 0.0, 1.0]], np.float32)
            elif i == 6:    # camera inside the grid volume's margin, steep roll
                fr["pose"] = synth.perturbed_pose(cfg, i, rng)
        frames.append(fr)
    return cfg, frames


def run_scene(name, spec):
    cfg, frames = scene_frames(spec)
    origin, nvox = cfg.grid()
    assert int(np.prod(nvox)) >= 60000
    assert_fma_regime(int(np.prod(nvox)))
    fake_clip = FakeClip(cfg.feature_dim)
    fake_seg = FakeSeg()
    t_origin, t_nvox = torch.from_numpy(origin), torch.from_numpy(nvox)
    if spec["cls"] == "ClipSeemFusion":
        vol = clip_seem_fusion.ClipSeemFusion(t_origin, cfg.voxel_size, t_nvox, cfg.trunc, False,
                                              cfg.patch_size, cfg.patch_stride, fake_clip, fake_seg)
    else:
        real_clip = clipfusion.Clip
        clipfusion.Clip = lambda model, pretraining: fake_clip
        try:
            vol = clipfusion.ClipFusion(t_origin, cfg.voxel_size, t_nvox, cfg.trunc, False, "fake", "fake",
                                        cfg.patch_size, cfg.patch_stride)
        finally:
            clipfusion.Clip = real_clip
    B = spec["batch"]
    counts = []
    mid_state = None
    for call in range(spec["calls"]):
        batch = frames[call * B:(call + 1) * B]
        fake_clip.next_table = torch.stack([torch.from_numpy(f["table"]) for f in batch])
        fake_seg.queue = [torch.from_numpy(f["seg"].astype(np.int64)) for f in batch]
        w0, tw0 = vol.weight.clone(), vol.tsdf_weight.clone()
        vol.integrate(torch.stack([torch.from_numpy(f["depth"]) for f in batch]),
                      torch.stack([torch.from_numpy(f["rgb"]) for f in batch]),
                      torch.stack([torch.from_numpy(f["pose"]) for f in batch]),
                      torch.stack([torch.from_numpy(f["K"]) for f in batch]))
        counts.append([int((vol.weight - w0).sum()), int((vol.tsdf_weight - tw0).sum())])
        if call == spec["calls"] // 2 - 1:
            mid_state = dict(mid_tsdf=vol.tsdf.numpy().copy(), mid_weight=vol.weight.numpy().copy(),
                             mid_tsdf_weight=vol.tsdf_weight.numpy().copy())
    out = dict(
        origin=origin, nvox=nvox, voxel_size=np.float64(cfg.voxel_size), trunc=np.float64(cfg.trunc),
        feature_dim=np.int64(cfg.feature_dim), batch=np.int64(B), cls=np.array(spec["cls"]),
        depth=np.stack([f["depth"] for f in frames]), rgb=np.stack([f["rgb"] for f in frames]),
        seg=np.stack([f["seg"] for f in frames]),
        table=np.stack([np.ascontiguousarray(f["table"]) for f in frames]),
        table_hwc=np.array([not f["table"].flags["C_CONTIGUOUS"] for f in frames]),
        pose=np.stack([f["pose"] for f in frames]), K=np.stack([f["K"] for f in frames]),
        counts=np.array(counts, np.int64),   # per call: sum of weight / tsdf_weight increments
        tsdf=vol.tsdf.numpy(), weight=vol.weight.numpy(), tsdf_weight=vol.tsdf_weight.numpy(),
        rgb_state=vol.rgb.numpy(), clip_feat=vol.clip_feat.numpy(),
        xyz_world_sha256=np.array(hashlib.sha256(vol.xyz_world.numpy().tobytes()).hexdigest()),
        torch_version=np.array(torch.__version__), torch_threads=np.int64(torch.get_num_threads()),
    )
    out.update(mid_state)
    if spec["cls"] == "ClipSeemFusion":
        lab = vol.labels_one_hot.numpy()
        nz = np.flatnonzero(lab)
        out["labels_nz_index"] = nz.astype(np.int64)
        out["labels_nz_value"] = lab.reshape(-1)[nz].astype(np.int32)
        out["n_classes"] = np.int64(lab.shape[1])
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "voxels", int(np.prod(nvox)), "counts", counts, "->", os.path.getsize(path) // 1024, "KiB")


def run_query_golden():
    rng = np.random.default_rng(77)
    M, C, T = 257, 32, 7
    F = rng.standard_normal((M, C)).astype(np.float32)
    F[5] = 0.0                      # an unobserved voxel: zero feature row
    Fn = torch.from_numpy(F)
    Fn = Fn / Fn.norm(dim=-1, keepdim=True)
    Fn = torch.nan_to_num(Fn)       # clip_seem_fusion.py:507-511
    X = rng.standard_normal((T, C)).astype(np.float32)
    X /= np.linalg.norm(X, axis=-1, keepdims=True)
    Xt = torch.from_numpy(X)

    class _Self:
        def text_inference(self, labels):
            return Xt

    relevance = clipfusion.Clip.run_query(_Self(), Fn, ["x"] * T)
    surgery = clipfusion.Clip.clip_feature_surgery(Fn[None], Xt)
    red = torch.from_numpy(rng.standard_normal((1, C)).astype(np.float32))
    surgery_red = clipfusion.Clip.clip_feature_surgery(Fn[None], Xt, redundant_feats=red)
    # row 0 all-zero variant (w == 1): clipfusion.py:913-915 reads row 0 only
    F0 = Fn.clone()
    F0[0] = 0
    surgery_row0_zero = clipfusion.Clip.clip_feature_surgery(F0[None], Xt)

    # extract_mesh_by_object (handy_utils.py:585-611) on a small random mesh
    V, Fc = 60, 90
    verts = rng.standard_normal((V, 3)).astype(np.float32)
    faces = rng.integers(0, V, size=(Fc, 3)).astype(np.int64)
    colors = rng.random((V, 3)).astype(np.float32)
    vidx = rng.integers(0, 3, size=V).astype(np.int64)
    import open3d as o3d  # stub module from _ref_loader

    class _Mesh:
        pass

    o3d.geometry = type("g", (), {"TriangleMesh": _Mesh})
    o3d.utility = type("u", (), {"Vector3dVector": staticmethod(lambda a: a), "Vector3iVector": staticmethod(lambda a: a)})
    handy_utils.o3d = o3d
    ov, of, oc, _ = handy_utils.extract_mesh_by_object(verts, faces.copy(), colors, vidx, 1)

    path = os.path.join(HERE, "query.npz")
    np.savez_compressed(path, F_raw=F, F=Fn.numpy(), X=X, relevance=relevance.numpy(), surgery=surgery.numpy(),
                        redundant=red.numpy(), surgery_red=surgery_red.numpy(),
                        surgery_row0_zero=surgery_row0_zero.numpy(),
                        mesh_verts=verts, mesh_faces=faces, mesh_colors=colors, mesh_vidx=vidx,
                        obj1_verts=ov, obj1_faces=of, obj1_colors=oc)
    print("query ->", os.path.getsize(path) // 1024, "KiB")


def run_mesh_golden():
    """extract_mesh of both reference classes on the final state of the seem_a / fusion_a goldens.
    skimage is not installed here, so the one library call inside extract_mesh
    (skimage.measure.marching_cubes, clip_seem_fusion.py:829) is served by oracle/mc.py's
    marching_cubes_raw; every other line is the unmodified reference running on torch-CPU."""
    from oracle import mc

    def fake_marching_cubes(volume, level=0):
        assert level == 0
        verts, faces = mc.marching_cubes_raw(volume)
        return verts, faces, None, None

    out = {}
    rng = np.random.default_rng(91)
    for name, mod in (("seem_a", clip_seem_fusion), ("fusion_a", clipfusion)):
        with np.load(os.path.join(HERE, name + ".npz")) as z:
            g = {k: z[k] for k in z.files}
        cfgC = int(g["feature_dim"])
        fake_clip = FakeClip(cfgC)
        t_origin, t_nvox = torch.from_numpy(g["origin"]), torch.from_numpy(g["nvox"])
        if name == "seem_a":
            vol = mod.ClipSeemFusion(t_origin, float(g["voxel_size"]), t_nvox, float(g["trunc"]), False, 0, 0,
                                     fake_clip, FakeSeg())
        else:
            real_clip = clipfusion.Clip
            clipfusion.Clip = lambda model, pretraining: fake_clip
            try:
                vol = mod.ClipFusion(t_origin, float(g["voxel_size"]), t_nvox, float(g["trunc"]), False, "fake",
                                     "fake", 0, 0)
            finally:
                clipfusion.Clip = real_clip
        vol.tsdf.copy_(torch.from_numpy(g["tsdf"]))
        vol.weight.copy_(torch.from_numpy(g["weight"]))
        vol.rgb.copy_(torch.from_numpy(g["rgb_state"]))
        vol.clip_feat.copy_(torch.from_numpy(g["clip_feat"]))
        mod.skimage.measure.marching_cubes = fake_marching_cubes
        if name == "seem_a":
            n = int(np.prod(g["nvox"]))
            obj = rng.integers(-1, 9, size=tuple(int(v) for v in g["nvox"])).astype(np.int64)
            seg_color = rng.random((n, 3)).astype(np.float32) * 1.2 - 0.1     # exercises the clamp
            vol.voxel_obj_idx = torch.from_numpy(obj)
            vol.objects_segmentation_color = torch.from_numpy(seg_color)
            verts, faces, colors, feats, vobj, vseg = vol.extract_mesh()
            out.update(seem_obj=obj, seem_seg_color=seg_color, seem_vertex_obj=vobj.numpy(),
                       seem_vertex_seg=vseg.numpy())
        else:
            verts, faces, colors, feats = vol.extract_mesh()
        key = name.split("_")[0]
        out.update({key + "_verts": np.asarray(verts), key + "_faces": np.asarray(faces),
                    key + "_colors": colors.numpy(), key + "_feats": feats.numpy()})
        print(name, "mesh:", np.asarray(verts).shape, np.asarray(faces).shape, np.asarray(verts).dtype,
              np.asarray(faces).dtype, colors.dtype, feats.dtype)
    path = os.path.join(HERE, "mesh.npz")
    np.savez_compressed(path, **out)
    print("mesh ->", os.path.getsize(path) // 1024, "KiB")


def run_objects_golden():
    """flood_fill_3d (handy_utils.py:295-480) on a random class grid with an untrained in-situ model."""
    rng = np.random.default_rng(123)
    shape = (18, 15, 13)
    grid = rng.choice(np.array([-1, 133, 0, 5, 17, 60]), size=shape, p=[0.45, 0.15, 0.1, 0.1, 0.1, 0.1]).astype(np.int64)
    grid[2:6, 3:7, 4:8] = 60           # one solid object
    grid[10:12, :, 6] = 133            # a null-class sheet cutting through

    class _Model:
        model_trained = False
        labels = ["null"]

    know, ids = handy_utils.flood_fill_3d(grid, None, None, None, _Model(), None)
    keys = list(know["unique_objects"].keys())
    out = dict(class_grid=grid, voxel_obj_ids=ids.astype(np.int32), obj_ids=np.array(keys),
               obj_class_id=np.array([know["unique_objects"][k]["class_id"] for k in keys], np.int64),
               obj_index=np.array([know["unique_objects"][k]["object_index"] for k in keys], np.int64),
               obj_size=np.array([len(know["unique_objects"][k]["voxels"]) for k in keys], np.int64),
               count_keys=np.array(list(know["object_counts"].keys())),
               count_vals=np.array(list(know["object_counts"].values()), np.int64),
               class_names=np.array(handy_utils.predefined_classes))
    path = os.path.join(HERE, "objects.npz")
    np.savez_compressed(path, **out)
    print("objects:", len(keys), "objects ->", os.path.getsize(path) // 1024, "KiB")


def run_bounds_golden():
    """backproject_pcd (clipfusion.py:510-572) + the bounds rule of clip_seem_fusion.py:276-288."""
    cfg = synth.SceneConfig(extent=(2.2, 2.0, 1.6), voxel_size=0.05, height=48, width=64, patch_size=32,
                            patch_stride=16, feature_dim=8, frames=12, missing_fraction=0.1, seed=33, name="bounds")
    frames = [synth.make_frame(cfg, i) for i in range(cfg.frames)]
    frames[3]["depth"][:, :20] = np.nan
    frames[5]["depth"][:] *= 3.0          # beyond max_depth

    class _Dataset(torch.utils.data.Dataset):
        imwidth, imheight = cfg.width, cfg.height

        def __len__(self):
            return len(frames)

        def __getitem__(self, i):
            f = frames[i]
            return (torch.from_numpy(f["rgb"]), torch.from_numpy(f["depth"]), torch.from_numpy(f["pose"]),
                    torch.from_numpy(f["K"]), i)

    max_depth = 4
    xyz, rgb = clipfusion.backproject_pcd(_Dataset(), batch_size=1, num_workers=0, device="cpu", max_depth=max_depth)
    trunc_m = 3 * cfg.voxel_size
    minbound = torch.tensor(np.percentile(xyz.cpu(), 1, axis=0)).float() - trunc_m
    maxbound = torch.tensor(np.percentile(xyz.cpu(), 99, axis=0)).float() + trunc_m
    nvox = ((maxbound - minbound) / cfg.voxel_size).round().int()
    path = os.path.join(HERE, "bounds.npz")
    np.savez_compressed(path, depth=np.stack([f["depth"] for f in frames]), pose=np.stack([f["pose"] for f in frames]),
                        K=np.stack([f["K"] for f in frames]), max_depth=np.float64(max_depth), xyz=xyz.numpy(),
                        origin=minbound.numpy(), nvox=nvox.numpy(), voxel_size=np.float64(cfg.voxel_size),
                        trunc_vox=np.int64(3))
    print("bounds:", tuple(xyz.shape), "origin", minbound.tolist(), "nvox", nvox.tolist())


def run_query_drivers_golden():
    """The drivers around the scorers, executed from the unmodified reference where it offers a callable:
      * InSituManager.clip_text_query (clip_seem_fusion.py:482-561), called unbound on a stand-in `self`: row
        normalisation + nan_to_num, clip_feature_surgery, mean-subtract / clip / min-max relevance;
      * segment() (eval_scannet_segmentation.py:546-561): clamp_min(0.1) normalisation, argsort of softmax(100 cos).
    query_mesh.py and hypersim_eval.py are scripts without functions: their expressions (query_mesh.py:24-25, 39,
    59-73; hypersim_eval.py:50-51, 80-89) are evaluated here with torch exactly as written there."""
    import importlib
    import tempfile
    import types
    rng = np.random.default_rng(2024)
    M, C, T = 400, 24, 9
    feats = (rng.standard_normal((M, C)) * rng.uniform(0.01, 3.0, size=(M, 1))).astype(np.float32)
    feats[0] = rng.standard_normal(C).astype(np.float32)
    feats[7] = 0.0           # unobserved vertex -> NaN after the division -> 0
    feats[11] *= 0.001       # norm below clamp_min's 0.1
    X = rng.standard_normal((T, C)).astype(np.float32)
    X /= np.linalg.norm(X, axis=-1, keepdims=True)
    Xt = torch.from_numpy(X)
    out = dict(feats=feats, text=X)

    # -- clip_text_query, unbound
    class _ClipModel:
        clip_feature_surgery = staticmethod(clipfusion.Clip.clip_feature_surgery)

        def encode_text_with_prompt_ensemble(self, texts, device, prompt_templates=None):
            assert prompt_templates == ["a photo of {}"]
            return Xt[: len(texts)]

    clip_seem_fusion.plt = types.SimpleNamespace(cm=types.SimpleNamespace(
        turbo=lambda r: np.stack([r, r, r, np.ones_like(r)], axis=1)))
    names = ["obj%d" % i for i in range(T - 1)]
    me = types.SimpleNamespace(control_objects=list(names), control_text_features=None, clip_model=_ClipModel(),
                               vert_clip_feat=feats.copy(), verts=[], faces=[], scene_knowledge=None)
    mesh_json = clip_seem_fusion.InSituManager.clip_text_query(me, "the query")
    colors = np.asarray(mesh_json["colors"], dtype=np.float64)
    out["text_query_relevance"] = colors[:, 0].astype(np.float32)       # turbo stand-in passes the relevance through
    out["text_query_alpha"] = colors[:, 3].astype(np.float32)           # relevance * 0.5
    out["text_query_column"] = np.int64(T - 1)

    # -- segment()
    ess = importlib.import_module("eval_scannet_segmentation")

    class _ClipSeg:
        def text_inference(self, prompts):
            return Xt

    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "vertex_clip_feats.npy")
        seg_feats = feats.copy()
        seg_feats[7] = rng.standard_normal(C).astype(np.float32) * 1e-3   # segment() raises on NaN rows
        np.save(path, seg_feats)
        out["segment_feats"] = seg_feats
        out["segment_labels"] = ess.segment(_ClipSeg(), path, ["p"] * T).numpy().astype(np.int64)

    # -- query_mesh.py:24-25, 39 (run_query relevance of the last of five labels) and :59-73 (surgery + outliers)
    F = torch.from_numpy(seg_feats.copy())
    F /= F.norm(dim=-1, keepdim=True)

    class _Self:
        def text_inference(self, labels):
            return Xt[:5]

    relevance = clipfusion.Clip.run_query(_Self(), F, ["x"] * 5)[:, -1]
    out["query_mesh_half"] = ((relevance - 0.5) * 2).clamp(0, 1).numpy()
    similarity = clipfusion.Clip.clip_feature_surgery(F[None], Xt)
    similarity = (similarity - similarity.min(1, keepdim=True)[0]) / (
        similarity.max(1, keepdim=True)[0] - similarity.min(1, keepdim=True)[0])
    out["query_mesh_minmax"] = similarity[0].numpy()
    clipped = []
    for n in range(T):
        rel = similarity[0, :, n]
        median, std = torch.median(rel), torch.std(rel)
        clipped.append(torch.where(rel > median + 2 * std, rel, torch.zeros_like(rel)))
    out["query_mesh_outliers"] = torch.stack(clipped, dim=1).numpy()

    # -- hypersim_eval.py:50-51, 80-89: 4 background texts + one target at a time, presence = max relevance
    G = torch.from_numpy(feats.copy())
    G /= torch.clamp_min(G.norm(dim=-1, keepdim=True), 0.1)
    bg, targets = Xt[:4], Xt[4:]
    thresholds = torch.linspace(0, 1, 101)
    pres, preds = [], []
    for i in range(len(targets)):
        tf = torch.cat((bg, targets[i, None]), dim=0)
        rel = (100 * (G @ tf.T)).softmax(dim=-1)[..., -1]
        pres.append(rel.max())
        preds.append(rel.max() > thresholds)
    out["hypersim_presence"] = torch.stack(pres).numpy()
    out["hypersim_preds"] = torch.stack(preds).numpy()
    path = os.path.join(HERE, "query_drivers.npz")
    np.savez_compressed(path, **out)
    print("query_drivers ->", os.path.getsize(path) // 1024, "KiB")


def run_tiled_golden():
    """Clip.get_patches / Clip.img_inference_tiled (clipfusion.py:789-839) of the unmodified reference with a
    deterministic stand-in for open_clip's encode_image (a fixed projection of the 224x224 tile pooled to 4x4)."""
    rng = np.random.default_rng(606)
    C = 16
    proj = torch.from_numpy(rng.standard_normal((3 * 16, C)).astype(np.float32))

    class _Encoder:
        @staticmethod
        def encode_image(x):
            pooled = torch.nn.functional.adaptive_avg_pool2d(x, 4).reshape(len(x), -1)
            return pooled @ proj

    clip = clipfusion.Clip.__new__(clipfusion.Clip)
    torch.nn.Module.__init__(clip)
    clip.clip = _Encoder()
    clip.feature_dim = C
    clip.channel_mean = torch.nn.Parameter(torch.tensor([0.48145466, 0.4578275, 0.40821073])[None, :, None, None],
                                           requires_grad=False)
    clip.channel_std = torch.nn.Parameter(torch.tensor([0.26862954, 0.26130258, 0.27577711])[None, :, None, None],
                                          requires_grad=False)
    rgb = torch.from_numpy(rng.random((2, 3, 96, 128)).astype(np.float32))
    patches = clip.get_patches(rgb, 64, 32)
    feat_img = clip.img_inference_tiled(rgb, 64, 32)
    path = os.path.join(HERE, "tiled.npz")
    np.savez_compressed(path, rgb=rgb.numpy(), proj=proj.numpy(), patch_size=np.int64(64), patch_stride=np.int64(32),
                        patches_shape=np.array(patches.shape), patches_sha256=np.array(
                            hashlib.sha256(np.ascontiguousarray(patches.numpy()).tobytes()).hexdigest()),
                        patch_1_2=patches[1, 0, 2].numpy(), feat_img=feat_img.numpy(),
                        feat_img_strides=np.array(feat_img.stride()))
    print("tiled:", tuple(patches.shape), tuple(feat_img.shape), "->", os.path.getsize(path) // 1024, "KiB")


def run_sensor_golden():
    """ScanNetDataset.__getitem__ (clipfusion.py:243-256) of the unmodified reference on a two-frame scan directory
    written here: 640x480 JPEG colour (decoded to uint8 by cv2) and 16-bit PNG depth in millimetres.  The depth
    image holds every uint16 value and the colour image every uint8 value, so the golden is the complete table of
    the dataset's conversions rgb = float(u8) / 255 and depth = float(u16) / 1000."""
    import tempfile

    import cv2
    rng = np.random.default_rng(31337)
    with tempfile.TemporaryDirectory() as d:
        for sub in ("color", "depth", "pose", "intrinsic"):
            os.makedirs(os.path.join(d, sub))
        np.savetxt(os.path.join(d, "intrinsic", "intrinsic_depth.txt"), np.eye(4))
        depth_in, rgb_in = [], []
        for i in range(2):
            dimg = rng.permutation(np.arange(640 * 480, dtype=np.int64) % 65536).astype(np.uint16).reshape(480, 640)
            cimg = rng.integers(0, 256, size=(480, 640, 3), dtype=np.uint8)
            cimg[0, :256, 0] = np.arange(256)
            cv2.imwrite(os.path.join(d, "depth", "%d.png" % i), dimg)
            cv2.imwrite(os.path.join(d, "color", "%d.jpg" % i), cimg)
            pose = np.eye(4)
            pose[0, 3] = i          # 1 m apart: both frames are key frames (clipfusion.py:226-233)
            np.savetxt(os.path.join(d, "pose", "%d.txt" % i), pose)
            depth_in.append(cv2.imread(os.path.join(d, "depth", "%d.png" % i), cv2.IMREAD_ANYDEPTH))
            rgb_in.append(cv2.cvtColor(cv2.imread(os.path.join(d, "color", "%d.jpg" % i)), cv2.COLOR_BGR2RGB))
        ds = clipfusion.ScanNetDataset(d)
        assert len(ds) == 2
        depth_lut = np.full(65536, np.nan, np.float32)
        rgb_lut = np.full(256, np.nan, np.float32)
        for i in range(2):
            rgb_f, depth_f, _, _, _ = ds[i]
            assert depth_in[i].dtype == np.uint16 and rgb_in[i].dtype == np.uint8
            depth_lut[depth_in[i].reshape(-1)] = depth_f.numpy().reshape(-1)
            rgb_lut[rgb_in[i].reshape(-1)] = rgb_f.numpy().reshape(-1)
            # consistency: one output value per input value
            assert np.array_equal(depth_lut[depth_in[i]], depth_f.numpy())
            assert np.array_equal(rgb_lut[rgb_in[i]], rgb_f.numpy())
    assert not np.isnan(depth_lut).any() and not np.isnan(rgb_lut).any()
    path = os.path.join(HERE, "sensor.npz")
    np.savez_compressed(path, depth_lut=depth_lut, rgb_lut=rgb_lut)
    print("sensor ->", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    torch.set_num_threads(8)
    only = sys.argv[1:]
    for name, spec in SCENES.items():
        if not only or name in only:
            run_scene(name, spec)
    if not only or "query" in only:
        run_query_golden()
    if not only or "mesh" in only:
        run_mesh_golden()
    if not only or "objects" in only:
        run_objects_golden()
    if not only or "bounds" in only:
        run_bounds_golden()
    if not only or "query_drivers" in only:
        run_query_drivers_golden()
    if not only or "tiled" in only:
        run_tiled_golden()
    if not only or "sensor" in only:
        run_sensor_golden()
