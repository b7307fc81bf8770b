"""Grid files and full-state checkpoints.

The reference persists only what its viewers need - ``voxel_rgb.npy`` ([nx,ny,nz,3]) and ``voxel_clip_feats.npy``
([nx,ny,nz,C]) via ``np.save`` (/root/reference/clip_seem_fusion.py:563-571, reloaded at :204-221) - and re-fuses a
scan from scratch for every new version.  `save_state` writes those two files in the same format and, next to
them, everything else integrate() carries (tsdf, weight, tsdf_weight, label histogram, geometry), so that a later
scan version can be integrated INTO the existing grid (`load_state` + more integrate() calls: the v00 -> v01
incremental re-fusion of BASELINE config 5).  Arrays are streamed through a bounded pinned buffer: a 24 M-voxel
x 768-d feature grid is 74 GB and must not be staged whole in host memory.
"""
import json
import os

import numpy as np
import torch

FORMAT_VERSION = 1
_CHUNK_BYTES = 256 << 20

# file name -> (buffer attribute, trailing shape taken from the buffer)
_STATE_FILES = {
    "voxel_rgb.npy": "rgb",                    # reference format (clip_seem_fusion.py:567)
    "voxel_clip_feats.npy": "clip_feat",       # reference format (clip_seem_fusion.py:568-571)
    "voxel_tsdf.npy": "tsdf",
    "voxel_weight.npy": "weight",
    "voxel_tsdf_weight.npy": "tsdf_weight",
    "voxel_labels_one_hot.npy": "labels_one_hot",
}


def _grid_shape(volume, tensor):
    nxs = len(volume.global_x_planes())
    return (nxs, volume.ny_local, volume._dims[2]) + tuple(tensor.shape[1:])


def _stream_out(tensor, path, shape):
    """tensor (any device) -> .npy file of `shape`, in bounded chunks."""
    flat = tensor.detach().reshape(-1)
    out = np.lib.format.open_memmap(path, mode="w+", dtype=np.dtype(str(flat.dtype).replace("torch.", "")), shape=shape)
    dst = torch.from_numpy(out.reshape(-1))
    step = max(1, _CHUNK_BYTES // flat.element_size())
    stage = torch.empty(min(step, flat.numel()), dtype=flat.dtype).pin_memory() if flat.is_cuda else None
    for lo in range(0, flat.numel(), step):
        hi = min(flat.numel(), lo + step)
        if stage is not None:
            stage[: hi - lo].copy_(flat[lo:hi])
            dst[lo:hi] = stage[: hi - lo]
        else:
            dst[lo:hi] = flat[lo:hi]
    out.flush()
    del out


def _stream_in(path, tensor):
    src = np.load(path, mmap_mode="r")
    flat = tensor.detach().reshape(-1)
    if src.size != flat.numel():
        raise ValueError("%s holds %d elements, the volume buffer %d" % (path, src.size, flat.numel()))
    if str(src.dtype) != str(flat.dtype).replace("torch.", ""):
        raise ValueError("%s is %s, the volume buffer %s" % (path, src.dtype, flat.dtype))
    src = src.reshape(-1)
    step = max(1, _CHUNK_BYTES // flat.element_size())
    for lo in range(0, flat.numel(), step):
        hi = min(flat.numel(), lo + step)
        flat[lo:hi].copy_(torch.from_numpy(np.array(src[lo:hi])))


def save_state(volume, directory):
    """Write the volume's grid files (reference formats) and the rest of its state into `directory`."""
    os.makedirs(directory, exist_ok=True)
    written = []
    for name, attr in _STATE_FILES.items():
        t = getattr(volume, attr, None)
        if t is None:
            continue
        t2 = t if t.dim() > 1 else t[:, None]
        shape = _grid_shape(volume, t2) if t.dim() > 1 else _grid_shape(volume, t2)[:3]
        _stream_out(t, os.path.join(directory, name), shape)
        written.append(name)
    origin = volume.origin.detach().cpu().tolist() if isinstance(volume.origin, torch.Tensor) else list(volume.origin)
    meta = dict(format_version=FORMAT_VERSION, cls=type(volume).__name__, origin=[float(v) for v in origin],
                voxel_size=float(volume.voxel_size), nvox=[int(v) for v in volume._dims], trunc=float(volume.trunc),
                x_begin=int(volume.x_begin), x_end=int(volume.x_end), x_span=int(volume.x_span),
                x_stride=int(volume.x_stride), y_ranks=int(volume.y_ranks), y_rank=int(volume.y_rank),
                feature_dim=int(volume.n_clip_feats), files=written, stats=volume.stats(check=False) if volume.tsdf.is_cuda else None)
    with open(os.path.join(directory, "volume_state.json"), "w") as f:
        json.dump(meta, f, indent=1, default=str)
    return meta


def load_state(volume, directory):
    """Fill an already constructed volume (same geometry, feature dim and slab) from `directory`."""
    with open(os.path.join(directory, "volume_state.json")) as f:
        meta = json.load(f)
    if meta.get("format_version") != FORMAT_VERSION:
        raise ValueError("unsupported checkpoint format %r" % meta.get("format_version"))
    mine = dict(nvox=[int(v) for v in volume._dims], x_begin=int(volume.x_begin), x_end=int(volume.x_end),
                x_span=int(volume.x_span), x_stride=int(volume.x_stride), y_ranks=int(volume.y_ranks),
                y_rank=int(volume.y_rank), feature_dim=int(volume.n_clip_feats))
    for key, val in mine.items():
        if meta.get(key, 0) != val:
            raise ValueError("checkpoint %s = %r does not match the volume's %r" % (key, meta[key], val))
    if abs(meta["voxel_size"] - float(volume.voxel_size)) > 1e-9 * max(1.0, abs(meta["voxel_size"])):
        raise ValueError("checkpoint voxel_size %r does not match the volume's %r" % (meta["voxel_size"], volume.voxel_size))
    for name in meta["files"]:
        attr = _STATE_FILES[name]
        t = getattr(volume, attr, None)
        if t is None:
            continue   # e.g. a ClipFusion volume loading a ClipSeemFusion checkpoint: no label histogram to fill
        _stream_in(os.path.join(directory, name), t)
    return meta


# ---- mesh and scene files (clip_seem_fusion.py:576-604) -------------------------------------------------------

def save_mesh_ply(path, verts, faces, vertex_colors=None):
    """Binary little-endian PLY with per-vertex RGBA, the layout ``trimesh.Trimesh(...).export("x.ply")`` writes for
    mesh_rgb / mesh_segmentation (clip_seem_fusion.py:576-595): float32 x y z, uint8 red green blue alpha, faces
    as (uint8 3, int32 x 3).  Colours are floats in [0,1] ([V,3] or [V,4]) or uint8."""
    verts = np.asarray(verts, np.float32).reshape(-1, 3)
    faces = np.asarray(faces, np.int64).reshape(-1, 3)
    header = ["ply", "format binary_little_endian 1.0", "comment spatially_aware_ai_b200",
              "element vertex %d" % len(verts), "property float x", "property float y", "property float z"]
    vdtype = [("x", "<f4"), ("y", "<f4"), ("z", "<f4")]
    rgba = None
    if vertex_colors is not None:
        c = vertex_colors.detach().cpu().numpy() if hasattr(vertex_colors, "detach") else np.asarray(vertex_colors)
        c = c.reshape(len(verts), -1) if len(verts) else c.reshape(0, 3)
        if c.dtype != np.uint8:
            c = np.rint(np.clip(c.astype(np.float64), 0, 1) * 255).astype(np.uint8)
        rgba = np.full((len(verts), 4), 255, np.uint8)
        rgba[:, : c.shape[1]] = c[:, :4]
        header += ["property uchar red", "property uchar green", "property uchar blue", "property uchar alpha"]
        vdtype += [("red", "u1"), ("green", "u1"), ("blue", "u1"), ("alpha", "u1")]
    header += ["element face %d" % len(faces), "property list uchar int vertex_indices", "end_header"]
    vrec = np.empty(len(verts), dtype=vdtype)
    vrec["x"], vrec["y"], vrec["z"] = verts[:, 0], verts[:, 1], verts[:, 2]
    if rgba is not None:
        vrec["red"], vrec["green"], vrec["blue"], vrec["alpha"] = rgba.T
    frec = np.empty(len(faces), dtype=[("n", "u1"), ("v", "<i4", (3,))])
    frec["n"] = 3
    frec["v"] = faces.astype(np.int32)
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(vrec.tobytes())
        f.write(frec.tobytes())


def load_mesh_ply(path):
    """Reads the files save_mesh_ply writes (and trimesh's binary PLY export of a coloured triangle mesh):
    returns (verts [V,3] f32, faces [F,3] i64, colors [V,4] u8 or None)."""
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    lines = data[:end].decode("ascii").splitlines()
    if "format binary_little_endian 1.0" not in lines:
        raise ValueError("only binary little-endian PLY is supported")
    nv = nf = 0
    props, element = {"vertex": [], "face": []}, None
    for ln in lines:
        tok = ln.split()
        if tok[:1] == ["element"]:
            element = tok[1]
            if element == "vertex":
                nv = int(tok[2])
            elif element == "face":
                nf = int(tok[2])
        elif tok[:1] == ["property"] and element in props:
            props[element].append(tok[1:])
    types = {"float": "<f4", "uchar": "u1", "int": "<i4", "uint": "<u4", "double": "<f8"}
    vdtype = np.dtype([(p[-1], types[p[0]]) for p in props["vertex"]])
    vrec = np.frombuffer(data, dtype=vdtype, count=nv, offset=end)
    lst = props["face"][0]
    if lst[0] != "list":
        raise ValueError("unsupported face element")
    fdtype = np.dtype([("n", types[lst[1]]), ("v", types[lst[2]], (3,))])
    frec = np.frombuffer(data, dtype=fdtype, count=nf, offset=end + nv * vdtype.itemsize)
    if nf and not (frec["n"] == 3).all():
        raise ValueError("only triangle meshes are supported")
    verts = np.stack([vrec["x"], vrec["y"], vrec["z"]], axis=1).astype(np.float32)
    colors = None
    if "red" in vrec.dtype.names:
        alpha = vrec["alpha"] if "alpha" in vrec.dtype.names else np.full(nv, 255, np.uint8)
        colors = np.stack([vrec["red"], vrec["green"], vrec["blue"], alpha], axis=1).astype(np.uint8)
    return verts, frec["v"].astype(np.int64), colors


def save_scene_knowledge(path, scene_knowledge):
    """clip_seem_fusion.py:597-598: ``json.dump(scene_knowledge, f, default=str)``."""
    with open(path, "w") as f:
        json.dump(scene_knowledge, f, default=str)
