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
    nxs = volume.x_end - volume.x_begin
    return (nxs, volume._dims[1], volume._dims[2]) + tuple(tensor.shape[1:])


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
                x_begin=int(volume.x_begin), x_end=int(volume.x_end), feature_dim=int(volume.n_clip_feats),
                files=written, stats=volume.stats(check=False) if volume.tsdf.is_cuda else None)
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
                feature_dim=int(volume.n_clip_feats))
    for key, val in mine.items():
        if meta[key] != val:
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
