"""Mesh-side interfaces of the path: ``extract_mesh_by_object`` (/root/reference/handy_utils.py:585-611),
``extract_mesh_by_id`` (/root/reference/extract_obj_mesh.py:12-36) and the volumes' ``extract_mesh``
(/root/reference/clip_seem_fusion.py:824-888, clipfusion.py:723-763).
"""
import ctypes

import numpy as np


def extract_mesh_by_object(vertices, faces, colors, vertex_indices, obj_idx):
    """Sub-mesh of the vertices labelled `obj_idx`: faces are kept when all three corners are
    selected and are re-indexed into the compacted vertex list (handy_utils.py:585-611).

    Returns (object_vertices, object_faces, object_colors, mesh) where `mesh` is an open3d
    TriangleMesh when open3d is importable and None otherwise.  The reference re-indexes with a
    Python dict loop over every face corner; here it is one cumulative-sum lookup."""
    vertices, faces, colors = np.asarray(vertices), np.asarray(faces), np.asarray(colors)
    selected = np.asarray(vertex_indices) == obj_idx
    object_indices = np.flatnonzero(selected)
    object_vertices = vertices[object_indices]
    object_colors = colors[object_indices]
    if faces.size:
        keep = selected[faces].all(axis=1)
        remap = np.cumsum(selected) - 1
        object_faces = remap[faces[keep]].astype(faces.dtype, copy=False)
    else:
        object_faces = faces.reshape(0, 3)
    return object_vertices, object_faces, object_colors, _to_open3d(object_vertices, object_faces, object_colors)


def extract_mesh_by_id(vertices, faces, colors, labels, object_id):
    """extract_obj_mesh.py:12-36: same selection, returns only the mesh object (or the arrays when
    open3d is not installed)."""
    v, f, c, mesh = extract_mesh_by_object(vertices, faces, colors, labels, object_id)
    return mesh if mesh is not None else (v, f, c)


def _to_open3d(vertices, faces, colors):
    try:
        import open3d as o3d
    except ImportError:
        return None
    mesh = o3d.geometry.TriangleMesh()
    mesh.vertices = o3d.utility.Vector3dVector(vertices)
    mesh.triangles = o3d.utility.Vector3iVector(faces)
    mesh.vertex_colors = o3d.utility.Vector3dVector(colors)
    return mesh


def _aligned_scratch(nbytes, device):
    import torch
    buf = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
    return buf, (buf.data_ptr() + 255) // 256 * 256


def _halo_ptr(halo, key, n_plane, dtype, device, channels=None):
    """Device pointer (and the tensor keeping it alive) of one halo-plane array, or (None, None)."""
    import torch
    if not halo or halo.get(key) is None:
        return None, None
    t = halo[key].to(device=device, dtype=dtype).contiguous()
    want = n_plane if channels is None else n_plane * channels
    if t.numel() != want:
        raise ValueError("halo[%r] has %d elements, the plane needs %d" % (key, t.numel(), want))
    return t.data_ptr(), t


def marching_cubes_device(volume, halo=None, return_edge_ids=False):
    """Marching cubes (level 0) over the volume's TSDF with unobserved voxels (weight == 0) treated as NaN, faces
    with a NaN vertex and unused vertices already removed (clip_seem_fusion.py:825-842).
    Returns device tensors (verts [V,3] f32 in voxel-index coordinates, verts_world [V,3] f32, faces [F,3] i64).
    halo: for an x-slab that has a successor, {"tsdf": [ny*nz], "weight": [ny*nz]} of the successor's first plane
    (slab.exchange_halo): the cells across the cut are meshed too and the vertices lying in that plane are emitted
    (they duplicate vertices of the successor's own mesh; slab.weld_slab_meshes removes the duplicates).
    return_edge_ids: also return each vertex's global grid-edge id [V] int64 (the identity that welding uses)."""
    import torch
    from . import _lib
    if not volume.tsdf.is_cuda:
        raise RuntimeError("extract_mesh runs on a CUDA (sm_100) device only; there is no CPU path")
    lib, dev = _lib.load(), volume.tsdf.device
    stream = torch.cuda.current_stream(dev).cuda_stream
    grid = volume._grid_desc()
    nbytes = ctypes.c_uint64()
    _lib.check(lib.saf_mesh_workspace_bytes(ctypes.byref(grid), ctypes.byref(nbytes)), "saf_mesh_workspace_bytes")
    scratch, base = _aligned_scratch(nbytes.value, dev)
    nv, nf = ctypes.c_uint64(), ctypes.c_uint64()
    n_plane = volume._dims[1] * volume._dims[2]
    h_tsdf, keep0 = _halo_ptr(halo, "tsdf", n_plane, torch.float32, dev)
    h_weight, keep1 = _halo_ptr(halo, "weight", n_plane, torch.int32, dev)
    with torch.cuda.device(dev):
        _lib.check(lib.saf_mesh_count(ctypes.byref(grid), volume.tsdf.data_ptr(), volume.weight.data_ptr(), h_tsdf, h_weight,
                                  base, nbytes.value, ctypes.byref(nv), ctypes.byref(nf), stream), "saf_mesh_count")
    verts = torch.empty((nv.value, 3), dtype=torch.float32, device=dev)
    verts_world = torch.empty((nv.value, 3), dtype=torch.float32, device=dev)
    faces = torch.empty((nf.value, 3), dtype=torch.int64, device=dev)
    edge_ids = torch.empty(nv.value, dtype=torch.int64, device=dev) if return_edge_ids else None
    if nv.value or nf.value:
        with torch.cuda.device(dev):
            _lib.check(lib.saf_mesh_emit(ctypes.byref(grid), volume.tsdf.data_ptr(), volume.weight.data_ptr(), h_tsdf,
                                     h_weight, base, nbytes.value, verts.data_ptr(), verts_world.data_ptr(),
                                     edge_ids.data_ptr() if return_edge_ids else None, faces.data_ptr(), stream),
                   "saf_mesh_emit")
    del scratch, keep0, keep1
    if return_edge_ids:
        return verts, verts_world, faces, edge_ids
    return verts, verts_world, faces


def sample_vertices(volume, verts, field, mode="bilinear", clamp01=False, halo_field=None):
    """torch.nn.functional.grid_sample of a per-voxel field at mesh vertices, as extract_mesh calls it
    (clip_seem_fusion.py:843-877): field [N] or [N,C] (rows of the volume's slab) -> [V,C] on the device.
    halo_field: the same field on the successor slab's first plane, [ny*nz] or [ny*nz, C]."""
    import torch
    from . import _lib
    dev = volume.tsdf.device
    n = volume.tsdf.shape[0]
    field = field.to(device=dev, dtype=torch.float32).reshape(n, -1).contiguous()
    out = torch.empty((verts.shape[0], field.shape[1]), dtype=torch.float32, device=dev)
    m = {"bilinear": _lib.SAF_SAMPLE_TRILINEAR, "nearest": _lib.SAF_SAMPLE_NEAREST}[mode]
    h_ptr, keep = _halo_ptr({"f": halo_field}, "f", volume._dims[1] * volume._dims[2], torch.float32, dev, field.shape[1])
    with torch.cuda.device(dev):
        _lib.check(_lib.load().saf_mesh_sample(ctypes.byref(volume._grid_desc()), verts.data_ptr(), verts.shape[0],
                                           field.data_ptr(), h_ptr, field.shape[1], m, int(bool(clamp01)),
                                           out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
               "saf_mesh_sample")
    del keep
    return out


def _h(halo, key):
    return None if not halo else halo.get(key)


def extract_mesh_fusion(volume, halo=None, return_edge_ids=False):
    """ClipFusion.extract_mesh (clipfusion.py:723-763): (verts_world, faces, vertex_colors, vertex_clip_feats);
    the first two as numpy arrays, the sampled attributes as device tensors, like the reference.
    halo: for x-slabs, the successor's first plane {"tsdf", "weight", "rgb", "clip_feat"} (slab.exchange_halo).
    return_edge_ids: append the vertices' global grid-edge ids (numpy int64 [V]) for slab.weld_slab_meshes."""
    verts, verts_world, faces, ids = marching_cubes_device(volume, halo, return_edge_ids=True)
    vertex_colors = sample_vertices(volume, verts, volume.rgb, "bilinear", clamp01=True, halo_field=_h(halo, "rgb"))
    vertex_clip_feats = sample_vertices(volume, verts, volume.clip_feat, "bilinear", halo_field=_h(halo, "clip_feat"))
    out = (verts_world.cpu().numpy(), faces.cpu().numpy(), vertex_colors, vertex_clip_feats)
    return out + (ids.cpu().numpy(),) if return_edge_ids else out


def extract_mesh_seem(volume, halo=None, return_edge_ids=False):
    """ClipSeemFusion.extract_mesh (clip_seem_fusion.py:824-888): adds vertex_obj_idx [V,1] and
    vertex_segment_color [V,3], nearest samples of the caller-set `voxel_obj_idx` and
    `objects_segmentation_color` attributes (clip_seem_fusion.py:349-372).
    halo: as in extract_mesh_fusion, plus "voxel_obj_idx" and "objects_segmentation_color" planes."""
    verts, verts_world, faces, ids = marching_cubes_device(volume, halo, return_edge_ids=True)
    vertex_colors = sample_vertices(volume, verts, volume.rgb, "bilinear", clamp01=True, halo_field=_h(halo, "rgb"))
    vertex_clip_feats = sample_vertices(volume, verts, volume.clip_feat, "bilinear", halo_field=_h(halo, "clip_feat"))
    vertex_obj_idx = sample_vertices(volume, verts, volume.voxel_obj_idx, "nearest",
                                     halo_field=_h(halo, "voxel_obj_idx"))
    vertex_segment_color = sample_vertices(volume, verts, volume.objects_segmentation_color, "nearest", clamp01=True,
                                           halo_field=_h(halo, "objects_segmentation_color"))
    out = (verts_world.cpu().numpy(), faces.cpu().numpy(), vertex_colors, vertex_clip_feats, vertex_obj_idx,
           vertex_segment_color)
    return out + (ids.cpu().numpy(),) if return_edge_ids else out
