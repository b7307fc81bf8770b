"""Mesh-side interfaces of the path: ``extract_mesh_by_object`` (/root/reference/handy_utils.py:585-611),
``extract_mesh_by_id`` (/root/reference/extract_obj_mesh.py:12-36) and the volumes' ``extract_mesh``
(/root/reference/clip_seem_fusion.py:824-888, clipfusion.py:723-763).
"""
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


def extract_mesh_seem(volume):
    raise NotImplementedError("extract_mesh (clip_seem_fusion.py:824-888) is listed under 'next' in DESIGN.md; "
                              "GPU marching cubes has not landed yet")


def extract_mesh_fusion(volume):
    raise NotImplementedError("extract_mesh (clipfusion.py:723-763) is listed under 'next' in DESIGN.md; "
                              "GPU marching cubes has not landed yet")
