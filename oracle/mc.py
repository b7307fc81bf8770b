"""CPU oracle of extract_mesh: marching cubes + the reference's face/vertex filtering and vertex sampling.

TEST INFRASTRUCTURE ONLY (same rule as oracle.py): imported by tests/ and tests/golden/make_golden.py.

Reference lines restated here (numpy, fp32):
  ClipSeemFusion.extract_mesh   /root/reference/clip_seem_fusion.py:824-888
  ClipFusion.extract_mesh       /root/reference/clipfusion.py:723-763
  torch grid_sample, 5-D input, bilinear / nearest, zeros padding, align_corners=False
                                (ATen grid_sampler_3d CPU kernel: unnormalise ((g+1)*size-1)/2,
                                 eight corner weights as products of three differences, taps added
                                 in the order tnw,tne,tsw,tse,bnw,bne,bsw,bse)

Parity status.  The reference delegates the triangulation itself to skimage.measure.marching_cubes
(Lewiner), a third-party dependency that is NOT installed in this image (scikit-image, unpinned in the
reference's environment.yml apart from the conda solve) - so the TRIANGULATION is "parity unpinned":
`marching_cubes_raw` below is a from-scratch marching cubes (case table derived in
tools/gen_mc_tables.py) that places vertices on the same grid edges by the same linear interpolation,
but its face connectivity in ambiguous cells and its vertex/face ORDER are its own.  Everything the
reference itself does around that call (NaN masking of unobserved voxels, dropping faces with NaN
vertices, compacting vertices, trilinear / nearest sampling, world transform) IS pinned: the golden file
tests/golden/mesh.npz was produced by running the unmodified reference extract_mesh with
`marching_cubes_raw` injected in place of the missing skimage function.
"""
import importlib.util
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("gen_mc_tables", os.path.join(os.path.dirname(_HERE), "tools",
                                                                            "gen_mc_tables.py"))
_gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_gen)
N_TRIS, TRI_TABLE, EDGE_CORNERS = _gen.build_tables()


def masked_tsdf(tsdf, weight, nvox):
    """clip_seem_fusion.py:826: NaN where the voxel was never observed (weight == 0)."""
    out = np.asarray(tsdf, np.float32).copy()
    out[np.asarray(weight) == 0] = np.nan
    return out.reshape([int(v) for v in nvox])


def marching_cubes_raw(vol, x_offset=0, return_ids=False):
    """All iso-crossings of `vol` [nx,ny,nz] at level 0, NaN-unaware like the library call the reference makes:
    returns (verts [V,3] f32 in index coordinates - NaN when an endpoint of the edge is NaN -, faces [F,3] i64).
    Order: vertices by (voxel flat index, axis) of their grid edge; faces by cell flat index, then table order.
    x_offset: global x index of the sub-volume's first plane (x-slabs keep global coordinates).
    return_ids: also return each vertex's grid-edge id ((x*ny + y)*nz + z)*3 + axis with the global x."""
    vol = np.asarray(vol, np.float32)
    nx, ny, nz = vol.shape
    with np.errstate(invalid="ignore"):
        inside = ~(vol > 0)                      # NaN counts as not above the level
    flat = np.arange(nx * ny * nz, dtype=np.int64).reshape(nx, ny, nz)
    # vertices: one per grid edge whose end points lie on different sides
    edge_ids, positions = [], []
    for axis in range(3):
        lo = [slice(None)] * 3
        hi = [slice(None)] * 3
        lo[axis], hi[axis] = slice(0, -1), slice(1, None)
        lo, hi = tuple(lo), tuple(hi)
        cross = inside[lo] != inside[hi]
        a, b = vol[lo][cross], vol[hi][cross]
        with np.errstate(invalid="ignore", divide="ignore"):
            t = (a / (a - b)).astype(np.float32)
        idx = np.argwhere(cross)
        idx[:, 0] += x_offset
        idx = idx.astype(np.float32)
        idx[:, axis] = idx[:, axis] + t          # fp32 add, as the kernel does
        edge_ids.append(flat[lo][cross] * 3 + axis)
        positions.append(idx)
    edge_ids = np.concatenate(edge_ids)
    positions = np.concatenate(positions).astype(np.float32)
    order = np.argsort(edge_ids, kind="stable")
    edge_ids, verts = edge_ids[order], positions[order]
    # faces
    global_ids = edge_ids + np.int64(x_offset) * ny * nz * 3
    if min(nx, ny, nz) < 2:
        return (verts, np.zeros((0, 3), np.int64), global_ids) if return_ids else (verts, np.zeros((0, 3), np.int64))
    case = np.zeros((nx - 1, ny - 1, nz - 1), np.int32)
    for c in range(8):
        dx, dy, dz = c & 1, (c >> 1) & 1, (c >> 2) & 1
        case |= inside[dx:nx - 1 + dx, dy:ny - 1 + dy, dz:nz - 1 + dz].astype(np.int32) << c
    cells = np.argwhere(N_TRIS[case] > 0)
    cell_case = case[cells[:, 0], cells[:, 1], cells[:, 2]]
    face_rows, face_keys = [], []
    for s in range(TRI_TABLE.shape[1]):
        m = N_TRIS[cell_case] > s
        if not m.any():
            continue
        cc, cs = cells[m], cell_case[m]
        ids = np.empty((len(cc), 3), np.int64)
        for k in range(3):
            e = TRI_TABLE[cs, s, k]
            axis = e // 4
            base = EDGE_CORNERS[e, 0]
            vx = cc[:, 0] + (base & 1)
            vy = cc[:, 1] + ((base >> 1) & 1)
            vz = cc[:, 2] + ((base >> 2) & 1)
            ids[:, k] = ((vx * ny + vy) * nz + vz) * 3 + axis
        face_rows.append(np.searchsorted(edge_ids, ids))
        face_keys.append(((cc[:, 0] * ny + cc[:, 1]) * nz + cc[:, 2]) * 8 + s)
    if not face_rows:
        return (verts, np.zeros((0, 3), np.int64), global_ids) if return_ids else (verts, np.zeros((0, 3), np.int64))
    faces = np.concatenate(face_rows)
    keys = np.concatenate(face_keys)
    faces = faces[np.argsort(keys, kind="stable")]
    return (verts, faces.astype(np.int64), global_ids) if return_ids else (verts, faces.astype(np.int64))


def filter_mesh(verts, faces, ids=None):
    """clip_seem_fusion.py:832-842: drop faces touching a NaN vertex, then vertices no face uses
    (`ids`, when given, are filtered along with the vertices)."""
    good = ~np.any(np.isnan(verts[faces]), axis=(1, 2))
    faces = faces[good]
    used = np.zeros(len(verts), bool)
    used[np.unique(faces.flatten())] = True
    reindex = np.cumsum(used) - 1
    if ids is not None:
        return verts[used], reindex[faces], ids[used]
    return verts[used], reindex[faces]


def _unnormalize(verts, nvox):
    """grid = (verts + 0.5) / nvox * 2 - 1 (clip_seem_fusion.py:843) - with `verts` a numpy array and `nvox` an
    int32 torch tensor this dispatches to Tensor.__rtruediv__, i.e. reciprocal(nvox) * (verts + 0.5) -, then
    ATen's unnormalisation ((g + 1) * size - 1) / 2; every step rounded to fp32."""
    n = np.asarray(nvox, np.float32)
    r = (np.float32(1) / n).astype(np.float32)
    g = ((r * (verts.astype(np.float32) + np.float32(0.5))).astype(np.float32) * np.float32(2)
         - np.float32(1)).astype(np.float32)
    return (((g + np.float32(1)) * n - np.float32(1)) / np.float32(2)).astype(np.float32)


def sample_trilinear(field, verts, nvox):
    """grid_sample(field.T.view(C,*nvox)[None], grid, bilinear, align_corners=False) -> [V,C].  field [N,C]."""
    nx, ny, nz = (int(v) for v in nvox)
    field = np.asarray(field, np.float32).reshape(nx * ny * nz, -1)
    p = _unnormalize(verts, nvox)                      # columns: D (x), H (y), W (z) source indices
    p0 = np.floor(p)
    i0 = p0.astype(np.int64)
    fr_hi = (p0 + np.float32(1) - p).astype(np.float32)   # (i0 + 1) - p   weight of the lower tap
    fr_lo = (p - p0).astype(np.float32)                   # p - i0         weight of the upper tap
    out = np.zeros((len(verts), field.shape[1]), np.float32)
    # ATen order: t/b = x (D) low/high, n/s = y (H), w/e = z (W); weight = (wz * wy) * wx
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                wz = fr_lo[:, 2] if dz else fr_hi[:, 2]
                wy = fr_lo[:, 1] if dy else fr_hi[:, 1]
                wx = fr_lo[:, 0] if dx else fr_hi[:, 0]
                w = ((wz * wy).astype(np.float32) * wx).astype(np.float32)
                x, y, z = i0[:, 0] + dx, i0[:, 1] + dy, i0[:, 2] + dz
                ok = (x >= 0) & (x < nx) & (y >= 0) & (y < ny) & (z >= 0) & (z < nz)
                idx = ((x * ny + y) * nz + z)[ok]
                out[ok] = (out[ok] + (field[idx] * w[ok, None]).astype(np.float32)).astype(np.float32)
    return out


def sample_nearest(field, verts, nvox):
    """grid_sample(..., mode='nearest', align_corners=False), zeros padding -> [V,C]."""
    nx, ny, nz = (int(v) for v in nvox)
    field = np.asarray(field, np.float32).reshape(nx * ny * nz, -1)
    p = np.rint(_unnormalize(verts, nvox)).astype(np.int64)
    ok = (p[:, 0] >= 0) & (p[:, 0] < nx) & (p[:, 1] >= 0) & (p[:, 1] < ny) & (p[:, 2] >= 0) & (p[:, 2] < nz)
    out = np.zeros((len(verts), field.shape[1]), np.float32)
    idx = ((p[:, 0] * ny + p[:, 1]) * nz + p[:, 2])[ok]
    out[ok] = field[idx]
    return out


def extract_mesh(tsdf, weight, rgb, clip_feat, nvox, voxel_size, origin, voxel_obj_idx=None, seg_color=None):
    """clip_seem_fusion.py:824-888 (6-tuple) / clipfusion.py:723-763 (first four entries)."""
    vol = masked_tsdf(tsdf, weight, nvox)
    verts, faces = filter_mesh(*marching_cubes_raw(vol))
    colors = np.clip(sample_trilinear(rgb, verts, nvox), 0, 1)
    feats = sample_trilinear(clip_feat, verts, nvox)
    obj = None if voxel_obj_idx is None else sample_nearest(np.asarray(voxel_obj_idx, np.float32).reshape(-1, 1),
                                                            verts, nvox)
    seg = None if seg_color is None else np.clip(sample_nearest(seg_color, verts, nvox), 0, 1)
    verts_world = (verts * np.float32(voxel_size) + np.asarray(origin, np.float32)).astype(np.float32)
    return verts_world, faces, colors, feats, obj, seg
