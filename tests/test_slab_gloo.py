"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU (x-slab) path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from spatially_aware_ai_b200 import slab


def test_slab_bounds_cover_grid():
    for nx in (1, 7, 304, 400):
        for world in (1, 2, 3, 8):
            if world > nx:
                continue
            cuts = [slab.slab_bounds(nx, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == nx
            assert all(a[1] == b[0] for a, b in zip(cuts[:-1], cuts[1:]))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        slab.slab_bounds(10, 2, 2)


def test_merge_topk_rule():
    s = torch.tensor([[[0.5, 0.5, 0.1]], [[0.5, 0.9, float("-inf")]]])       # [ranks=2, T=1, n=3]
    i = torch.tensor([[[7, 3, 9]], [[1, 20, -1]]])
    ms, mi = slab.merge_topk(s, i, 4)
    assert mi.tolist() == [[20, 1, 3, 7]]
    assert np.allclose(ms.numpy(), [[0.9, 0.5, 0.5, 0.5]])
    ms, mi = slab.merge_topk(s, i, 6)
    assert mi.tolist() == [[20, 1, 3, 7, 9, -1]]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, F, X, k, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        M = F.shape[0]
        a, b = slab.slab_bounds(M, world, rank)          # rows stand in for voxels of an x-slab
        local = O.normalize_rows(F[a:b]) @ X.T
        li = O.topk_indices(local, min(k, b - a)) + a    # global indices
        ls = np.take_along_axis(local.T, li - a, axis=1)
        pad = k - li.shape[1]
        if pad:
            li = np.pad(li, ((0, 0), (0, pad)), constant_values=-1)
            ls = np.pad(ls, ((0, 0), (0, pad)), constant_values=-np.inf)
        gs, gi = slab.gather_topk(torch.from_numpy(ls.astype(np.float32)), torch.from_numpy(li), k)
        row0 = torch.from_numpy(local[0].copy()) if rank == 0 else torch.zeros(X.shape[0])
        row0 = slab.broadcast_row0_scores(row0, 0)
        verts = np.full((rank + 2, 3), float(rank), np.float32)
        faces = np.arange(3 * (rank + 1), dtype=np.int64).reshape(-1, 3) % (rank + 2)
        mesh = slab.gather_mesh(verts, faces, dst=0)
        np.savez(os.path.join(out_dir, "r%d.npz" % rank), gs=gs.numpy(), gi=gi.numpy(), row0=row0.numpy(),
                 mv=mesh[0] if mesh else np.zeros(0), mf=mesh[1] if mesh else np.zeros(0))
    finally:
        dist.destroy_process_group()


def test_gather_topk_row0_and_mesh_world2(tmp_path):
    rng = np.random.default_rng(3)
    M, C, T, k = 301, 16, 5, 9
    F = rng.standard_normal((M, C)).astype(np.float32)
    F[40] = F[250]      # a tie across the two slabs
    X = rng.standard_normal((T, C)).astype(np.float32)
    mp.spawn(_worker, args=(2, _free_port(), F, X, k, str(tmp_path)), nprocs=2, join=True)
    full = O.normalize_rows(F) @ X.T
    want = O.topk_indices(full, k)
    for r in range(2):
        z = np.load(tmp_path / ("r%d.npz" % r))
        assert np.array_equal(z["gi"], want)
        assert np.allclose(z["gs"], np.take_along_axis(full.T, want, axis=1), atol=1e-6)
        assert np.allclose(z["row0"], full[0], atol=1e-6)
    z0 = np.load(tmp_path / "r0.npz")
    assert z0["mv"].shape == (2 + 3, 3) and z0["mf"].shape == (1 + 2, 3)
    assert z0["mf"][1:].min() >= 2      # rank 1's faces re-based past rank 0's vertices


# ---- seam cells for slab meshes: halo exchange (gloo) and welding (numpy) --------------------------------------

def _mesh_sets(verts, faces):
    """Order-independent description of a mesh: the set of vertex positions and of faces as position triples
    (rotation-normalised so that the winding still counts)."""
    vset = {tuple(v) for v in verts.tolist()}
    fset = set()
    for f in faces:
        tri = [tuple(verts[i].tolist()) for i in f]
        k = tri.index(min(tri))
        fset.add(tuple(tri[k:] + tri[:k]))
    return vset, fset


def test_weld_slab_meshes_equals_full_grid_mesh():
    """Three slabs meshed on their own cells plus the cells across each cut (halo plane), welded: the same
    vertices and faces as the mesh of the whole grid (oracle marching cubes on both sides)."""
    from oracle import mc
    rng = np.random.default_rng(21)
    nx, ny, nz = 17, 9, 8
    vol = rng.uniform(-1, 1, (nx, ny, nz)).astype(np.float32)
    vol[rng.random(vol.shape) < 0.25] = np.nan
    full_v, full_f = mc.filter_mesh(*mc.marching_cubes_raw(vol))
    cuts = [0, 5, 11, nx]
    vol[6, 3, 2] = vol[11, 4, 4] = 0.0                        # exact zeros: coincident vertices of different edges
    full_v, full_f = mc.filter_mesh(*mc.marching_cubes_raw(vol))
    parts = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        hi = min(b + 1, nx)                                   # the successor's first plane, when there is one
        v, f, ids = mc.filter_mesh(*mc.marching_cubes_raw(vol[a:hi], x_offset=a, return_ids=True))
        parts.append((v, f, ids, (v[:, :1].copy(),)))         # one attribute: the vertex's own x
    assert sum(len(p[0]) for p in parts) > len(full_v)        # the cut planes' vertices exist twice before welding
    v, f, (attr,) = slab.weld_slab_meshes(parts)
    assert len(v) == len(full_v) and len(f) == len(full_f)
    assert _mesh_sets(v, f) == _mesh_sets(full_v, full_f)
    assert np.array_equal(v, full_v) and np.array_equal(f, full_f)      # even the order is the single-grid one
    assert np.array_equal(attr[:, 0], v[:, 0])


def _halo_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import spatially_aware_ai_b200 as saf
        from tests.helpers import FakeClip, FakeSeg
        nvox = (6, 4, 3)
        a, b = slab.slab_bounds(nvox[0], world, rank)
        vol = saf.ClipSeemFusion(torch.zeros(3), 0.5, torch.tensor(nvox), 1.0, False, 0, 0, FakeClip(5), FakeSeg(),
                                 x_begin=a, x_end=b)
        n = (b - a) * 12
        base = a * 12
        vol.tsdf.copy_(torch.arange(base, base + n, dtype=torch.float32))
        vol.weight.copy_(torch.arange(base, base + n, dtype=torch.int32))
        vol.rgb.copy_(torch.arange(base * 3, (base + n) * 3, dtype=torch.float32).view(n, 3))
        vol.clip_feat.copy_(torch.arange(base * 5, (base + n) * 5, dtype=torch.float32).view(n, 5))
        vol.voxel_obj_idx = torch.arange(base, base + n, dtype=torch.int32).view(b - a, 4, 3)
        halo = slab.exchange_halo(vol, extra=("voxel_obj_idx",))
        out = {"rank": np.array([rank])}
        if halo is not None:
            out.update({k: v.numpy() for k, v in halo.items()})
        np.savez(os.path.join(out_dir, "h%d.npz" % rank), **out)
    finally:
        dist.destroy_process_group()


def test_exchange_halo_world2(tmp_path):
    mp.spawn(_halo_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    h0, h1 = np.load(tmp_path / "h0.npz"), np.load(tmp_path / "h1.npz")
    first = 3 * 12                                            # rank 1's first voxel
    assert np.array_equal(h0["tsdf"], np.arange(first, first + 12, dtype=np.float32))
    assert np.array_equal(h0["weight"], np.arange(first, first + 12, dtype=np.int32))
    assert np.array_equal(h0["rgb"], np.arange(first * 3, (first + 12) * 3, dtype=np.float32).reshape(12, 3))
    assert np.array_equal(h0["clip_feat"], np.arange(first * 5, (first + 12) * 5, dtype=np.float32).reshape(12, 5))
    assert np.array_equal(h0["voxel_obj_idx"].reshape(-1), np.arange(first, first + 12, dtype=np.int32))
    assert set(h1.files) == {"rank"}                                    # the last rank has no successor


def _worker_cyclic(rank, world, port, F, X, k, nx, plane, out_dir):
    """Block-cyclic rows: rank-local top-k -> global voxel ids -> merged over the ranks; plus the bench's
    interleaved frame sharing."""
    import types
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        kw = slab.cyclic_slab(nx, world, rank)
        xs = []
        for start in range(kw["x_begin"], kw["x_end"], kw["x_stride"]):
            xs.extend(range(start, min(start + kw["x_span"], kw["x_end"])))
        vol = types.SimpleNamespace(_dims=[nx, plane, 1], global_x_planes=lambda: xs)
        rows = (np.array(xs)[:, None] * plane + np.arange(plane)[None]).reshape(-1)
        local = O.normalize_rows(F[rows]) @ X.T
        li = O.topk_indices(local, k)
        ls = np.take_along_axis(local.T, li, axis=1)
        gi_local = slab.local_to_global_rows(vol, torch.from_numpy(li))
        gs, gi = slab.gather_topk(torch.from_numpy(ls.astype(np.float32)), gi_local, k)
        # interleaved sharing: rank r holds items r, r + world, ...
        n_total, per = 11, (11 + world - 1) // world
        items = torch.arange(n_total * 6, dtype=torch.int16).reshape(n_total, 2, 3)
        mine = torch.stack([items[min(n_total - 1, j * world + rank)] for j in range(per)])
        whole = bench.share_interleaved(mine, n_total, world, dist.all_gather_into_tensor)
        np.savez(os.path.join(out_dir, "c%d.npz" % rank), gs=gs.numpy(), gi=gi.numpy(), whole=whole.numpy())
    finally:
        dist.destroy_process_group()


def test_cyclic_topk_and_frame_sharing_world2(tmp_path):
    rng = np.random.default_rng(8)
    nx, plane, C, T, k = 44, 7, 12, 4, 6
    F = rng.standard_normal((nx * plane, C)).astype(np.float32)
    X = rng.standard_normal((T, C)).astype(np.float32)
    mp.spawn(_worker_cyclic, args=(2, _free_port(), F, X, k, nx, plane, str(tmp_path)), nprocs=2, join=True)
    full = O.normalize_rows(F) @ X.T
    want = O.topk_indices(full, k)
    for r in range(2):
        z = np.load(tmp_path / ("c%d.npz" % r))
        assert np.array_equal(z["gi"], want)
        assert np.allclose(z["gs"], np.take_along_axis(full.T, want, axis=1), atol=1e-6)
        assert np.array_equal(z["whole"], np.arange(11 * 6, dtype=np.int16).reshape(11, 2, 3))


def _worker_redistribute(rank, world, port, nx, ny, nz, C, out_dir):
    import spatially_aware_ai_b200 as saf
    from spatially_aware_ai_b200 import synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        vol = saf.ClipSeemFusion(torch.zeros(3), 0.05, torch.tensor([nx, ny, nz]), 0.1, False, 0, 0, synth.FakeClip(C),
                                 synth.FakeSeg(), **slab.cyclic_slab(nx, world, rank))
        plane = ny * nz
        gid = (torch.tensor(vol.global_x_planes())[:, None] * plane + torch.arange(plane)[None]).reshape(-1)
        vol.tsdf.copy_(gid.float() * 0.5)
        vol.weight.copy_(gid.int())
        vol.tsdf_weight.copy_(gid.int() + 7)
        vol.rgb.copy_(gid.float()[:, None] + torch.tensor([0.0, 0.25, 0.5]))
        vol.clip_feat.copy_(gid.float()[:, None] * 2 + torch.arange(C))
        vol.labels_one_hot.copy_((gid[:, None] + torch.arange(vol.n_classes)).int())
        new = slab.redistribute_to_contiguous(vol)
        xb, xe = slab.slab_bounds(nx, world, rank)
        assert (new.x_begin, new.x_end, new.x_span) == (xb, xe, 0)
        want = torch.arange(xb * plane, xe * plane)
        assert torch.equal(new.weight, want.int()) and torch.equal(new.tsdf, want.float() * 0.5)
        assert torch.equal(new.tsdf_weight, want.int() + 7)
        assert torch.equal(new.rgb, want.float()[:, None] + torch.tensor([0.0, 0.25, 0.5]))
        assert torch.equal(new.clip_feat, want.float()[:, None] * 2 + torch.arange(C))
        assert torch.equal(new.labels_one_hot, (want[:, None] + torch.arange(vol.n_classes)).int())
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_redistribute_cyclic_to_contiguous(tmp_path, world):
    mp.spawn(_worker_redistribute, args=(world, _free_port(), 44, 5, 3, 6, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))
