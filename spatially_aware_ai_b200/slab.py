"""x-slab partitioning of the voxel grid across GPUs (one process per GPU).

The fusion path shards naturally: every voxel's update depends only on the frame, never on other
voxels, so rank r integrates ALL frames into its own slab [x_begin, x_end) and no collective runs in
the fusion loop.  The flat index (x*ny + y)*nz + z makes an x-slab a
contiguous range of every buffer, and the kernels keep global x indices so slabs concatenate to
exactly the single-GPU grid.  torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests)
is used only to combine small results:

  * query: each rank's top-k [T,k] (score, global voxel index) is all-gathered and merged;
  * surgery weights depend on global row 0: its owner broadcasts the T scores of that row;
  * meshes: every rank but the last receives its successor's first plane (tsdf, weight and the sampled fields)
    so that it can mesh the cells across the cut; per-slab vertex/face arrays are then gathered to rank 0
    (sizes first, then payloads) and the vertices duplicated on the cut planes are welded.
"""
import numpy as np
import torch
import torch.distributed as dist


def slab_bounds(nx, world_size, rank):
    """Contiguous, balanced split of nx x-planes: returns (x_begin, x_end) of `rank`."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, rem = divmod(int(nx), int(world_size))
    x_begin = rank * base + min(rank, rem)
    return x_begin, x_begin + base + (1 if rank < rem else 0)


def slab_index_base(nvox, world_size, rank):
    """Global flat index of the slab's first voxel."""
    x_begin, _ = slab_bounds(nvox[0], world_size, rank)
    return x_begin * int(nvox[1]) * int(nvox[2])


def cyclic_slab(nx, world_size, rank, span=8):
    """Block-cyclic split: rank r holds the stripes [r*span + k*world*span, ... + span) of the nx
    x-planes.  Returns the constructor keywords of the fusion classes.  A camera sees a compact part of the grid
    per frame; dealing the planes out in stripes gives every rank the same share of every view."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    if rank * span >= nx:
        raise ValueError("grid of %d planes has no stripe for rank %d (span %d)" % (nx, rank, span))
    return dict(x_begin=rank * span, x_end=int(nx), x_span=span, x_stride=span * world_size)


def sheared_slab(world_size, rank):
    """Sheared block-column split: column (bx, by) of 8 x 8 x nz voxels lives on rank (bx + by) % world_size.
    Returns the constructor keywords of the fusion classes.  Unlike x-stripes it also spreads a surface that
    lies in a single x-plane (a wall of an axis-aligned room) over all ranks."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    return dict(y_ranks=world_size, y_rank=rank) if world_size > 1 else {}


def local_to_global_rows(volume, rows):
    """Slab-local flat voxel indices (rows of the volume's buffers; -1 = padding) -> global flat indices
    (x*ny + y)*nz + z of the whole grid, for contiguous, block-cyclic and sheared slabs alike."""
    if getattr(volume, "y_ranks", 0) > 1:
        table = volume.global_rows(rows.device)
        return torch.where(rows < 0, rows, table[rows.clamp(min=0)])
    plane = volume._dims[1] * volume._dims[2]
    xs = torch.as_tensor(volume.global_x_planes(), dtype=torch.int64, device=rows.device)
    lx = torch.div(rows.clamp(min=0), plane, rounding_mode="floor")
    out = xs[lx] * plane + rows.clamp(min=0) % plane
    return torch.where(rows < 0, rows, out)


def redistribute_to_contiguous(volume, group=None):
    """Block-cyclic slabs -> contiguous slabs (rank r gets slab_bounds(nx, world, r)): the volume of every rank is
    replaced by a new one of the same class holding the contiguous slab, filled from the ranks that fused its
    planes (one all_to_all per state buffer, NCCL over NVLink on GPUs).  Fusion balances best on block-cyclic
    slabs; extract_mesh / label_objects want contiguous ones.  Returns the new volume."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nx, ny, nz = volume._dims
    plane = ny * nz
    xs = volume.global_x_planes()
    bounds = [slab_bounds(nx, world, q) for q in range(world)]
    dest_of = lambda x: next(q for q, (a, b) in enumerate(bounds) if a <= x < b)   # noqa: E731
    send_planes = [0] * world
    for x in xs:                       # local planes are in ascending x: the planes for a destination are one run
        send_planes[dest_of(x)] += 1
    # what every rank holds, derived from the layout (no communication): planes of my new range per source rank
    if not volume.x_span:
        raise ValueError("redistribute_to_contiguous expects block-cyclic slabs")

    def planes_of(r):
        out = []
        for start in range(r * volume.x_span, nx, volume.x_stride):
            out.extend(range(start, min(start + volume.x_span, nx)))
        return out
    xb, xe = bounds[rank]
    recv_x = [[x for x in planes_of(r) if xb <= x < xe] for r in range(world)]
    recv_planes = [len(v) for v in recv_x]
    order = torch.as_tensor(np.argsort(np.concatenate([np.asarray(v, np.int64) for v in recv_x]), kind="stable"),
                            device=volume.tsdf.device)
    kw = dict(x_begin=xb, x_end=xe)
    cls = type(volume)
    if hasattr(volume, "segmentation_model"):
        new = cls(volume.origin, volume.voxel_size, volume.nvox, volume.trunc, volume.scale_patches_by_depth,
                  volume.clip_patch_size, volume.clip_patch_stride, volume.clip, volume.segmentation_model, **kw)
    else:
        new = cls(volume.origin, volume.voxel_size, volume.nvox, volume.trunc, volume.scale_patches_by_depth, volume.clip,
                  None, volume.clip_patch_size, volume.clip_patch_stride, **kw)
    new = new.to(volume.tsdf.device)
    new.feature_source = volume.feature_source
    for name in ("tsdf", "weight", "tsdf_weight", "rgb", "clip_feat", "labels_one_hot"):
        src = getattr(volume, name, None)
        if src is None:
            continue
        width = src[0].numel() if src.dim() > 1 else 1
        flat = src.reshape(len(xs), plane * width)
        out = torch.empty((sum(recv_planes), plane * width), dtype=src.dtype, device=src.device)
        dist.all_to_all_single(out, flat.contiguous(), output_split_sizes=recv_planes, input_split_sizes=send_planes,
                               group=group)
        getattr(new, name).copy_(out.index_select(0, order).reshape(getattr(new, name).shape))
        del out
    return new


def merge_topk(scores, indices, k):
    """Merge candidate lists [..., T, n] -> top-k per text, descending score, ties to the lower index;
    entries with index < 0 are padding."""
    scores = scores.reshape(-1, scores.shape[-2], scores.shape[-1]).permute(1, 0, 2).reshape(scores.shape[-2], -1)
    indices = indices.reshape(-1, indices.shape[-2], indices.shape[-1]).permute(1, 0, 2).reshape(indices.shape[-2], -1)
    pad = indices < 0
    s = torch.where(pad, torch.full_like(scores, float("-inf")), scores)
    big = torch.iinfo(torch.int64).max
    i = torch.where(pad, torch.full_like(indices, big), indices)
    # sort by index, then stable sort by descending score: ties keep ascending index order
    order = torch.argsort(i, dim=1, stable=True)
    s, i = torch.gather(s, 1, order), torch.gather(i, 1, order)
    order = torch.argsort(s, dim=1, descending=True, stable=True)
    s, i = torch.gather(s, 1, order)[:, :k], torch.gather(i, 1, order)[:, :k]
    i = torch.where(i == big, torch.full_like(i, -1), i)
    return s, i


def gather_topk(local_scores, local_indices, k, group=None):
    """All ranks receive the global top-k [T,k] built from every rank's local top-k (global indices)."""
    world = dist.get_world_size(group)
    s_all = [torch.empty_like(local_scores) for _ in range(world)]
    i_all = [torch.empty_like(local_indices) for _ in range(world)]
    dist.all_gather(s_all, local_scores.contiguous(), group=group)
    dist.all_gather(i_all, local_indices.contiguous(), group=group)
    return merge_topk(torch.stack(s_all), torch.stack(i_all), k)


def broadcast_row0_scores(local_row0_scores, owner_rank, group=None):
    """clip_feature_surgery weights use row 0 of the GLOBAL feature matrix (clipfusion.py:913-915):
    the rank holding voxel 0 passes its [T] scores, the others pass a same-shaped buffer."""
    buf = local_row0_scores.clone()
    dist.broadcast(buf, src=owner_rank, group=group)
    return buf


def first_plane(volume, extra=()):
    """The arrays of the volume's first x-plane that its predecessor needs for seam cells and vertex sampling."""
    n_plane = volume._dims[1] * volume._dims[2]
    out = {"tsdf": volume.tsdf[:n_plane], "weight": volume.weight[:n_plane], "rgb": volume.rgb[:n_plane],
           "clip_feat": volume.clip_feat[:n_plane]}
    for name in extra:   # e.g. "voxel_obj_idx", "objects_segmentation_color" (clip_seem_fusion.py:349-372)
        t = getattr(volume, name)
        out[name] = t.reshape(-1, *t.shape[3:])[:n_plane] if t.dim() >= 3 else t.reshape(t.shape[0], -1)[:n_plane]
    return out


def exchange_halo(volume, extra=(), group=None):
    """Every rank sends its first plane to rank - 1 and receives rank + 1's: returns the `halo` dict for
    extract_mesh(halo=...), or None on the last rank.  Slabs must be in rank order."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = {k: v.contiguous() for k, v in first_plane(volume, extra).items()}
    halo = {k: torch.empty_like(v) for k, v in mine.items()} if rank + 1 < world else None
    ops = []
    for k in sorted(mine):
        if rank > 0:
            ops.append(dist.P2POp(dist.isend, mine[k], rank - 1, group))
        if halo is not None:
            ops.append(dist.P2POp(dist.irecv, halo[k], rank + 1, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return halo


def weld_slab_meshes(parts):
    """Join per-slab meshes (in slab order) into one: parts[r] = (verts [V,3], faces [F,3] i64, edge_ids [V] i64,
    attrs) with attrs a tuple of [V,c] arrays and edge_ids the vertices' global grid-edge ids
    (mesh.marching_cubes_device(..., return_edge_ids=True)).  A vertex that a slab emitted on its successor's first
    plane is a copy of the successor's vertex with the same edge id when the successor has it: faces are re-pointed
    to the successor's vertex (whose attributes were sampled with its own rows) and the copy is dropped."""
    n = len(parts)
    verts = [np.asarray(p[0], np.float32).reshape(-1, 3) for p in parts]
    faces = [np.asarray(p[1], np.int64).reshape(-1, 3) for p in parts]
    ids = [np.asarray(p[2], np.int64).reshape(-1) for p in parts]
    attrs = [tuple(np.asarray(a) for a in p[3]) for p in parts]
    keep = [np.ones(len(v), bool) for v in verts]
    target = [np.full(len(v), -1, np.int64) for v in verts]   # for dropped copies: the successor's local vertex
    for r in range(n - 1):
        if not len(ids[r]) or not len(ids[r + 1]):
            continue
        order = np.argsort(ids[r + 1], kind="stable")
        pos = np.searchsorted(ids[r + 1], ids[r], sorter=order)
        pos = np.minimum(pos, len(order) - 1)
        hit = ids[r + 1][order[pos]] == ids[r]
        keep[r][hit] = False
        target[r][hit] = order[pos[hit]]
    base, new_index = 0, []
    for r in range(n):
        new_index.append(np.cumsum(keep[r]) - 1 + base)
        base += int(keep[r].sum())
    out_faces = []
    for r in range(n):
        remap = new_index[r].copy()
        dropped = ~keep[r]
        if dropped.any():
            remap[dropped] = new_index[r + 1][target[r][dropped]]
        out_faces.append(remap[faces[r]])
    if not n:
        return np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int64), ()
    out_verts = np.concatenate([verts[r][keep[r]] for r in range(n)])
    out_faces = np.concatenate(out_faces)
    out_attrs = tuple(np.concatenate([attrs[r][k][keep[r]] for r in range(n)]) for k in range(len(attrs[0])))
    # canonical order of the single-grid mesh: vertices by ascending edge id (a copy kept because its successor
    # does not use that vertex sits at the end of its slab's list, but belongs among the successor's first plane)
    all_ids = np.concatenate([ids[r][keep[r]] for r in range(n)])
    order = np.argsort(all_ids, kind="stable")
    inverse = np.empty_like(order)
    inverse[order] = np.arange(len(order))
    return out_verts[order], inverse[out_faces], tuple(a[order] for a in out_attrs)


def extract_mesh_distributed(volume, extra=(), dst=0, group=None):
    """extract_mesh of a grid sharded into x-slabs (one volume per rank, slabs in rank order): exchanges the first
    planes, lets every rank mesh its slab plus the cells across its cut, gathers the pieces on `dst` and welds
    them.  Returns on `dst` the tuple extract_mesh() returns for the whole grid (numpy arrays; vertices in the
    single-grid order), None elsewhere.  `extra`: per-voxel attributes set by the caller that the volume's
    extract_mesh samples too ("voxel_obj_idx", "objects_segmentation_color" for ClipSeemFusion)."""
    rank = dist.get_rank(group)
    halo = exchange_halo(volume, extra, group)
    out = volume.extract_mesh(halo=halo, return_edge_ids=True)
    dev = volume.tsdf.device
    # every array travels as a device tensor through gather_rows (sizes first, then padded payloads: NCCL over
    # NVLink on GPUs) - per-vertex features are [V, C] floats, far too much for a pickled object gather
    arrays = [torch.as_tensor(np.asarray(out[0], np.float32)), torch.as_tensor(np.asarray(out[1], np.int64)),
              torch.as_tensor(np.asarray(out[-1], np.int64))] + [a.detach() for a in out[2:-1]]
    gathered = [gather_rows(a.to(dev), dst=dst, group=group) for a in arrays]
    if rank != dst:
        return None
    world = dist.get_world_size(group)
    pieces = [(gathered[0][r].cpu().numpy(), gathered[1][r].cpu().numpy(), gathered[2][r].cpu().numpy(),
               tuple(g[r].cpu().numpy() for g in gathered[3:])) for r in range(world)]
    wv, wf, wattrs = weld_slab_meshes(pieces)
    return (wv, wf) + tuple(wattrs)


def gather_rows(t, dst=0, group=None):
    """Variable-length gather of one tensor per rank ([n_r, ...], same trailing shape and dtype everywhere) to rank
    `dst`: the row counts are all-gathered, the payloads padded to the longest and gathered as tensors on the
    device the group communicates on.  Returns the list of per-rank tensors on `dst`, None elsewhere."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    counts = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    nmax = max(1, max(counts))
    pad = torch.zeros((nmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return [bufs[r][: counts[r]] for r in range(world)]


def gather_mesh(verts, faces, dst=0, group=None):
    """Variable-length gather of per-slab meshes (numpy [V,3] float32, [F,3] int64 with slab-local vertex
    ids) to rank `dst`, which gets the concatenation with face indices re-based; others get None."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    sizes = torch.tensor([len(verts), len(faces)], dtype=torch.int64, device=dev)
    all_sizes = [torch.empty_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = torch.stack(all_sizes).cpu().numpy()
    vmax, fmax = int(all_sizes[:, 0].max()), int(all_sizes[:, 1].max())
    vbuf = torch.zeros((max(vmax, 1), 3), dtype=torch.float32, device=dev)
    fbuf = torch.zeros((max(fmax, 1), 3), dtype=torch.int64, device=dev)
    if len(verts):
        vbuf[: len(verts)] = torch.as_tensor(np.asarray(verts, np.float32)).to(dev)
    if len(faces):
        fbuf[: len(faces)] = torch.as_tensor(np.asarray(faces, np.int64)).to(dev)
    vall = [torch.empty_like(vbuf) for _ in range(world)] if rank == dst else None
    fall = [torch.empty_like(fbuf) for _ in range(world)] if rank == dst else None
    dist.gather(vbuf, vall, dst=dst, group=group)
    dist.gather(fbuf, fall, dst=dst, group=group)
    if rank != dst:
        return None
    out_v, out_f, base = [], [], 0
    for r in range(world):
        nv, nf = int(all_sizes[r, 0]), int(all_sizes[r, 1])
        out_v.append(vall[r][:nv].cpu().numpy())
        out_f.append(fall[r][:nf].cpu().numpy() + base)
        base += nv
    return np.concatenate(out_v), np.concatenate(out_f)
