"""x-slab partitioning of the voxel grid across GPUs (one process per GPU).

The fusion path shards naturally: every voxel's update depends only on the frame, never on other
voxels, so rank r integrates ALL frames into its own slab [x_begin, x_end) and no collective runs in
the fusion loop.  The flat index (x*ny + y)*nz + z makes an x-slab a
contiguous range of every buffer, and the kernels keep global x indices so slabs concatenate to
exactly the single-GPU grid.  torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests)
is used only to combine small results:

  * query: each rank's top-k [T,k] (score, global voxel index) is all-gathered and merged;
  * surgery weights depend on global row 0: its owner broadcasts the T scores of that row;
  * meshes: per-slab vertex/face arrays are gathered to rank 0 (sizes first, then payloads).
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
