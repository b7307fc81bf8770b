"""Shared helpers for the parity tests: golden loading and oracle replay."""
import os

import numpy as np

from oracle import oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FUSION_GOLDENS = ["seem_a", "fusion_a", "fusion_b2", "seem_edge"]


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        g = {k: z[k] for k in z.files}
    if "cls" in g:
        g["cls"] = str(g["cls"])
        g["batch"] = int(g["batch"])
        g["feature_dim"] = int(g["feature_dim"])
        g["voxel_size"] = float(g["voxel_size"])
        g["trunc"] = float(g["trunc"])
    return g


def golden_table(g, i):
    """Frame i's feature image [C,npy,npx] in the memory layout the generator used."""
    t = g["table"][i]
    if g["table_hwc"][i]:
        t = np.ascontiguousarray(t.transpose(1, 2, 0)).transpose(2, 0, 1)
    return t


def golden_labels(g):
    n = int(np.prod(g["nvox"]))
    lab = np.zeros(n * int(g["n_classes"]), np.int32)
    lab[g["labels_nz_index"]] = g["labels_nz_value"]
    return lab.reshape(n, int(g["n_classes"]))


def golden_calls(g):
    """Yield per-integrate-call batches (depth, rgb, seg, table list, pose, K)."""
    B = g["batch"]
    n_calls = len(g["counts"])
    for c in range(n_calls):
        sl = slice(c * B, (c + 1) * B)
        tables = [golden_table(g, i) for i in range(c * B, (c + 1) * B)]
        yield g["depth"][sl], g["rgb"][sl], g["seg"][sl], tables, g["pose"][sl], g["K"][sl]


def replay_oracle(g, x_begin=0, x_end=None, upto=None, num_threads=2):
    with_labels = g["cls"] == "ClipSeemFusion"
    vol = O.OracleVolume(g["origin"], g["voxel_size"], g["nvox"], g["trunc"], g["feature_dim"],
                         with_labels=with_labels, x_begin=x_begin, x_end=x_end, num_threads=num_threads)
    counts = []
    for c, (depth, rgb, seg, tables, pose, K) in enumerate(golden_calls(g)):
        if upto is not None and c >= upto:
            break
        cnt = vol.integrate(depth, rgb, pose, K, np.stack(tables), seg if with_labels else None)
        counts.append(cnt.sum(axis=0))
    return vol, np.array(counts)


def cosine_rows(a, b):
    num = (a.astype(np.float64) * b.astype(np.float64)).sum(-1)
    den = np.linalg.norm(a.astype(np.float64), axis=-1) * np.linalg.norm(b.astype(np.float64), axis=-1)
    out = np.ones_like(num)
    nzm = den > 0
    out[nzm] = num[nzm] / den[nzm]
    return out


# ---- GPU-side replay through the drop-in classes ------------------------------------------------

from spatially_aware_ai_b200.synth import FakeClip, FakeSeg  # noqa: E402,F401


def make_gpu_volume(g, device="cuda", x_begin=0, x_end=None, **slab_kw):
    import torch
    import spatially_aware_ai_b200 as saf
    clip, seg = FakeClip(g["feature_dim"]), FakeSeg()
    origin, nvox = torch.from_numpy(g["origin"]), torch.from_numpy(g["nvox"])
    if g["cls"] == "ClipSeemFusion":
        vol = saf.ClipSeemFusion(origin, g["voxel_size"], nvox, g["trunc"], False, 0, 0, clip, seg,
                                 x_begin=x_begin, x_end=x_end, **slab_kw)
    else:
        vol = saf.ClipFusion(origin, g["voxel_size"], nvox, g["trunc"], False, clip, None, 0, 0,
                             x_begin=x_begin, x_end=x_end, **slab_kw)
    return vol.to(device), clip, seg


def replay_gpu(g, device="cuda", x_begin=0, x_end=None, upto=None, seg_dtype=None, pose_on_device=True):
    import torch
    vol, clip, seg = make_gpu_volume(g, device, x_begin, x_end)
    counts = []
    for c, (depth, rgb, segs, tables, pose, K) in enumerate(golden_calls(g)):
        if upto is not None and c >= upto:
            break
        clip.next_table = torch.stack([torch.from_numpy(np.ascontiguousarray(t)) for t in tables]).to(device)
        # keep the generator's memory layout (hwc-backed views) per frame when the batch is 1
        if len(tables) == 1 and not tables[0].flags["C_CONTIGUOUS"]:
            t = torch.from_numpy(np.ascontiguousarray(tables[0].transpose(1, 2, 0))).to(device)
            clip.next_table = t.permute(2, 0, 1)[None]
        dt = seg_dtype or torch.int64
        seg.queue = [torch.from_numpy(s.astype(np.int64)).to(device=device, dtype=dt) for s in segs]
        p, k = torch.from_numpy(pose), torch.from_numpy(K)
        if pose_on_device:
            p, k = p.to(device), k.to(device)
        vol.integrate(torch.from_numpy(depth).to(device), torch.from_numpy(rgb).to(device), p, k)
        st = vol.stats()
        B = len(tables)
        counts.append([sum(st["last_valid"][:B]), sum(st["last_tsdf_valid"][:B])])
    return vol, np.array(counts)


def replay_gpu_sequence(g, device="cuda", x_begin=0, x_end=None):
    """All frames of a batch-1 golden through ONE integrate_sequence call (window mode)."""
    import torch
    assert g["batch"] == 1
    vol, clip, seg = make_gpu_volume(g, device, x_begin, x_end)
    n = len(g["counts"])
    clip.next_table = torch.stack([torch.from_numpy(np.ascontiguousarray(golden_table(g, i))) for i in range(n)]).to(device)
    seg.queue = [torch.from_numpy(g["seg"][i].astype(np.int64)).to(device) for i in range(n)]
    vol.integrate_sequence(torch.from_numpy(g["depth"][:n]).to(device), torch.from_numpy(g["rgb"][:n]).to(device),
                           torch.from_numpy(g["pose"][:n]).to(device), torch.from_numpy(g["K"][:n]).to(device))
    return vol
