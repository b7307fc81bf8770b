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
