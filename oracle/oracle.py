"""numpy front end of the CPU oracle (oracle/saf_oracle.c) plus the query-path restatement.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

Parity status: pinned against the unmodified reference executed on torch-CPU in the build
container (tests/golden/*.npz, generator tests/golden/make_golden.py).

Reference lines restated here (numpy, fp32):
  Clip.run_query                      /root/reference/clipfusion.py:899-904
  Clip.clip_feature_surgery           /root/reference/clipfusion.py:906-934
  query normalisation/post-processing /root/reference/clip_seem_fusion.py:507-533,
                                      query_mesh.py:24-25,39,59-73,
                                      eval_scannet_segmentation.py:546-561,
                                      hypersim_eval.py:50-51,80-89
  extract_mesh_by_object              /root/reference/handy_utils.py:585-611
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libsaf_oracle.so")
_lib = None

N_CLASSES = 133 + 10  # clip_seem_fusion.py:655


def build(force=False):
    src = os.path.join(_HERE, "saf_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.saf_oracle_integrate.restype = ctypes.c_int
        _lib.saf_oracle_integrate_ex.restype = ctypes.c_int
        _lib.saf_oracle_label_argmax.restype = None
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def depth_from_mm(depth_u16):
    """clipfusion.py:187-188 / 253-254 / 357-362: depth = float32(uint16 millimetres) / 1000 (true fp32 division)."""
    return np.asarray(depth_u16, np.uint16).astype(np.float32) / np.float32(1000)


def rgb_from_u8(rgb_u8):
    """clipfusion.py:185 / 245 / 355: rgb = float32(uint8) / 255 (true fp32 division)."""
    return np.asarray(rgb_u8, np.uint8).astype(np.float32) / np.float32(255)


def check_sensor_conversions():
    """Mismatches of the kernels' division-free conversion against the true divisions, over all inputs."""
    lib().saf_oracle_check_sensor_conversions.restype = ctypes.c_int
    return int(lib().saf_oracle_check_sensor_conversions())


class OracleVolume:
    """State + integrate() of ClipSeemFusion (with_labels=True, bilinear rgb) or ClipFusion
    (with_labels=False, nearest rgb) on an x-slab of the grid.  clip_seem_fusion.py:612-674."""

    def __init__(self, origin, voxel_size, nvox, trunc, feature_dim, with_labels=True, rgb_mode=None,
                 x_begin=0, x_end=None, num_threads=1):
        self.origin = np.asarray(origin, dtype=np.float32).copy()
        self.voxel_size = float(voxel_size)
        self.nvox = np.asarray(nvox, dtype=np.int32).copy()
        self.trunc = float(trunc)
        self.C = int(feature_dim)
        self.with_labels = with_labels
        self.rgb_mode = (1 if with_labels else 0) if rgb_mode is None else int(rgb_mode)
        self.x_begin = int(x_begin)
        self.x_end = int(self.nvox[0]) if x_end is None else int(x_end)
        self.num_threads = num_threads
        n = (self.x_end - self.x_begin) * int(self.nvox[1]) * int(self.nvox[2])
        self.n = n
        self.tsdf = np.zeros(n, np.float32)
        self.rgb = np.zeros((n, 3), np.float32)
        self.clip_feat = np.zeros((n, self.C), np.float32)
        self.weight = np.zeros(n, np.int32)
        self.tsdf_weight = np.zeros(n, np.int32)
        self.labels_one_hot = np.zeros((n, N_CLASSES), np.int32) if with_labels else None
        self.last_valid = None
        self.last_tsdf_valid = None
        self.last_counts = None

    def integrate(self, depth, rgb, poses, K, table, seg=None, want_masks=True, table_mode=0):
        """depth [B,H,W], rgb [B,H,W,3], poses [B,4,4], K [B,3,3], table [B,C,npy,npx] (any strides),
        seg [B,H,W] class ids (any dtype) or None.  table_mode 1: table [B,C,1,n_segments], the sample is the row
        of the voxel's class id (north_star's segment table; saf_oracle.c saf_oracle_integrate_ex)."""
        depth = np.ascontiguousarray(depth, np.float32)
        rgb = np.ascontiguousarray(rgb, np.float32)
        poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 16)
        K = np.ascontiguousarray(K, np.float32).reshape(-1, 9)
        B, H, W = depth.shape
        table = np.asarray(table, np.float32)
        assert table.ndim == 4 and table.shape[0] == B
        _, C, npy, npx = table.shape
        assert C >= self.C
        es = table.itemsize
        sb, sc, sy, sx = (s // es for s in table.strides)
        if npx > 1 and npy > 1 and sy != sx * npx:
            table = np.ascontiguousarray(table)
            sb, sc, sy, sx = (s // es for s in table.strides)
        sr = sx if npx > 1 else (sy if npy > 1 else 1)
        segf = None
        if self.with_labels or table_mode == 1:
            assert seg is not None
            segf = np.ascontiguousarray(seg, np.float32)  # pano_seg.float(), clip_seem_fusion.py:760
        valid = np.zeros((B, self.n), np.uint8) if want_masks else None
        tvalid = np.zeros((B, self.n), np.uint8) if want_masks else None
        counts = np.zeros(2 * B, np.int64)
        rc = lib().saf_oracle_integrate_ex(
            _p(self.origin), ctypes.c_float(self.voxel_size), _p(self.nvox), ctypes.c_int(self.x_begin),
            ctypes.c_int(self.x_end), ctypes.c_float(self.trunc), ctypes.c_int(B), ctypes.c_int(H),
            ctypes.c_int(W), _p(depth), _p(rgb), _p(segf), _p(table), ctypes.c_ssize_t(sb),
            ctypes.c_ssize_t(sc), ctypes.c_ssize_t(sr), ctypes.c_int(npy), ctypes.c_int(npx),
            ctypes.c_int(self.C), _p(poses), _p(K), ctypes.c_int(self.rgb_mode), ctypes.c_int(N_CLASSES),
            _p(self.tsdf), _p(self.tsdf_weight), _p(self.weight), _p(self.rgb), _p(self.clip_feat),
            _p(self.labels_one_hot), _p(valid), _p(tvalid), _p(counts), ctypes.c_int(self.num_threads),
            ctypes.c_int(table_mode))
        if rc < 0:
            raise ValueError("saf_oracle_integrate: bad argument (%d)" % rc)
        if rc == 1:
            raise RuntimeError("Class values must be smaller than num_classes.")  # torch one_hot's error
        self.last_valid = None if valid is None else valid.astype(bool)
        self.last_tsdf_valid = None if tvalid is None else tvalid.astype(bool)
        self.last_counts = counts.reshape(B, 2)
        return self.last_counts

    def label_argmax(self):
        out = np.empty(self.n, np.int64)
        lib().saf_oracle_label_argmax(_p(self.labels_one_hot), ctypes.c_int64(self.n),
                                      ctypes.c_int(N_CLASSES), _p(out))
        return out


# ----------------------------------------------------------------------------------------------
# query path (numpy fp32)
# ----------------------------------------------------------------------------------------------

def _softmax(x, axis=-1):
    x = x - x.max(axis=axis, keepdims=True)
    e = np.exp(x)
    return e / e.sum(axis=axis, keepdims=True)


def normalize_rows(feats, mode="nan_to_num"):
    """clip_seem_fusion.py:507-511 (divide by the norm, NaN -> 0) and
    eval_scannet_segmentation.py / hypersim_eval.py:50-51 (norm clamped from below)."""
    feats = np.asarray(feats, np.float32)
    norm = np.linalg.norm(feats, axis=-1, keepdims=True).astype(np.float32)
    if mode == "nan_to_num":
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.nan_to_num(feats / norm, nan=0.0, posinf=0.0, neginf=0.0).astype(np.float32)
    return (feats / np.maximum(norm, np.float32(0.1))).astype(np.float32)


def run_query(img_feats, text_feats):
    """clipfusion.py:899-904 with the text encoder's output passed in: softmax(100 * F @ X^T)."""
    c = img_feats.shape[-1]
    logits = np.float32(100.0) * (np.asarray(img_feats, np.float32) @ np.asarray(text_feats, np.float32)[:, :c].T)
    return _softmax(logits.astype(np.float32), -1).astype(np.float32)


def clip_feature_surgery(image_features, text_features, redundant_feats=None):
    """clipfusion.py:906-934, computed through the GEMM identity
    sim[m,t] = w_t S[m,t] - mean_s(w_s S[m,s]),  S = F X^T,  w = T softmax(2 S[0,:])."""
    F = np.asarray(image_features, np.float32)
    X = np.asarray(text_features, np.float32)
    if redundant_feats is not None:
        return F @ (X - np.asarray(redundant_feats, np.float32)).T
    prob = _softmax((F[:, :1, :] @ X.T) * np.float32(2.0), -1)
    w = prob / prob.mean(-1, keepdims=True)           # [b,1,T]
    S = F @ X.T                                       # [b,M,T]
    ws = S * w
    return (ws - ws.mean(-1, keepdims=True)).astype(np.float32)


def clip_feature_surgery_literal(image_features, text_features):
    """Same function following the reference's own evaluation order (materialises [b,M,T,C]);
    small inputs only."""
    F = np.asarray(image_features, np.float32)
    X = np.asarray(text_features, np.float32)
    prob = _softmax((F[:, :1, :] @ X.T) * np.float32(2.0), -1)
    w = prob / prob.mean(-1, keepdims=True)
    feats = F[:, :, None, :] * X[None, None, :, :]
    feats = feats * w.reshape(1, 1, -1, 1)
    feats = feats - feats.mean(2, keepdims=True)
    return feats.sum(-1).astype(np.float32)


def relevance_minmax(sim_col):
    """clip_seem_fusion.py:527-533: subtract the mean, clip to [0,1], min-max normalise."""
    rel = np.asarray(sim_col, np.float32).copy()
    rel = rel - rel.mean()
    rel = np.clip(rel, 0, 1)
    return (rel - rel.min()) / (rel.max() - rel.min())


def relevance_half(rel):
    """query_mesh.py:39: ((rel - 0.5) * 2).clamp(0, 1)."""
    return np.clip((np.asarray(rel, np.float32) - np.float32(0.5)) * np.float32(2), 0, 1)


def minmax_per_text(sim):
    """query_mesh.py:59-61: min-max over the rows, per text.  sim [M,T]."""
    sim = np.asarray(sim, np.float32)
    mn, mx = sim.min(axis=0, keepdims=True), sim.max(axis=0, keepdims=True)
    return (sim - mn) / (mx - mn)


def relevance_outliers(rel):
    """query_mesh.py:63-73: keep entries above median + 2 sigma (torch.median = lower median, torch.std unbiased)."""
    rel = np.asarray(rel, np.float32)
    med = np.sort(rel)[(len(rel) - 1) // 2]
    thr = med + np.float32(2) * rel.std(ddof=1, dtype=np.float32)
    return np.where(rel > thr, rel, np.float32(0)).astype(np.float32)


def segment_labels(feats, text):
    """eval_scannet_segmentation.py:546-561: clamp_min(0.1) normalisation, argsort(descending) of softmax(100 cos);
    ties towards the lower text index."""
    F = normalize_rows(feats, "clamp_min")
    rel = _softmax(np.float32(100) * (F @ np.asarray(text, np.float32).T), -1)
    return np.argsort(-rel.astype(np.float64), axis=-1, kind="stable")


def presence_scores(feats, background_text, target_text):
    """hypersim_eval.py:50-51, 80-89: per target, max over rows of softmax(100 cos([background.., target]))[-1]."""
    F = normalize_rows(feats, "clamp_min")
    out = []
    for i in range(len(target_text)):
        tf = np.concatenate([background_text, target_text[i:i + 1]], 0).astype(np.float32)
        out.append(_softmax(np.float32(100) * (F @ tf.T), -1)[:, -1].max())
    return np.array(out, np.float32)


def topk_indices(scores, k):
    """Top-k rows per text column, ties broken towards the lower row index.  scores [M,T] -> [T,k]."""
    M, T = scores.shape
    out = np.empty((T, k), np.int64)
    for t in range(T):
        order = np.lexsort((np.arange(M), -scores[:, t].astype(np.float64)))
        out[t] = order[:k]
    return out


def extract_mesh_by_object(vertices, faces, colors, vertex_indices, obj_idx):
    """handy_utils.py:585-611 (arrays only; the reference also wraps them in an open3d mesh)."""
    sel = np.where(vertex_indices == obj_idx)[0]
    mask = np.zeros(len(vertices), bool)
    mask[sel] = True
    keep = mask[faces].all(axis=1)
    remap = np.cumsum(mask) - 1
    return vertices[sel], remap[faces[keep]], colors[sel]
