"""Times K1 / K2 / K3 of the window path (and the frame-by-frame path) separately on the cfg2 workload.
Usage: python tools/prof_window.py [window] [n_windows]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spatially_aware_ai_b200 as saf  # noqa: E402
from spatially_aware_ai_b200 import _lib, synth  # noqa: E402
from spatially_aware_ai_b200.synth import FakeClip, FakeSeg  # noqa: E402

window = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_windows = int(sys.argv[2]) if len(sys.argv) > 2 else 6
stride = int(sys.argv[3]) if len(sys.argv) > 3 else 5
cfg = synth.baseline_config("cfg2")
origin, nvox = cfg.grid()
dev = torch.device("cuda:0")
clip, seg = FakeClip(cfg.feature_dim), FakeSeg()
vol = saf.ClipSeemFusion(torch.from_numpy(origin), cfg.voxel_size, torch.from_numpy(nvox), cfg.trunc, False,
                         cfg.patch_size, cfg.patch_stride, clip, seg).to(dev)
lib = _lib.load()
P = window * n_windows
host = [synth.make_frame(cfg, (i * stride) % cfg.frames, table_layout="hwc") for i in range(P)]
d_depth = torch.stack([torch.from_numpy(f["depth"]) for f in host]).to(dev)
d_rgb = torch.stack([torch.from_numpy(f["rgb"]) for f in host]).to(dev)
d_seg = torch.stack([torch.from_numpy(f["seg"]) for f in host]).to(dev)
d_table = torch.stack([torch.from_numpy(np.ascontiguousarray(f["table"].transpose(1, 2, 0))) for f in host]).to(dev)
npy, npx = cfg.npatches
H, W, C = cfg.height, cfg.width, cfg.feature_dim
frames = (_lib.Frame * P)()
for i in range(P):
    f = frames[i]
    f.depth, f.rgb, f.seg, f.table = d_depth[i].data_ptr(), d_rgb[i].data_ptr(), d_seg[i].data_ptr(), d_table[i].data_ptr()
    f.seg_dtype, f.table_stride_c, f.table_stride_r, f.npy, f.npx = _lib.SAF_SEG_U8, 1, C, npy, npx
    f.pose[:] = host[i]["pose"].reshape(-1).tolist()
    f.K[:] = host[i]["K"].reshape(-1).tolist()
ws = vol._workspace(max(window, 1), npy * npx * C)
g, v = vol._grid_desc(), vol._volume_desc()
st = torch.cuda.current_stream(dev).cuda_stream
trunc = float(cfg.trunc)
# warm: one pass over everything
_lib.check(lib.saf_integrate_sequence(ctypes.byref(g), ctypes.byref(v), frames, P, H, W, trunc, 1, ctypes.byref(ws), st), "seq")
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
tot = np.zeros(3)
upd = 0
for wi in range(n_windows):
    fr = ctypes.cast(ctypes.byref(frames, wi * window * ctypes.sizeof(_lib.Frame)), ctypes.POINTER(_lib.Frame))
    s0 = vol.stats()["total_valid"]
    ev[0].record()
    _lib.check(lib.saf_frustum_cull(ctypes.byref(g), fr, window, H, W, trunc, ctypes.byref(ws), st), "k1")
    ev[1].record()
    if window > 1:
        _lib.check(lib.saf_tsdf_update_window(ctypes.byref(g), ctypes.byref(v), fr, window, H, W, trunc, ctypes.byref(ws), st), "k2w")
    else:
        _lib.check(lib.saf_tsdf_update(ctypes.byref(g), ctypes.byref(v), fr, 1, H, W, trunc, ctypes.byref(ws), None, None, st), "k2")
    ev[2].record()
    if window > 1:
        _lib.check(lib.saf_feature_accumulate_window(ctypes.byref(g), ctypes.byref(v), fr, window, H, W, 1, ctypes.byref(ws), st), "k3w")
    else:
        _lib.check(lib.saf_feature_accumulate(ctypes.byref(g), ctypes.byref(v), fr, 1, 0, H, W, 1, ctypes.byref(ws), st), "k3")
    ev[3].record()
    torch.cuda.synchronize()
    t = np.array([ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(3)])
    tot += t
    s1 = vol.stats()
    upd += s1["total_valid"] - s0
    print("window %d: K1 %.1f us  K2 %.1f us  K3 %.1f us   updates %d  blocks %d processed %d" %
          (wi, t[0], t[1], t[2], s1["total_valid"] - s0, s1["last_blocks"], s1["last_processed"]))
print("avg per window: K1 %.1f K2 %.1f K3 %.1f us; updates/window %.0f; K3 alone %.1f Mupd/s; serial %.1f Mupd/s" %
      (tot[0] / n_windows, tot[1] / n_windows, tot[2] / n_windows, upd / n_windows, upd / tot[2], upd / tot.sum()))
