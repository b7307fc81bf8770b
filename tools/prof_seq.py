"""Times saf_integrate_sequence step by step on consecutive cfg2 frames: overlapped path (100 frames per call) vs
serial path (8 frames per call).  Usage: python tools/prof_seq.py [pool] [stride]"""
import ctypes, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spatially_aware_ai_b200 as saf
from spatially_aware_ai_b200 import _lib, synth
from tests.helpers import FakeClip, FakeSeg

P = int(sys.argv[1]) if len(sys.argv) > 1 else 400
stride = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = synth.baseline_config("cfg2")
origin, nvox = cfg.grid()
dev = torch.device("cuda:0")
clip, seg = FakeClip(cfg.feature_dim), FakeSeg()
vol = saf.ClipSeemFusion(torch.from_numpy(origin), cfg.voxel_size, torch.from_numpy(nvox), cfg.trunc, False, 0, 0, clip, seg).to(dev)
lib = _lib.load()
host = [synth.make_frame(cfg, (i * stride) % cfg.frames, table_layout="hwc") for i in range(P)]
d_depth = torch.stack([torch.from_numpy(f["depth"]) for f in host]).to(dev)
d_rgb = torch.stack([torch.from_numpy(f["rgb"]) for f in host]).to(dev)
d_seg = torch.stack([torch.from_numpy(f["seg"]) for f in host]).to(dev)
d_table = torch.stack([torch.from_numpy(np.ascontiguousarray(f["table"].transpose(1, 2, 0))) for f in host]).to(dev)
npy, npx = cfg.npatches; H, W, C = cfg.height, cfg.width, cfg.feature_dim
frames = (_lib.Frame * P)()
for i in range(P):
    f = frames[i]
    f.depth, f.rgb, f.seg, f.table = d_depth[i].data_ptr(), d_rgb[i].data_ptr(), d_seg[i].data_ptr(), d_table[i].data_ptr()
    f.seg_dtype, f.table_stride_c, f.table_stride_r, f.npy, f.npx = _lib.SAF_SEG_U8, 1, C, npy, npx
    f.pose[:] = host[i]["pose"].reshape(-1).tolist(); f.K[:] = host[i]["K"].reshape(-1).tolist()
ws = vol._workspace(8, npy * npx * C)
g, v = vol._grid_desc(), vol._volume_desc()
st = torch.cuda.current_stream(dev).cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

def seq(lo, n):
    fr = ctypes.cast(ctypes.byref(frames, lo * ctypes.sizeof(_lib.Frame)), ctypes.POINTER(_lib.Frame))
    _lib.check(lib.saf_integrate_sequence(ctypes.byref(g), ctypes.byref(v), fr, n, H, W, float(cfg.trunc), 1, ctypes.byref(ws), st), "seq")

for rep in range(2):
    for lo in range(0, P, 100):
        s0 = vol.stats()
        torch.cuda.synchronize(); e0.record(); seq(lo, 100); e1.record(); torch.cuda.synchronize()
        s1 = vol.stats()
        print("rep %d overlapped frames %4d..%4d: %.2f ms  updates %d blocks %d cull %d" % (rep, lo, lo + 99, e0.elapsed_time(e1),
              s1["total_valid"] - s0["total_valid"], s1["total_blocks"] - s0["total_blocks"], s1["depth_cull_on"]))
for lo in range(0, P, 100):
    torch.cuda.synchronize(); e0.record()
    for k in range(lo, lo + 100, 8):
        seq(k, min(8, lo + 100 - k))
    e1.record(); torch.cuda.synchronize()
    print("serial     frames %4d..%4d: %.2f ms  cull %d" % (lo, lo + 99, e0.elapsed_time(e1), vol.stats()["depth_cull_on"]))
