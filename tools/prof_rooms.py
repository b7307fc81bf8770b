"""Times saf_integrate_sequence on the N-room multi-GPU workload as seen by one rank (emulated on one GPU).
Usage: python tools/prof_rooms.py [rooms] [own] [frames_per_room]"""
import ctypes, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spatially_aware_ai_b200 as saf
from spatially_aware_ai_b200 import _lib, synth
from tests.helpers import FakeClip, FakeSeg

rooms = int(sys.argv[1]) if len(sys.argv) > 1 else 2
own = int(sys.argv[2]) if len(sys.argv) > 2 else 0
fpr = int(sys.argv[3]) if len(sys.argv) > 3 else 48
cfg = synth.baseline_config("cfg2")
origin, nvox_room = cfg.grid()
wall = int(round(0.2 / cfg.voxel_size)) if rooms > 1 else 0
slab_nx = int(nvox_room[0]) + wall
nvox = nvox_room.copy(); nvox[0] = slab_nx * rooms
dev = torch.device("cuda:0")
clip, seg = FakeClip(cfg.feature_dim), FakeSeg()
vol = saf.ClipSeemFusion(torch.from_numpy(origin), cfg.voxel_size, torch.from_numpy(nvox), cfg.trunc, False, 0, 0, clip, seg,
                         x_begin=own * slab_nx, x_end=(own + 1) * slab_nx).to(dev)
lib = _lib.load()
P = fpr * rooms
host = []
for i in range(P):
    fr = synth.make_frame(cfg, ((i // rooms) * 5) % cfg.frames, table_layout="hwc")
    fr["pose"] = fr["pose"].copy(); fr["pose"][0, 3] += (i % rooms) * slab_nx * cfg.voxel_size
    host.append(fr)
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
for rep in range(3):
    s0 = vol.stats()
    torch.cuda.synchronize(); e0.record()
    _lib.check(lib.saf_integrate_sequence(ctypes.byref(g), ctypes.byref(v), frames, P, H, W, float(cfg.trunc), 1, ctypes.byref(ws), st), "seq")
    e1.record(); torch.cuda.synchronize()
    s1 = vol.stats()
    print("rooms %d own %d: %d frames in %.2f ms; updates %d (%.1f M/s); blocks/call sum %d; cull_on %d" %
          (rooms, own, P, e0.elapsed_time(e1), s1["total_valid"] - s0["total_valid"],
           (s1["total_valid"] - s0["total_valid"]) / e0.elapsed_time(e1) / 1e3, s1["total_blocks"] - s0["total_blocks"], s1["depth_cull_on"]))
