"""Where an end-to-end step of the bench goes (cfg2, one GPU): H2D alone, fusion alone (public API, resident inputs),
both overlapped as bench.py's e2e leg does.  usage: python tools/prof_e2e.py [frames_per_step] [steps]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spatially_aware_ai_b200 as saf  # noqa: E402
from spatially_aware_ai_b200 import synth  # noqa: E402
from spatially_aware_ai_b200.synth import FakeClip, FakeSeg  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 100
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cfg = synth.baseline_config("cfg2")
origin, nvox = cfg.grid()
dev = torch.device("cuda:0")
vol = saf.ClipSeemFusion(torch.from_numpy(origin), cfg.voxel_size, torch.from_numpy(nvox), cfg.trunc, False,
                         cfg.patch_size, cfg.patch_stride, FakeClip(cfg.feature_dim), FakeSeg()).to(dev)
P = 2 * F
from concurrent.futures import ThreadPoolExecutor
with ThreadPoolExecutor(16) as ex:
    host = list(ex.map(lambda i: synth.make_frame(cfg, i % cfg.frames, table_layout="hwc"), range(P)))
hd = torch.from_numpy(np.stack([f["depth_mm"] for f in host])).pin_memory()
hr = torch.from_numpy(np.stack([f["rgb_u8"] for f in host])).pin_memory()
d_seg = torch.from_numpy(np.stack([f["seg"] for f in host])).to(dev)
d_table = torch.from_numpy(np.stack([np.ascontiguousarray(f["table"].transpose(1, 2, 0)) for f in host])).to(dev).permute(0, 3, 1, 2)
poses = torch.from_numpy(np.stack([f["pose"] for f in host]))
Ks = torch.from_numpy(np.stack([f["K"] for f in host]))
copy_stream = torch.cuda.Stream(dev)


def sl(s):
    k0 = (s * F) % P
    return slice(k0, k0 + F)


PREALLOC = os.environ.get("E2E_PREALLOC", "1") == "1"
stage_d = [torch.empty_like(hd[:F], device=dev) for _ in range(3)]
stage_r = [torch.empty_like(hr[:F], device=dev) for _ in range(3)]


def upload(s):
    with torch.cuda.stream(copy_stream):
        if PREALLOC:
            dd, rr = stage_d[s % 3], stage_r[s % 3]
            dd.copy_(hd[sl(s)], non_blocking=True)
            rr.copy_(hr[sl(s)], non_blocking=True)
        else:
            dd = hd[sl(s)].to(dev, non_blocking=True)
            rr = hr[sl(s)].to(dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(copy_stream)
    return dd, rr, ev


def fuse(s, dd, rr):
    vol.integrate_sequence(dd, rr, poses[sl(s)], Ks[sl(s)], clip_feat_img=d_table[sl(s)], seg_maps=d_seg[sl(s)])


# warm
dd, rr, ev = upload(0); torch.cuda.synchronize(); fuse(0, dd, rr); vol.stats()
# 1. H2D alone
t0 = time.perf_counter()
for s in range(steps):
    dd, rr, ev = upload(s)
    torch.cuda.synchronize()
t_h2d = (time.perf_counter() - t0) / steps
bytes_step = F * (hd[0].numel() * 2 + hr[0].numel())
# 2. fusion alone on matching inputs (uploaded before the clock starts), one stats() read per step
t_fuse = t_call = 0.0
for s in range(steps):
    dd, rr, ev = upload(s)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fuse(s, dd, rr)
    t1 = time.perf_counter()
    vol.stats()
    t_fuse += time.perf_counter() - t0
    t_call += t1 - t0
t_fuse /= steps
t_call /= steps
# 3. overlapped as in bench.py
nxt = upload(0)
torch.cuda.synchronize()
s0 = vol.stats()["total_valid"]
t0 = time.perf_counter()
for s in range(steps):
    dd, rr, ev = nxt
    nxt = upload(s + 1)
    torch.cuda.current_stream(dev).wait_event(ev)
    fuse(s, dd, rr)
    if not PREALLOC:
        dd.record_stream(torch.cuda.current_stream(dev)); rr.record_stream(torch.cuda.current_stream(dev))
    st = vol.stats()
t_both = (time.perf_counter() - t0) / steps
upd = (st["total_valid"] - s0) / steps
print("frames/step %d: H2D alone %.3f ms (%.1f GB/s), fusion alone %.3f ms (call returns after %.3f ms), overlapped %.3f ms; "
      "%.3e updates/step -> e2e %.3e updates/s, kernels-only bound %.3e" %
      (F, t_h2d * 1e3, bytes_step / t_h2d / 1e9, t_fuse * 1e3, t_call * 1e3, t_both * 1e3, upd, upd / t_both, upd / t_fuse))
