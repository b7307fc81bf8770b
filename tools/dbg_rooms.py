import sys, ctypes, numpy as np, torch
sys.path.insert(0, "/root/repo")
import spatially_aware_ai_b200 as saf
from spatially_aware_ai_b200 import synth
from tests.helpers import FakeClip, FakeSeg
cfg = synth.baseline_config("cfg2")
origin, nvox_room = cfg.grid()
n_rooms = int(sys.argv[1]) if len(sys.argv) > 1 else 2
wall = 10 if n_rooms > 1 else 0
slab_nx = int(nvox_room[0]) + wall
nvox = nvox_room.copy(); nvox[0] = slab_nx * n_rooms
clip, seg = FakeClip(cfg.feature_dim), FakeSeg()
vol = saf.ClipSeemFusion(torch.from_numpy(origin), cfg.voxel_size, torch.from_numpy(nvox), cfg.trunc, False, 0, 0, clip, seg, x_begin=0, x_end=slab_nx).cuda()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(24):
    fr = synth.make_frame(cfg, i * 41, table_layout="hwc")
    fr["pose"][0, 3] += (i % n_rooms) * slab_nx * cfg.voxel_size
    clip.next_table = torch.from_numpy(np.ascontiguousarray(fr["table"].transpose(1, 2, 0))).cuda().permute(2, 0, 1)[None]
    seg.queue = [torch.from_numpy(fr["seg"]).cuda()]
    args = (torch.from_numpy(fr["depth"]).cuda()[None], torch.from_numpy(fr["rgb"]).cuda()[None], torch.from_numpy(fr["pose"])[None], torch.from_numpy(fr["K"])[None])
    torch.cuda.synchronize(); e0.record()
    vol.integrate(*args)
    e1.record(); torch.cuda.synchronize()
    st = vol.stats()
    print("frame %2d room %d: %6.1f us  blocks %5d processed %5d valid %6d tv %7d  cull_on %d" % (i, i % n_rooms, e0.elapsed_time(e1) * 1e3, st["last_blocks"], st["last_processed"], st["last_valid"][0], st["last_tsdf_valid"][0], st["depth_cull_on"]))
