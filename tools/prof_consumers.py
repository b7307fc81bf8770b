"""Times the consumers of the fused grid at BASELINE config-2 scale: label argmax, object labelling, marching cubes
and vertex sampling (CUDA events), with the bytes each one has to move.  Usage: python tools/prof_consumers.py [frames]"""
import ctypes, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spatially_aware_ai_b200 as saf
from spatially_aware_ai_b200 import mesh, synth
from tests.helpers import FakeClip, FakeSeg

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 200
cfg = synth.baseline_config("cfg2")
origin, nvox = cfg.grid()
dev = torch.device("cuda:0")
clip, seg = FakeClip(cfg.feature_dim), FakeSeg()
vol = saf.ClipSeemFusion(torch.from_numpy(origin), cfg.voxel_size, torch.from_numpy(nvox), cfg.trunc, False, 0, 0, clip, seg).to(dev)
for lo in range(0, n_frames, 40):
    frames = [synth.make_frame(cfg, (i * 5) % cfg.frames, table_layout="hwc") for i in range(lo, min(n_frames, lo + 40))]
    clip.next_table = torch.stack([torch.from_numpy(np.ascontiguousarray(f["table"].transpose(1, 2, 0))) for f in frames]).to(dev).permute(0, 3, 1, 2)
    seg.queue = [torch.from_numpy(f["seg"]).to(dev) for f in frames]
    vol.integrate_sequence(torch.stack([torch.from_numpy(f["depth"]) for f in frames]).to(dev),
                           torch.stack([torch.from_numpy(f["rgb"]) for f in frames]).to(dev),
                           torch.stack([torch.from_numpy(f["pose"]) for f in frames]), torch.stack([torch.from_numpy(f["K"]) for f in frames]))
torch.cuda.synchronize()
N, C = vol.tsdf.numel(), cfg.feature_dim
print("grid %s = %.1f M voxels, %d frames fused, %d observed voxels" % (nvox.tolist(), N / 1e6, n_frames, int((vol.weight > 0).sum())))

def timed(fn, reps=3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return out, best

grid, ms = timed(lambda: vol.label_argmax())
print("label_argmax: %.3f ms, %.0f GB/s (reads N*143*4 B, writes N*8 B)" % (ms, N * (143 * 4 + 8) / ms / 1e6))
grid3 = grid.view(*[int(v) for v in nvox])
(ids, n_obj), ms = timed(lambda: saf.label_objects(grid3))
print("label_objects: %d objects, %.3f ms (incl. scratch allocation and the count read-back)" % (n_obj, ms))
(verts, verts_world, faces), ms = timed(lambda: mesh.marching_cubes_device(vol))
print("marching cubes: %d verts, %d faces, %.3f ms (incl. scratch allocation and the count read-back; tsdf+weight %.0f MB)" %
      (len(verts), len(faces), ms, N * 8 / 1e6))
feats, ms = timed(lambda: mesh.sample_vertices(vol, verts, vol.clip_feat, "bilinear"))
V = len(verts)
print("vertex features [%d,%d]: %.3f ms, %.0f GB/s counting 2 rows read + 1 written per vertex" % (V, C, ms, V * 3 * C * 4 / ms / 1e6))
cols, ms = timed(lambda: mesh.sample_vertices(vol, verts, vol.rgb, "bilinear", clamp01=True))
print("vertex colours: %.3f ms" % ms)
t0 = time.perf_counter(); out = vol.__class__.__mro__[1].__dict__  # noqa
vol.voxel_obj_idx, vol.objects_segmentation_color = ids, vol.rgb
t0 = time.perf_counter(); res = vol.extract_mesh(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("extract_mesh() end to end (6-tuple, verts/faces copied to the host): %.1f ms" % (dt * 1e3))
