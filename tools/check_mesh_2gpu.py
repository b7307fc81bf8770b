"""torchrun --nproc-per-node N tools/check_mesh_2gpu.py : slab.extract_mesh_distributed over NCCL against the mesh of
the same grid held by one volume (rank 0)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spatially_aware_ai_b200 import slab
from tests import helpers as Hh

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
nvox, C = (64, 48, 40), 64
rng = np.random.default_rng(5)
n = int(np.prod(nvox))
g = dict(cls="ClipFusion", feature_dim=C, origin=np.array([-1.0, 0.5, 2.0], np.float32), nvox=np.asarray(nvox, np.int32),
         voxel_size=0.03, trunc=0.09)
state = dict(tsdf=rng.uniform(-1, 1, n).astype(np.float32), weight=rng.integers(0, 4, n).astype(np.int32),
             rgb=rng.uniform(0, 1, (n, 3)).astype(np.float32), clip_feat=rng.standard_normal((n, C)).astype(np.float32))
a, b = slab.slab_bounds(nvox[0], world, rank)
vol, _, _ = Hh.make_gpu_volume(g, x_begin=a, x_end=b)
pl = nvox[1] * nvox[2]
for k, t in state.items():
    getattr(vol, k).copy_(torch.from_numpy(t[a * pl:b * pl]))
out = slab.extract_mesh_distributed(vol)
if rank == 0:
    full, _, _ = Hh.make_gpu_volume(g)
    for k, t in state.items():
        getattr(full, k).copy_(torch.from_numpy(t))
    fv, ff, fc, ffeat = full.extract_mesh()
    wv, wf, wc, wfeat = out
    ok = np.array_equal(wv, fv) and np.array_equal(wf, ff)
    print("distributed mesh over %d ranks: %d verts %d faces; identical to single-volume mesh: %s; max attr diff %.2e / %.2e"
          % (world, len(wv), len(wf), ok, np.abs(wc - fc.cpu().numpy()).max(), np.abs(wfeat - ffeat.cpu().numpy()).max()))
    assert ok and np.abs(wfeat - ffeat.cpu().numpy()).max() <= 1e-5
dist.destroy_process_group()
