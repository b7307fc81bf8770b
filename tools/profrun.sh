python -m pytest tests/test_objects_gpu.py -q 2>&1 | tail -n 12
python - <<'PY'
import torch, time, numpy as np, sys
sys.path.insert(0, '.')
import spatially_aware_ai_b200 as saf
rng = np.random.default_rng(0)
grid = torch.from_numpy(rng.choice(np.array([-1, 133, 0, 1, 2, 3]), size=(304, 304, 154), p=[0.5, 0.1, 0.1, 0.1, 0.1, 0.1]).astype(np.int64)).cuda()
saf.label_objects(grid); torch.cuda.synchronize()
t0 = time.perf_counter(); ids, n = saf.label_objects(grid); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("label_objects 304x304x154 noise: %d objects in %.2f ms" % (n, dt * 1e3))
PY
