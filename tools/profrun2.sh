python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/b_2gpu.json 2> gpurun_out/b_2gpu.err
tail -n 4 gpurun_out/b_2gpu.err | cut -c1-300; cut -c1-400 gpurun_out/b_2gpu.json
python - <<'PY'
import torch, time
h = torch.empty(256<<20, dtype=torch.uint8).pin_memory()
d = torch.empty(256<<20, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
for n in (1<<20, 4<<20, 39<<20, 256<<20):
    t0=time.perf_counter()
    for _ in range(20): d[:n].copy_(h[:n], non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t0
    print("H2D %d MiB: %.1f GB/s" % (n>>20, 20*n/dt/1e9))
PY
