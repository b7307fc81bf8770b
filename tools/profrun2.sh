N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/b_${N}gpu.json 2> gpurun_out/b_${N}gpu.err
grep "\[bench\]" gpurun_out/b_${N}gpu.err | cut -c1-200; cut -c1-200 gpurun_out/b_${N}gpu.json
