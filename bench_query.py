#!/usr/bin/env python
"""bench_query.py -- language-query scoring throughput (BASELINE.json configs[3]):
T CLIP text embeddings scored by cosine against an [M, C] fused feature matrix.

    python bench_query.py [--rows M] [--texts T] [--dim C] [--k K] [--precision tf32|fp32]
    torchrun --nproc-per-node N bench_query.py --rows M ...     # M rows split into N x-slabs, one per rank

Prints one JSON line: scores kernel GB/s and TFLOP/s (CUDA events, median of --iters), top-k time,
and the numpy oracle's time on a bounded row sample for scale.  Under torchrun every rank scores its own slab of
the rows and the per-rank top-k lists are combined with an NCCL all_gather (slab.gather_topk): the only
collective of the query path; the line then reports the global top-k time (max over ranks) as well.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=24_000_000)
    ap.add_argument("--texts", type=int, default=256)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--precision", default="tf32")
    ap.add_argument("--no-topk", action="store_true")
    ap.add_argument("--cpu-rows", type=int, default=200_000)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import spatially_aware_ai_b200 as saf
    from spatially_aware_ai_b200 import slab
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    json_fd = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        sys.stdout.flush()
        json_fd = os.dup(1)       # NCCL prints its version banner to fd 1: keep stdout to the one JSON line
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    M_total, T, C = args.rows, args.texts, args.dim
    r_begin, r_end = slab.slab_bounds(M_total, world, rank)      # rows stand for voxels: a contiguous slab per rank
    M = r_end - r_begin
    g = torch.Generator(device=dev).manual_seed(rank)
    F = torch.empty((M, C), dtype=torch.float32, device=dev)
    step = 1 << 20
    for r0 in range(0, M, step):
        F[r0:r0 + step].normal_(generator=g)
    X = torch.randn((T, C), generator=torch.Generator(device=dev).manual_seed(12345), device=dev)
    X = X / X.norm(dim=1, keepdim=True)
    out = torch.empty((M, T), dtype=torch.float32, device=dev)
    peaks = {}
    ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(ppath):
        peaks = json.load(open(ppath))

    def timed(fn, iters):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    ms = timed(lambda: saf.query_scores(F, X, norm="nan_to_num", mode="dot", precision=args.precision, out=out), args.iters)
    read_b = M * C * 4 + T * C * 4
    write_b = M * T * 4
    flops = 2.0 * M * C * T
    res = {"metric": "query_scores", "rows": M, "rows_total": M_total, "n_gpus": world, "texts": T, "dim": C,
           "precision": args.precision,
           "scores_ms": ms, "read_GBps": read_b / ms / 1e6, "read_plus_write_GBps": (read_b + write_b) / ms / 1e6,
           "tflops": flops / ms / 1e9, "hbm_peak_GBps": peaks.get("hbm_gbs"),
           "frac_hbm_rw": (read_b + write_b) / ms / 1e6 / peaks["hbm_gbs"] if peaks else None,
           "tf32_dense_peak_tflops_nominal": 1125.0, "frac_tensor_nominal": flops / ms / 1e9 / 1125.0}
    if not args.no_topk:
        tk = timed(lambda: saf.query_topk(F, X, args.k, norm="nan_to_num", mode="dot", precision=args.precision), 2)
        res["topk_ms"] = tk
        res["topk_k"] = args.k
        res["topk_read_GBps"] = read_b / tk / 1e6
        if world > 1:
            def global_topk():
                ts, ti = saf.query_topk(F, X, args.k, norm="nan_to_num", mode="dot", precision=args.precision,
                                        index_base=r_begin)
                return slab.gather_topk(ts, ti, args.k)

            gs, gi = global_topk()
            dist.barrier()
            tg = torch.tensor([timed(global_topk, 2)], dtype=torch.float64, device=dev)
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
            res["global_topk_ms"] = float(tg.item())
            res["global_topk_rows_per_s"] = M_total / (float(tg.item()) * 1e-3)
            # every rank holds the same merged list; its winners come from all slabs
            owners = torch.bucketize(gi[:, 0].contiguous(), torch.tensor(
                [slab.slab_bounds(M_total, world, r)[1] for r in range(world)], device=dev), right=True)
            res["top1_owner_ranks"] = sorted(set(owners.cpu().tolist()))
    # CPU: numpy (MKL/OpenBLAS sgemm, all cores) on a row sample
    from oracle import oracle as O
    Fs = F[: args.cpu_rows].cpu().numpy()
    Xs = X.cpu().numpy()
    t0 = time.perf_counter()
    ref = O.normalize_rows(Fs) @ Xs.T
    cpu_s = time.perf_counter() - t0
    got = out[: args.cpu_rows].cpu().numpy()
    res["cpu_rows_per_s"] = args.cpu_rows / cpu_s
    res["gpu_rows_per_s"] = M / (ms * 1e-3)
    res["max_abs_err_vs_oracle"] = float(np.abs(got - ref).max())
    res["cpu_cores"] = os.cpu_count()
    if rank == 0:
        if json_fd is not None:
            os.write(json_fd, (json.dumps(res) + "\n").encode())
        else:
            print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
