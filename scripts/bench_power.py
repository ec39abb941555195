#!/usr/bin/env python
"""BASELINE config 3: PowerMethodGNN spectral-radius estimate, 100 iterations, FEM heat-equation
2-D operator on an N x N grid (default 8192: 67 M rows, 604 M nnz), fp32, row-partitioned over the
ranks of one box (torchrun) or on one GPU through the drop-in layer.

    python scripts/bench_power.py [--grid 8192] [--iters 100] [--dtype f32]
    python -m torch.distributed.run --nproc-per-node 8 ... scripts/bench_power.py
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import glab_b200 as G  # noqa: E402
from glab_b200 import dist as gd  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=8192)
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--reps", type=int, default=3)
    # NOTE: capturing the iteration (fused kernels + NCCL all-reduce) in a CUDA graph was tried and
    # dead-locked at 2 ranks (the in-kernel flag acquire and NCCL's own kernels end up waiting on each
    # other inside one graph launch); the iteration is therefore launched eagerly.
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dt = torch.float32 if args.dtype == "f32" else torch.float64
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N = args.grid
    n = N * N
    part = gd.RowPartition(n, world, align=256)
    r0, r1 = part.bounds(rank)
    ei, ev = G.generators.heat_fem_2d((N + 1, N + 1), (1.0, 1.0), dt, dev, rows=(r0, r1))
    gen = torch.Generator().manual_seed(24601 + rank)
    b0 = torch.rand(r1 - r0, 1, generator=gen, dtype=dt).to(dev)
    if world == 1:
        ea = torch.cat([ev, torch.zeros_like(ev)], 1)
        va = torch.cat([b0, torch.zeros_like(b0)], 1)
        layer = G.PowerMethodGNN.PowerMethodGNN(args.iters)
        g0 = torch.zeros(3, dtype=dt, device=dev)
        times = []
        for _ in range(args.reps + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _, _, g = layer(va, ei, ea, g0, None)
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        plan = G.get_plan(ei, n)
        z = plan.nnz
        lam = g[2].item()
        t = min(times[1:])
        extra = {"api": "PowerMethodGNN(100).forward on device tensors (includes the returned edge_attr column)"}
    else:
        halo = gd.HaloPlan.build(part, rank, ei[1])
        lei = torch.stack([ei[0] - r0, halo.local_columns(ei[1])]).contiguous()
        del ei
        op = gd.DistOperator(lei, ev.contiguous(), halo, k=1, engine=os.environ.get("GLAB_DIST_ENGINE", "peer"))
        zt = torch.tensor([op.plan.nnz], dtype=torch.int64, device=dev)
        dist.all_reduce(zt)
        z = int(zt.item())
        times = []
        graph = None
        for _ in range(args.reps + 1):
            if graph is None:
                op.load("v0", b0)
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            if graph is None:
                res, bout, yout = op.power_method(args.iters, "v0")
            else:
                graph.replay()
            torch.cuda.synchronize()
            dist.barrier()
            times.append(time.perf_counter() - t0)
        lam = res[0].item()
        t = min(times[1:])
        tt = torch.tensor([t], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = tt.item()
        extra = {"engine": op.engine, "halo_rows": halo.n_halo}
        op.close()
    if rank == 0:
        spmvs = args.iters + 1
        print(json.dumps({"workload": "H%d heat-eqn FEM 9-pt, PowerMethodGNN(%d), %s" % (N, args.iters, args.dtype),
                          "n_gpus": world, "rows": n, "nnz": z, "seconds": t, "ms_per_iteration": t / spmvs * 1e3,
                          "Gnnz_per_s": spmvs * z / t / 1e9, "lambda": lam, **extra}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
