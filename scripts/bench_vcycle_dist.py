#!/usr/bin/env python
"""BASELINE config 5 on N GPUs (torchrun): two-grid V-cycle (3+3 Jacobi sweeps, Chebyshev-4
coarse solve) with k right-hand-side columns on the row-partitioned 5-point Laplacian.

    python -m torch.distributed.run --nproc-per-node 8 ... scripts/bench_vcycle_dist.py --grid 8192 --k 8
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
from glab_b200.dist_vcycle import DistTwoGrid  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=8192)
    ap.add_argument("--k", type=int, default=8)
    ap.add_argument("--cycles", type=int, default=3)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    N, k = args.grid, args.k
    t0 = time.perf_counter()
    ei, ev = G.UtilsGNN.laplacianfun_torch(N, device=dev)
    tg = DistTwoGrid(ei, ev, k, rank, world, engine=os.environ.get("GLAB_DIST_ENGINE", "peer"))
    del ei, ev
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0
    f0, f1 = tg.fine.bounds(rank)
    gen = torch.Generator().manual_seed(24601 + rank)
    b = torch.rand(f1 - f0, k, generator=gen).to(dev)
    tg.load_x(torch.rand(f1 - f0, k, generator=gen).to(dev))

    def rnorm():
        r = tg.residual_local(b)
        s = (r.double() ** 2).sum(0)
        dist.all_reduce(s)
        return torch.sqrt(s)

    norms = [rnorm()]
    tg.cycle(b)                      # warm-up
    norms.append(rnorm())
    times = []
    for _ in range(args.cycles):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        tg.cycle(b)
        torch.cuda.synchronize()
        dist.barrier()
        times.append(time.perf_counter() - t0)
        norms.append(rnorm())
    t = torch.tensor([min(times)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t = t.item()
    work = 7 * tg.nnz["A"] + 2 * tg.nnz["P"] + 4 * tg.nnz["Ac"]
    if rank == 0:
        print(json.dumps({"workload": "row-partitioned two-grid V-cycle, L%d, k=%d fp32" % (N, k), "n_gpus": world,
                          "rows": N * N, "nnz_A": tg.nnz["A"], "nnz_P": tg.nnz["P"], "nnz_Ac": tg.nnz["Ac"],
                          "setup_s_replicated": t_setup, "cycle_ms": t * 1e3, "spmv_nnz_per_cycle": work,
                          "Gnnz_per_s": work / t / 1e9, "Gnnz_x_columns_per_s": work * k / t / 1e9,
                          "residual_norm_col0": [float(v[0]) for v in norms]}), flush=True)
    tg.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
