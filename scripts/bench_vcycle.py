#!/usr/bin/env python
"""BASELINE config 5 (single-GPU part): two-grid V-cycle exactly as VCycle.py:193-237 (3 + 3
weighted-Jacobi sweeps w = 0.7, classical SOC theta = 0.25, direct interpolation, Galerkin
coarse operator, Chebyshev degree 4 coarse solve with d = -4, c = -3.4, deterministic
every-other C/F splitting) on the N x N 5-point Laplacian with k right-hand-side columns.

    python scripts/bench_vcycle.py [--grid 4096] [--k 8] [--cycles 3]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import glab_b200 as G  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=4096)
    ap.add_argument("--k", type=int, default=8)
    ap.add_argument("--cycles", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    V = G.VCycle
    N, k = args.grid, args.k
    n = N * N
    ei, ev = G.UtilsGNN.laplacianfun_torch(N, device=dev)
    A = torch.sparse_coo_tensor(ei, ev.flatten(), dtype=torch.float)
    torch.manual_seed(24601)
    b = torch.rand(n, k, device=dev)
    x = torch.rand(n, k, device=dev)
    r0 = torch.norm(V.runResidual(A, b, x), dim=0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x = V.runVCycle(A, b, x, 3, 3, 5, True)          # first call builds + caches the hierarchy
    torch.cuda.synchronize()
    t_first = time.perf_counter() - t0
    norms = [r0, torch.norm(V.runResidual(A, b, x), dim=0)]
    times = []
    for _ in range(args.cycles):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x = V.runVCycle(A, b, x, 3, 3, 5, True)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        norms.append(torch.norm(V.runResidual(A, b, x), dim=0))
    op = V._operator(A)
    tg = V._two_grid(A, None)
    z, zp, zc = ei.shape[1], tg.plan_P.nnz, tg.Ac._nnz()
    work = (3 + 3 + 1) * z + 2 * zp + 4 * zc          # SpMV-bearing steps of one cycle (nnz)
    t = min(times)
    print(json.dumps({
        "workload": "two-grid V-cycle, L%d, k=%d fp32" % (N, k), "rows": n, "nnz_A": z, "nnz_P": zp, "nnz_Ac": zc,
        "coarse_rows": tg.P.shape[1], "first_cycle_incl_setup_ms": t_first * 1e3, "cycle_ms": t * 1e3,
        "spmv_nnz_per_cycle": work, "Gnnz_per_s": work / t / 1e9, "Gnnz_x_columns_per_s": work * k / t / 1e9,
        "residual_norm_col0": [float(v[0]) for v in norms],
        "residual_reduction_per_cycle_col0": [float(norms[i + 1][0] / norms[i][0]) for i in range(len(norms) - 1)]}))


if __name__ == "__main__":
    main()
