#!/usr/bin/env python
"""Where a multilevel cycle spends its time: VCycle.runVCycleML on the N x N Laplacian with k right-hand sides,
W-cycle vs V-cycle (gamma), timed with CUDA events; run it under different GLAB_MS / GLAB_ML_GRAPH settings to
separate the cost of the coarse level visits.  Single GPU.

    python scripts/probe_multilevel.py [--grid 4096] [--k 8]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import glab_b200 as G  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=4096)
    ap.add_argument("--k", type=int, default=8)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    N, k = args.grid, args.k
    n = N * N
    ei, ev = G.UtilsGNN.laplacianfun_torch(N, device=dev)
    A = torch.sparse_coo_tensor(ei, ev.flatten().float(), (n, n))
    torch.manual_seed(1)
    b = torch.rand(n, k, device=dev)
    out = {"grid": N, "k": k, "GLAB_MS": os.environ.get("GLAB_MS"), "GLAB_ML_GRAPH": os.environ.get("GLAB_ML_GRAPH")}
    for gamma in (2, 1):
        x = torch.zeros(n, k, device=dev)
        for _ in range(2):
            x = G.VCycle.runVCycleML(A, b, x, 3, 3, gamma)
        ts = []
        for _ in range(3):
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            x = G.VCycle.runVCycleML(A, b, x, 3, 3, gamma)
            e.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(e))
        out["gamma_%d_ms_per_cycle" % gamma] = min(ts)
    info = G.VCycle.hierarchy_info(A, "multilevel")
    out["rows_per_level"] = info["rows_per_level"]
    out["visits_per_level"] = info["visits_per_level"]
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
