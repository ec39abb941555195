#!/usr/bin/env python
"""Jacobi sweeps on a SMALL operator (default L1448 = one rank's share of L4096 at 8 GPUs, whose
CSR + vectors fit the 126 MB L2): time per sweep with ping-pong buffers, for comparing with the
HBM-roofline time of the same bytes and as the ncu target for the L2 hit rate of repeated sweeps.

    python scripts/probe_small.py --grid 1448 --sweeps 40
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("GLAB_MS", "2")      # time the multi-sweep kernel on every size
import torch  # noqa: E402
import glab_b200 as G  # noqa: E402

rt = G.runtime


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=1448)
    ap.add_argument("--sweeps", type=int, default=40)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-multi", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    N = args.grid
    n = N * N
    ei, ev = G.generators.laplacian_2d(N, torch.float32, dev)
    plan = G.Plan.from_coo(ei.contiguous(), n)
    vals = rt.get_vals(plan, ev.contiguous())
    torch.manual_seed(24601)
    b = torch.rand(n, 1, device=dev)
    xa = torch.rand(n, 1, device=dev)
    xb = torch.empty_like(xa)
    diag = torch.full((n,), -4.0, device=dev)
    w = torch.tensor([0.7], device=dev)

    def sweeps():
        a, c = xa, xb
        for _ in range(args.sweeps):
            rt.jacobi(plan, vals, diag, b, a, c, w)
            a, c = c, a

    def multi():
        rt.jacobi_sweeps(plan, vals, diag, b, xa, xb, w, args.sweeps)

    z = plan.nnz
    moved = z * (plan.index_bytes + 4) + 4 * (n + 1) + 4 * n * 4
    for name, fn in (("one launch per sweep", sweeps), ("multi-sweep kernel", multi)):
        if name.startswith("multi") and args.no_multi:
            continue
        fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / args.sweeps)
        print(json.dumps({"grid": N, "mode": name, "rows": n, "nnz": z, "index_bytes": plan.index_bytes,
                          "sweeps_per_pass": args.sweeps, "ms_per_sweep": best, "moved_MB": moved / 1e6,
                          "moved_GBps": moved / best / 1e6, "Gnnz_per_s": z / best / 1e6}), flush=True)


if __name__ == "__main__":
    main()
