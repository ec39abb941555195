#!/usr/bin/env python
"""AMG setup on the device timed per stage with CUDA events on the
N x N 5-point Laplacian: classical SOC, PMIS C/F splitting, direct interpolation, sparse
prolongator assembly, transpose plan and the two SpGEMMs of the Galerkin operator, next to
torch.sparse.mm (cuSPARSE, the library path the glue used before) for the same product.

    python scripts/bench_setup.py [--grid 4096] [--reps 3] [--split pmis|alternating]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import glab_b200 as G  # noqa: E402


def timed(fn, reps):
    best, out = None, None
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        best = ms if best is None else min(best, ms)
    return best, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--split", default="alternating", choices=["pmis", "alternating"])
    ap.add_argument("--no-library", action="store_true")
    ap.add_argument("--profile", action="store_true", help="kernel list (torch.profiler) of the two SpGEMMs")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    rt, V = G.runtime, G.VCycle
    N = args.grid
    n = N * N
    ei, ev = G.UtilsGNN.laplacianfun_torch(N, device=dev)
    ev = ev.float().contiguous()
    A = torch.sparse_coo_tensor(ei, ev.flatten(), (n, n))
    op = V._operator(A)
    z, zo = op.edge_index.shape[1], op.off_index.shape[1]
    plan_off = rt.get_plan(op.off_index, n)
    vals_off = rt.get_vals(plan_off, op.off_attr)
    plan_A = rt.get_plan(op.edge_index, n)
    vals_A = rt.get_vals(plan_A, op.edge_attr)
    res = {"workload": "AMG setup, L%d fp32, %s splitting" % (N, args.split), "rows": n, "nnz_A": z, "reps": args.reps}

    t, S = timed(lambda: rt.soc_classic(plan_off, vals_off, 0.25), args.reps)
    res["soc_classic_ms"] = t
    Sf = (S > 0).float()
    t, (cflag, rounds) = timed(lambda: rt.cf_split_pmis(plan_off, S, 1), args.reps)
    res["pmis_ms"], res["pmis_rounds"], res["pmis_coarse_fraction"] = t, rounds, float(cflag.mean().item())
    if args.split == "alternating":
        cflag = V.default_splitting(n, dev)
    diag = op.diag.reshape(-1).float().contiguous()
    t, w = timed(lambda: rt.direct_interp(plan_off, vals_off, Sf, diag, cflag), args.reps)
    res["direct_interp_ms"] = t
    mode = 1 if args.split == "pmis" else 0
    t, (pi, pv, nc) = timed(lambda: rt.interp_assemble(plan_off, w, cflag, mode), args.reps)
    res["prolongator_assembly_ms"], res["nnz_P"], res["coarse_rows"] = t, int(pv.numel()), nc
    plan_P = rt.Plan.from_coo(pi, n, nc)
    ti = torch.stack([pi[1], pi[0]]).contiguous()

    def transpose():
        p = rt.Plan.from_coo(ti, nc, n)
        return p, rt.get_vals(p, pv.view(-1, 1).clone())
    t, (plan_PT, vals_PT) = timed(transpose, args.reps)
    res["transpose_plan_ms"] = t
    t, (ap_i, ap_v) = timed(lambda: rt.spgemm(plan_A, vals_A, plan_P, pv), args.reps)
    res["spgemm_AP_ms"], res["nnz_AP"], res["spgemm_AP_info"] = t, int(ap_v.numel()), dict(rt.spgemm.last)
    plan_AP = rt.Plan.from_coo(ap_i, n, nc)
    t, (ac_i, ac_v) = timed(lambda: rt.spgemm(plan_PT, vals_PT, plan_AP, ap_v), args.reps)
    res["spgemm_PtAP_ms"], res["nnz_Ac"], res["spgemm_PtAP_info"] = t, int(ac_v.numel()), dict(rt.spgemm.last)
    res["galerkin_ms"] = res["spgemm_AP_ms"] + res["spgemm_PtAP_ms"]
    res["setup_total_ms"] = sum(res[k] for k in ("soc_classic_ms", "direct_interp_ms", "prolongator_assembly_ms",
                                                 "transpose_plan_ms", "galerkin_ms")) + (
        res["pmis_ms"] if args.split == "pmis" else 0.0)
    if not args.no_library:
        P = torch.sparse_coo_tensor(pi, pv, (n, nc), is_coalesced=True)
        t, Ac_lib = timed(lambda: torch.sparse.mm(P.t(), torch.sparse.mm(A, P)).coalesce(), args.reps)
        res["library_torch_sparse_mm_galerkin_ms"] = t
        mine = torch.sparse_coo_tensor(ac_i, ac_v, (nc, nc), is_coalesced=True)
        d = (mine - Ac_lib).coalesce().values().abs().max().item() if Ac_lib._nnz() else 0.0
        res["galerkin_max_abs_diff_vs_library"] = d
        res["nnz_Ac_library"] = int(Ac_lib._nnz())
    print(json.dumps(res))
    if args.profile:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            rt.spgemm(plan_A, vals_A, plan_P, pv)
            torch.cuda.synchronize()
            rt.spgemm(plan_PT, vals_PT, plan_AP, ap_v)
            torch.cuda.synchronize()
        rows = sorted((e.time_range.start, e.name[:100], e.device_time) for e in prof.events()
                      if e.device_type == torch.autograd.DeviceType.CUDA)
        for st, nm, us in rows:
            print("%9.1f us  %8.1f us  %s" % (st - rows[0][0], us, nm))


if __name__ == "__main__":
    main()
