#!/usr/bin/env python
"""Per-kernel roofline table on one B200: every fused layer step and AMG setup kernel on the
BASELINE.json operators, timed with CUDA events (after warm-up, inputs larger than L2),
algorithmic bytes from SURVEY.md section 8d.  Writes JSON lines to stdout.

    python scripts/bench_kernels.py [--quick] > gpurun_out/kernels.jsonl
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import glab_b200 as G  # noqa: E402

rt = G.runtime
PEAK = 6499.0
if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


INDEX_BYTES = 4   # bytes of column index the pipeline kernels stream per nonzero for the current plan


def report(name, op, ms, nbytes, z, extra=None, streams_index=True):
    """alg_MB / GBps / frac_of_measured_peak use SURVEY 8d's algorithmic bytes (4-byte column indices);
    moved_frac subtracts the index bytes the kernel does not stream when the plan holds 16-bit
    row-relative indices."""
    d = {"kernel": name, "operator": op, "ms": round(ms, 4), "alg_MB": round(nbytes / 1e6, 1),
         "GBps": round(nbytes / ms / 1e6, 1), "frac_of_measured_peak": round(nbytes / ms / 1e6 / PEAK, 3),
         "Gnnz_per_s": round(z / ms / 1e6, 1)}
    if streams_index:
        moved = nbytes - (4 - INDEX_BYTES) * z
        d.update(index_bytes=INDEX_BYTES, moved_frac=round(moved / ms / 1e6 / PEAK, 3))
    if extra:
        d.update(extra)
    print(json.dumps(d), flush=True)


def layer_kernels(op_name, ei, ev, dt, ks=(1,)):
    dev = ei.device
    n = int(ei[0].max().item()) + 1
    t0 = time.perf_counter()
    plan = G.Plan.from_coo(ei, n)
    torch.cuda.synchronize()
    t_plan = (time.perf_counter() - t0) * 1e3
    vals = rt.get_vals(plan, ev)
    z, s = plan.nnz, ev.element_size()
    base = z * (4 + s) + 4 * (n + 1)
    global INDEX_BYTES
    INDEX_BYTES = plan.index_bytes
    report("plan_create (COO int64 -> CSR int32)", op_name, t_plan, z * 16 + z * 4 + 4 * n, z, streams_index=False)
    ws = torch.zeros(4, dtype=torch.float64, device=dev)
    for k in ks:
        x = torch.rand(n, k, dtype=dt, device=dev)
        b = torch.rand(n, k, dtype=dt, device=dev)
        y = torch.empty_like(x)
        y2 = torch.empty_like(x)
        r = torch.rand(n, k, dtype=dt, device=dev)
        diag = torch.full((n,), -4.0, dtype=dt, device=dev)
        w = torch.tensor([0.7], dtype=dt, device=dev)
        tag = " k=%d %s" % (k, str(dt)[6:])
        report("spmm" + tag, op_name, timeit(lambda: rt.spmm(plan, vals, x, y)), base + 2 * n * k * s, z)
        report("residual" + tag, op_name, timeit(lambda: rt.residual(plan, vals, x, b, y)), base + 3 * n * k * s, z)
        report("jacobi" + tag, op_name, timeit(lambda: rt.jacobi(plan, vals, diag, b, x, y, w)),
               base + (3 * k + 1) * n * s, z)
        report("cheby_first" + tag, op_name, timeit(lambda: rt.cheby_first(plan, vals, b, x, y, r, y2, w)),
               base + 5 * n * k * s, z)
        report("cheby_next" + tag, op_name, timeit(lambda: rt.cheby_next(plan, vals, x, y, r, y2, w, w, w)),
               base + 6 * n * k * s, z)
        if k == 1:
            report("power_step" + tag, op_name, timeit(lambda: rt.power_step(plan, vals, x, y, None, ws[0:2])),
                   base + 2 * n * s, z)
            report("rayleigh" + tag, op_name, timeit(lambda: rt.rayleigh(plan, vals, x, y, y2, None, ws[2:4])),
                   base + 3 * n * s, z)
            out = torch.empty(z, 2, dtype=dt, device=dev)
            report("edge_messages (optional output)" + tag, op_name,
                   timeit(lambda: rt.edge_messages(plan, vals, x, out, 1), reps=5), z * (4 + 2 * s) + n * s, z,
                   streams_index=False)
            del out
        del x, b, y, y2, r
    return plan


def amg_kernels(op_name, ei, ev, dt):
    dev = ei.device
    n = int(ei[0].max().item()) + 1
    diag = G.generators.diagonal_of(ei, ev, n).reshape(-1).contiguous()
    eo, ao = G.UtilsGNN.remove_diag_entries(ei, ev)
    eo, ao = eo.contiguous(), ao.contiguous()
    plan = G.Plan.from_coo(eo, n)
    vals = rt.get_vals(plan, ao)
    z, s = plan.nnz, ev.element_size()
    tag = " " + str(dt)[6:]
    global INDEX_BYTES
    INDEX_BYTES = 4                     # the per-edge-output kernels stream int32 column indices
    report("soc_classic" + tag, op_name, timeit(lambda: rt.soc_classic(plan, vals, 0.25)), 2 * z * s + 4 * (n + 1), z)
    report("soc_sa" + tag, op_name, timeit(lambda: rt.soc_sa(plan, vals, diag)), z * (4 + 2 * s) + 4 * (n + 1) + n * s, z)
    S = (rt.soc_classic(plan, vals, 0.25) > 0).to(dt)
    C = torch.zeros(n, dtype=dt, device=dev)
    C[0::2] = 1
    report("direct_interp" + tag, op_name, timeit(lambda: rt.direct_interp(plan, vals, S, diag, C)),
           z * (4 + 3 * s) + 4 * (n + 1) + 3 * n * s, z)
    strong = int((S > 0).sum().item())
    print(json.dumps({"info": "soc mask", "operator": op_name, "strong_edges": strong, "edges": z}), flush=True)


def main():
    quick = "--quick" in sys.argv
    only_amg = "--amg" in sys.argv
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    if only_amg:
        for dt in (torch.float32, torch.float64):
            ei, ev = G.generators.constant_diffusion_fem(1.0, 0.01, 4096, dt, dev)
            amg_kernels("D4096 anisotropic periodic FEM", ei, ev.contiguous(), dt)
            del ei, ev
            torch.cuda.empty_cache()
        ei, ev = G.generators.laplacian_2d(4096, torch.float32, dev)
        amg_kernels("L4096 5-pt Laplacian", ei, ev.contiguous(), torch.float32)
        return
    NL = 2048 if quick else 4096
    ei, ev = G.generators.laplacian_2d(NL, torch.float64, dev)
    layer_kernels("L%d 5-pt Laplacian" % NL, ei, ev.float().contiguous(), torch.float32, ks=(1, 8))
    layer_kernels("L%d 5-pt Laplacian" % NL, ei, ev.contiguous(), torch.float64, ks=(1,))
    del ei, ev
    torch.cuda.empty_cache()
    if not quick:
        ei, ev = G.generators.laplacian_2d(8192, torch.float64, dev)
        layer_kernels("L8192 5-pt Laplacian (67M rows)", ei, ev.float().contiguous(), torch.float32, ks=(1,))
        del ei, ev
        torch.cuda.empty_cache()
    NH = 2048 if quick else 8192
    ei, ev = G.generators.heat_fem_2d((NH + 1, NH + 1), (1.0, 1.0), torch.float32, dev)
    layer_kernels("H%d 9-pt heat-eqn FEM" % NH, ei, ev.contiguous(), torch.float32, ks=(1,))
    del ei, ev
    torch.cuda.empty_cache()
    ND = 1024 if quick else 4096
    ei, ev = G.generators.constant_diffusion_fem(1.0, 0.01, ND, torch.float32, dev)
    amg_kernels("D%d anisotropic periodic FEM" % ND, ei, ev.contiguous(), torch.float32)
    ei, ev = G.generators.constant_diffusion_fem(1.0, 0.01, ND, torch.float64, dev)
    amg_kernels("D%d anisotropic periodic FEM" % ND, ei, ev.contiguous(), torch.float64)


if __name__ == "__main__":
    main()
