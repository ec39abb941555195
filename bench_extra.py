"""Sub-results of bench.py's one JSON line (key "extra"): the other BASELINE.json configs and the
north_star's 67 M-row single-GPU target, each with its own roofline figure, parity block and (N = 1)
a bounded CPU-port baseline.  Every function returns a plain dict; bench.py wraps each call so that
a failure is reported as {"error": ...} and never takes the headline down.

    L8192_layers    fused SpMV / Jacobi / Chebyshev steps on the 8192 x 8192 Laplacian (67.1 M rows), 1 GPU
    config3_power   PowerMethodGNN(100) on the 8192 x 8192 heat-equation FEM operator, N GPUs
    config4_amg     SOCClassicGNN + SOCSAGNN + DirectInterpGNN on the anisotropic periodic FEM operator (16.7 M rows)
    config5_vcycle  V-cycle (Jacobi smoother + direct interpolation), 8 RHS columns, 67 M-row Laplacian, N GPUs

Parity blocks never use oracle/: they compare against independent formulations evaluated with plain
torch on the GPU (fp64 shifted-slice stencils, element-wise torch restatements of the reference
formulas) or, at N > 1, against a single-GPU recompute on rank 0.
"""
import os
import time

import torch

TOL32 = 1e-5      # north_star tolerance for fp32 outputs


# ------------------------------------------------------------------------------------ helpers
def hashed_uniform(i0, i1, seed, device, cols=1, dtype=torch.float32):
    """Deterministic pseudo-random [i1 - i0, cols] in [0, 1) that depends only on the GLOBAL index
    (partition-independent: every rank generates exactly its slab of the same global vector)."""
    i = torch.arange(i0, i1, dtype=torch.int64, device=device).view(-1, 1) * cols + \
        torch.arange(cols, dtype=torch.int64, device=device).view(1, -1)
    x = (i * 2654435761 + (seed + 1) * 40503) & 0xFFFFFFFF
    x = x ^ (x >> 16)
    x = (x * 0x45D9F3B) & 0xFFFFFFFF
    x = x ^ (x >> 16)
    x = (x * 0x45D9F3B) & 0xFFFFFFFF
    x = x ^ (x >> 16)
    return (x.to(torch.float64) / 4294967296.0).to(dtype)


def timed_ms(fn, reps=10, warm=3):
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


def roof(ms, alg_bytes, z, peak, index_bytes=4):
    moved = alg_bytes - (4 - index_bytes) * z
    return {"ms": ms, "gnnz_per_s": z / ms / 1e6, "algorithmic_GBps": alg_bytes / ms / 1e6,
            "frac": alg_bytes / ms / 1e6 / peak, "moved_frac": moved / ms / 1e6 / peak,
            "frac_of_nominal_8TBps": alg_bytes / ms / 1e6 / 8000.0,       # north_star quotes ~8 TB/s
            "bytes_per_launch": int(alg_bytes), "index_bytes_streamed": index_bytes}


def relerr(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float(((a - b).norm() / b.norm().clamp_min(1e-300)).item())


def grid_stencil_5pt(x):
    """(A x) for laplacianfun_torch(N) (UtilsGNN.py:53-67: diag -4, neighbours +1, Dirichlet truncation)
    on an [N, N, k] fp64 grid, written with shifted slices -- independent of every CSR code path."""
    y = -4.0 * x
    y[:, 1:] += x[:, :-1]
    y[:, :-1] += x[:, 1:]
    y[1:, :] += x[:-1, :]
    y[:-1, :] += x[1:, :]
    return y


def grid_stencil_9pt(x, c0, ce, cn, cc):
    """9-point constant stencil with zero (eliminated Dirichlet) boundary on an [N, N] fp64 grid."""
    y = c0 * x
    y[:, 1:] += ce * x[:, :-1]
    y[:, :-1] += ce * x[:, 1:]
    y[1:, :] += cn * x[:-1, :]
    y[:-1, :] += cn * x[1:, :]
    y[1:, 1:] += cc * x[:-1, :-1]
    y[1:, :-1] += cc * x[:-1, 1:]
    y[:-1, 1:] += cc * x[1:, :-1]
    y[:-1, :-1] += cc * x[1:, 1:]
    return y


def cores():
    return os.cpu_count() or 1


def _free():
    import gc
    gc.collect()
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------ L8192 layers
def l8192_layers(G, dev, peak, N=8192):
    """north_star target: fused SpMV / Jacobi / Chebyshev on the 67 M-row 5-point Laplacian at 1 GPU."""
    rt = G.runtime
    n = N * N
    ei, ev = G.generators.laplacian_2d(N, torch.float32, dev)
    ei, ev = ei.contiguous(), ev.contiguous()
    plan = G.Plan.from_coo(ei, n)
    vals = rt.get_vals(plan, ev)
    z, s, ib = plan.nnz, 4, plan.index_bytes
    x = hashed_uniform(0, n, 1, dev)
    b = hashed_uniform(0, n, 2, dev)
    y, y2, r = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    diag = torch.full((n,), -4.0, device=dev)
    w = torch.tensor([0.7], device=dev)
    base = z * (4 + s) + 4 * (n + 1)
    out = {"workload": "L%d: %dx%d 5-point Laplacian, fp32, k = 1, 1 GPU" % (N, N, N), "rows": n, "nnz": z,
           "kernels": {}}
    k = out["kernels"]
    k["spmm"] = roof(timed_ms(lambda: rt.spmm(plan, vals, x, y)), base + 2 * n * s, z, peak, ib)
    k["jacobi"] = roof(timed_ms(lambda: rt.jacobi(plan, vals, diag, b, x, y, w)), base + 4 * n * s, z, peak, ib)
    k["cheby_first"] = roof(timed_ms(lambda: rt.cheby_first(plan, vals, b, x, y, r, y2, w)), base + 5 * n * s, z, peak, ib)
    k["cheby_next"] = roof(timed_ms(lambda: rt.cheby_next(plan, vals, x, y, r, y2, w, w, w)), base + 6 * n * s, z, peak, ib)
    xa, xb = x.clone(), torch.empty_like(x)
    ms10 = timed_ms(lambda: rt.jacobi_sweeps(plan, vals, diag, b, xa, xb, w, 10), reps=3, warm=1) / 10
    k["jacobi_10_sweeps_one_launch"] = roof(ms10, base + 4 * n * s, z, peak, ib)
    out["target"] = {"north_star": ">= 0.70 of the HBM roofline for fused SpMV / Jacobi / Chebyshev on this operator",
                     "min_frac": min(k[q]["frac"] for q in ("spmm", "jacobi", "cheby_first", "cheby_next")),
                     "min_moved_frac": min(k[q]["moved_frac"] for q in ("spmm", "jacobi", "cheby_first", "cheby_next"))}
    # ---- parity at full size against independent formulas
    ones = torch.ones(n, 1, device=dev)
    idx = torch.arange(n, device=dev)
    gy, gx = idx // N, idx % N
    expected = -(((gy == 0) | (gy == N - 1)).float() + ((gx == 0) | (gx == N - 1)).float())
    rowsum_exact = bool(torch.equal(rt.spmm(plan, vals, ones).view(-1), expected))
    del ones, idx, gy, gx, expected
    yx = rt.spmm(plan, vals, x)
    ref = grid_stencil_5pt(x.double().view(N, N, 1)).view(n, 1)
    e_spmm = relerr(yx, ref)
    xo = rt.jacobi(plan, vals, diag, b, x, torch.empty_like(x), w)
    ref_j = x.double() + (0.7 * (b.double() - ref)) / -4.0
    e_jac = relerr(xo, ref_j)
    # Chebyshev degree 4 vs the same recurrence in fp64 on the grid (ChebyGNN.py:117-283 op order)
    c_, d_ = -3.4, -4.0
    v, _, _ = G.ChebyGNN.ChebyRelaxGNN(4)(torch.cat([b, x], 1), ei, ev, torch.tensor([c_, d_]))
    xs, bs = x.double().view(N, N, 1), b.double().view(N, N, 1)
    rr = bs - grid_stencil_5pt(xs)
    alpha = 1.0 / d_
    p = rr.clone()
    xs = xs + alpha * p
    beta = 0.0
    for it in range(2, 5):
        rr = rr - alpha * grid_stencil_5pt(p)
        beta = 0.5 * (c_ * alpha) ** 2 if it == 2 else ((c_ * alpha) / 2) ** 2
        alpha = 1.0 / (d_ - beta / alpha)
        p = rr + beta * p
        xs = xs + alpha * p
    e_cheb = relerr(v[:, 1:2], xs.view(n, 1))
    out["parity"] = {"ok": bool(rowsum_exact and max(e_spmm, e_jac, e_cheb) <= TOL32), "tolerance": TOL32,
                     "row_sums_exact": rowsum_exact, "spmm_rel_err": e_spmm, "jacobi_rel_err": e_jac,
                     "chebyshev4_rel_err": e_cheb,
                     "against": "fp64 shifted-slice stencil formulations of A x, the Jacobi update and the degree-4 "
                                "Chebyshev recurrence on the [N, N] grid (no CSR code involved); A*1 vs the analytic row sums"}
    del plan, vals, ei, ev, x, b, y, y2, r, xa, xb, yx, ref, xo, ref_j, v, xs, bs, rr, p
    _free()
    return out


# ------------------------------------------------------------------------------------ config 3
HEAT_C = (8.0 / 3.0, -1.0 / 3.0, -1.0 / 3.0, -1.0 / 3.0)     # centre 8/3, eight neighbours -1/3


def _grid_power(b0_grid, iters):
    b = b0_grid.double()
    nrm = None
    for _ in range(iters):
        b = grid_stencil_9pt(b, *HEAT_C)
        nrm = b.norm()
        b = b / nrm
    ab = grid_stencil_9pt(b, *HEAT_C)
    return float(((b * ab).sum() / (b * b).sum()).item()), float(nrm.item()), b


def config3_power(G, dev, rank, world, peak, N=8192, iters=100, cpu_baseline=True, dtype=torch.float32):
    rt = G.runtime
    from glab_b200 import dist as gd
    import torch.distributed as dist
    n = N * N
    part = gd.RowPartition(n, world, align=256)
    r0, r1 = part.bounds(rank)
    t0 = time.perf_counter()
    f64 = dtype == torch.float64
    tol = 1e-12 if f64 else TOL32
    ei, ev = G.generators.heat_fem_2d((N + 1, N + 1), (1.0, 1.0), dtype, dev, rows=(r0, r1))
    ei, ev = ei.contiguous(), ev.contiguous()
    b0 = hashed_uniform(r0, r1, 3, dev).to(dtype)          # the same fp32-representable start vector in both precisions
    va = torch.cat([b0, torch.zeros_like(b0)], 1)
    ea = torch.cat([ev, torch.zeros_like(ev)], 1)
    g0 = torch.zeros(3, device=dev, dtype=dtype)
    layer = G.PowerMethodGNN.PowerMethodGNN(iters)
    graph = ei if world == 1 else gd.PartitionedGraph(ei, n, part, rank, world)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0

    def run():
        return layer(va, graph, ea, g0, None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    v, e, g = run()                       # warm-up (builds the plan / partitioned operator)
    del e
    times = []
    for _ in range(3):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        v, e, g = run()
        b.record()
        barrier()
        times.append(a.elapsed_time(b))
        del e
    ms = min(times)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        zt = torch.tensor([ei.shape[1]], dtype=torch.int64, device=dev)
        dist.all_reduce(zt)
        z = int(zt.item())
    else:
        z = int(ei.shape[1])
    spmvs = iters + 1
    s = 8 if f64 else 4
    sfx = "f64" if f64 else "f32"
    out = {"workload": "H%d: PowerMethodGNN(%d) on the %dx%d heat-equation FEM operator (9-point), %s"
                       % (N, iters, N, N, "fp64" if f64 else "fp32"),
           "api": "PowerMethodGNN(%d).forward(vertex_attr, %s, edge_attr, g) incl. the returned edge_attr column"
                  % (iters, "edgeij_pair" if world == 1 else "dist.PartitionedGraph"),
           "n_gpus": world, "rows": n, "nnz": z, "ms_total": ms, "ms_per_iteration": ms / spmvs,
           "value_nnz_per_s": spmvs * z / (ms * 1e-3), "lambda": float(g[2].item()), "norm": float(g[0].item()),
           "setup_s": t_setup}
    # roofline of the dominant kernel: glab_power_step on this rank's rows (kernel-only loop)
    if world == 1:
        plan = rt.get_plan(ei, n)
        vals = rt.get_vals(plan, ea, 0, dtype)
        xa, xb = b0.view(-1).clone(), torch.empty(n, device=dev, dtype=dtype)
        ss = torch.zeros(4, dtype=torch.float64, device=dev)
        ms_k = timed_ms(lambda: rt.power_step(plan, vals, xa, xb, None, ss[0:2]), reps=10)
        out["roofline"] = dict(kernel="glab_power_step_" + sfx, bound="hbm", peak=peak, unit="GB/s",
                               **roof(ms_k, z * (4 + s) + 4 * (n + 1) + 2 * n * s, z, peak,
                                      4 if f64 else plan.index_bytes))   # fp64 9-point kernels keep int32 indices
        del plan, vals, xa, xb
    else:
        nl, zl = r1 - r0, int(ei.shape[1])
        per_iter = ms / spmvs
        out["roofline"] = dict(kernel="glab_power_step_halo_%s (whole iteration incl. the 2-scalar reduction over ranks)" % sfx,
                               bound="hbm", peak=peak, unit="GB/s",
                               **roof(per_iter, zl * (4 + s) + 4 * (nl + 1) + 2 * nl * s, zl, peak, 4))
    # ---- parity: independent fp64 power iteration on the grid (rank 0), lambda / norm / iterate
    if world > 1:
        full = [torch.empty(part.bounds(q)[1] - part.bounds(q)[0], device=dev, dtype=dtype) for q in range(world)] \
            if rank == 0 else None
        dist.gather(v[:, 0].contiguous(), full, dst=0)
    else:
        full = [v[:, 0]]
    if rank == 0:
        b0_all = hashed_uniform(0, n, 3, dev)
        lam_ref, nrm_ref, b_ref = _grid_power(b0_all.view(N, N), iters)
        e_vec = relerr(torch.cat(full), b_ref)
        e_lam = abs(float(g[2].item()) - lam_ref) / abs(lam_ref)
        e_nrm = abs(float(g[0].item()) - nrm_ref) / abs(nrm_ref)
        out["parity"] = {"ok": bool(max(e_lam, e_nrm, e_vec) <= tol), "tolerance": tol, "lambda_rel_err": e_lam,
                         "norm_rel_err": e_nrm, "iterate_rel_err": e_vec, "lambda_reference": lam_ref,
                         "against": "independent fp64 power iteration applying the closed-form 9-point stencil with "
                                    "shifted grid slices on rank 0 (all %d rows)" % n}
        del b0_all, b_ref
    if rank == 0 and world == 1 and cpu_baseline:
        out["cpu_baseline"] = _cpu_power(1024, 5)
    del v, g, va, ea, ei, ev, b0, graph, full
    _free()
    return out


def _cpu_power(N, iters):
    from oracle import port
    torch.set_num_threads(cores())
    n = N * N
    ei, ev = _cpu_heat(N)
    b0 = torch.rand(n, 1, generator=torch.Generator().manual_seed(24601))
    va = torch.cat([b0, torch.zeros_like(b0)], 1)
    ea = torch.cat([ev, torch.zeros_like(ev)], 1)
    t0 = time.perf_counter()
    port.power_method(iters, va, ei, ea, torch.zeros(3))
    dt = time.perf_counter() - t0
    z = ei.shape[1]
    return {"value": (iters + 1) * z / dt, "unit": "nnz/s", "cores": cores(), "kind": "port",
            "sample": "PowerMethodGNN(%d) on the %dx%d heat-equation operator (%d nnz), oracle/port.py" % (iters, N, N, z)}


def _cpu_heat(N):
    import glab_b200 as G
    ei, ev = G.generators.heat_fem_2d((N + 1, N + 1), (1.0, 1.0), torch.float32, "cpu")
    return ei.contiguous(), ev.contiguous()


# ------------------------------------------------------------------------------------ config 4
def config4_amg(G, dev, peak, N=4096, cpu_baseline=True):
    """SOCClassicGNN + SOCSAGNN + DirectInterpGNN on the anisotropic periodic FEM operator."""
    rt = G.runtime
    dt = torch.float32
    n = N * N
    ei, ev = G.generators.constant_diffusion_fem(1.0, 0.01, N, dtype=dt, device=dev)
    diag = G.generators.diagonal_of(ei, ev, n)
    keep = ei[0] != ei[1]
    eo, ao = ei[:, keep].contiguous(), ev[keep].contiguous()
    del ei, ev, keep
    z, s = int(eo.shape[1]), 4
    theta = 0.25
    split = torch.zeros(n, 1, dtype=dt, device=dev)
    split[0::2] = 1
    # through the drop-in layers (what a user calls)
    S = G.SOCClassicGNN.SOCClassicGNN(theta)(torch.zeros(n, 1, dtype=dt, device=dev), eo, ao)
    sa = G.SOCSAGNN.SOCSAGNN()(diag, eo, ao)[1][:, 1]
    Sflag = (S.reshape(-1, 1) > 0).to(dt)
    w = G.DirectInterpGNN.DirectInterpGNN()(torch.hstack([diag, split]), eo, torch.hstack([ao, Sflag]))
    plan = rt.get_plan(eo, n)
    vals = rt.get_vals(plan, ao)
    d1 = diag.reshape(-1).contiguous()
    S1, c1 = Sflag.reshape(-1).contiguous(), split.reshape(-1).contiguous()
    ib = getattr(plan, "edge_index_bytes", 4)
    out = {"workload": "D%d: SOCClassicGNN(0.25) + SOCSAGNN + DirectInterpGNN on the %dx%d periodic anisotropic FEM "
                       "operator (alpha = 1, beta = 0.01), fp32" % (N, N, N),
           "rows": n, "edges_off_diagonal": z, "strong_edges": int((S > 0).sum().item()), "kernels": {}}
    k = out["kernels"]
    k["soc_classic"] = roof(timed_ms(lambda: rt.soc_classic(plan, vals, theta)), 2 * z * s + 4 * (n + 1), z, peak, 4)
    k["soc_sa"] = roof(timed_ms(lambda: rt.soc_sa(plan, vals, d1)), z * (4 + 2 * s) + 4 * (n + 1) + n * s, z, peak, ib)
    k["direct_interp"] = roof(timed_ms(lambda: rt.direct_interp(plan, vals, S1, d1, c1)),
                              z * (4 + 3 * s) + 4 * (n + 1) + 3 * n * s, z, peak, ib)
    total_ms = sum(k[q]["ms"] for q in k)
    out["value_edges_per_s"] = 3 * z / (total_ms * 1e-3)
    # ---- parity: element-wise torch restatement of the reference formulas on the GPU, bit for bit
    # (every row of this operator has exactly 8 off-diagonal edges, in ascending column order)
    A8 = ao.view(n, 8)
    v = (-1 * A8).max(1, keepdim=True).values                                        # SOCClassicGNN.py:69
    S_ref = torch.relu(((-1 * A8) / v) - theta).reshape(-1)                          # :125
    ok_S = bool(torch.equal(S, S_ref))
    dj = diag.reshape(-1)[eo[1]].view(n, 8)
    sa_ref = ((A8 * A8) / (diag.view(n, 1) * dj)).reshape(-1)                        # SOCSAGNN.py:67
    ok_sa = bool(torch.equal(sa, sa_ref))
    S8 = Sflag.view(n, 8)
    C8 = split.reshape(-1)[eo[1]].view(n, 8)
    num = torch.zeros(n, dtype=dt, device=dev)
    den = torch.zeros(n, dtype=dt, device=dev)
    for j in range(8):                                                               # sequential, like scatter_add_
        num = num + A8[:, j]
        den = den + (A8[:, j] * S8[:, j]) * C8[:, j]                                 # DirectInterpGNN.py:89-94
    alpha = (1 / diag.reshape(-1)) * (num / den)                                     # :127
    w_ref = ((1 - split.reshape(-1)).view(n, 1) * ((-A8) * alpha.view(n, 1))).reshape(-1)   # :150
    same_w = (w == w_ref) | (torch.isnan(w) & torch.isnan(w_ref))
    ok_w = bool(same_w.all())
    out["parity"] = {"ok": bool(ok_S and ok_sa and ok_w and out["strong_edges"] == 6 * n), "bit_exact": True,
                     "soc_classic_values_and_mask": ok_S, "soc_sa_values": ok_sa,
                     "direct_interp_values_and_nan_pattern": ok_w, "nan_weights": int(torch.isnan(w).sum().item()),
                     "against": "the reference's element-wise formulas restated with plain torch ops on the GPU (fp32, "
                                "sequential per-row sums), all %d edges, bit for bit incl. the NaN pattern" % z}
    if cpu_baseline:
        out["cpu_baseline"] = _cpu_amg(512)
    del A8, v, S_ref, dj, sa_ref, S8, C8, num, den, alpha, w_ref, same_w, S, sa, w, plan, vals, eo, ao
    _free()
    return out


def _cpu_amg(N):
    from oracle import port
    import glab_b200 as G
    torch.set_num_threads(cores())
    dt = torch.float32
    n = N * N
    ei, ev = G.generators.constant_diffusion_fem(1.0, 0.01, N, dtype=dt, device="cpu")
    diag = G.generators.diagonal_of(ei, ev, n)
    keep = ei[0] != ei[1]
    eo, ao = ei[:, keep].contiguous(), ev[keep].contiguous()
    split = torch.zeros(n, 1, dtype=dt)
    split[0::2] = 1
    t0 = time.perf_counter()
    S = port.soc_classic(0.25, torch.zeros(n, 1, dtype=dt), eo, ao)
    port.soc_sa(diag, eo, ao)
    port.direct_interp(torch.hstack([diag, split]), eo, torch.hstack([ao, (S.reshape(-1, 1) > 0).to(dt)]))
    dtm = time.perf_counter() - t0
    z = eo.shape[1]
    return {"value": 3 * z / dtm, "unit": "edges/s", "cores": cores(), "kind": "port",
            "sample": "SOC classic + SA + direct interpolation on the %dx%d periodic operator (%d edges), oracle/port.py" % (N, N, z)}


# ------------------------------------------------------------------------------------ config 5
def config5_vcycle(G, dev, rank, world, peak, N=8192, k=8, cycles=3, cpu_baseline=True, multilevel=None):
    """V-cycle with Jacobi smoother + direct interpolation, k RHS columns, on the N x N Laplacian."""
    import torch.distributed as dist
    V = G.VCycle
    n = N * N
    t0 = time.perf_counter()
    out = {"workload": "L%d: V-cycle (3 + 3 Jacobi sweeps w = 0.7, classical SOC 0.25, direct interpolation, Galerkin "
                       "coarse operators), %d RHS columns, fp32" % (N, k), "n_gpus": world, "rows": n, "rhs_columns": k}
    if world == 1:
        ei, ev = G.UtilsGNN.laplacianfun_torch(N, device=dev)
        A = torch.sparse_coo_tensor(ei, ev.flatten().float(), (n, n))
        z = int(ei.shape[1])
        del ei, ev
        b = hashed_uniform(0, n, 5, dev, cols=k)
        x = torch.zeros(n, k, device=dev)
        variants = {}
        for name in (["two_grid", "multilevel"] if multilevel is not False and hasattr(V, "runVCycleML") else ["two_grid"]):
            if name == "two_grid":
                def cyc(xx):
                    return V.runVCycle(A, b, xx, 3, 3, 5, True)
            else:
                def cyc(xx):
                    return V.runVCycleML(A, b, xx, 3, 3)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            xx = cyc(x)                                   # builds + caches the hierarchy
            torch.cuda.synchronize()
            t_first = time.perf_counter() - t1
            norms = [torch.norm(V.runResidual(A, b, x), dim=0), torch.norm(V.runResidual(A, b, xx), dim=0)]
            ts = []
            for _ in range(cycles):
                a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                xx = cyc(xx)
                e.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(e))
                norms.append(torch.norm(V.runResidual(A, b, xx), dim=0))
            red = [float((norms[i + 1] / norms[i]).max().item()) for i in range(len(norms) - 1)]
            info = V.hierarchy_info(A, name) if hasattr(V, "hierarchy_info") else {}
            work = info.get("spmv_nnz_per_cycle")
            variants[name] = {"ms_per_cycle": min(ts), "first_cycle_incl_setup_ms": t_first * 1e3,
                              "residual_norm_col0": [float(v_[0]) for v_ in norms],
                              "worst_column_reduction_per_cycle": red, **info}
            if work:
                variants[name]["value_nnz_columns_per_s"] = work * k / (min(ts) * 1e-3)
            if name == "two_grid":
                x_tg = xx
            else:
                x_ml = xx
        out["variants"] = variants
        out["nnz_A"] = z
        # ---- parity (two-grid cycle): independent fp64 residual of the returned x, and column == single-RHS run
        r_gpu = V.runResidual(A, b, x_tg)
        r_ref = b.double().view(N, N, k) - grid_stencil_5pt(x_tg.double().view(N, N, k))
        e_res = relerr(r_gpu, r_ref.view(n, k))
        x1 = torch.zeros(n, 1, device=dev)
        b1 = b[:, 3:4].contiguous()
        for _ in range(cycles + 1):
            x1 = V.runVCycle(A, b1, x1, 3, 3, 5, True)
        col_equal = bool(torch.equal(x_tg[:, 3:4], x1))
        out["parity"] = {"ok": bool(e_res <= TOL32 and col_equal), "tolerance": TOL32, "residual_rel_err": e_res,
                         "column_equals_single_rhs_run_bitwise": col_equal,
                         "against": "fp64 shifted-slice residual b - A x of the returned iterate (all rows, all columns); "
                                    "column 3 of the 8-column run vs the same cycles run on that column alone"}
        if "multilevel" in variants:
            # (the iterate of a random right-hand side is ~N^2 larger than b: every row sum -4 x_i + sum x_j
            # cancels ~7 digits, so the fp32 residual is compared on the scale |A| |x| = 8 |x| the sum works on)
            r_gpu = V.runResidual(A, b, x_ml)
            r_ref = b.double().view(N, N, k) - grid_stencil_5pt(x_ml.double().view(N, N, k))
            e_ml = float(((r_gpu.double() - r_ref.view(n, k)).norm() / (8.0 * x_ml.double().norm())).item())
            rate = max(variants["multilevel"]["worst_column_reduction_per_cycle"][1:])
            out["parity"]["multilevel_residual_err_rel_to_normA_normx"] = e_ml
            out["parity"]["multilevel_worst_reduction_after_first_cycle"] = rate
            out["parity"]["ok"] = bool(out["parity"]["ok"] and e_ml <= TOL32 and rate <= 0.5)
            del x_ml
        if cpu_baseline:
            out["cpu_baseline"] = _cpu_vcycle(64)
        del A, b, x, xx, x_tg, r_gpu, r_ref, x1, b1
        G.VCycle._operators.clear()
    else:
        from glab_b200.dist_vcycle import DistTwoGrid
        from glab_b200.dist_multilevel import DistMultilevel
        ei, ev = G.UtilsGNN.laplacianfun_torch(N, device=dev)
        A_full = torch.sparse_coo_tensor(ei, ev.flatten().float(), (n, n))
        engine = os.environ.get("GLAB_DIST_ENGINE", "peer")
        variants, par = {}, {}
        for name in (["two_grid", "multilevel"] if multilevel is not False else ["two_grid"]):
            t1 = time.perf_counter()
            if name == "two_grid":
                tg = DistTwoGrid(ei, ev, k, rank, world, engine=engine)
                part0 = tg.fine
                work = 7 * tg.nnz["A"] + 2 * tg.nnz["P"] + 4 * tg.nnz["Ac"]
                extra_info = {"nnz_A": tg.nnz["A"], "nnz_P": tg.nnz["P"], "nnz_Ac": tg.nnz["Ac"]}
            else:
                tg = DistMultilevel(A_full, k, rank, world, engine=engine)
                part0 = tg.parts[0]
                extra_info = tg.info()
                work = extra_info["spmv_nnz_per_cycle"]
            torch.cuda.synchronize()
            setup_s = time.perf_counter() - t1
            f0, f1 = part0.bounds(rank)
            b = hashed_uniform(f0, f1, 5, dev, cols=k)
            tg.load_x(torch.zeros(f1 - f0, k, device=dev))

            def rnorm():
                r = tg.residual_local(b)
                s_ = (r.double() ** 2).sum(0)
                dist.all_reduce(s_)
                return torch.sqrt(s_)

            norms = [rnorm()]
            tg.cycle(b)
            norms.append(rnorm())
            ts = []
            for _ in range(cycles):
                torch.cuda.synchronize()
                dist.barrier()
                a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                tg.cycle(b)
                e.record()
                torch.cuda.synchronize()
                dist.barrier()
                ts.append(a.elapsed_time(e))
                norms.append(rnorm())
            t = torch.tensor([min(ts)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            variants[name] = {"ms_per_cycle": ms, "spmv_nnz_per_cycle": int(work),
                              "value_nnz_columns_per_s": work * k / (ms * 1e-3), "setup_s": setup_s,
                              "setup_breakdown_s": getattr(tg, "setup_times", None),
                              "residual_norm_col0": [float(v_[0]) for v_ in norms],
                              "worst_column_reduction_per_cycle":
                                  [float((norms[i + 1] / norms[i]).max().item()) for i in range(len(norms) - 1)],
                              **extra_info}
            # ---- parity: the partitioned iterate after these cycles == the single-GPU cycle on rank 0, bit for bit
            xl = tg.x_local().contiguous()
            full = [torch.empty(part0.bounds(q)[1] - part0.bounds(q)[0], k, device=dev) for q in range(world)] \
                if rank == 0 else None
            dist.gather(xl, full, dst=0)
            tg.check()
            tg.close()
            del tg
            _free()
            if rank == 0:
                b_all = hashed_uniform(0, n, 5, dev, cols=k)
                xr = torch.zeros(n, k, device=dev)
                for _ in range(cycles + 1):
                    xr = G.VCycle.runVCycle(A_full, b_all, xr, 3, 3, 5, True) if name == "two_grid" else \
                        G.VCycle.runVCycleML(A_full, b_all, xr, 3, 3)
                got = torch.cat(full)
                par[name] = {"ok": bool(torch.equal(got, xr)), "bit_exact": True,
                             "max_abs": float((got - xr).abs().max().item()),
                             "against": "VCycle.%s on ONE GPU (rank 0) for the same %d cycles, all rows and columns"
                                        % ("runVCycle" if name == "two_grid" else "runVCycleML", cycles + 1)}
                del b_all, xr, got
                G.VCycle._operators.clear()
                _free()
            dist.barrier()
        out["variants"] = variants
        if rank == 0:
            out["parity"] = {"ok": all(p_["ok"] for p_ in par.values()), **par}
        del ei, ev, b, A_full
    _free()
    return out


def _cpu_vcycle(N):
    from oracle import port
    torch.set_num_threads(cores())
    n = N * N
    ei, ev = port.laplacian_2d(N)
    b = torch.rand(n, 1, generator=torch.Generator().manual_seed(24601))
    x = torch.zeros(n, 1)
    split = torch.zeros(n)
    split[0::2] = 1
    t0 = time.perf_counter()
    port.two_grid_vcycle(ei, ev, b, x, split)
    dt = time.perf_counter() - t0
    z = ei.shape[1]
    work = 7 * z + 2 * (z // 2) + 4 * (z // 2)
    return {"value": work / dt, "unit": "nnz*columns/s", "cores": cores(), "kind": "port",
            "sample": "one two-grid V-cycle (incl. its setup, as the reference recomputes it every call) on the %dx%d "
                      "Laplacian, 1 RHS column, oracle/port.py; nnz of P and A_c estimated as nnz(A)/2" % (N, N)}
