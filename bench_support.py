"""Workload drivers used by bench.py (kept apart so bench.py stays a readable contract)."""
import os
import time

import torch


class _Staged:
    # staging depth of step_e2e: with 3 buffers the upload of step i+1 does not have to wait for the layer
    # calls of step i-1 (measured at 8 GPUs: the H2D stream idled 0.3 of every 2.15 ms with 2 buffers)
    NBUF = 3


class _Trace:
    """Optional per-step timeline of step_e2e (events on the three streams + host clock); a no-op
    unless the smoother's `trace` attribute is a list (bench.e2e_timeline sets it AFTER the timed region)."""

    def __init__(self, sink):
        self.sink = sink
        if sink is not None:
            self.ev, self.t0 = [], time.perf_counter()

    def mark(self, stream):
        if self.sink is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream)
            self.ev.append(e)

    def done(self):
        if self.sink is not None:
            self.sink.append((self.ev, self.t0, time.perf_counter()))


class SingleGpuSmoother(_Staged):
    """BASELINE config 2 on one GPU: 10 Jacobi sweeps + Chebyshev degree 4, fp32, k = 1."""

    N_JACOBI, CHEB_DEG, OMEGA, CHEB_C, CHEB_D = 10, 4, 0.7, -3.4, -4.0

    def __init__(self, G, N, dev):
        self.G, self.rt, self.dev = G, G.runtime, dev
        rt = self.rt
        n = N * N
        t0 = time.perf_counter()
        ei, ev = G.UtilsGNN.laplacianfun_torch(N, device=dev)
        self.ei = ei.contiguous()
        self.ev = ev.float().contiguous()           # cast like JacobiGNN.py:161
        del ei, ev
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        self.plan = G.get_plan(self.ei, n)
        self.vals = rt.get_vals(self.plan, self.ev)
        import os
        if os.environ.get("GLAB_L2_PERSIST", "0") == "1":      # experiment knob (L2-sized operators)
            self.vals = rt.adopt_vals(self.plan, self.vals)
            rt.l2_persist(self.plan, True)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        self.n_local = self.n = n
        self.nnz_local = self.nnz_global = self.plan.nnz
        self.setup_info = {"generate_ms": (t1 - t0) * 1e3, "plan_build_ms": (t2 - t1) * 1e3,
                           "plan_identity_perm": self.plan.identity, "max_row_nnz": self.plan.max_row_nnz,
                           "index_bytes_streamed": self.plan.index_bytes}
        self.index_bytes = self.plan.index_bytes
        from bench_extra import hashed_uniform
        self.N = N
        self.b_host = hashed_uniform(0, n, 1, "cpu").pin_memory()
        self.x_host = hashed_uniform(0, n, 2, "cpu").pin_memory()
        self.va_host = torch.cat([-4 * torch.ones(n, 1), self.b_host, self.x_host], 1).pin_memory()
        self.out_host = torch.empty(n, 1).pin_memory()
        self.diag = torch.full((n,), -4.0, device=dev)
        self.b = self.b_host.to(dev).contiguous()
        self.x0 = self.x_host.to(dev).contiguous()
        self.xa, self.xb = torch.empty_like(self.x0), torch.empty_like(self.x0)
        self.r, self.p, self.p2 = (torch.empty_like(self.x0) for _ in range(3))
        self.w = torch.tensor([self.OMEGA], device=dev)
        rows, self.g_out = G.ChebyGNN._recurrence(self.CHEB_DEG, torch.tensor([self.CHEB_C, self.CHEB_D]))
        self.table = torch.stack([torch.stack(r) for r in rows]).to(dev).contiguous()
        # operator-side inputs of the layer API (step-invariant)
        self.ea2 = torch.cat([self.ev, torch.zeros_like(self.ev)], 1)
        self.gw = torch.tensor(self.OMEGA).reshape(-1)
        self.gc = torch.tensor([self.CHEB_C, self.CHEB_D])
        self.jac = G.JacobiGNN.JacobiGNN()
        self.cheb = G.ChebyGNN.ChebyRelaxGNN(self.CHEB_DEG)
        self.h2d_bytes = self.va_host.numel() * 4
        self.d2h_bytes = self.out_host.numel() * 4

    def step_kernels(self, time_jacobi=False):
        rt, plan, vals = self.rt, self.plan, self.vals
        ev = None
        if time_jacobi:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        src, dst = self.x0, self.xa
        for i in range(self.N_JACOBI):
            rt.jacobi(plan, vals, self.diag, self.b, src, dst, self.w)
            src, dst = dst, (self.xb if dst is self.xa else self.xa)
        if time_jacobi:
            ev[1].record()
        t = self.table
        rt.cheby_first(plan, vals, self.b, src, dst, self.r, self.p, t[0, 1:2])
        p, p2 = self.p, self.p2
        for it in range(1, self.CHEB_DEG):
            rt.cheby_next(plan, vals, p, p2, self.r, dst, t[it, 0:1], t[it, 1:2], t[it, 2:3])
            p, p2 = p2, p
        self.result = dst
        return ev

    def layer_pass(self, va):
        """The drop-in layer calls on device-resident vertex_attr = [A_ii, b, x]."""
        x1 = self.jac(self.N_JACOBI, va, self.ei, self.ea2, self.gw)
        v, e, g = self.cheb(self.rt.pack([va[:, 1:2].contiguous(), x1]), self.ei, self.ev, self.gc)
        return v[:, 1:2]

    def parity(self):
        """The layer pass at full size against an independent fp64 formulation: 10 Jacobi sweeps and the
        degree-4 Chebyshev recurrence applied with shifted grid slices (no CSR code involved)."""
        from bench_extra import grid_stencil_5pt, relerr
        N, n, dev = self.N, self.n, self.dev
        va = self.va_host.to(dev)
        x1 = self.jac(self.N_JACOBI, va, self.ei, self.ea2, self.gw)
        v, e, g = self.cheb(self.rt.pack([va[:, 1:2].contiguous(), x1]), self.ei, self.ev, self.gc)
        del e
        b = va[:, 1].double().view(N, N, 1)
        x = va[:, 2].double().view(N, N, 1)
        for _ in range(self.N_JACOBI):
            x = x + (self.OMEGA * (b - grid_stencil_5pt(x))) / -4.0
        e_jac = relerr(x1, x.view(n, 1))
        c_, d_ = self.CHEB_C, self.CHEB_D
        r = b - grid_stencil_5pt(x)
        alpha = 1.0 / d_
        p = r.clone()
        x = x + alpha * p
        for it in range(2, self.CHEB_DEG + 1):
            r = r - alpha * grid_stencil_5pt(p)
            beta = 0.5 * (c_ * alpha) ** 2 if it == 2 else ((c_ * alpha) / 2) ** 2
            alpha = 1.0 / (d_ - beta / alpha)
            p = r + beta * p
            x = x + alpha * p
        e_x = relerr(v[:, 1:2], x.view(n, 1))
        e_r = relerr(v[:, 2:3], r.view(n, 1))
        return {"ok": bool(max(e_jac, e_x) <= 1e-5), "tolerance": 1e-5, "jacobi10_rel_err": e_jac,
                "chebyshev4_x_rel_err": e_x, "chebyshev4_r_rel_err_informative": e_r, "rows_checked": n,
                "against": "fp64 shifted-slice stencil recompute of the same 10 Jacobi sweeps + degree-4 Chebyshev "
                           "recurrence on the [N, N] grid (independent of every CSR code path)"}

    def _e2e_setup(self):
        """Double-buffered staging so that consecutive steps overlap their PCIe copies with compute:
        step i+1's H2D and step i-1's D2H run on copy streams while step i computes.  Every step
        still uploads its own inputs from pinned host memory and downloads its own result."""
        dev, n = self.dev, self.n
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.va_dev = [torch.empty(n, 3, device=dev) for _ in range(self.NBUF)]
        self.res_dev = [torch.empty(n, 1, device=dev) for _ in range(self.NBUF)]
        self.out_hosts = [torch.empty(n, 1).pin_memory() for _ in range(self.NBUF)]
        self.ev_in = [torch.cuda.Event() for _ in range(self.NBUF)]
        self.ev_comp = [torch.cuda.Event() for _ in range(self.NBUF)]
        self.ev_out = [torch.cuda.Event() for _ in range(self.NBUF)]
        self.e2e_i = 0

    def step_e2e(self):
        if not hasattr(self, "s_in"):
            self._e2e_setup()
        i = self.e2e_i % self.NBUF
        self.e2e_i += 1
        tr = _Trace(getattr(self, "trace", None))
        cur = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(self.ev_comp[i])            # the step that last read this buffer is done
            tr.mark(self.s_in)
            self.va_dev[i].copy_(self.va_host, non_blocking=True)   # H2D of this step's inputs
            tr.mark(self.s_in)
            self.ev_in[i].record(self.s_in)
        cur.wait_event(self.ev_in[i])
        cur.wait_event(self.ev_out[i])                       # this result buffer has been downloaded
        tr.mark(cur)
        va = self.va_dev[i]
        x1 = self.jac(self.N_JACOBI, va, self.ei, self.ea2, self.gw)
        v, e, g = self.cheb(self.rt.pack([va[:, 1:2].contiguous(), x1]), self.ei, self.ev, self.gc)
        self.res_dev[i].copy_(v[:, 1:2])
        tr.mark(cur)
        self.ev_comp[i].record(cur)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_comp[i])
            tr.mark(self.s_out)
            self.out_hosts[i].copy_(self.res_dev[i], non_blocking=True)   # D2H of this step's result
            tr.mark(self.s_out)
            self.ev_out[i].record(self.s_out)
        tr.done()
        return self.out_hosts[i]


class PartitionedSmoother(_Staged):
    """The same smoothing pass on an operator row-block partitioned over `world` GPUs (strong
    scaling): each rank builds only its slab of the stencil and wraps it in a dist.PartitionedGraph,
    the handle the drop-in layers take in place of `edgeij_pair`.  Halo rows travel over NVLink peer
    memory from inside the fused kernels.
      step_kernels  the pass on resident vectors through the partitioned operator, replayed as CUDA graphs
      step_e2e      JacobiGNN.forward + ChebyRelaxGNN.forward on the PartitionedGraph: per step this rank's
                    slab of vertex_attr = [A_ii, b, x] is uploaded from pinned host memory (the same three
                    vectors as at N = 1) and its slab of the result is downloaded"""

    N_JACOBI, CHEB_DEG, OMEGA, CHEB_C, CHEB_D = 10, 4, 0.7, -3.4, -4.0

    def __init__(self, G, N, dev, rank, world, engine="peer", use_graph=True):
        import torch.distributed as dist
        from glab_b200 import dist as gd
        self.G, self.rt, self.dev, self.rank, self.world, self.N = G, G.runtime, dev, rank, world, N
        n = N * N
        self.n = n
        t0 = time.perf_counter()
        self.part = gd.RowPartition(n, world, align=256)
        r0, r1 = self.part.bounds(rank)
        self.r0, self.r1 = r0, r1
        ei, ev = G.generators.laplacian_2d(N, torch.float64, dev, rows=(r0, r1))
        self.ev = ev.float().contiguous()
        self.pg = gd.PartitionedGraph(ei, n, self.part, rank, world, engine=engine)
        self.halo = self.pg.halo
        del ei
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        # operator-side inputs of the layer API (step-invariant); one edge_attr tensor for both layers so
        # that they share one partitioned operator (its plan, values and peer-mapped vectors)
        self.ea2 = torch.cat([self.ev, torch.zeros_like(self.ev)], 1)
        self.op = self.pg.operator(self.ea2, 1, torch.float32)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        nl = self.halo.n_local
        self.n_local, self.nnz_local = nl, self.op.plan.nnz
        z = torch.tensor([self.nnz_local], dtype=torch.int64, device=dev)
        dist.all_reduce(z)
        self.nnz_global = int(z.item())
        self.setup_info = {"generate_ms": (t1 - t0) * 1e3, "plan_build_ms": (t2 - t1) * 1e3, "engine": engine,
                           "rows_local": nl, "halo_rows": self.halo.n_halo, "interior": [self.op.lo, self.op.hi],
                           "cuda_graph": bool(use_graph),
                           "index16_tiles": [self.op.plan.index16_tiles, self.op.plan.tiles]}
        # effective bytes of column index streamed per nonzero: 2 in the tiles flagged for 16-bit indices
        halo16 = os.environ.get("GLAB_IDX16_HALO", "1") != "0"
        frac16 = self.op.plan.index16_tiles / max(self.op.plan.tiles, 1)
        self.index_bytes = 4 - 2.0 * frac16 if (halo16 or engine != "peer") else 4.0
        self.setup_info["halo_kernels_use_index16"] = bool(halo16 and engine == "peer")
        # glab_jacobi_sweeps_halo_* runs all sweeps in one launch only for row blocks of <= 4096 tiles (kMsAutoTiles)
        ms = getattr(self.op, "multi_sweep", False) and getattr(self.op, "fused", False) and \
            (self.op.plan.tiles <= 4096 or os.environ.get("GLAB_MS", "1") == "2") and os.environ.get("GLAB_MS", "1") != "0"
        self.setup_info["jacobi_sweeps_per_launch"] = "multi-sweep kernel" if ms else "one launch per sweep"
        from bench_extra import hashed_uniform
        self.b_host = hashed_uniform(r0, r1, 1, "cpu").pin_memory()
        self.x_host = hashed_uniform(r0, r1, 2, "cpu").pin_memory()
        self.va_host = torch.cat([-4 * torch.ones(nl, 1), self.b_host, self.x_host], 1).pin_memory()
        self.diag = torch.full((nl,), -4.0, device=dev)
        self.b = self.b_host.to(dev).contiguous()
        self.w = torch.tensor([self.OMEGA], device=dev)
        rows, _ = G.ChebyGNN._recurrence(self.CHEB_DEG, torch.tensor([self.CHEB_C, self.CHEB_D]))
        self.table = torch.stack([torch.stack(r) for r in rows]).to(dev).contiguous()
        self.x = torch.empty(nl, 1, device=dev)
        self.r = torch.empty(nl, 1, device=dev)
        self.gw = torch.tensor(self.OMEGA).reshape(-1)
        self.gc = torch.tensor([self.CHEB_C, self.CHEB_D])
        self.jac = G.JacobiGNN.JacobiGNN()
        self.cheb = G.ChebyGNN.ChebyRelaxGNN(self.CHEB_DEG)
        self.op.load("v0", self.x_host.to(dev))
        torch.cuda.synchronize()
        dist.barrier()
        self.h2d_bytes = self.va_host.numel() * 4
        self.d2h_bytes = nl * 4
        self.graph = None
        self.jac_graph = None
        self.use_graph = use_graph and engine in ("peer", "peer-split")
        self.setup_info["cuda_graph"] = self.use_graph
        if self.use_graph:
            self._capture()

    def _sweeps(self):
        op = self.op
        op.publish("v0")
        return op.jacobi(self.N_JACOBI, self.diag, self.b, self.w, "v0")

    def _pass(self):
        cur = self._sweeps()
        self.op.chebyshev(self.CHEB_DEG, self.b, self.table, cur, self.x, self.r)

    def _capture(self):
        import torch.distributed as dist
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._pass()          # warm every kernel (attribute setup, lazy init) before capture
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        dist.barrier()
        c0 = self.rt.launch_count
        self.jac_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.jac_graph):
            self._jac_result = self._sweeps()
        c1 = self.rt.launch_count
        self.cheb_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.cheb_graph):
            self.op.chebyshev(self.CHEB_DEG, self.b, self.table, self._jac_result, self.x, self.r)
        self.graph_launches = (c1 - c0, self.rt.launch_count - c1)   # kernels inside each graph
        torch.cuda.synchronize()
        dist.barrier()

    def step_kernels(self, time_jacobi=False):
        ev = None
        if time_jacobi:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        if self.use_graph:
            self.jac_graph.replay()
            self.rt.launch_count += sum(self.graph_launches)
        else:
            cur = self._sweeps()
        if time_jacobi:
            ev[1].record()
        if self.use_graph:
            self.cheb_graph.replay()
        else:
            self.op.chebyshev(self.CHEB_DEG, self.b, self.table, cur, self.x, self.r)
        return ev

    def layer_pass(self, va):
        """The drop-in layer calls on this rank's slab (va = [A_ii, b, x] on the device)."""
        x1 = self.jac(self.N_JACOBI, va, self.pg, self.ea2, self.gw)
        v, e, g = self.cheb(self.rt.pack([va[:, 1:2].contiguous(), x1]), self.pg, self.ea2, self.gc)
        return v[:, 1:2]

    def _e2e_setup(self):
        dev, nl = self.dev, self.n_local
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.va_dev = [torch.empty(nl, 3, device=dev) for _ in range(self.NBUF)]
        self.res_dev = [torch.empty(nl, 1, device=dev) for _ in range(self.NBUF)]
        self.out_hosts = [torch.empty(nl, 1).pin_memory() for _ in range(self.NBUF)]
        self.ev_in = [torch.cuda.Event() for _ in range(self.NBUF)]
        self.ev_comp = [torch.cuda.Event() for _ in range(self.NBUF)]
        self.ev_out = [torch.cuda.Event() for _ in range(self.NBUF)]
        self.e2e_i = 0

    def step_e2e(self):
        """Per step: this rank's slab of vertex_attr is uploaded from pinned host memory, the layers
        run on the PartitionedGraph, the slab of the result is downloaded; triple-buffered staging
        lets the copies of neighbouring steps overlap the compute."""
        if not hasattr(self, "s_in"):
            self._e2e_setup()
        i = self.e2e_i % self.NBUF
        self.e2e_i += 1
        tr = _Trace(getattr(self, "trace", None))
        cur = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(self.ev_comp[i])
            tr.mark(self.s_in)
            self.va_dev[i].copy_(self.va_host, non_blocking=True)
            tr.mark(self.s_in)
            self.ev_in[i].record(self.s_in)
        cur.wait_event(self.ev_in[i])
        cur.wait_event(self.ev_out[i])
        tr.mark(cur)
        self.res_dev[i].copy_(self.layer_pass(self.va_dev[i]))
        tr.mark(cur)
        self.ev_comp[i].record(cur)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_comp[i])
            tr.mark(self.s_out)
            self.out_hosts[i].copy_(self.res_dev[i], non_blocking=True)
            tr.mark(self.s_out)
            self.ev_out[i].record(self.s_out)
        tr.done()
        return self.out_hosts[i]

    def parity(self):
        """Collective.  The partitioned layer pass on the deterministic global inputs, gathered on rank 0
        and compared BIT FOR BIT with the same layer calls on the unpartitioned operator on one GPU."""
        import torch.distributed as dist
        G, dev = self.G, self.dev
        va = self.va_host.to(dev)
        mine = self.layer_pass(va).contiguous()
        self.op.check()
        full = [torch.empty(self.part.bounds(q)[1] - self.part.bounds(q)[0], 1, device=dev)
                for q in range(self.world)] if self.rank == 0 else None
        dist.gather(mine, full, dst=0)
        if self.rank != 0:
            return None
        from bench_extra import hashed_uniform
        n, N = self.n, self.N
        ei, ev = G.UtilsGNN.laplacianfun_torch(N, device=dev)
        ev = ev.float().contiguous()
        b, x = hashed_uniform(0, n, 1, dev), hashed_uniform(0, n, 2, dev)
        vg = torch.cat([-4 * torch.ones(n, 1, device=dev), b, x], 1)
        x1 = self.jac(self.N_JACOBI, vg, ei, torch.cat([ev, torch.zeros_like(ev)], 1), self.gw)
        ref = self.cheb(torch.cat([b, x1], 1), ei, ev, self.gc)[0][:, 1:2]
        got = torch.cat(full)
        ok = bool(torch.equal(got, ref))
        return {"ok": ok, "bit_exact": True, "max_abs": float((got - ref).abs().max().item()), "rows_checked": n,
                "against": "the same JacobiGNN.forward(10) + ChebyRelaxGNN(4).forward calls on the UNPARTITIONED "
                           "operator on one GPU (rank 0); every row of the gathered result, bit for bit"}

    def close(self):
        self.op.close()
