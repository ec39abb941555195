"""Workload drivers used by bench.py (kept apart so bench.py stays a readable contract)."""
import os
import time

import torch


class SingleGpuSmoother:
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
        torch.manual_seed(24601)
        self.b_host = torch.rand(n, 1).pin_memory()
        self.x_host = torch.rand(n, 1).pin_memory()
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

    def _e2e_setup(self):
        """Double-buffered staging so that consecutive steps overlap their PCIe copies with compute:
        step i+1's H2D and step i-1's D2H run on copy streams while step i computes.  Every step
        still uploads its own inputs from pinned host memory and downloads its own result."""
        dev, n = self.dev, self.n
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.va_dev = [torch.empty(n, 3, device=dev) for _ in range(2)]
        self.res_dev = [torch.empty(n, 1, device=dev) for _ in range(2)]
        self.out_hosts = [torch.empty(n, 1).pin_memory() for _ in range(2)]
        self.ev_in = [torch.cuda.Event() for _ in range(2)]
        self.ev_comp = [torch.cuda.Event() for _ in range(2)]
        self.ev_out = [torch.cuda.Event() for _ in range(2)]
        self.e2e_i = 0

    def step_e2e(self):
        if not hasattr(self, "s_in"):
            self._e2e_setup()
        i = self.e2e_i % 2
        self.e2e_i += 1
        cur = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(self.ev_comp[i])            # the step that last read this buffer is done
            self.va_dev[i].copy_(self.va_host, non_blocking=True)   # H2D of this step's inputs
            self.ev_in[i].record(self.s_in)
        cur.wait_event(self.ev_in[i])
        cur.wait_event(self.ev_out[i])                       # this result buffer has been downloaded
        va = self.va_dev[i]
        x1 = self.jac(self.N_JACOBI, va, self.ei, self.ea2, self.gw)
        v, e, g = self.cheb(self.rt.pack([va[:, 1:2].contiguous(), x1]), self.ei, self.ev, self.gc)
        self.res_dev[i].copy_(v[:, 1:2])
        self.ev_comp[i].record(cur)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_comp[i])
            self.out_hosts[i].copy_(self.res_dev[i], non_blocking=True)   # D2H of this step's result
            self.ev_out[i].record(self.s_out)
        return self.out_hosts[i]


class PartitionedSmoother:
    """The same smoothing pass on an operator row-block partitioned over `world` GPUs (strong
    scaling): each rank builds only its slab of the stencil, halo rows travel over NVLink peer
    memory (glab_halo_push / glab_halo_wait), and the whole pass is replayed as one CUDA graph."""

    N_JACOBI, CHEB_DEG, OMEGA, CHEB_C, CHEB_D = 10, 4, 0.7, -3.4, -4.0

    def __init__(self, G, N, dev, rank, world, engine="peer", use_graph=True):
        import torch.distributed as dist
        from glab_b200 import dist as gd
        self.G, self.rt, self.dev, self.rank, self.world = G, G.runtime, dev, rank, world
        n = N * N
        t0 = time.perf_counter()
        self.part = gd.RowPartition(n, world, align=256)
        r0, r1 = self.part.bounds(rank)
        ei, ev = G.generators.laplacian_2d(N, torch.float64, dev, rows=(r0, r1))
        ev = ev.float().contiguous()
        self.halo = gd.HaloPlan.build(self.part, rank, ei[1])
        lei = torch.stack([ei[0] - r0, self.halo.local_columns(ei[1])]).contiguous()
        del ei
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        self.op = gd.DistOperator(lei, ev, self.halo, k=1, engine=engine)
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
        # effective bytes of column index streamed per nonzero: the fused halo kernels (engine "peer") stay on
        # int32 unless GLAB_IDX16=3; the other engines launch the plain kernels, which stream 2 bytes in the
        # tiles flagged for 16-bit indices
        halo16 = os.environ.get("GLAB_IDX16", "2") == "3"
        frac16 = self.op.plan.index16_tiles / max(self.op.plan.tiles, 1)
        self.index_bytes = 4 - 2.0 * frac16 if (halo16 or engine != "peer") else 4.0
        self.setup_info["halo_kernels_use_index16"] = bool(halo16 and engine == "peer")
        g = torch.Generator().manual_seed(24601 + rank)
        self.b_host = torch.rand(nl, 1, generator=g).pin_memory()
        self.x_host = torch.rand(nl, 1, generator=g).pin_memory()
        self.out_host = torch.empty(nl, 1).pin_memory()
        self.diag = torch.full((nl,), -4.0, device=dev)
        self.b = self.b_host.to(dev).contiguous()
        self.x_stage = torch.empty(nl, 1, device=dev)
        self.w = torch.tensor([self.OMEGA], device=dev)
        rows, _ = G.ChebyGNN._recurrence(self.CHEB_DEG, torch.tensor([self.CHEB_C, self.CHEB_D]))
        self.table = torch.stack([torch.stack(r) for r in rows]).to(dev).contiguous()
        self.x = torch.empty(nl, 1, device=dev)
        self.r = torch.empty(nl, 1, device=dev)
        self.op.load("v0", self.x_host.to(dev))
        torch.cuda.synchronize()
        dist.barrier()
        self.h2d_bytes = 2 * nl * 4
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

    def _e2e_setup(self):
        dev, nl = self.dev, self.n_local
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.b_stage = [torch.empty(nl, 1, device=dev) for _ in range(2)]
        self.x_stage2 = [torch.empty(nl, 1, device=dev) for _ in range(2)]
        self.res_dev = [torch.empty(nl, 1, device=dev) for _ in range(2)]
        self.out_hosts = [torch.empty(nl, 1).pin_memory() for _ in range(2)]
        self.ev_in = [torch.cuda.Event() for _ in range(2)]
        self.ev_comp = [torch.cuda.Event() for _ in range(2)]
        self.ev_out = [torch.cuda.Event() for _ in range(2)]
        self.e2e_i = 0

    def step_e2e(self):
        """Per step: this rank's slab of b and x0 is uploaded from pinned host memory, the captured
        smoothing pass runs, the slab of the result is downloaded; double-buffered staging lets the
        copies of neighbouring steps overlap the compute."""
        if not hasattr(self, "s_in"):
            self._e2e_setup()
        i = self.e2e_i % 2
        self.e2e_i += 1
        cur = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(self.ev_comp[i])
            self.b_stage[i].copy_(self.b_host, non_blocking=True)
            self.x_stage2[i].copy_(self.x_host, non_blocking=True)
            self.ev_in[i].record(self.s_in)
        cur.wait_event(self.ev_in[i])
        cur.wait_event(self.ev_out[i])
        self.b.copy_(self.b_stage[i])
        self.op.vec["v0"][:self.n_local].copy_(self.x_stage2[i])
        self.step_kernels()
        self.res_dev[i].copy_(self.x)
        self.ev_comp[i].record(cur)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_comp[i])
            self.out_hosts[i].copy_(self.res_dev[i], non_blocking=True)
            self.ev_out[i].record(self.s_out)
        return self.out_hosts[i]

    def close(self):
        self.op.close()
