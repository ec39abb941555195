"""Workload drivers used by bench.py (kept apart so bench.py stays a readable contract)."""
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
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        self.n_local = self.n = n
        self.nnz_local = self.nnz_global = self.plan.nnz
        self.setup_info = {"generate_ms": (t1 - t0) * 1e3, "plan_build_ms": (t2 - t1) * 1e3,
                           "plan_identity_perm": self.plan.identity, "max_row_nnz": self.plan.max_row_nnz}
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

    def step_e2e(self):
        dev = self.dev
        va = self.va_host.to(dev, non_blocking=True)
        x1 = self.jac(self.N_JACOBI, va, self.ei, self.ea2, self.gw)
        v, e, g = self.cheb(torch.cat([va[:, 1:2], x1], 1), self.ei, self.ev, self.gc)
        self.out_host.copy_(v[:, 1:2])          # D2H of the step's result (synchronous)
        return self.out_host
