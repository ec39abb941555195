"""Recursive multilevel cycle (BASELINE.json configs[4]: "VCycle multilevel (Jacobi smoother +
interpolation)") -- an EXTENSION of the reference's two-grid VCycle.py:175-237, whose Chebyshev
coarse "solve" stops contracting on large grids (residual reduction 0.999 per cycle on the 67 M-row
Laplacian).  Every stage is one of the reference's layer formulas, run by the same fused kernels:

  strength   S_ij = (A_ij*A_ij)/(A_ii*A_jj)            SOCSAGNN.py:67        glab_soc_sa_*
  rho        power iteration + Rayleigh quotient on D^-1 A   PowerMethodGNN.py:296-334  glab_power_step_* / glab_rayleigh_*
  smoother   x + (w*(b - A x))/A_ii, w = 1.4 / rho      JacobiGNN.py:119      glab_jacobi_sweeps_*
  residual   b - A x                                    GNNResidual.py:115    glab_residual_*
  transfer   P^T r,  x + P x_c                          VCycle.py:215,226     glab_spmm_* / glab_spmm_add_*
  coarse op  A_c = P^T (A P)                            VCycle.py:209         glab_spgemm_*

The prolongator is the smoothed-aggregation one (Vanek / Mandel / Brezina): aggregates are rooted at a
distance-2 maximal independent set of the strength graph (the device PMIS kernel on H*H, H = strong
edges + diagonal, formed with the device SpGEMM), every other vertex joins the root with the largest
index among its strong neighbours, else among its distance-2 neighbours;
P = (I - (omega_p / rho) D^-1 A) P_tentative as one more SpGEMM.  The coarsest operator (<= coarsest_n
rows) is inverted densely once at setup and applied with glab_spmm_* as a dense-row operator.  With
W-cycling on the coarse levels the cycle contracts the residual of the 5-point Laplacian by ~0.33 per
cycle independently of the grid size (oracle/ml_sa.py is the CPU restatement; parity unpinned against
the reference, which has no multilevel cycle).
"""
import torch

from . import _runtime as rt
from . import generators
from .UtilsGNN import coo_to_gnn_input

DEFAULTS = dict(theta=0.08, omega_p=4.0 / 3.0, jacobi_weight=1.4, power_iters=15, coarsest_n=400, max_levels=25,
                seed=0)


def start_vector(n, dtype, device):
    i = torch.arange(n, dtype=torch.int64, device=device)
    x = (i * 2654435761 + 12345) & 0xFFFFFFFF
    return ((x.to(torch.float64) + 1.0) / 4294967297.0).to(dtype)


class Level:
    """One operator of the hierarchy with everything its part of the cycle needs, resident on the device."""

    def __init__(self, edge_index, edge_val, n, opts):
        dev, dt = edge_val.device, edge_val.dtype
        self.n, self.dtype, self.device = n, dt, dev
        self.edge_index = edge_index.contiguous()
        self.edge_val = edge_val.reshape(-1, 1).contiguous()
        self.plan = rt.Plan.from_coo(self.edge_index, n)
        self.vals = rt.get_vals(self.plan, self.edge_val)
        self.nnz = self.plan.nnz
        self.diag = generators.diagonal_of(self.edge_index, self.edge_val, n).reshape(-1).contiguous()
        self.rho = self._rho(opts["power_iters"])
        self.w = torch.tensor([opts["jacobi_weight"] / self.rho], dtype=dt, device=dev)
        self.P = self.PT = None
        self.inv = None

    def _rho(self, iters):
        """|Rayleigh quotient| of D^-1 A after `iters` power iterations (deferred normalisation)."""
        scaled = (self.edge_val.reshape(-1) / self.diag[self.edge_index[0]]).contiguous()
        sv = rt.get_vals(self.plan, scaled.view(-1, 1))
        cur = start_vector(self.n, self.dtype, self.device)
        nxt = torch.empty_like(cur)
        sums = torch.zeros(2 * (iters + 2), dtype=torch.float64, device=self.device)
        prev = None
        for it in range(iters):
            ss = sums[2 * it:2 * it + 2]
            rt.power_step(self.plan, sv, cur, nxt, prev, ss)
            cur, nxt = nxt, cur
            prev = ss
        ray = sums[2 * iters:2 * iters + 2]
        rt.rayleigh(self.plan, sv, cur, nxt, torch.empty_like(cur), prev, ray)
        return abs(float((ray[0] / ray[1]).item()))


def _aggregate(lev, opts):
    """(agg int64 [n], n_agg, root flags) on the device; integer logic identical to oracle/ml_sa.py."""
    dev, dt, n = lev.device, lev.dtype, lev.n
    ei = lev.edge_index
    S = rt.soc_sa(lev.plan, lev.vals, lev.diag)         # every edge incl. the diagonal (S_ii = 1), caller's edge order
    th = torch.tensor(opts["theta"], dtype=dt, device=dev)
    keep = (S >= th * th) | (ei[0] == ei[1])
    h_idx = ei[:, keep].contiguous()
    ones = torch.ones(h_idx.shape[1], 1, dtype=dt, device=dev)
    plan_h = rt.Plan.from_coo(h_idx, n)
    hv = rt.get_vals(plan_h, ones)
    h2_idx, _ = rt.spgemm(plan_h, hv, plan_h, hv)                         # distance <= 2 pattern
    off2 = h2_idx[0] != h2_idx[1]
    h2_off = h2_idx[:, off2].contiguous()
    plan_h2 = rt.Plan.from_coo(h2_off, n)
    root, _ = rt.cf_split_pmis(plan_h2, torch.ones(h2_off.shape[1], dtype=dt, device=dev), opts["seed"])
    root = root > 0
    rid = torch.cumsum(root.to(torch.int64), 0) - 1
    neg = torch.full((n,), -1, dtype=torch.int64, device=dev)
    m1 = (h_idx[0] != h_idx[1]) & root[h_idx[1]]
    cand = neg.clone().scatter_reduce_(0, h_idx[0][m1], h_idx[1][m1], "amax", include_self=True)
    m2 = root[h2_off[1]]
    cand2 = neg.clone().scatter_reduce_(0, h2_off[0][m2], h2_off[1][m2], "amax", include_self=True)
    me = torch.arange(n, dtype=torch.int64, device=dev)
    pick = torch.where(root, me, torch.where(cand >= 0, cand, cand2))
    agg = torch.where(pick >= 0, rid[pick.clamp_min(0)], neg)
    left = torch.nonzero(agg < 0).reshape(-1)
    na = int(root.sum().item())
    agg[left] = na + torch.arange(left.numel(), dtype=torch.int64, device=dev)
    return agg, na + int(left.numel()), root


def _prolongator(lev, agg, na, opts):
    """P = (I - (omega_p / rho) D^-1 A) P_tentative as one SpGEMM; returns (edge_index, values) of P."""
    dev, dt, n = lev.device, lev.dtype, lev.n
    ei = lev.edge_index
    scale = torch.tensor(opts["omega_p"] / lev.rho, dtype=dt, device=dev)
    a = lev.edge_val.reshape(-1)
    m = -(scale / lev.diag[ei[0]]) * a
    m = torch.where(ei[0] == ei[1], m + 1, m)
    mv = rt.get_vals(lev.plan, m.view(-1, 1).contiguous())
    pt_idx = torch.stack([torch.arange(n, dtype=torch.int64, device=dev), agg]).contiguous()
    plan_pt = rt.Plan.from_coo(pt_idx, n, na)
    ptv = rt.get_vals(plan_pt, torch.ones(n, 1, dtype=dt, device=dev))
    return rt.spgemm(lev.plan, mv, plan_pt, ptv)


class Hierarchy:
    def __init__(self, A, **kw):
        self.opts = dict(DEFAULTS, **kw)
        dev = rt.compute_device(A)
        ei, ea = coo_to_gnn_input(A)
        ei, ea = rt.to_device(ei, dev).contiguous(), rt.to_device(ea, dev).contiguous()
        self.levels = []
        n = A.shape[0]
        while True:
            lev = Level(ei, ea, n, self.opts)
            self.levels.append(lev)
            if n <= self.opts["coarsest_n"] or len(self.levels) >= self.opts["max_levels"]:
                break
            agg, na, root = _aggregate(lev, self.opts)
            if na >= 0.9 * n:
                break
            pi, pv = _prolongator(lev, agg, na, self.opts)
            lev.agg, lev.n_agg, lev.root = agg, na, root
            lev.P_index, lev.P_vals = pi, pv
            lev.plan_P = rt.Plan.from_coo(pi, n, na)
            lev.vals_P = rt.get_vals(lev.plan_P, pv.view(-1, 1))
            ti = torch.stack([pi[1], pi[0]]).contiguous()
            lev.plan_PT = rt.Plan.from_coo(ti, na, n)
            lev.vals_PT = rt.get_vals(lev.plan_PT, pv.view(-1, 1))
            lev._keep = (ti,)
            ap_i, ap_v = rt.spgemm(lev.plan, lev.vals, lev.plan_P, lev.vals_P)
            plan_ap = rt.Plan.from_coo(ap_i, n, na)
            ei, ea = rt.spgemm(lev.plan_PT, lev.vals_PT, plan_ap, ap_v)
            ea = ea.view(-1, 1)
            n = na
        last = self.levels[-1]
        dense = torch.sparse_coo_tensor(last.edge_index, last.edge_val.reshape(-1).double(), (last.n, last.n)).to_dense()
        inv = torch.linalg.inv(dense).to(last.dtype).contiguous()          # setup only; applied by glab_spmm_*
        nn_ = last.n
        rowptr = torch.arange(0, nn_ * nn_ + 1, nn_, dtype=torch.int32, device=last.device)
        colidx = torch.arange(nn_, dtype=torch.int32, device=last.device).repeat(nn_)
        last.inv_plan = rt.Plan.from_csr(rowptr, colidx, nn_, nn_)
        last.inv_vals = inv.reshape(-1)
        self.buffers = {}
        self.graphs = {}

    # ------------------------------------------------------------------ the cycle
    def _bufs(self, k):
        b = self.buffers.get(k)
        if b is None:
            b = []
            for lev in self.levels:
                mk = lambda: torch.empty(lev.n, k, dtype=lev.dtype, device=lev.device)   # noqa: E731
                b.append({"x": [mk(), mk()], "b": mk(), "r": mk()})
            self.buffers[k] = b
        return b

    def _visit(self, l, bufs, n_pre, n_post, gamma, zero_guess):
        """One visit of level l: right-hand side in bufs[l]["b"], iterate in bufs[l]["x"][0] (result there too)."""
        lev, B = self.levels[l], bufs[l]
        if l == len(self.levels) - 1:
            rt.spmm(lev.inv_plan, lev.inv_vals, B["b"], B["x"][0])
            return
        xs = B["x"]
        if zero_guess:
            xs[0].zero_()
        cur = rt.jacobi_sweeps(lev.plan, lev.vals, lev.diag, B["b"], xs[0], xs[1], lev.w, n_pre)
        other = xs[1] if cur is xs[0] else xs[0]
        rt.residual(lev.plan, lev.vals, cur, B["b"], B["r"])
        rt.spmm(lev.plan_PT, lev.vals_PT, B["r"], bufs[l + 1]["b"])
        visits = gamma if l + 1 < len(self.levels) - 1 else 1
        for g in range(visits):
            self._visit(l + 1, bufs, n_pre, n_post, gamma, zero_guess=(g == 0))
        rt.spmm_add(lev.plan_P, lev.vals_P, bufs[l + 1]["x"][0], cur, other)
        res = rt.jacobi_sweeps(lev.plan, lev.vals, lev.diag, B["b"], other, cur, lev.w, n_post)
        if res is not xs[0]:
            xs[0].copy_(res)

    def run_from(self, l0, k, n_pre, n_post, gamma, first_zero):
        """The part of a cycle rooted at level l0 (right-hand side in the level's "b" buffer, result in
        its "x"[0] buffer), as many visits as the W-cycle makes there.  The coarse levels are launch-bound
        -- a W-cycle on 7 levels is ~50 level visits of ~8 tiny kernels -- so the whole sequence is captured
        once per (l0, k, sweeps, gamma) in a CUDA graph and replayed (GLAB_ML_GRAPH=0: launch one by one)."""
        import os
        bufs = self._bufs(k)
        visits = (gamma if l0 < len(self.levels) - 1 else 1) if l0 > 0 else 1

        def body():
            for g in range(visits):
                self._visit(l0, bufs, n_pre, n_post, gamma, zero_guess=(first_zero and g == 0))

        if os.environ.get("GLAB_ML_GRAPH", "1") == "0":
            body()
            return bufs[l0]["x"][0]
        key = (l0, k, n_pre, n_post, gamma, first_zero)
        gr = self.graphs.get(key)
        if gr is None:
            save_b = bufs[l0]["b"].clone()
            save_x = bufs[l0]["x"][0].clone()
            side = torch.cuda.Stream(self.levels[0].device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                body()                      # warm-up: lazy attribute setup of every kernel variant involved
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            bufs[l0]["b"].copy_(save_b)
            bufs[l0]["x"][0].copy_(save_x)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                body()
            self.graphs[key] = gr
            bufs[l0]["b"].copy_(save_b)
            bufs[l0]["x"][0].copy_(save_x)
        gr.replay()
        return bufs[l0]["x"][0]

    def cycle(self, b, x, n_pre=3, n_post=3, gamma=2):
        """One cycle on k = b.shape[1] right-hand sides; returns the new iterate (a fresh tensor)."""
        k = b.shape[1]
        bufs = self._bufs(k)
        bufs[0]["b"].copy_(b)
        bufs[0]["x"][0].copy_(x)
        return self.run_from(0, k, n_pre, n_post, gamma, first_zero=False).clone()

    # ------------------------------------------------------------------ bookkeeping
    def info(self, n_pre=3, n_post=3, gamma=2):
        lv = self.levels
        vis = self._visit_counts(gamma)
        work = 0
        for l, lev in enumerate(lv[:-1]):
            work += vis[l] * ((n_pre + n_post + 1) * lev.nnz + 2 * lev.plan_P.nnz)
        work += vis[-1] * lv[-1].n * lv[-1].n
        return {"levels": len(lv), "rows_per_level": [l.n for l in lv], "nnz_per_level": [l.nnz for l in lv],
                "rho_per_level": [round(l.rho, 4) for l in lv], "visits_per_level": vis,
                "operator_complexity": sum(l.nnz for l in lv) / lv[0].nnz, "spmv_nnz_per_cycle": int(work),
                "cycle": "W on the coarse levels" if gamma == 2 else "V", "options": self.opts}

    def _visit_counts(self, gamma):
        """How often one cycle visits each level (mirrors _visit)."""
        L = len(self.levels)
        counts = [0] * L

        def go(l):
            counts[l] += 1
            if l == L - 1:
                return
            for _ in range(gamma if l + 1 < L - 1 else 1):
                go(l + 1)

        go(0)
        return counts
