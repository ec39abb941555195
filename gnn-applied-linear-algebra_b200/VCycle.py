"""Drop-in for pytorch/VCycle.py: the two-grid V-cycle glue around the GNN layers
(Jacobi pre/post-smoothing, classical SOC, direct interpolation, Galerkin coarse operator,
residual, Chebyshev coarse solve, correction).

Same function names and call order as the reference's free functions (VCycle.py:58-237).
Differences, all on the glue side (the layer calls are the reference's):
  * nothing runs at import time (the reference executes its N=5 demo on import, :239-277);
  * the prolongator P = [I + W](:, C) is assembled SPARSE on the device -- the reference builds a
    dense n x n matrix (:129-134), which caps it at n ~ 1e4;
  * the AMG hierarchy (SOC, P, P^T, A_c and their CSR plans) is cached per operator instead of
    being rebuilt inside every runVCycle call;
  * the diagonal A_ii is read from A (the reference hard-codes -4, :117,165; identical for
    laplacianfun_torch matrices);
  * pyamg's CLJP (:114) is not available: pass `splitting=` (any 0/1 vector) or get the
    reference's own deterministic alternative C(1:2:end)=1 (matlab/test_vcycle.m:66-67).
    The splitting is an input of the hot path, not part of it.
"""
import weakref

import torch

from . import _runtime as rt
from . import generators
from .ChebyGNN import ChebyRelaxGNN
from .DirectInterpGNN import DirectInterpGNN
from .GNNResidual import GNNResidual
from .JacobiGNN import JacobiGNN
from .SOCClassicGNN import SOCClassicGNN
from .UtilsGNN import coo_to_gnn_input, remove_diag_entries

theta = 0.25     # SOC threshold           (VCycle.py:248)
cheb_deg = 4     # coarse Chebyshev degree  (VCycle.py:249)

SOCGNN = SOCClassicGNN(theta)
DIGNN = DirectInterpGNN()
ResidualGNN = GNNResidual()
ChebyGNN = ChebyRelaxGNN(cheb_deg)
JacGNN = JacobiGNN()


def default_splitting(n, device="cpu"):
    s = torch.zeros(n, dtype=torch.float32, device=device)
    s[0::2] = 1
    return s


class _Operator:
    """A's COO on the compute device, kept alive so the plan/value caches stay valid."""

    def __init__(self, A):
        device = rt.compute_device(A)
        ei, ea = coo_to_gnn_input(A)
        self.n = A.shape[0]
        self.edge_index = rt.to_device(ei, device).contiguous()
        self.edge_attr = rt.to_device(ea, device).contiguous()
        self.diag = generators.diagonal_of(self.edge_index, self.edge_attr, self.n)
        self.off_index, self.off_attr = remove_diag_entries(self.edge_index, self.edge_attr)
        self.off_index = self.off_index.contiguous()
        self.off_attr = self.off_attr.contiguous()
        self.device = device
        self.hierarchy = {}


_operators = {}


def _operator(A):
    key = id(A)
    hit = _operators.get(key)
    if hit is not None and hit[0]() is A and hit[2] == A._version:
        return hit[1]
    op = _Operator(A)
    if len(_operators) > 8:
        _operators.clear()
    _operators[key] = (weakref.ref(A), op, A._version)
    return op


def _place(op, t):
    return rt.to_device(t, op.device)


def _back(like, t):
    return t if like.is_cuda else t.cpu()


def runResidual(A, b, x):
    """r = b - A x  (VCycle.py:58-70)."""
    op = _operator(A)
    r = ResidualGNN(torch.cat([_place(op, b), _place(op, x)], 1), op.edge_index, op.edge_attr)
    return _back(b, r)


def runSOC(A):
    """Boolean strong-connection flags of the off-diagonal edges, [z_off, 1]  (VCycle.py:72-92)."""
    op = _operator(A)
    placeholder = torch.zeros((op.n, 1), dtype=op.off_attr.dtype, device=op.device)
    S = SOCGNN(placeholder, op.off_index, op.off_attr).reshape(-1, 1) > 0
    return S if A.is_cuda else S.cpu()


def runDirectInterp(A, S, N=None, splitting=None):
    """Prolongator P (sparse [n, n_coarse]) from direct interpolation  (VCycle.py:94-137)."""
    op = _operator(A)
    n = op.n
    split = default_splitting(n, op.device) if splitting is None else _place(op, splitting).reshape(-1)
    split = split.to(op.off_attr.dtype)
    e_di = torch.hstack([op.off_attr, _place(op, S).reshape(-1, 1).to(op.off_attr.dtype)])
    v_di = torch.hstack([op.diag.to(op.off_attr.dtype), split.view(-1, 1)])
    w = DIGNN(v_di, op.off_index, e_di, None)
    coarse = split > 0
    new_id = torch.cumsum(coarse.to(torch.int64), 0) - 1
    rows, cols = op.off_index[0], op.off_index[1]
    keep = coarse[cols] & ((w != 0) | torch.isnan(w))   # .to_sparse() of the reference drops exact zeros
    cidx = torch.nonzero(coarse).reshape(-1)
    p_rows = torch.cat([rows[keep], cidx])
    p_cols = torch.cat([new_id[cols[keep]], new_id[cidx]])
    p_vals = torch.cat([w[keep], torch.ones(cidx.numel(), dtype=w.dtype, device=op.device)])
    P = torch.sparse_coo_tensor(torch.stack([p_rows, p_cols]), p_vals, (n, int(cidx.numel()))).coalesce()
    return P if A.is_cuda else P.cpu()


def runCheby(A, b, x, c, d):
    """Chebyshev relaxation of degree cheb_deg  (VCycle.py:139-154)."""
    op = _operator(A)
    g = torch.tensor([c, d])
    v, _, _ = ChebyGNN(torch.cat([_place(op, b), _place(op, x)], 1), op.edge_index, op.edge_attr, g)
    k = b.shape[1]
    return _back(b, v[:, k:2 * k].reshape(b.shape[0], -1))


def runJacobi(n_iters, w, A, b, x):
    """n_iters weighted-Jacobi sweeps  (VCycle.py:156-173)."""
    op = _operator(A)
    dt = op.edge_attr.dtype
    e2 = op.hierarchy.get("jacobi_edge_attr")
    if e2 is None:
        e2 = torch.cat([op.edge_attr, torch.zeros_like(op.edge_attr)], 1)
        op.hierarchy["jacobi_edge_attr"] = e2
    v = torch.cat([op.diag.to(dt), _place(op, b).to(dt), _place(op, x).to(dt)], 1)
    g = torch.tensor(w).reshape(-1)
    return _back(b, JacGNN(n_iters, v, op.edge_index, e2, g))


class _TwoGrid:
    def __init__(self, A, splitting):
        op = _operator(A)
        S = runSOC(A)
        P = runDirectInterp(A, S, None, splitting)
        P = rt.to_device(P, op.device).coalesce()
        Ad = rt.to_device(A, op.device).coalesce().to(P.dtype)
        self.Ac = torch.sparse.mm(P.t(), torch.sparse.mm(Ad, P)).coalesce()      # VCycle.py:209
        self.P = P
        pi, pv = P.indices().contiguous(), P.values().contiguous()
        self.plan_P = rt.Plan.from_coo(pi, P.shape[0], P.shape[1])
        self.vals_P = rt.get_vals(self.plan_P, pv.view(-1, 1))
        ti = torch.stack([pi[1], pi[0]]).contiguous()
        self.plan_PT = rt.Plan.from_coo(ti, P.shape[1], P.shape[0])
        self.vals_PT = rt.get_vals(self.plan_PT, pv.view(-1, 1))
        self._keep = (pi, pv, ti)
        self.fast = {}


def _two_grid(A, splitting):
    op = _operator(A)
    key = "two_grid" if splitting is None else ("two_grid", splitting.data_ptr(), splitting._version)
    tg = op.hierarchy.get(key)
    if tg is None:
        tg = _TwoGrid(A, splitting)
        op.hierarchy[key] = tg
    return tg


def _jacobi_inplace(plan, vals, diag, b, x, scratch, w_dev, n_iters):
    """n_iters fused sweeps; returns the buffer that holds the result (x or scratch)."""
    cur, other = x, scratch
    for _ in range(n_iters):
        rt.jacobi(plan, vals, diag, b, cur, other, w_dev)
        cur, other = other, cur
    return cur


def runVCycle(A, b, x, n_presmooth, n_postsmooth, n_coarsesolve, use_jacobi=True, splitting=None):
    """Two-grid V-cycle with a Chebyshev coarse solve  (VCycle.py:175-237); returns the new x.
    `n_coarsesolve` is accepted and unused, as in the reference.

    The cycle issues exactly the fused kernels that runJacobi / runResidual / runCheby issue
    (same arithmetic, bit-identical results), but on buffers and scalar tables cached with the
    hierarchy, so a cycle is 7 + 2 + 4 SpMV-bearing launches and no layout glue."""
    op = _operator(A)
    if not use_jacobi:
        return _runVCycle_layers(A, b, x, n_presmooth, n_postsmooth, n_coarsesolve, use_jacobi, splitting)
    tg = _two_grid(A, splitting)                                      # :203-209 (cached)
    dt = op.edge_attr.dtype
    dev = op.device
    k = b.shape[1]
    fast = tg.fast.get(k)
    if fast is None:
        plan_A = rt.get_plan(op.edge_index, op.n)
        copA = _operator(tg.Ac)
        plan_C = rt.get_plan(copA.edge_index, copA.n)
        from .ChebyGNN import _recurrence
        rows, _ = _recurrence(cheb_deg, torch.tensor([-3.4, -4.0]))            # :221-222
        fast = dict(plan_A=plan_A, vals_A=rt.get_vals(plan_A, op.edge_attr), plan_C=plan_C,
                    vals_C=rt.get_vals(plan_C, copA.edge_attr), diag=op.diag.to(dt).reshape(-1).contiguous(),
                    w=torch.tensor(0.7).reshape(-1).to(device=dev, dtype=dt),   # fp32 0.7 like VCycle.py:195,171
                    table=torch.stack([torch.stack(r_) for r_ in rows]).to(device=dev, dtype=dt).contiguous(),
                    xs=[torch.empty(op.n, k, dtype=dt, device=dev) for _ in range(2)],
                    r=torch.empty(op.n, k, dtype=dt, device=dev),
                    c=[torch.empty(copA.n, k, dtype=dt, device=dev) for _ in range(5)])
        tg.fast[k] = fast
    f = fast
    bd = rt.dense(_place(op, b).to(dt))
    xs = f["xs"]
    xs[0].copy_(_place(op, x).to(dt))
    cur = _jacobi_inplace(f["plan_A"], f["vals_A"], f["diag"], bd, xs[0], xs[1], f["w"], n_presmooth)   # :194-196
    other = xs[1] if cur is xs[0] else xs[0]
    rt.residual(f["plan_A"], f["vals_A"], cur, bd, f["r"])                                               # :212
    rc, xc0, xc, rr, p, p2 = (f["c"][0], None, f["c"][1], f["c"][2], f["c"][3], f["c"][4])
    rt.spmm(tg.plan_PT, tg.vals_PT, f["r"], rc)                                                          # :215  P^T r
    xc0 = torch.zeros_like(rc)                                                                           # :218
    t = f["table"]
    rt.cheby_first(f["plan_C"], f["vals_C"], rc, xc0, xc, rr, p, t[0, 1:2])                              # :221-223
    for it in range(1, cheb_deg):
        rt.cheby_next(f["plan_C"], f["vals_C"], p, p2, rr, xc, t[it, 0:1], t[it, 1:2], t[it, 2:3])
        p, p2 = p2, p
    rt.spmm_add(tg.plan_P, tg.vals_P, xc, cur, other)                                                    # :226  x + P xc
    res = _jacobi_inplace(f["plan_A"], f["vals_A"], f["diag"], bd, other, cur, f["w"], n_postsmooth)     # :229-231
    return _back(b, res.clone())


def _runVCycle_layers(A, b, x, n_presmooth, n_postsmooth, n_coarsesolve, use_jacobi=True, splitting=None):
    """The same cycle written with the public run* functions (kept for the Chebyshev-smoother
    variant and as the cross-check of the cached fast path)."""
    op = _operator(A)
    if use_jacobi:
        x = runJacobi(n_presmooth, 0.7, A, b, x)                      # :194-196
    else:
        x = runCheby(A, b, x, -3.461, -4.0)                          # :198-200
    tg = _two_grid(A, splitting)                                      # :203-209 (cached)
    r = _place(op, runResidual(A, b, x))                              # :212
    rc = rt.spmm(tg.plan_PT, tg.vals_PT, rt.dense(r.to(tg.vals_PT.dtype)))      # :215  P^T r
    xc = torch.zeros_like(rc)                                         # :218
    xc = runCheby(tg.Ac, rc, xc, -3.4, -4.0)                          # :221-223
    xd = rt.dense(_place(op, x).to(tg.vals_P.dtype))
    xd = rt.spmm_add(tg.plan_P, tg.vals_P, rt.dense(xc), xd)          # :226  x + P xc
    x = _back(b, xd)
    if use_jacobi:
        x = runJacobi(n_postsmooth, 0.7, A, b, x)                     # :229-231
    else:
        x = runCheby(A, b, x, -3.4, -4.0)                            # :233-235
    return x
