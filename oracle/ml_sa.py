"""ORACLE (test infrastructure, never on the product path): CPU restatement (numpy / scipy) of the
recursive smoothed-aggregation hierarchy and W-cycle that glab_b200.multilevel builds on the device.

PARITY UNPINNED against the reference: the reference's cycle is two-grid only (VCycle.py:175-237)
and its coarse "solve" does not converge on large grids; BASELINE.json's configs[4] asks for a
multilevel cycle, which is therefore an EXTENSION assembled from the reference's own layer formulas:

  strength     S_ij = (A_ij*A_ij)/(A_ii*A_jj) on every edge             (SOCSAGNN.py:67), strong <=> S >= theta^2
  smoother     x <- x + (w*(b - A x))/A_ii, w = jacobi_weight / rho     (JacobiGNN.py:119)
  rho          Rayleigh quotient after `power_iters` power iterations on D^-1 A  (PowerMethodGNN.py:296-334)
  residual     r = b - A x                                              (GNNResidual.py:115)
  transfer     r_c = P^T r, x += P x_c (matvec blocks)                  (MatVecGNN.py:109-114, VCycle.py:215,226)
  coarse op    A_c = P^T (A P)                                          (VCycle.py:209)

and the standard smoothed-aggregation prolongator (Vanek, Mandel, Brezina 1996) with MIS(2) roots:
aggregates = PMIS (oracle/cf_split.py) on the distance-2 strength graph, every other vertex joins the
root with the largest index among its strong neighbours (else among its distance-2 neighbours),
P = (I - (omega_p/rho) D^-1 A) P_tentative.  All index arithmetic is integer and deterministic, so the
device must reproduce aggregates bit for bit; floating-point stages are compared to tolerance.

What IS pinned: every building block above against oracle/port.py (itself pinned bit for bit to the unmodified
reference layers): strength values and mask, Jacobi sweeps, residual, restriction / prolongation through the
scatter-sum seam, the Galerkin product and the power-method estimate of rho
(tests/test_oracle_multilevel.py::test_*_reference_*).  What stays unpinned is the composition: the aggregation
rule, the prolongator smoothing and the recursion.
"""
import numpy as np
import scipy.sparse as sp

from .cf_split import pmis

DEFAULTS = dict(theta=0.08, omega_p=4.0 / 3.0, jacobi_weight=1.4, power_iters=15, coarsest_n=400, max_levels=25,
                seed=0)


def start_vector(n, dtype):
    """Deterministic start vector of the power iteration (same integers on the device)."""
    i = np.arange(n, dtype=np.uint64)
    x = (i * np.uint64(2654435761) + np.uint64(12345)) & np.uint64(0xFFFFFFFF)
    return ((x.astype(np.float64) + 1.0) / 4294967297.0).astype(dtype)


def rho_dinv_a(A, iters, dtype):
    """|Rayleigh quotient| of D^-1 A after `iters` normalised power iterations."""
    d = A.diagonal()
    M = sp.diags(1.0 / d) @ A
    x = start_vector(A.shape[0], np.float64)
    for _ in range(iters):
        y = M @ x
        x = y / np.linalg.norm(y)
    y = M @ x
    return float(abs((x @ y) / (x @ x)))


def strength_mask(A, theta, dtype):
    """Per edge of A (row-major sorted COO incl. the diagonal): S >= theta^2, diagonal always kept."""
    C = A.tocoo()
    r, c = C.row, C.col
    v = C.data.astype(dtype)
    d = A.diagonal().astype(dtype)
    with np.errstate(divide="ignore", invalid="ignore"):
        S = (v * v) / (d[r] * d[c])
    th = dtype(theta) * dtype(theta)      # dtype: numpy scalar type (np.float32 / np.float64)
    return r, c, (S >= th) | (r == c)


def aggregates(n, r, c, keep, seed=0):
    """(agg [n] int64, n_agg, root flags): MIS(2) roots + nearest-root assignment (see module doc)."""
    rs, cs = r[keep], c[keep]
    H = sp.csr_matrix((np.ones(rs.size, dtype=np.float64), (rs, cs)), shape=(n, n))
    H2 = (H @ H).tocoo()
    off2 = H2.row != H2.col
    r2, c2 = H2.row[off2], H2.col[off2]
    root, _ = pmis(n, r2, c2, np.ones(r2.size, dtype=bool), seed)
    root = root.astype(bool)
    rid = np.cumsum(root) - 1
    cand = np.full(n, -1, dtype=np.int64)
    off1 = rs != cs
    m = off1 & root[cs]
    np.maximum.at(cand, rs[m], cs[m])
    cand2 = np.full(n, -1, dtype=np.int64)
    m2 = root[c2]
    np.maximum.at(cand2, r2[m2], c2[m2])
    pick = np.where(root, np.arange(n), np.where(cand >= 0, cand, cand2))
    agg = np.where(pick >= 0, rid[np.maximum(pick, 0)], -1)
    left = np.nonzero(agg < 0)[0]
    na = int(root.sum())
    agg[left] = na + np.arange(left.size)
    return agg.astype(np.int64), na + int(left.size), root


def prolongator(A, agg, na, rho, omega_p, dtype):
    n = A.shape[0]
    Pt = sp.csr_matrix((np.ones(n, dtype=dtype), (np.arange(n), agg)), shape=(n, na))
    d = A.diagonal()
    M = sp.eye(n, format="csr") - sp.diags((omega_p / rho) / d) @ A
    return (M @ Pt).tocsr().astype(dtype)


def build(A, dtype=np.float64, **kw):
    """levels: list of dicts A, d, w, rho [, P, agg, n_agg]."""
    o = dict(DEFAULTS, **kw)
    A = sp.csr_matrix(A).astype(dtype)
    levels = []
    while True:
        n = A.shape[0]
        rho = rho_dinv_a(A, o["power_iters"], dtype)
        lev = {"A": A, "d": A.diagonal(), "rho": rho, "w": o["jacobi_weight"] / rho}
        levels.append(lev)
        if n <= o["coarsest_n"] or len(levels) >= o["max_levels"]:
            break
        r, c, keep = strength_mask(A, o["theta"], np.dtype(dtype).type)
        agg, na, root = aggregates(n, r, c, keep, o["seed"])
        if na >= 0.9 * n:
            break
        P = prolongator(A, agg, na, rho, o["omega_p"], dtype)
        lev.update(P=P, agg=agg, n_agg=na, root=root)
        A = (P.T @ (A @ P)).tocsr()
        A.sort_indices()
    return levels


def cycle(levels, b, x, n_pre=3, n_post=3, gamma=2, level=0):
    """One multilevel cycle (gamma = 1: V, 2: W on the coarse levels); b, x are [n, k]."""
    L = levels[level]
    A = L["A"]
    if level == len(levels) - 1:
        return np.linalg.solve(A.toarray().astype(np.float64), b.astype(np.float64)).astype(b.dtype)
    d = L["d"].reshape(-1, 1)
    for _ in range(n_pre):
        x = x + (L["w"] * (b - A @ x)) / d
    rc = L["P"].T @ (b - A @ x)
    xc = np.zeros_like(rc)
    for _ in range(gamma if level + 1 < len(levels) - 1 else 1):
        xc = cycle(levels, rc, xc, n_pre, n_post, gamma, level + 1)
    x = x + L["P"] @ xc
    for _ in range(n_post):
        x = x + (L["w"] * (b - A @ x)) / d
    return x
