"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference's hot path.

This file is the checker that travels to the GPU box (``/root/reference`` does not).
It restates, in plain CPU torch and in the reference's own operation order, the
edge-wise message passing behind sandialabs/gnn-applied-linear-algebra's PyTorch layers,
so that on the same inputs it reproduces the reference **bit for bit** on CPU.

PINNING: tests/test_oracle_pinning.py compares every function below against the
unmodified reference files (loaded by oracle/ref_loader.py) when /root/reference is
present, and tests/test_oracle_golden.py compares it against the committed fixtures
under tests/golden/ (generated from the reference by tests/golden/make_golden.py)
everywhere.  The one unpinned piece is the C/F splitting (pyamg CLJP is absent and no
reference test fixes its output): the splitting is an *input* here.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package never does.

All citations are relative to /root/reference/pytorch/.
"""
import torch

# --------------------------------------------------------------------------- primitives


def scatter_sum(src, index, n):
    """torch_scatter.scatter(src, index, dim=0, dim_size=n, reduce="sum")
    == zeros.scatter_add_ (edge-order sequential accumulation).  MatVecGNN.py:60."""
    if src.dim() == 1:
        idx = index
    else:
        idx = index.view(-1, *([1] * (src.dim() - 1))).expand(src.size())
    return torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype).scatter_add_(0, idx, src)


def scatter_max(src, index, n):
    """reduce="max": segment maximum, 0 for rows without edges.  SOCClassicGNN.py:69."""
    out = torch.zeros(n, dtype=src.dtype)
    out.scatter_reduce_(0, index, src, reduce="amax", include_self=False)
    return out


def gn_block(x, edge_index, edge_attr, u, edge_fn=None, vertex_fn=None, global_fn=None):
    """One graph-network block in torch_geometric.nn.MetaLayer order (spec twin:
    matlab/gnn.m:47-54): edge update on (x[row], x[col]) -> vertex update -> global
    update; a stage whose callback is None is skipped."""
    row, col = edge_index[0], edge_index[1]
    if edge_fn is not None:
        edge_attr = edge_fn(x[row], x[col], edge_attr, u)
    if vertex_fn is not None:
        x = vertex_fn(x, edge_index, edge_attr, u)
    if global_fn is not None:
        u = global_fn(x, edge_index, edge_attr, u)
    return x, edge_attr, u


# --------------------------------------------------------------------------- MatVecGNN.py


def matvec(x, edge_index, A_ij):
    """MatVecGNN.py:66-114 -- returns (vertex_attr=[x|y], edge_attr=[A_ij|c_ij])."""
    def edge(vi, vj, e, g):
        return torch.cat([e, e * vj], 1)                                   # :80-84

    def vertex(v, ei, e, g):
        c_ij = e[:, 1:1 + v.shape[1]]                                      # :105
        return torch.cat([v, scatter_sum(c_ij, ei[0], v.shape[0])], 1)     # :109-114

    v, e, _ = gn_block(x, edge_index, A_ij, None, edge, vertex)
    return v, e


# --------------------------------------------------------------------------- GNNResidual.py


def residual(vertex_attr, edge_index, A_ij):
    """GNNResidual.py:64-132 -- vertex_attr=[b,x] -> r = b - A x as [n,1]."""
    def edge(vi, vj, e, g):
        return torch.cat([e, e * vj[:, 1].view(-1, 1)], 1)                 # :77-86

    def vertex(v, ei, e, g):
        b = v[:, 0].view(-1, 1)
        x = v[:, 1].view(-1, 1)
        cbar = scatter_sum(e[:, 1].view(-1, 1), ei[0], v.shape[0])         # :109-114
        return torch.cat([b, x, b - cbar], 1)                              # :115-118

    v, _, _ = gn_block(vertex_attr, edge_index, A_ij, None, edge, vertex)
    return v[:, 2].view(-1, 1)                                             # :131


# --------------------------------------------------------------------------- JacobiGNN.py


def jacobi_iterate(vertex_attr, edge_index, edge_attr, g):
    """JacobiGNN.py:71-135 -- one sweep; vertex_attr=[A_ii,b,x], edge_attr=[A_ij,c_ij], g=[w]."""
    def edge(vi, vj, e, g_):
        A = e[:, 0].view(-1, 1)
        return torch.cat([A, A * vj[:, 2].view(-1, 1)], 1)                 # :81-88

    def vertex(v, ei, e, g_):
        d = v[:, 0].view(-1, 1)
        b = v[:, 1].view(-1, 1)
        x = v[:, 2].view(-1, 1)
        w = g_[0]
        cbar = scatter_sum(e[:, 1], ei[0], v.shape[0]).reshape(-1, 1)      # :63-69
        x = x + w * (b - cbar) / d                                         # :119
        return torch.cat([d, b, x], 1)

    return gn_block(vertex_attr, edge_index, edge_attr, g, edge, vertex)


def jacobi(n_iters, vertex_attr, edge_index, edge_attr, g):
    """JacobiGNN.py:138-148 -- returns x [n,1]."""
    for _ in range(n_iters):
        vertex_attr, edge_attr, g = jacobi_iterate(vertex_attr, edge_index, edge_attr, g)
    return vertex_attr[:, 2].reshape(-1, 1)


# --------------------------------------------------------------------------- ChebyGNN.py


def _cheby_zbar(ei, e, n):
    return scatter_sum(e[:, 1], ei[0], n).view(-1, 1)                      # :73-89


def chebyshev(deg, vertex_attr, edge_index, edge_attr, g):
    """ChebyGNN.py:287-353 -- ChebyRelaxGNN(deg).forward.
    in : vertex_attr=[b,x], edge_attr=[A_ij], g=[c,d]
    out: vertex_attr=[b,x,r,p], edge_attr=[A_ij,z_ij], g=[c,d,alpha,beta]."""
    def l1_edge_first(vi, vj, e, g_):
        return torch.cat([e, e * vj[:, 1].view(-1, 1)], 1)                 # :62-70

    def l1_vertex_first(v, ei, e, g_):
        b = v[:, 0].view(-1, 1)
        x = v[:, 1].view(-1, 1)
        return torch.cat([b, x, b - _cheby_zbar(ei, e, v.shape[0])], 1)    # :108-121

    def l1_global_first(v, ei, e, g_):
        return torch.hstack([g_[0], g_[1], 1 / g_[1]])                     # :133-139

    def l2_vertex_first(v, ei, e, g_):
        b = v[:, 0].view(-1, 1)
        x = v[:, 1].view(-1, 1)
        r = v[:, 2].view(-1, 1)
        p = r
        x = x + g_[2] * p                                                  # :160-161
        return torch.cat([b, x, r, p], 1)

    def l1_edge_next(vi, vj, e, g_):
        A = e[:, 0].view(-1, 1)
        return torch.cat([A, A * vj[:, 3].view(-1, 1)], 1)                 # :177-183

    def l1_vertex_next(v, ei, e, g_):
        b = v[:, 0].view(-1, 1)
        x = v[:, 1].view(-1, 1)
        r = v[:, 2].view(-1, 1)
        p = v[:, 3].view(-1, 1)
        r = r - g_[2] * _cheby_zbar(ei, e, v.shape[0])                     # :208-214 (alpha of prev iter)
        return torch.cat([b, x, r, p], 1)

    def l2_vertex_next(v, ei, e, g_):
        b = v[:, 0].view(-1, 1)
        x = v[:, 1].view(-1, 1)
        r = v[:, 2].view(-1, 1)
        p = v[:, 3].view(-1, 1)
        p = r + g_[3] * p                                                  # :240
        x = x + g_[2] * p                                                  # :241
        return torch.cat([b, x, r, p], 1)

    def l1_global_second(v, ei, e, g_):
        c, d, alpha = g_[0], g_[1], g_[2]
        beta = 0.5 * (c * alpha) ** 2                                      # :262
        alpha = 1 / (d - beta / alpha)                                     # :263
        return torch.hstack([c, d, alpha, beta])

    def l1_global_later(v, ei, e, g_):
        c, d, alpha = g_[0], g_[1], g_[2]
        beta = ((c * alpha) / 2) ** 2                                      # :282
        alpha = 1 / (d - beta / alpha)                                     # :283
        return torch.hstack([c, d, alpha, beta])

    layers = []
    if deg > 0:
        layers += [(l1_edge_first, l1_vertex_first, l1_global_first), (None, l2_vertex_first, None)]
    if deg > 1:
        layers += [(l1_edge_next, l1_vertex_next, l1_global_second), (None, l2_vertex_next, None)]
    for _ in range(deg - 2):
        layers += [(l1_edge_next, l1_vertex_next, l1_global_later), (None, l2_vertex_next, None)]
    for ef, vf, gf in layers:
        vertex_attr, edge_attr, g = gn_block(vertex_attr, edge_index, edge_attr, g, ef, vf, gf)
    return vertex_attr, edge_attr, g


# --------------------------------------------------------------------------- PowerMethodGNN.py


def power_method(num_iter, vertex_attr, edge_index, edge_attr, g):
    """PowerMethodGNN.py:296-334 -- vertex_attr=[b,y], edge_attr=[A_ij,c_ij], g=[n,n_A,lambda]."""
    def edge(vi, vj, e, g_):
        A = e[:, 0].view(-1, 1)
        return torch.cat([A, A * vj[:, 0].view(-1, 1)], 1)                 # :100-106

    def cbar_of(v, ei, e):
        return scatter_sum(e[:, 1], ei[0], v.shape[0]).reshape(-1, 1)      # :64-83

    def it1_vertex(v, ei, e, g_):
        return torch.cat([cbar_of(v, ei, e), v[:, 1].view(-1, 1)], 1)      # :148-158

    def square_vertex(v, ei, e, g_):
        b = v[:, 0].view(-1, 1)
        return torch.cat([b, b * b], 1)                                    # :121-126

    def it2_global(v, ei, e, g_):
        n = torch.sqrt(torch.sum(v[:, 1]))                                 # :180-183
        return torch.tensor([n, g_[1], g_[2]])                             # :185

    def it3_vertex(v, ei, e, g_):
        return torch.cat([v[:, 0].view(-1, 1) / g_[0], v[:, 1].view(-1, 1)], 1)   # :202-207

    def ray1_vertex(v, ei, e, g_):
        b = v[:, 0].view(-1, 1)
        return torch.cat([b, b * cbar_of(v, ei, e)], 1)                    # :229-237

    def ray1_global(v, ei, e, g_):
        return torch.tensor([g_[0], torch.sum(v[:, 1]), g_[2]])            # :259-266

    def ray2_global(v, ei, e, g_):
        return torch.tensor([g_[0], g_[1], g_[1] / torch.sum(v[:, 1])])    # :288-294

    layers = []
    for _ in range(num_iter):
        layers += [(edge, it1_vertex, None), (None, square_vertex, it2_global), (None, it3_vertex, None)]
    layers += [(edge, ray1_vertex, ray1_global), (None, square_vertex, ray2_global)]
    for ef, vf, gf in layers:
        vertex_attr, edge_attr, g = gn_block(vertex_attr, edge_index, edge_attr, g, ef, vf, gf)
    return vertex_attr, edge_attr, g


# --------------------------------------------------------------------------- SOCClassicGNN.py


def soc_classic(theta, vertex_attr, edge_index, edge_attr):
    """SOCClassicGNN.py:131-147 -- off-diagonal edges only; returns S_ij [z] (1-D),
    S_ij = relu(-A_ij / max_{k!=i}(-A_ik) - theta)."""
    def l1_vertex(v, ei, e, g_):
        return scatter_max(-1 * e[:, 0], ei[0], v.shape[0]).reshape(-1, 1)  # :69-72

    def l2_edge(vi, vj, e, g_):
        v_i = vi[:, 0].view(-1, 1)
        A = e[:, 0].view(-1, 1)
        S = torch.relu(-1 * A / v_i - theta)                               # :125
        return torch.cat([A, S], 1)

    v, e, g = gn_block(vertex_attr, edge_index, edge_attr, None, None, l1_vertex)
    v, e, g = gn_block(v, edge_index, e, g, l2_edge)
    return e[:, 1]


# --------------------------------------------------------------------------- SOCSAGNN.py


def soc_sa(diag, edge_index, edge_attr):
    """SOCSAGNN.py:49-71 used as MetaLayer(EdgeUpdate()) (:91) -- vertex_attr=[A_ii];
    returns edge_attr=[A_ij, S_ij], S_ij = (A_ij*A_ij)/(A_ii*A_jj)."""
    def edge(vi, vj, e, g_):
        return torch.cat([e, (e * e) / (vi * vj)], 1)                      # :67-71

    _, e, _ = gn_block(diag, edge_index, edge_attr, None, edge)
    return e


# --------------------------------------------------------------------------- DirectInterpGNN.py


def direct_interp(vertex_attr, edge_index, edge_attr):
    """DirectInterpGNN.py:155-174 -- vertex_attr=[A_ii,C_i], off-diagonal
    edge_attr=[A_ij,S_ij]; returns w_ij [z] (NaN where a C row has no strong C neighbour:
    0*inf, exactly as the reference)."""
    def l1_edge(vi, vj, e, g_):
        return torch.cat([e, vj[:, 1].view(-1, 1)], 1)                     # :61-69

    def l1_vertex(v, ei, e, g_):
        d = v[:, 0].view(-1, 1)
        C = v[:, 1].view(-1, 1)
        n = d.shape[0]
        A, S, w = e[:, 0], e[:, 1], e[:, 2]
        num = scatter_sum(A, ei[0], n)                                     # :89
        den = scatter_sum(A * S * w, ei[0], n)                             # :92
        gamma = (num / den).reshape(-1, 1)                                 # :94-97
        return torch.cat([d, C, (1 / d) * gamma], 1)                       # :127-131

    def l2_edge(vi, vj, e, g_):
        A = e[:, 0].view(-1, 1)
        S = e[:, 1].view(-1, 1)
        C_i = vi[:, 1].view(-1, 1)
        alpha_i = vi[:, 2].view(-1, 1)
        return torch.cat([A, S, (1 - C_i) * (-A * alpha_i)], 1)            # :150-152

    v, e, _ = gn_block(vertex_attr, edge_index, edge_attr, None, l1_edge, l1_vertex)
    v, e, _ = gn_block(v, edge_index, e, None, l2_edge)
    return e[:, 2]


# --------------------------------------------------------------------------- MatrixWeightedNorm.py


def matrix_weighted_norm(x, edge_index, W_ij):
    """MatrixWeightedNorm.py:49-161 -- sqrt(x^T W x) as one block."""
    def edge(vi, vj, e, g_):
        return torch.cat([e, e * vj], 1)                                   # :62-68

    def vertex(v, ei, e, g_):
        cbar = scatter_sum(e[:, 1], ei[0], v.shape[0]).reshape(-1, 1)      # :86
        return torch.cat([v, v * cbar], 1)                                 # :107-109

    def glob(v, ei, e, g_):
        return torch.sqrt(torch.sum(v[:, 1]))                              # :127-129,138-140

    _, _, u = gn_block(x, edge_index, W_ij, None, edge, vertex, glob)
    return u


# --------------------------------------------------------------------------- UtilsGNN.py


def laplacian_2d(N, dtype=torch.float64):
    """UtilsGNN.py:53-67 -- 5-point (negative) Laplacian on an N x N grid as
    (edge_index int64 [2,z] row-major sorted incl. diagonal, values [z,1]):
    diag -4, off-diag +1, Dirichlet truncation.  Restated without scipy."""
    idx = torch.arange(N * N, dtype=torch.int64)
    gy, gx = idx // N, idx % N
    rows, cols, vals = [], [], []
    # scipy COO->CSR-sorted order per row: column ascending = (i-N, i-1, i, i+1, i+N)
    for dy, dx, val in ((-1, 0, 1.0), (0, -1, 1.0), (0, 0, -4.0), (0, 1, 1.0), (1, 0, 1.0)):
        ok = (gy + dy >= 0) & (gy + dy < N) & (gx + dx >= 0) & (gx + dx < N)
        rows.append(torch.where(ok, idx, -1))
        cols.append(idx + dy * N + dx)
        vals.append(torch.full((N * N,), val, dtype=dtype))
    rows = torch.stack(rows, 1).reshape(-1)
    cols = torch.stack(cols, 1).reshape(-1)
    vals = torch.stack(vals, 1).reshape(-1)
    keep = rows >= 0
    return torch.stack([rows[keep], cols[keep]]), vals[keep].reshape(-1, 1)


def remove_diag_entries(edge_index, edge_val):
    """UtilsGNN.py:69-72 (torch_geometric remove_self_loops: mask row != col, order kept)."""
    keep = edge_index[0] != edge_index[1]
    return edge_index[:, keep], edge_val[keep]


def coo_to_gnn_input(A):
    """UtilsGNN.py:74-78."""
    A = A.coalesce()
    return A.indices(), A.values().reshape(-1, 1)


# --------------------------------------------------------------------------- VCycle.py


def two_grid_vcycle(edge_index, A_val, b, x, splitting, n_pre=3, n_post=3, theta=0.25,
                    cheb_deg=4, w=0.7, diag_value=-4.0, coarse_c=-3.4, coarse_d=-4.0):
    """VCycle.py:175-237 runVCycle(use_jacobi=True) restated for any operator given as
    row-major-sorted COO (edge_index, A_val [z,1]) with the C/F ``splitting`` supplied
    (VCycle.py:114 calls pyamg CLJP; unpinned -> input).  fp32 like the reference
    (diag_vals/vertex_attr are created as torch.float, :87,117,165)."""
    n = b.shape[0]
    A = torch.sparse_coo_tensor(edge_index, A_val.flatten(), (n, n), dtype=torch.float)
    ei, ea = coo_to_gnn_input(A)

    def run_jacobi(k, xx):                                                 # :156-173
        e2 = torch.cat([ea, torch.zeros_like(ea)], 1)
        dv = diag_value * torch.ones((n, 1), dtype=torch.float)
        return jacobi(k, torch.cat([dv, b, xx], 1), ei, e2, torch.tensor(w).reshape(-1))

    x = run_jacobi(n_pre, x)                                               # :194-196
    eo, ao = remove_diag_entries(ei, ea)                                   # :80
    S = soc_classic(theta, torch.zeros((n, 1), dtype=torch.float), eo, ao).reshape(-1, 1) > 0   # :87-90
    e_di = torch.hstack([ao, S])                                           # :103
    dv = diag_value * torch.ones((n, 1), dtype=torch.float)
    v_di = torch.hstack([dv, splitting.reshape(-1, 1)])                    # :120
    w_ij = direct_interp(v_di, eo, e_di)                                   # :123
    W = torch.sparse_coo_tensor(eo, w_ij, (n, n), dtype=torch.float)
    W = (torch.eye(n) + W).to_dense()                                      # :129-131
    P = W[:, splitting.flatten() > 0].to_sparse()                          # :133-137
    Ac = P.t() @ (A @ P)                                                   # :209
    r = residual(torch.cat([b, x], 1), ei, ea)                             # :212
    rc = P.t() @ r                                                         # :215
    xc = torch.zeros_like(rc)
    eci, eca = coo_to_gnn_input(Ac)
    vc, _, _ = chebyshev(cheb_deg, torch.cat([rc, xc], 1), eci, eca,
                         torch.tensor([coarse_c, coarse_d]))               # :221-223,139-154
    x = x + P @ vc[:, 1].reshape(-1, 1)                                    # :226
    return run_jacobi(n_post, x)                                           # :229-231


def prolongator(edge_index_off, w_ij, splitting, n, coarse_rows_identity=False):
    """VCycle.py:126-137 restated on its own: P = (eye(n) + W)[:, splitting > 0].to_sparse(),
    through the reference's DENSE n x n intermediate (small n only).  With
    ``coarse_rows_identity`` the W entries of coarse rows are zeroed first -- the MATLAB twin's
    rule (matlab/test_direct_interpolation.m:130-132); the Python reference keeps them (NaN when a
    coarse row has no strong coarse neighbour)."""
    w = w_ij.clone()
    if coarse_rows_identity:
        w[splitting.flatten()[edge_index_off[0]] > 0] = 0
    W = torch.sparse_coo_tensor(edge_index_off, w, (n, n), dtype=w.dtype)
    W = (torch.eye(n, dtype=w.dtype) + W).to_dense()                        # :129-131
    return W[:, splitting.flatten() > 0].to_sparse().coalesce()            # :133-137


def galerkin(A, P):
    """VCycle.py:209 -- Ac = P^T (A P) on torch.sparse COO tensors (CPU)."""
    return (P.t() @ (A @ P)).coalesce()
