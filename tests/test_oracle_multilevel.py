"""CPU: the multilevel oracle (oracle/ml_sa.py) -- structural properties of the aggregation and the
convergence of the W-cycle it defines.  The device implementation (glab_b200.multilevel) is compared
with it in tests/test_multilevel_gpu.py."""
import numpy as np
import scipy.sparse as sp

from oracle import ml_sa


def laplacian(N):
    T = sp.diags([1.0, -2.0, 1.0], [-1, 0, 1], shape=(N, N))
    return (sp.kron(sp.eye(N), T) + sp.kron(T, sp.eye(N))).tocsr()


def test_aggregates_partition_the_vertices_and_roots_are_distance_two_independent():
    A = laplacian(40)
    n = A.shape[0]
    r, c, keep = ml_sa.strength_mask(A, 0.08, np.float64)
    agg, na, root = ml_sa.aggregates(n, r, c, keep)
    assert agg.min() == 0 and agg.max() == na - 1 and np.unique(agg).size == na
    H = sp.csr_matrix((np.ones(keep.sum()), (r[keep], c[keep])), shape=(n, n))
    H2 = (H @ H).tocoo()
    off = H2.row != H2.col
    assert not (root[H2.row[off]] & root[H2.col[off]]).any()          # no two roots within distance 2
    sizes = np.bincount(agg)
    assert sizes.max() <= 13 and sizes.mean() > 4                     # aggregates of a 5-point grid: ~ 3 x 3 patches
    agg2, na2, _ = ml_sa.aggregates(n, r, c, keep)
    assert na2 == na and np.array_equal(agg, agg2)                    # deterministic


def test_w_cycle_contracts_independently_of_the_grid_size():
    rates = {}
    for N in (32, 64, 96):
        A = laplacian(N)
        lv = ml_sa.build(A)
        assert lv[-1]["A"].shape[0] <= 400 and len(lv) >= 2
        b = np.random.default_rng(1).random((N * N, 2))
        x = np.zeros_like(b)
        norms = [np.linalg.norm(b, axis=0)]
        for _ in range(6):
            x = ml_sa.cycle(lv, b, x)
            norms.append(np.linalg.norm(b - A @ x, axis=0))
        red = [float((norms[i + 1] / norms[i]).max()) for i in range(2, 6)]
        rates[N] = max(red)
    assert all(v <= 0.40 for v in rates.values()), rates
    # the reference's two-grid cycle (Chebyshev-4 coarse "solve") is at 0.98+ on these sizes: see DESIGN.md


# ---------------------------------------------------------------------------------------------------
# The multilevel cycle has no counterpart in the reference ("parity unpinned" as a whole), but every building
# block of oracle/ml_sa.py is a formula of a reference layer.  These tests pin the blocks to oracle/port.py --
# which tests/test_oracle_pinning.py / test_oracle_golden.py pin bit for bit to the unmodified reference layers --
# so that what stays unpinned is exactly the composition (aggregation rule, prolongator smoothing, recursion).
def _coo(A):
    import torch
    C = A.tocoo()
    order = np.lexsort((C.col, C.row))
    ei = torch.tensor(np.stack([C.row[order], C.col[order]]), dtype=torch.int64)
    ev = torch.tensor(C.data[order], dtype=torch.float64).view(-1, 1)
    return ei, ev


def test_strength_values_are_the_reference_sa_strength_layer():
    """ml_sa.strength_mask thresholds S_ij = (A_ij * A_ij) / (A_ii * A_jj): SOCSAGNN.py:67 via port.soc_sa."""
    import torch
    from oracle import port
    A = laplacian(12) + sp.diags(np.linspace(0.0, 0.5, 144))           # non-constant diagonal
    A = sp.csr_matrix(A)
    ei, ev = _coo(A)
    diag = torch.tensor(A.diagonal()).view(-1, 1)
    S_ref = port.soc_sa(diag, ei, ev)[:, 1].numpy()                    # reference formula on every edge: [A_ij, S_ij]
    r, c, keep = ml_sa.strength_mask(A, 0.3, np.float64)
    assert np.array_equal(r, ei[0].numpy()) and np.array_equal(c, ei[1].numpy())
    want = (S_ref >= np.float64(0.3) * np.float64(0.3)) | (r == c)
    assert np.array_equal(keep, want)


def test_smoother_residual_and_transfers_are_the_reference_layers():
    """One level visit of ml_sa.cycle restated with the pinned port: Jacobi sweeps (JacobiGNN.py:119), residual
    (GNNResidual.py:115), restriction / prolongation as matvec blocks (MatVecGNN.py:109-114), to 1e-13."""
    import torch
    from oracle import port
    N = 10
    A = sp.csr_matrix(laplacian(N))
    lv = ml_sa.build(A, np.float64, coarsest_n=30)
    assert len(lv) >= 2
    L = lv[0]
    n = A.shape[0]
    rng = np.random.default_rng(3)
    b, x = rng.random((n, 1)), rng.random((n, 1))
    ei, ev = _coo(A)
    tb, tx = torch.tensor(b), torch.tensor(x)
    diag = torch.tensor(L["d"]).view(-1, 1)
    ea = torch.cat([ev, torch.zeros_like(ev)], 1)
    g = torch.tensor([L["w"]], dtype=torch.float64)
    # three sweeps
    x_np = x.copy()
    for _ in range(3):
        x_np = x_np + (L["w"] * (b - A @ x_np)) / L["d"].reshape(-1, 1)
    x_port = port.jacobi(3, torch.cat([diag, tb, tx], 1), ei, ea, g)
    assert np.allclose(x_np, x_port.numpy(), rtol=1e-13, atol=1e-14)
    # residual
    r_np = b - A @ x_np
    r_port = port.residual(torch.cat([tb, x_port], 1), ei, ev)
    assert np.allclose(r_np, r_port.numpy().reshape(n, -1)[:, -1:], rtol=1e-12, atol=1e-13)
    # restriction and prolongation: matvec blocks with P^T and P
    P = L["P"].tocsr()
    pi, pv = _coo(P)
    ti, tv = _coo(P.T.tocsr())
    # (rectangular operators: the MatVec block's edge update c_ij = A_ij * x_j and its scatter-sum seam,
    #  MatVecGNN.py:84 and :60, applied with the row count of the target space)
    rc_np = P.T @ r_np
    rc_port = port.scatter_sum(tv * torch.tensor(r_np)[ti[1]], ti[0], P.shape[1])
    assert np.allclose(rc_np, rc_port.numpy(), rtol=1e-12, atol=1e-13)
    xc = rng.random((P.shape[1], 1))
    px_port = port.scatter_sum(pv * torch.tensor(xc)[pi[1]], pi[0], n)
    assert np.allclose(P @ xc, px_port.numpy(), rtol=1e-12, atol=1e-13)
    # square case through the whole MatVec block
    y_port = port.matvec(tx, ei, ev)[0][:, 1:2]
    assert np.allclose(A @ x, y_port.numpy(), rtol=1e-12, atol=1e-13)
    # Galerkin operator: A_c = P^T (A P)   (VCycle.py:209)
    Ac = lv[1]["A"]
    assert abs(Ac - (P.T @ (A @ P))).max() < 1e-12


def test_spectral_radius_estimate_is_the_reference_power_method():
    """ml_sa.rho_dinv_a = |Rayleigh quotient| after power_iters iterations: PowerMethodGNN.py:296-334 through
    port.power_method on the operator D^-1 A with the same start vector."""
    import torch
    from oracle import port
    A = sp.csr_matrix(laplacian(9))
    n = A.shape[0]
    M = sp.csr_matrix(sp.diags(1.0 / A.diagonal()) @ A)
    ei, ev = _coo(M)
    x0 = torch.tensor(ml_sa.start_vector(n, np.float64)).view(-1, 1)
    out = port.power_method(15, torch.cat([x0, torch.zeros_like(x0)], 1), ei, torch.cat([ev, torch.zeros_like(ev)], 1),
                            torch.zeros(3, dtype=torch.float64))
    g = out[2] if isinstance(out, (tuple, list)) else out
    lam_port = float(g.reshape(-1)[2])
    rho = ml_sa.rho_dinv_a(A, 15, np.float64)
    # (the reference builds its global attribute with torch.tensor([...]): fp32 -- PowerMethodGNN.py:185,266,294)
    assert abs(rho - abs(lam_port)) <= 1e-6 * abs(lam_port), (rho, lam_port)
