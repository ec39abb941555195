"""CPU: operator generators against the oracle / literal restatements of the reference
generators (UtilsGNN.py:53-67, matlab/heateqnfem2dfun.m:52-172, DiffCoeffs/FEM.py:184-198)."""
import numpy as np
import torch

from oracle import port


def test_laplacian_matches_oracle(G):
    for N in (1, 2, 5, 9):
        ei, ev = G.UtilsGNN.laplacianfun_torch(N)
        oi, ov = port.laplacian_2d(N)
        assert torch.equal(ei, oi) and torch.equal(ev, ov) and ev.dtype == torch.float64
    ei, ev = G.UtilsGNN.laplacianfun_torch(64)
    assert ei.shape[1] == 20224


def _heat_literal(ncx, ncy, hx, hy):
    """Literal restatement of the element loop of heateqnfem2dfun.m:91-121 + bcs=[2,2] (:161-165)."""
    a = hy / hx
    ev = (1 / 6 / a) * np.array([2 * a * a + 2, -2 * a * a + 1, a * a - 2, -1 - a * a])
    xn = np.array([[0, 1, 0, 0], [1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]])
    yn = np.array([[0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0], [1, 0, 0, 0]])
    cn = np.array([[0, 0, 1, 0], [0, 0, 0, 1], [1, 0, 0, 0], [0, 1, 0, 0]])
    EK = ev[0] * np.eye(4) + ev[1] * xn + ev[2] * yn + ev[3] * cn
    npts = (ncx + 1) * (ncy + 1)
    K = np.zeros((npts, npts))
    for Y in range(1, ncy + 1):
        for X in range(1, ncx + 1):
            base = (ncx + 1) * (Y - 1) + X
            Gi = np.array([base, base + 1, base + ncx + 2, base + ncx + 1]) - 1
            K[np.ix_(Gi, Gi)] += EK
    idx = np.arange(npts)
    gy, gx = idx // (ncx + 1), idx % (ncx + 1)
    interior = (gx > 0) & (gx < ncx) & (gy > 0) & (gy < ncy)
    return K[np.ix_(interior, interior)]


def test_heat_fem(G):
    for ncx, ncy, hx, hy in ((5, 5, 1, 1), (6, 4, 1, 2), (4, 7, 0.5, 3.5)):
        K = _heat_literal(ncx, ncy, hx, hy)
        ei, ev = G.generators.heat_fem_2d((ncx, ncy), (hx, hy))
        n = (ncx - 1) * (ncy - 1)
        A = torch.sparse_coo_tensor(ei, ev.flatten(), (n, n)).to_dense().numpy()
        assert np.abs(A - K).max() < 1e-14 and (A != 0).sum() == (K != 0).sum()
        assert torch.all(ei[0][1:] >= ei[0][:-1])
    w = G.generators.heat_fem_stencil(1.0, 1.0)
    assert abs(w[(0, 0)] - 8 / 3) < 1e-15 and abs(w[(1, 1)] + 1 / 3) < 1e-15


def test_constant_diffusion_fem(G, golden):
    ei, ev = G.generators.constant_diffusion_fem(1.0, 0.01, 6)
    assert ei.shape == (2, 324)
    A = torch.sparse_coo_tensor(ei, ev.flatten(), (36, 36)).to_dense()
    assert torch.allclose(A, A.T) and A.sum(1).abs().max() < 1e-13     # symmetric, zero row sums (periodic)
    assert (A[0, 6] > 0) and (A[0, 1] < 0)                             # positive N/S coupling for beta << alpha
    c = torch.sparse_coo_tensor(ei, ev.flatten(), (36, 36)).coalesce()
    assert torch.equal(c.indices(), ei)                               # already coalesced / row-major sorted
