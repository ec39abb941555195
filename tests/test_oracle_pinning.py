"""CPU, build container only: the oracle restatement against the UNMODIFIED reference files
loaded from /root/reference (skipped where the reference is absent, e.g. on the GPU box --
there the committed golden vectors of test_oracle_golden.py carry the pin)."""
import pytest
import torch

from conftest import same
from oracle import port, ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference sources not present")


@pytest.fixture(scope="module")
def R():
    return ref_loader.load()


@pytest.mark.parametrize("N", [3, 12])
@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_layers_bit_exact(R, N, dt):
    torch.manual_seed(1000 + N)
    n = N * N
    ei, ev = R.UtilsGNN.laplacianfun_torch(N)
    ev = ev.to(dt)
    # perturb the values so products are not exactly representable
    ev = ev * (1 + 0.1 * torch.rand_like(ev))
    x, b, batch = torch.rand(n, 1, dtype=dt), torch.rand(n, 1, dtype=dt), torch.zeros(n)
    mv = R.MetaLayer(R.MatVecGNN.EdgeUpdate(), R.MatVecGNN.VertexUpdate(R.MatVecGNN.edge_to_vertex_aggregation))
    v, e, _ = mv(x, ei, ev, None, batch=batch)
    pv, pe = port.matvec(x, ei, ev)
    assert same(v, pv) and same(e, pe)
    assert same(R.GNNResidual.GNNResidual()(torch.cat([b, x], 1), ei, ev), port.residual(torch.cat([b, x], 1), ei, ev))
    diag = torch.rand(n, 1, dtype=dt) + 3
    va = torch.cat([diag, b, x], 1)
    ea = torch.cat([ev, torch.zeros_like(ev)], 1)
    g = torch.tensor(0.6).reshape(-1)
    assert same(R.JacobiGNN.JacobiGNN()(7, va, ei, ea, g), port.jacobi(7, va, ei, ea, g))
    gc = torch.tensor([-3.4, -4.0])
    for deg in (1, 2, 5):
        for a, b_ in zip(R.ChebyGNN.ChebyRelaxGNN(deg)(torch.cat([b, x], 1), ei, ev, gc),
                         port.chebyshev(deg, torch.cat([b, x], 1), ei, ev, gc)):
            assert same(a, b_)
    vp = torch.cat([x, torch.zeros_like(x)], 1)
    for a, b_ in zip(R.PowerMethodGNN.PowerMethodGNN(6)(vp, ei, ea, torch.zeros(3, dtype=dt), batch),
                     port.power_method(6, vp, ei, ea, torch.zeros(3, dtype=dt))):
        assert same(a, b_)
    eo, ao = R.UtilsGNN.remove_diag_entries(ei, ev)
    S = R.SOCClassicGNN.SOCClassicGNN(0.3)(torch.zeros(n, 1, dtype=dt), eo, ao)
    assert same(S, port.soc_classic(0.3, torch.zeros(n, 1, dtype=dt), eo, ao))
    assert same(R.MetaLayer(R.SOCSAGNN.EdgeUpdate())(diag, eo, ao, batch=batch)[1], port.soc_sa(diag, eo, ao))
    split = (torch.rand(n, 1, dtype=dt) > 0.5).to(dt)
    ed = torch.hstack([ao, S.reshape(-1, 1) > 0])
    assert same(R.DirectInterpGNN.DirectInterpGNN()(torch.hstack([diag, split]), eo, ed, None),
                port.direct_interp(torch.hstack([diag, split]), eo, ed))


def test_vcycle_bit_exact(R):
    V = R.VCycle
    N = 6
    V.N = N
    torch.manual_seed(5)
    ei, ev = R.UtilsGNN.laplacianfun_torch(N)
    A = torch.sparse_coo_tensor(ei, ev.flatten(), dtype=torch.float)
    x, b = torch.rand(N * N, 1), torch.rand(N * N, 1)
    split = torch.zeros(N * N)
    split[0::2] = 1
    xr, xp = x.clone(), x.clone()
    for _ in range(3):
        xr = V.runVCycle(A, b, xr, 3, 3, 5, True)
        xp = port.two_grid_vcycle(ei, ev, b, xp, split)
        assert same(xr, xp)
