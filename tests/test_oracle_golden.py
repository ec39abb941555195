"""CPU: the oracle restatement (oracle/port.py) against the committed golden vectors that were
produced by the unmodified reference (tests/golden/make_golden.py) and against the reference's
own known answers.  Bit-exact."""
import torch

from conftest import same
from oracle import port


def _cases(golden):
    return golden["layers"]


def test_laplacian_generator(golden):
    for c in _cases(golden):
        ei, ev = port.laplacian_2d(c["N"])
        assert torch.equal(ei, c["edge_index"]) and same(ev, c["edge_val64"])


def test_matvec_residual(golden):
    for c in _cases(golden):
        ev = c["edge_val64"].to(c["dtype"])
        for key, x in (("matvec", c["x"]), ("matvec2", c["x2"])):
            v, e = port.matvec(x, c["edge_index"], ev)
            assert same(v, c[key][0]) and same(e, c[key][1])
        r = port.residual(torch.cat([c["b"], c["x"]], 1), c["edge_index"], ev)
        assert same(r, c["residual"])


def test_jacobi(golden):
    for c in _cases(golden):
        dt, n = c["dtype"], c["N"] ** 2
        ev = c["edge_val64"].to(dt)
        va = torch.cat([-4 * torch.ones(n, 1, dtype=dt), c["b"], c["x"]], 1)
        ea = torch.cat([ev, torch.zeros_like(ev)], 1)
        g = torch.tensor(0.7).reshape(-1)
        assert same(port.jacobi(10, va, c["edge_index"], ea, g), c["jacobi10"])
        for a, b in zip(port.jacobi_iterate(va, c["edge_index"], ea, g), c["jacobi_iterate"]):
            assert same(a, b)


def test_chebyshev(golden):
    for c in _cases(golden):
        ev = c["edge_val64"].to(c["dtype"])
        g = torch.tensor([-3.46, -4.0])
        for deg, ref in c["cheby"].items():
            out = port.chebyshev(deg, torch.cat([c["b"], c["x"]], 1), c["edge_index"], ev, g)
            for a, b in zip(out, ref):
                assert same(a, b)


def test_power_method(golden):
    for c in _cases(golden):
        dt = c["dtype"]
        ev = c["edge_val64"].to(dt)
        ea = torch.cat([ev, torch.zeros_like(ev)], 1)
        vp = torch.cat([c["x"], torch.zeros_like(c["x"])], 1)
        out = port.power_method(10, vp, c["edge_index"], ea, torch.zeros(3, dtype=dt))
        for a, b in zip(out, c["power10"]):
            assert same(a, b)


def test_amg_setup(golden):
    for c in _cases(golden):
        dt, n = c["dtype"], c["N"] ** 2
        ev = c["edge_val64"].to(dt)
        eo, ao = port.remove_diag_entries(c["edge_index"], ev)
        assert torch.equal(eo, c["off_index"]) and same(ao, c["off_val"])
        S = port.soc_classic(c["theta"], torch.zeros(n, 1, dtype=dt), eo, ao)
        assert same(S, c["soc_classic"])
        dv = -4 * torch.ones(n, 1, dtype=dt)
        assert same(port.soc_sa(dv, eo, ao), c["soc_sa"])
        ed = torch.hstack([ao, (S.reshape(-1, 1) > 0)])
        w = port.direct_interp(torch.hstack([dv, c["splitting"]]), eo, ed)
        assert same(w, c["direct_interp"])
        if c["N"] == 5:  # the 0*inf rows of the reference are reproduced (40 of 80 edges)
            assert torch.isnan(w).sum().item() == 40
        assert same(port.matrix_weighted_norm(c["x"], c["edge_index"], -ev), c["mwnorm"])


def test_vcycle(golden):
    for c in golden["vcycle"]:
        N = c["N"]
        ei, ev = port.laplacian_2d(N)
        split = torch.zeros(N * N)
        split[0::2] = 1
        x = c["x0"].clone()
        for ref_x in c["xs"]:
            x = port.two_grid_vcycle(ei, ev, c["b"], x, split)
            assert same(x, ref_x)


def test_known_answers(golden):
    k = golden["known"]["mv3"]
    v, _ = port.matvec(k["x"], k["edge_index"], k["A_ij"])
    assert v[:, 1].tolist() == [20.0, 301.0, 1030.0]                       # MatVecGNN.py:118-137
    v2, _ = port.matvec(k["x2"], k["edge_index"], k["A_ij"])
    assert v2[:, -2:].tolist() == [[20.0, 140.0], [301.0, 2107.0], [1030.0, 7210.0]]
    assert same(v[:, 1:], k["y"]) and same(v2[:, -2:], k["y2"])
    A = torch.sparse_coo_tensor(k["edge_index"], k["A_ij"].flatten()).to_dense()
    r = port.residual(torch.cat([A @ k["x"], k["x"]], 1), k["edge_index"], k["A_ij"])
    assert torch.count_nonzero(r) == 0                                     # GNNResidual.py:157-170
    p = golden["known"]["power3"]
    out = port.power_method(10, p["vertex_attr"], p["edge_index"], p["edge_attr"], torch.zeros(3))
    assert same(out[2], p["out"][2]) and abs(out[2][2].item() - 3.0) < 1e-3  # PowerMethodGNN.py:338-383
    # SOCSAGNN.py:77-98: every off-diagonal S_ij of laplacianfun_torch(5) is 0.0625
    ei, ev = port.laplacian_2d(5)
    eo, ao = port.remove_diag_entries(ei, ev)
    S = port.soc_sa(-4 * torch.ones(25, 1, dtype=torch.float64), eo, ao)[:, 1]
    assert torch.all(S == 0.0625)
    # SOCClassicGNN.py:151-186: all S_ij = 0.75 for theta = 0.25
    assert torch.all(port.soc_classic(0.25, torch.zeros(25, 1, dtype=torch.float64), eo, ao) == 0.75)
    # analytic: extreme eigenvalue of laplacianfun_torch(5) is -4 - 2*sqrt(3)
    x = torch.rand(25, 1, dtype=torch.float64, generator=torch.Generator().manual_seed(24601))
    ea = torch.cat([ev, torch.zeros_like(ev)], 1)
    lam = port.power_method(100, torch.cat([x, torch.zeros_like(x)], 1), ei, ea,
                            torch.zeros(3, dtype=torch.float64))[2][2].item()
    assert abs(lam - (-4 - 2 * 3 ** 0.5)) < 1e-6   # 100 iterations, gap ratio 0.93: not fully converged
