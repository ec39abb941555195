"""CPU: the oracle pieces behind the AMG-setup kernels.  port.prolongator / port.galerkin are
pinned to port.two_grid_vcycle (itself pinned bit-for-bit to the reference's runVCycle by
test_oracle_pinning.py / test_oracle_golden.py); oracle/cf_split.py (parity unpinned against the
reference, see its header) is checked for its defining properties and known hash values."""
import numpy as np
import torch

from oracle import cf_split, port


def test_mix32_known_answers_and_bijection():
    # murmur3 fmix32 test values
    assert cf_split.mix32(np.array([0, 1, 2], dtype=np.uint32)).tolist() == [0, 0x514E28B7, 0x30F4C306]
    h = cf_split.mix32(np.arange(1 << 16, dtype=np.uint32))
    assert np.unique(h).size == 1 << 16


def _laplace_strength(N):
    n = N * N
    ei, ev = port.laplacian_2d(N)
    ev = ev.float()
    eo, ao = port.remove_diag_entries(ei, ev)
    S = port.soc_classic(0.25, torch.zeros(n, 1), eo, ao)
    return n, ei, ev, eo, ao, S


def test_pmis_properties_and_determinism():
    n, ei, ev, eo, ao, S = _laplace_strength(21)
    strong = (S > 0).numpy()
    r, c = eo[0].numpy(), eo[1].numpy()
    a, rounds = cf_split.pmis(n, r, c, strong, 0)
    b, _ = cf_split.pmis(n, r, c, strong, 0)
    assert np.array_equal(a, b) and rounds >= 1
    assert not np.array_equal(a, cf_split.pmis(n, r, c, strong, 1)[0])      # the seed matters
    cm = a.astype(bool)
    assert not (cm[r[strong]] & cm[c[strong]]).any()                       # independent set
    covered = np.zeros(n, dtype=bool)
    covered[r[strong][cm[c[strong]]]] = True
    assert covered[~cm].all()                                              # every F has a strong C
    # edge order does not matter (max / counts are order independent)
    p = np.random.default_rng(0).permutation(r.size)
    assert np.array_equal(a, cf_split.pmis(n, r[p], c[p], strong[p], 0)[0])
    # no strong edges at all: every vertex is isolated -> coarse, one round
    z, rounds = cf_split.pmis(5, np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(0, dtype=bool), 0)
    assert z.tolist() == [1] * 5 and rounds == 1
    assert cf_split.pmis(0, r[:0], c[:0], strong[:0])[0].size == 0


def test_prolongator_and_galerkin_helpers_match_the_pinned_cycle():
    N = 7
    n, ei, ev, eo, ao, S = _laplace_strength(N)
    split = torch.zeros(n)
    split[0::2] = 1
    torch.manual_seed(3)
    b, x = torch.rand(n, 1), torch.rand(n, 1)
    ref = port.two_grid_vcycle(ei, ev, b, x, split)
    # the same cycle from the separate helpers
    dv = -4.0 * torch.ones(n, 1)
    e2 = torch.cat([ev, torch.zeros_like(ev)], 1)
    w7 = torch.tensor(0.7).reshape(-1)
    x1 = port.jacobi(3, torch.cat([dv, b, x], 1), ei, e2, w7)
    wij = port.direct_interp(torch.hstack([dv, split.view(-1, 1)]), eo, torch.hstack([ao, (S.reshape(-1, 1) > 0)]))
    P = port.prolongator(eo, wij, split, n)
    A = torch.sparse_coo_tensor(ei, ev.flatten(), (n, n))
    Ac = port.galerkin(A, P)
    r = port.residual(torch.cat([b, x1], 1), ei, ev)
    rc = P.t() @ r
    eci, eca = port.coo_to_gnn_input(Ac)
    vc, _, _ = port.chebyshev(4, torch.cat([rc, torch.zeros_like(rc)], 1), eci, eca, torch.tensor([-3.4, -4.0]))
    x2 = x1 + P @ vc[:, 1].reshape(-1, 1)
    out = port.jacobi(3, torch.cat([dv, b, x2], 1), ei, e2, w7)
    assert torch.equal(out, ref)
    # MATLAB-twin rule only changes coarse rows
    P1 = port.prolongator(eo, wij, split, n, coarse_rows_identity=True).to_dense()
    Pd = P.to_dense()
    fine = split == 0
    assert torch.equal(P1[fine], Pd[fine])
    assert torch.equal(P1[~fine], torch.eye(n)[~fine][:, ~fine])


def _golden_setup():
    import os
    from conftest import ROOT
    return torch.load(os.path.join(ROOT, "tests", "golden", "reference_setup.pt"), weights_only=False)


def test_oracle_setup_helpers_match_the_reference_golden():
    """oracle port (SOC flags, prolongator, Galerkin operator) == the intermediates the UNMODIFIED
    reference produced (tests/golden/make_golden_setup.py), bit for bit."""
    for c in _golden_setup()["cases"]:
        N = c["N"]
        n, ei, ev, eo, ao, S = _laplace_strength(N)
        assert torch.equal(S.reshape(-1, 1) > 0, c["S"])
        split = torch.zeros(n)
        split[0::2] = 1
        dv = -4.0 * torch.ones(n, 1)
        wij = port.direct_interp(torch.hstack([dv, split.view(-1, 1)]), eo, torch.hstack([ao, c["S"]]))
        P = port.prolongator(eo, wij, split, n)
        assert tuple(P.shape) == c["P_shape"]
        assert torch.equal(P.indices(), c["P_indices"]) and torch.equal(P.values(), c["P_values"])
        A = torch.sparse_coo_tensor(ei, ev.flatten(), (n, n))
        Ac = port.galerkin(A, P)
        assert tuple(Ac.shape) == c["Ac_shape"]
        assert torch.equal(Ac.indices(), c["Ac_indices"]) and torch.equal(Ac.values(), c["Ac_values"])
