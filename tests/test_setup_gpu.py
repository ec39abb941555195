"""GPU parity tests of the AMG-setup kernels: sparse prolongator
assembly, the SpGEMM behind the Galerkin operator and the PMIS coarse/fine splitting, each against
its CPU oracle on the same inputs.

Bars: sparsity patterns, C/F flags and the prolongator values bit-exact; SpGEMM values (both the
row-local and the expand-sort-compress path) bit-exact against a sequential restatement of the
documented summation order and <= 1e-5 / 1e-12 against torch.sparse on the CPU (the reference's
VCycle.py:209 path, whose summation order differs).
"""
import numpy as np
import pytest
import torch

from conftest import relerr, same
from oracle import cf_split, port

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-5, torch.float64: 1e-12}


def _setup_operator(G, kind, N, dt):
    """(edge_index, edge_attr) incl. diagonal, off-diagonal edges, diagonal vector -- on the CPU."""
    if kind == "laplace":
        ei, ev = port.laplacian_2d(N)
    elif kind == "aniso":
        ei, ev = G.generators.constant_diffusion_fem(1.0, 0.01, N, dtype=torch.float64)
    else:
        ei, ev = G.generators.heat_fem_2d((N + 1, N + 1), (1.0, 2.0), torch.float64, "cpu")
    ev = ev.to(dt)
    n = int(ei.max().item()) + 1
    eo, ao = port.remove_diag_entries(ei, ev)
    diag = G.generators.diagonal_of(ei, ev, n)
    return n, ei, ev, eo.contiguous(), ao.contiguous(), diag


def _strength(n, eo, ao, dt, theta=0.25):
    return port.soc_classic(theta, torch.zeros(n, 1, dtype=dt), eo, ao).reshape(-1, 1) > 0


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
@pytest.mark.parametrize("kind,N", [("laplace", 5), ("laplace", 12), ("aniso", 10), ("heat", 9)])
def test_prolongator_assembly_bit_exact(G, dev, dt, kind, N):
    """P = [I + W](:, C) from glab_interp_* == the reference's dense construction
    (VCycle.py:126-137), indices and values bit for bit, NaN entries included; both coarse-row
    rules; alternating, random and all-coarse / all-fine splittings."""
    rt = G.runtime
    n, ei, ev, eo, ao, diag = _setup_operator(G, kind, N, dt)
    S = _strength(n, eo, ao, dt)
    g = torch.Generator().manual_seed(5)
    splits = {"alternating": torch.zeros(n, dtype=dt), "random": (torch.rand(n, generator=g) < 0.4).to(dt),
              "all_coarse": torch.ones(n, dtype=dt), "all_fine": torch.zeros(n, dtype=dt)}
    splits["alternating"][0::2] = 1
    plan_off = rt.Plan.from_coo(eo.to(dev), n)
    for name, split in splits.items():
        w_ref = port.direct_interp(torch.hstack([diag, split.view(-1, 1)]), eo, torch.hstack([ao, S.to(dt)]))
        w_dev = w_ref.to(dev)     # the layer itself is covered by test_amg_setup_bit_exact
        for mode, ident in ((0, False), (1, True)):
            P_ref = port.prolongator(eo, w_ref, split, n, coarse_rows_identity=ident)
            pi, pv, nc = rt.interp_assemble(plan_off, w_dev, split.to(dev), mode)
            assert nc == int((split > 0).sum()), name
            assert tuple(P_ref.shape) == (n, nc)
            assert torch.equal(pi.cpu(), P_ref.indices()), (name, mode)
            assert same(pv.cpu(), P_ref.values()), (name, mode)
    # through the drop-in function, host tensors in -> host sparse tensor out
    A = torch.sparse_coo_tensor(ei, ev.flatten(), (n, n))
    P = G.VCycle.runDirectInterp(A, S, None, splits["alternating"])
    w_ref = port.direct_interp(torch.hstack([diag, splits["alternating"].view(-1, 1)]), eo,
                               torch.hstack([ao, S.to(dt)]))
    P_ref = port.prolongator(eo, w_ref, splits["alternating"], n)
    assert not P.is_cuda and P.is_coalesced()
    assert torch.equal(P.indices(), P_ref.indices()) and same(P.values(), P_ref.values())


def _random_sparse(m, k, density, dt, seed, empty_rows=()):
    g = torch.Generator().manual_seed(seed)
    mask = torch.rand(m, k, generator=g) < density
    for r in empty_rows:
        mask[r] = False
    vals = (torch.rand(m, k, generator=g, dtype=torch.float64) - 0.5).to(dt)
    idx = mask.nonzero().t().contiguous()                 # row-major sorted, unique
    return idx, vals[mask].contiguous(), mask, torch.where(mask, vals, torch.zeros_like(vals))


def _sequential_product(Xm, Xd, Ym, Yd):
    """Dense emulation of the documented summation order: for every output entry the products
    x_ij * y_jl over the structurally present j in ascending order, added one at a time in T."""
    m, k = Xd.shape
    n = Yd.shape[1]
    acc = torch.zeros(m, n, dtype=Xd.dtype)
    seen = torch.zeros(m, n, dtype=torch.bool)
    for j in range(k):
        present = Xm[:, j:j + 1] & Ym[j:j + 1, :]
        prod = Xd[:, j:j + 1] * Yd[j:j + 1, :]
        acc = torch.where(present & ~seen, prod, torch.where(present, acc + prod, acc))
        seen |= present
    return seen, acc


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_spgemm_bit_exact_and_vs_torch_sparse(G, dev, dt):
    rt = G.runtime
    paths = set()
    for (m, k, n, dx, dy, seed) in ((40, 30, 50, 0.15, 0.2, 1), (200, 200, 200, 0.03, 0.03, 2), (7, 1, 9, 0.9, 0.9, 3),
                                    (64, 80, 3, 0.1, 0.5, 4), (40, 30, 300, 0.5, 0.5, 5), (300, 70, 64, 0.2, 0.9, 6)):
        xi, xv, Xm, Xd = _random_sparse(m, k, dx, dt, seed, empty_rows=(0, m - 1))
        yi, yv, Ym, Yd = _random_sparse(k, n, dy, dt, seed + 100, empty_rows=(k // 2,) if k > 1 else ())
        px = rt.Plan.from_coo(xi.to(dev), m, k)
        py = rt.Plan.from_coo(yi.to(dev), k, n)
        zi, zv = rt.spgemm(px, xv.to(dev), py, yv.to(dev))
        paths.add(rt.spgemm.last["path"])
        seen, acc = _sequential_product(Xm, Xd, Ym, Yd)
        ref_idx = seen.nonzero().t().contiguous()
        assert torch.equal(zi.cpu(), ref_idx)                              # pattern + (row, col) order
        assert same(zv.cpu(), acc[seen])                                   # values, bit for bit
        ts = (torch.sparse_coo_tensor(xi, xv, (m, k)) @ torch.sparse_coo_tensor(yi, yv, (k, n))).to_dense()
        mine = torch.sparse_coo_tensor(zi.cpu(), zv.cpu(), (m, n)).to_dense()
        assert relerr(mine, ts) <= TOL[dt]
    assert {"row-local", "esc"} <= paths, paths        # both paths ran (and agree with the same emulation)
    # empty operands
    pe = rt.Plan.from_coo(torch.zeros(2, 0, dtype=torch.int64, device=dev), 5, 4)
    py = rt.Plan.from_coo(torch.tensor([[0, 3], [1, 2]], device=dev), 4, 6)
    zi, zv = rt.spgemm(pe, torch.zeros(0, dtype=dt, device=dev), py, torch.ones(2, dtype=dt, device=dev))
    assert zi.shape == (2, 0) and zv.numel() == 0
    # arbitrary (unsorted, duplicated) edge order of X: duplicates are separate products, summed
    xi = torch.tensor([[2, 0, 2, 1, 2], [1, 0, 1, 3, 0]])
    xv = torch.tensor([1.5, 2.0, 0.25, -1.0, 4.0], dtype=dt)
    yi = torch.tensor([[0, 1, 3], [1, 1, 2]])
    yv = torch.tensor([10.0, 1000.0, 100.0], dtype=dt)
    px = rt.Plan.from_coo(xi.to(dev), 3, 4)
    assert not px.identity
    py2 = rt.Plan.from_coo(yi.to(dev), 4, 6)
    xv_dev = xv.view(-1, 1).to(dev)
    zi, zv = rt.spgemm(px, rt.get_vals(px, xv_dev), py2, yv.to(dev))
    assert torch.equal(zi.cpu(), torch.tensor([[0, 1, 2], [1, 2, 1]]))
    assert torch.equal(zv.cpu(), torch.tensor([20.0, -100.0, 1500.0 + 250.0 + 40.0], dtype=dt))
    # argument errors come back as codes
    import ctypes
    n_prod = ctypes.c_int64()
    assert G.lib.glab_spgemm_products(py.handle, py.handle, None, ctypes.byref(n_prod), ctypes.byref(n_prod),
                                      None) == -1                          # inner dimensions 6 != 4


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_galerkin_operator_vs_oracle(G, dev, dt):
    """A_c = P^T A P of the cached two-grid hierarchy against torch.sparse on the CPU
    (VCycle.py:209): identical pattern, values within the tolerance."""
    N = 12
    n, ei, ev, eo, ao, diag = _setup_operator(G, "laplace", N, dt)
    A = torch.sparse_coo_tensor(ei, ev.flatten(), (n, n))
    split = torch.zeros(n, dtype=dt)
    split[0::2] = 1
    S = _strength(n, eo, ao, dt)
    w = port.direct_interp(torch.hstack([diag, split.view(-1, 1)]), eo, torch.hstack([ao, S.to(dt)]))
    P_ref = port.prolongator(eo, w, split, n)
    Ac_ref = port.galerkin(A, P_ref)
    tg = G.VCycle._two_grid(A.to(dev), None)
    Ac = tg.Ac.cpu()
    assert Ac.is_coalesced() and tuple(Ac.shape) == tuple(Ac_ref.shape)
    assert relerr(Ac.to_dense(), Ac_ref.to_dense()) <= TOL[dt]
    ref_pattern = (P_ref.to_dense() != 0).double().t() @ (A.to_dense() != 0).double() @ (P_ref.to_dense() != 0).double() > 0
    mine = torch.zeros_like(ref_pattern)
    mine[Ac.indices()[0], Ac.indices()[1]] = True
    assert torch.equal(mine, ref_pattern)


def _check_split_properties(n, rs, cs, cflag):
    c = cflag.astype(bool)
    assert not (c[rs] & c[cs]).any()                       # coarse points are independent in S
    has_c = np.zeros(n, dtype=bool)
    has_c[rs[c[cs]]] = True
    assert has_c[~c].all()                                 # every fine point depends on a coarse one


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_cf_split_pmis_bit_exact(G, dev, dt):
    rt = G.runtime
    cases = [("laplace", 33, 0.25), ("aniso", 24, 0.25), ("heat", 17, 0.25), ("aniso", 12, 0.0)]
    for kind, N, theta in cases:
        n, ei, ev, eo, ao, diag = _setup_operator(G, kind, N, dt)
        S = port.soc_classic(theta, torch.zeros(n, 1, dtype=dt), eo, ao)
        plan_off = rt.Plan.from_coo(eo.to(dev), n)
        for seed in (0, 24601, 2 ** 32 - 1):
            ref, rounds_ref = cf_split.pmis(n, eo[0].numpy(), eo[1].numpy(), (S > 0).numpy(), seed)
            cflag, rounds = rt.cf_split_pmis(plan_off, S.to(dev), seed)
            assert cflag.dtype == dt
            assert np.array_equal(cflag.cpu().numpy().astype(np.uint8), ref), (kind, N, seed)
            assert rounds == rounds_ref
            strong = (S > 0).numpy()
            _check_split_properties(n, eo[0].numpy()[strong], eo[1].numpy()[strong], ref)
    # random directed strength graph with isolated vertices (they become coarse)
    g = torch.Generator().manual_seed(9)
    n = 500
    rows = torch.randint(0, n - 10, (3000,), generator=g)
    cols = torch.randint(0, n - 10, (3000,), generator=g)
    keep = rows != cols
    key = torch.unique(rows[keep] * n + cols[keep])
    eo = torch.stack([key // n, key % n])
    S = (torch.rand(eo.shape[1], generator=g) - 0.4).to(dt)
    plan_off = rt.Plan.from_coo(eo.to(dev), n)
    ref, rounds_ref = cf_split.pmis(n, eo[0].numpy(), eo[1].numpy(), (S > 0).numpy(), 7)
    cflag, rounds = rt.cf_split_pmis(plan_off, S.to(dev), 7)
    assert np.array_equal(cflag.cpu().numpy().astype(np.uint8), ref) and rounds == rounds_ref
    assert ref[n - 10:].all()
    # through the drop-in function
    n, ei, ev, eo, ao, diag = _setup_operator(G, "laplace", 20, dt)
    A = torch.sparse_coo_tensor(ei, ev.flatten(), (n, n))
    Sb = G.VCycle.runSOC(A)
    split = G.VCycle.runCFSplit(A, Sb, seed=3)
    ref, _ = cf_split.pmis(n, eo[0].numpy(), eo[1].numpy(), Sb.flatten().numpy(), 3)
    assert split.shape == (n, 1) and np.array_equal(split.flatten().numpy().astype(np.uint8), ref)


def test_setup_at_scale_properties(G, dev):
    """1 M-row Laplacian (4.2 M off-diagonal edges): the device PMIS equals the numpy oracle; the
    prolongator built on it is finite and reproduces constants on interior rows (direct
    interpolation rows sum to 1); and the Galerkin operator from the device SpGEMM satisfies
    A_c v == P^T (A (P v)) computed with the SpMV kernels -- a size-independent check of the
    expand-sort-compress product."""
    V, rt = G.VCycle, G.runtime
    N = 1024
    n = N * N
    ei, ev = G.UtilsGNN.laplacianfun_torch(N, device=dev)
    A = torch.sparse_coo_tensor(ei, ev.flatten().float(), (n, n))
    S = V.runSOC(A)
    split = V.runCFSplit(A, S, seed=1)
    op = V._operator(A)
    ref, _ = cf_split.pmis(n, op.off_index[0].cpu().numpy(), op.off_index[1].cpu().numpy(), S.flatten().cpu().numpy(), 1)
    assert np.array_equal(split.flatten().cpu().numpy().astype(np.uint8), ref)
    frac = float(split.mean().item())
    assert 0.2 < frac < 0.5, frac
    tg = V._two_grid(A, split, coarse_rows="identity")
    P = tg.P
    nc = P.shape[1]
    assert nc == int(ref.sum()) and bool(torch.isfinite(P.values()).all())
    ones_c = torch.ones(nc, 1, device=dev)
    p1 = rt.spmm(tg.plan_P, tg.vals_P, ones_c).view(N, N)
    assert float((p1[1:-1, 1:-1] - 1).abs().max().item()) <= 1e-6
    torch.manual_seed(24601)
    v = torch.rand(nc, 1, device=dev)
    cop = V._operator(tg.Ac)
    plan_C = rt.get_plan(cop.edge_index, nc)
    lhs = rt.spmm(plan_C, rt.get_vals(plan_C, cop.edge_attr), v)
    plan_A = rt.get_plan(op.edge_index, n)
    apv = rt.spmm(plan_A, rt.get_vals(plan_A, op.edge_attr), rt.spmm(tg.plan_P, tg.vals_P, v))
    rhs = rt.spmm(tg.plan_PT, tg.vals_PT, apv)
    assert relerr(lhs, rhs) <= 1e-5
    # the coarse operator is symmetric in pattern (A and the strength graph are)
    ci = cop.edge_index
    fwd = ci[0] * nc + ci[1]
    bwd = torch.sort(ci[1] * nc + ci[0]).values
    assert torch.equal(fwd, bwd)


def test_setup_against_reference_golden(G, dev):
    """runSOC / runDirectInterp / the Galerkin operator of the cached hierarchy against the
    intermediates of the UNMODIFIED reference two-grid cycle (tests/golden/reference_setup.pt):
    strength flags and the prolongator bit for bit, P^T A P with the identical pattern and values
    within the fp32 tolerance (the reference's torch.sparse product sums in a different order)."""
    import os
    from conftest import ROOT
    V = G.VCycle
    for c in torch.load(os.path.join(ROOT, "tests", "golden", "reference_setup.pt"), weights_only=False)["cases"]:
        N = c["N"]
        ei, ev = G.UtilsGNN.laplacianfun_torch(N)
        A = torch.sparse_coo_tensor(ei, ev.flatten(), dtype=torch.float)
        S = V.runSOC(A)
        assert torch.equal(S, c["S"])
        P = V.runDirectInterp(A, S, N)
        assert tuple(P.shape) == c["P_shape"]
        assert torch.equal(P.indices(), c["P_indices"]) and same(P.values(), c["P_values"])
        Ac = V._two_grid(A, None).Ac.cpu()
        assert tuple(Ac.shape) == c["Ac_shape"] and torch.equal(Ac.indices(), c["Ac_indices"])
        assert relerr(Ac.values(), c["Ac_values"]) <= 1e-5
