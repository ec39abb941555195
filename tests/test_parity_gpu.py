"""GPU parity tests: every drop-in layer (CUDA kernels behind the C ABI) against the CPU oracle
(oracle/port.py, pinned to the reference) on the same seeded inputs, and against the committed
golden vectors produced by the unmodified reference.

Bars (BASELINE.json north_star): bit-exact strength masks / interpolation patterns; <= 1e-5
relative error in fp32 and <= 1e-12 in fp64 for matvec, smoother, eigenvalue and V-cycle
outputs.  Because the kernels accumulate in edge order without FMA contraction, most outputs
are in fact compared BIT FOR BIT; only reductions over all vertices (norms, Rayleigh quotient)
and the torch.sparse glue of the V-cycle use the tolerances.
"""
import ctypes

import os

import pytest
import torch

from conftest import relerr, same
from oracle import port

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-5, torch.float64: 1e-12}


def lap(N, dt):
    ei, ev = port.laplacian_2d(N)
    return ei, ev.to(dt)


def rand_problem(n, z, dt, seed, with_diag=True):
    """Random sparse operator in ARBITRARY edge order with duplicate (i,j) pairs, empty rows and
    one long row (exercises the stable sort, the permutation and the chunked staging path)."""
    g = torch.Generator().manual_seed(seed)
    rows = torch.randint(0, n, (z,), generator=g)
    cols = torch.randint(0, n, (z,), generator=g)
    rows[rows == 3] = 4                                    # row 3 is empty
    long_cols = torch.randint(0, n, (min(6000, 4 * n),), generator=g)
    rows = torch.cat([rows, torch.full_like(long_cols, 7), rows[:50]])   # long row 7 + duplicates
    cols = torch.cat([cols, long_cols, cols[:50]])
    if with_diag:
        d = torch.arange(n)
        d = d[d != 3]
        rows, cols = torch.cat([rows, d]), torch.cat([cols, d])
    perm = torch.randperm(rows.numel(), generator=g)
    ei = torch.stack([rows[perm], cols[perm]])
    ev = (torch.rand(ei.shape[1], 1, generator=g, dtype=torch.float64) - 0.5).to(dt)
    return ei, ev


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
@pytest.mark.parametrize("N", [5, 17, 64])
@pytest.mark.parametrize("k", [1, 2, 4, 8])
def test_matvec_bit_exact(G, dev, dt, N, k):
    torch.manual_seed(24601)
    ei, ev = lap(N, dt)
    x = torch.rand(N * N, k, dtype=dt)
    ref_v, ref_e = port.matvec(x, ei, ev)
    layer = G.MetaLayer(G.MatVecGNN.EdgeUpdate(), G.MatVecGNN.VertexUpdate(G.MatVecGNN.edge_to_vertex_aggregation))
    v, e, _ = layer(x.to(dev), ei.to(dev), ev.to(dev), None, batch=torch.zeros(N * N))
    assert v.is_cuda and same(v.cpu(), ref_v) and same(e.cpu(), ref_e)


def test_matvec_config1(G, dev):
    """BASELINE config 1: laplacianfun_torch(64), fp64, seed 24601, k = 1 and 2 (tolerance 1e-12)."""
    torch.manual_seed(24601)
    ei, ev = G.UtilsGNN.laplacianfun_torch(64)
    assert ei.shape[1] == 20224
    for k in (1, 2):
        x = torch.rand(4096, k, dtype=torch.float64)
        ref_v, _ = port.matvec(x, ei, ev)
        v, _, _ = G.MatVecGNN.MatVecGNN()(x.to(dev), ei.to(dev), ev.to(dev), None, None)
        assert relerr(v.cpu()[:, k:], ref_v[:, k:]) <= 1e-12
        A = torch.sparse_coo_tensor(ei, ev.flatten(), (4096, 4096)).to_sparse_csr()
        assert relerr(v.cpu()[:, k:], A @ x) <= 1e-12


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_host_tensors_in_host_tensors_out(G, dt):
    """The reference's users hold CPU tensors: the drop-in uploads, computes on the GPU and
    returns CPU tensors."""
    torch.manual_seed(3)
    ei, ev = lap(9, dt)
    b, x = torch.rand(81, 1, dtype=dt), torch.rand(81, 1, dtype=dt)
    r = G.GNNResidual.GNNResidual()(torch.cat([b, x], 1), ei, ev)
    assert not r.is_cuda and same(r, port.residual(torch.cat([b, x], 1), ei, ev))


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
@pytest.mark.parametrize("N", [5, 33])
def test_jacobi_bit_exact(G, dev, dt, N):
    torch.manual_seed(24601)
    n = N * N
    ei, ev = lap(N, dt)
    ev = ev * (1 + 0.05 * torch.rand_like(ev))
    va = torch.cat([torch.rand(n, 1, dtype=dt) + 3.5, torch.rand(n, 1, dtype=dt), torch.rand(n, 1, dtype=dt)], 1)
    ea = torch.cat([ev, torch.zeros_like(ev)], 1)
    g = torch.tensor(0.7).reshape(-1)
    J = G.JacobiGNN.JacobiGNN()
    for iters in (0, 1, 2, 10):
        out = J(iters, va.to(dev), ei.to(dev), ea.to(dev), g)
        assert same(out.cpu(), port.jacobi(iters, va, ei, ea, g)), iters
    out = J.iterate(va.to(dev), ei.to(dev), ea.to(dev), g.to(dev))
    ref = port.jacobi_iterate(va, ei, ea, g)
    assert same(out[0].cpu(), ref[0]) and same(out[1].cpu(), ref[1])


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
@pytest.mark.parametrize("k", [2, 8])
def test_jacobi_cheby_multi_rhs_extension(G, dev, dt, k):
    """k right-hand sides at once (config 5 extension): oracle = the reference run per column."""
    torch.manual_seed(9)
    N = 12
    n = N * N
    ei, ev = lap(N, dt)
    diag = -4 * torch.ones(n, 1, dtype=dt)
    b, x = torch.rand(n, k, dtype=dt), torch.rand(n, k, dtype=dt)
    ea = torch.cat([ev, torch.zeros_like(ev)], 1)
    g = torch.tensor(0.7).reshape(-1)
    out = G.JacobiGNN.JacobiGNN()(3, torch.cat([diag, b, x], 1).to(dev), ei.to(dev), ea.to(dev), g).cpu()
    gc = torch.tensor([-3.4, -4.0])
    cv, _, _ = G.ChebyGNN.ChebyRelaxGNN(4)(torch.cat([b, x], 1).to(dev), ei.to(dev), ev.to(dev), gc)
    rr = G.GNNResidual.GNNResidual()(torch.cat([b, x], 1).to(dev), ei.to(dev), ev.to(dev)).cpu()
    for c in range(k):
        ref = port.jacobi(3, torch.cat([diag, b[:, c:c + 1], x[:, c:c + 1]], 1), ei, ea, g)
        assert same(out[:, c:c + 1], ref)
        rv, _, _ = port.chebyshev(4, torch.cat([b[:, c:c + 1], x[:, c:c + 1]], 1), ei, ev, gc)
        for j in range(4):   # [b | x | r | p] blocks of k columns
            assert same(cv.cpu()[:, j * k + c], rv[:, j])
        assert same(rr[:, c:c + 1], port.residual(torch.cat([b[:, c:c + 1], x[:, c:c + 1]], 1), ei, ev))


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
@pytest.mark.parametrize("deg", [1, 2, 3, 4, 8])
def test_chebyshev_bit_exact(G, dev, dt, deg):
    torch.manual_seed(24601)
    N = 21
    ei, ev = lap(N, dt)
    b, x = torch.rand(N * N, 1, dtype=dt), torch.rand(N * N, 1, dtype=dt)
    g = torch.tensor([-3.46, -4.0])
    ref = port.chebyshev(deg, torch.cat([b, x], 1), ei, ev, g)
    out = G.ChebyGNN.ChebyRelaxGNN(deg)(torch.cat([b, x], 1).to(dev), ei.to(dev), ev.to(dev), g)
    for a, r in zip(out, ref):
        assert same(a.cpu(), r)


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_power_method(G, dev, dt):
    torch.manual_seed(24601)
    for N, iters in ((5, 10), (24, 100)):
        ei, ev = lap(N, dt)
        x = torch.rand(N * N, 1, dtype=dt)
        ea = torch.cat([ev, torch.zeros_like(ev)], 1)
        va = torch.cat([x, torch.zeros_like(x)], 1)
        rv, re_, rg = port.power_method(iters, va, ei, ea, torch.zeros(3, dtype=dt))
        ov, oe, og = G.PowerMethodGNN.PowerMethodGNN(iters)(va.to(dev), ei.to(dev), ea.to(dev),
                                                           torch.zeros(3, dtype=dt), torch.zeros(N * N))
        assert og.dtype == rg.dtype and ov.shape == rv.shape and oe.shape == re_.shape
        assert relerr(og.cpu(), rg) <= TOL[dt]
        assert relerr(ov.cpu(), rv) <= TOL[dt] and relerr(oe.cpu(), re_) <= TOL[dt], (relerr(ov.cpu(), rv), relerr(oe.cpu(), re_))
    # the reference's own 3x3 example: lambda -> 3 (PowerMethodGNN.py:338-383)
    A = torch.tensor([[1., 2., 0], [-2., 1., 2.], [1., 3., 1.]], dtype=dt)
    eij = torch.tensor([[i, j] for i in range(3) for j in range(3)]).T
    ea = torch.stack([A.flatten(), torch.zeros(9, dtype=dt)], 1)
    va = torch.cat([torch.rand(3, 1, dtype=dt), torch.zeros(3, 1, dtype=dt)], 1)
    _, _, g = G.PowerMethodGNN.PowerMethodGNN(10)(va.to(dev), eij.to(dev), ea.to(dev), torch.zeros(3, dtype=dt), None)
    assert abs(g[2].item() - 3.0) < 1e-3


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_amg_setup_bit_exact(G, dev, dt):
    """Classical SOC, SA SOC and direct interpolation on an ANISOTROPIC periodic FEM operator
    (config 4 family: positive N/S couplings, corner/E-W ratio 0.2538 next to theta = 0.25)."""
    N = 24
    n = N * N
    ei, ev = G.generators.constant_diffusion_fem(1.0, 0.01, N, dtype=torch.float64)
    ev = ev.to(dt)
    eo, ao = port.remove_diag_entries(ei, ev)
    diag = G.generators.diagonal_of(ei, ev, n)
    for theta in (0.25, 0.3, 0.0):
        S_ref = port.soc_classic(theta, torch.zeros(n, 1, dtype=dt), eo, ao)
        S = G.SOCClassicGNN.SOCClassicGNN(theta)(torch.zeros(n, 1, dtype=dt, device=dev), eo.to(dev), ao.to(dev))
        assert same(S.cpu(), S_ref)
        assert torch.equal(S.cpu() > 0, S_ref > 0)
    S_ref = port.soc_classic(0.25, torch.zeros(n, 1, dtype=dt), eo, ao)
    assert 0 < (S_ref > 0).sum() < S_ref.numel()          # the mask is non-trivial
    _, e_sa, _ = G.SOCSAGNN.SOCSAGNN()(diag.to(dev), eo.to(dev), ao.to(dev), batch=None)
    assert same(e_sa.cpu(), port.soc_sa(diag, eo, ao))
    split = torch.zeros(n, 1, dtype=dt)
    split[0::2] = 1
    ed = torch.hstack([ao, S_ref.reshape(-1, 1) > 0])
    w_ref = port.direct_interp(torch.hstack([diag, split]), eo, ed)
    w = G.DirectInterpGNN.DirectInterpGNN()(torch.hstack([diag, split]).to(dev), eo.to(dev), ed.to(dev), None)
    assert same(w.cpu(), w_ref)
    for pat in (torch.isnan, lambda t: t == 0, lambda t: t != 0):
        assert torch.equal(pat(w.cpu()), pat(w_ref))


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_arbitrary_edge_order_duplicates_empty_and_long_rows(G, dev, dt):
    n = 300
    ei, ev = rand_problem(n, 2500, dt, 77)
    x = torch.rand(n, 2, dtype=dt)
    ref_v, ref_e = port.matvec(x, ei, ev)
    v, e, _ = G.MatVecGNN.MatVecGNN()(x.to(dev), ei.to(dev), ev.to(dev), None, None)
    assert same(v.cpu(), ref_v) and same(e.cpu(), ref_e)
    plan = G.get_plan(ei.to(dev), n)
    assert not plan.identity and plan.max_row_nnz >= 1200
    # aggregation seam alone (scatter sum), 1-D and 2-D sources
    c = torch.rand(ei.shape[1], 4, dtype=dt)
    agg = G.MatVecGNN.edge_to_vertex_aggregation(ei.to(dev), c.to(dev), n)
    assert same(agg.cpu(), port.scatter_sum(c, ei[0], n))
    agg1 = G.MatVecGNN.edge_to_vertex_aggregation(ei.to(dev), c[:, 0].contiguous().to(dev), n)
    assert same(agg1.cpu(), port.scatter_sum(c[:, 0].contiguous(), ei[0], n))
    # Jacobi + residual + Chebyshev on the same messy operator
    diag = torch.rand(n, 1, dtype=dt) + 1
    b, x1 = torch.rand(n, 1, dtype=dt), torch.rand(n, 1, dtype=dt)
    ea = torch.cat([ev, torch.zeros_like(ev)], 1)
    g = torch.tensor(0.5).reshape(-1)
    out = G.JacobiGNN.JacobiGNN().iterate(torch.cat([diag, b, x1], 1).to(dev), ei.to(dev), ea.to(dev), g)
    ref = port.jacobi_iterate(torch.cat([diag, b, x1], 1), ei, ea, g)
    assert same(out[0].cpu(), ref[0]) and same(out[1].cpu(), ref[1])
    gc = torch.tensor([0.3, 2.0])
    outc = G.ChebyGNN.ChebyRelaxGNN(3)(torch.cat([b, x1], 1).to(dev), ei.to(dev), ev.to(dev), gc)
    refc = port.chebyshev(3, torch.cat([b, x1], 1), ei, ev, gc)
    for a, r in zip(outc, refc):
        assert same(a.cpu(), r)
    # AMG setup kernels with a permutation: outputs must come back in the CALLER's edge order
    eo, ao = rand_problem(n, 2500, dt, 78, with_diag=False)
    keep = eo[0] != eo[1]
    eo, ao = eo[:, keep], ao[keep]
    S_ref = port.soc_classic(0.25, torch.zeros(n, 1, dtype=dt), eo, ao)
    S = G.SOCClassicGNN.SOCClassicGNN(0.25)(torch.zeros(n, 1, dtype=dt), eo.to(dev), ao.to(dev))
    assert same(S.cpu(), S_ref)
    split = (torch.rand(n, 1) > 0.5).to(dt)
    ed = torch.hstack([ao, (S_ref.reshape(-1, 1) > 0).to(dt)])
    w = G.DirectInterpGNN.DirectInterpGNN()(torch.hstack([diag, split]).to(dev), eo.to(dev), ed.to(dev))
    assert same(w.cpu(), port.direct_interp(torch.hstack([diag, split]), eo, ed))
    _, e_sa, _ = G.SOCSAGNN.SOCSAGNN()(diag.to(dev), eo.to(dev), ao.to(dev))
    assert same(e_sa.cpu(), port.soc_sa(diag, eo, ao))


def test_empty_and_tiny_inputs(G, dev):
    ei = torch.zeros(2, 0, dtype=torch.long)
    ev = torch.zeros(0, 1)
    x = torch.rand(5, 1)
    v, e, _ = G.MatVecGNN.MatVecGNN()(x.to(dev), ei.to(dev), ev.to(dev), None, None)
    assert same(v.cpu(), torch.cat([x, torch.zeros(5, 1)], 1)) and e.shape == (0, 2)
    S = G.SOCClassicGNN.SOCClassicGNN(0.25)(torch.zeros(5, 1), ei.to(dev), ev.to(dev))
    assert S.shape == (0,)
    ei1 = torch.tensor([[0], [0]])
    v, _, _ = G.MatVecGNN.MatVecGNN()(torch.tensor([[3.0]]).to(dev), ei1.to(dev), torch.tensor([[2.0]]).to(dev), None, None)
    assert v.cpu().tolist() == [[3.0, 6.0]]
    with pytest.raises(G.GlabError):
        G.Plan.from_coo(torch.tensor([[0], [9]]).to(dev), 5)           # column out of range
    with pytest.raises(G.GlabError):
        G.MatVecGNN.MatVecGNN()(torch.rand(5, 3).to(dev), ei1.to(dev), torch.rand(1, 1).to(dev), None, None)


def test_golden_vectors(G, dev, golden):
    """The committed outputs of the unmodified reference, without any oracle code in between."""
    for c in golden["layers"]:
        dt, n = c["dtype"], c["N"] ** 2
        ei = c["edge_index"].to(dev)
        ev = c["edge_val64"].to(dt).to(dev)
        v, e, _ = G.MatVecGNN.MatVecGNN()(c["x2"].to(dev), ei, ev, None, None)
        assert same(v.cpu(), c["matvec2"][0]) and same(e.cpu(), c["matvec2"][1])
        assert same(G.GNNResidual.GNNResidual()(torch.cat([c["b"], c["x"]], 1).to(dev), ei, ev).cpu(), c["residual"])
        va = torch.cat([-4 * torch.ones(n, 1, dtype=dt), c["b"], c["x"]], 1).to(dev)
        ea = torch.cat([ev, torch.zeros_like(ev)], 1)
        assert same(G.JacobiGNN.JacobiGNN()(10, va, ei, ea, torch.tensor(0.7).reshape(-1)).cpu(), c["jacobi10"])
        for deg, ref in c["cheby"].items():
            out = G.ChebyGNN.ChebyRelaxGNN(deg)(torch.cat([c["b"], c["x"]], 1).to(dev), ei, ev, torch.tensor([-3.46, -4.0]))
            assert all(same(a.cpu(), r) for a, r in zip(out, ref))
        vp = torch.cat([c["x"], torch.zeros_like(c["x"])], 1).to(dev)
        _, _, g = G.PowerMethodGNN.PowerMethodGNN(10)(vp, ei, ea, torch.zeros(3, dtype=dt), None)
        assert relerr(g.cpu(), c["power10"][2]) <= TOL[dt]
        eo, ao = c["off_index"].to(dev), c["off_val"].to(dev)
        S = G.SOCClassicGNN.SOCClassicGNN(c["theta"])(torch.zeros(n, 1, dtype=dt, device=dev), eo, ao)
        assert same(S.cpu(), c["soc_classic"])
        dv = -4 * torch.ones(n, 1, dtype=dt, device=dev)
        assert same(G.SOCSAGNN.SOCSAGNN()(dv, eo, ao)[1].cpu(), c["soc_sa"])
        ed = torch.hstack([ao, (S.reshape(-1, 1) > 0)])
        w = G.DirectInterpGNN.DirectInterpGNN()(torch.hstack([dv, c["splitting"].to(dev)]), eo, ed, None)
        assert same(w.cpu(), c["direct_interp"])
        nrm = G.MatrixWeightedNorm.MatrixWeightedNorm()(c["x"].to(dev), ei, -ev)
        assert relerr(nrm.cpu(), c["mwnorm"]) <= TOL[dt]
    k = golden["known"]["mv3"]
    v, _, _ = G.MatVecGNN.MatVecGNN()(k["x"].to(dev), k["edge_index"].to(dev), k["A_ij"].to(dev), None, None)
    assert v.cpu()[:, 1].tolist() == [20.0, 301.0, 1030.0]


def test_vcycle_golden(G, dev, golden):
    """Two-grid cycle against the reference's own runVCycle outputs (VCycle.py:262-277).  The
    layer calls are bit-exact; P^T A P / P^T r go through different sparse kernels than
    torch's CPU ones, hence the fp32 tolerance on x and on the residual norms."""
    V = G.VCycle
    for c in golden["vcycle"]:
        N = c["N"]
        ei, ev = G.UtilsGNN.laplacianfun_torch(N, device=dev)
        A = torch.sparse_coo_tensor(ei, ev.flatten(), dtype=torch.float)
        x, b = c["x0"].to(dev), c["b"].to(dev)
        for ref_x, ref_norm in zip(c["xs"], c["residual_norms"]):
            x = V.runVCycle(A, b, x, 3, 3, 5, True)
            assert relerr(x.cpu(), ref_x) <= 1e-5
            rn = torch.norm(V.runResidual(A, b, x)).item()
            assert abs(rn - ref_norm) <= 1e-5 * max(ref_norm, 1e-2)
    # the cached fast path issues the same kernels as the public run* functions: bit-identical
    c = golden["vcycle"][1]
    ei, ev = G.UtilsGNN.laplacianfun_torch(c["N"], device=dev)
    A = torch.sparse_coo_tensor(ei, ev.flatten(), dtype=torch.float)
    xa = V.runVCycle(A, c["b"].to(dev), c["x0"].to(dev), 3, 3, 5, True)
    xb = V._runVCycle_layers(A, c["b"].to(dev), c["x0"].to(dev), 3, 3, 5, True)
    assert torch.equal(xa, xb)
    # host-tensor entry, like the reference script
    c = golden["vcycle"][0]
    ei, ev = G.UtilsGNN.laplacianfun_torch(c["N"])
    A = torch.sparse_coo_tensor(ei, ev.flatten(), dtype=torch.float)
    x = V.runVCycle(A, c["b"], c["x0"], 3, 3, 5, True)
    assert not x.is_cuda and relerr(x, c["xs"][0]) <= 1e-5


def test_c_abi_direct(G, dev):
    """Straight through the C ABI with raw pointers (what a non-Python host would do)."""
    lib = G.lib
    torch.manual_seed(1)
    N = 40
    n = N * N
    ei, ev = lap(N, torch.float32)
    eid, vals = ei.to(dev).contiguous(), ev.to(dev).flatten().contiguous()
    x = torch.rand(n, device=dev)
    y = torch.empty(n, device=dev)
    plan = ctypes.c_void_p()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = ctypes.c_void_p
    assert lib.glab_plan_create(n, n, eid.shape[1], P(eid[0].data_ptr()), P(eid[1].data_ptr()), st,
                                ctypes.byref(plan)) == 0
    mx, ident = ctypes.c_int32(), ctypes.c_int32()
    assert lib.glab_plan_info(plan, None, None, None, ctypes.byref(mx), ctypes.byref(ident)) == 0
    assert mx.value == 5 and ident.value == 1
    assert lib.glab_spmm_f32(plan, P(vals.data_ptr()), P(x.data_ptr()), 1, P(y.data_ptr()), 0, n, st) == 0
    ref, _ = port.matvec(x.cpu().view(-1, 1), ei, ev)
    assert same(y.cpu().view(-1, 1), ref[:, 1:])
    # row ranges (used for interior/boundary overlap): two halves == whole
    y2 = torch.full((n,), float("nan"), device=dev)
    assert lib.glab_spmm_f32(plan, P(vals.data_ptr()), P(x.data_ptr()), 1, P(y2.data_ptr()), 0, 777, st) == 0
    assert lib.glab_spmm_f32(plan, P(vals.data_ptr()), P(x.data_ptr()), 1, P(y2.data_ptr()), 777, n, st) == 0
    assert same(y2.cpu(), y.cpu())
    # argument errors
    assert lib.glab_spmm_f32(plan, P(vals.data_ptr()), P(x.data_ptr()), 3, P(y.data_ptr()), 0, n, st) == -1
    assert lib.glab_spmm_f32(plan, P(vals.data_ptr()), P(x.data_ptr()), 1, P(y.data_ptr()), 0, n + 1, st) == -1
    assert lib.glab_jacobi_f32(plan, P(vals.data_ptr()), P(x.data_ptr()), P(x.data_ptr()), P(x.data_ptr()),
                               P(x.data_ptr()), P(x.data_ptr()), 1, 0, n, st) == -1   # in-place sweep refused
    assert lib.glab_plan_destroy(plan) == 0


@pytest.mark.parametrize("dt,N", [(torch.float32, 4096), (torch.float64, 2048)])
def test_large_operator_properties(G, dev, dt, N):
    """Size-independent checks at sizes the CPU oracle cannot reach in seconds -- fp32 at the FULL
    size of BASELINE config 2 (4096^2 = 16.7 M rows, 83.9 M nnz), fp64 at 2048^2:
    A*1 equals the analytic row sums exactly, the fused Jacobi sweep equals its definition built
    from the fused residual (bit-exact), linearity, and an independent fp64 torch.sparse CSR
    product agrees within the tolerance."""
    n = N * N
    ei, ev = G.UtilsGNN.laplacianfun_torch(N, device=dev)
    ev = ev.to(dt)
    plan = G.get_plan(ei, n)
    assert plan.identity and plan.nnz == 5 * n - 4 * N
    vals = G.runtime.get_vals(plan, ev)
    ones = torch.ones(n, 1, dtype=dt, device=dev)
    y = G.runtime.spmm(plan, vals, ones)
    idx = torch.arange(n, device=dev)
    gy, gx = idx // N, idx % N
    expected = -(((gy == 0) | (gy == N - 1)).to(dt) + ((gx == 0) | (gx == N - 1)).to(dt))
    assert torch.equal(y.view(-1), expected)
    torch.manual_seed(24601)
    x = torch.rand(n, 1, dtype=dt, device=dev)
    b = torch.rand(n, 1, dtype=dt, device=dev)
    diag = torch.full((n,), -4.0, dtype=dt, device=dev)
    w = torch.tensor([0.7], dtype=dt, device=dev)
    xo = G.runtime.jacobi(plan, vals, diag, b, x, torch.empty_like(x), w)
    r = G.runtime.residual(plan, vals, x, b)
    assert torch.equal(xo, x + (w * r) / diag.view(-1, 1))
    A = torch.sparse_coo_tensor(ei, ev.flatten().double(), (n, n)).to_sparse_csr()
    yx = G.runtime.spmm(plan, vals, x)
    assert relerr(yx, A @ x.double()) <= TOL[dt]
    y2 = G.runtime.spmm(plan, vals, 2 * x)
    assert torch.equal(y2, 2 * yx)                                   # scaling by 2 is exact
    # power-method norms: deterministic (bitwise reproducible) reduction
    va = torch.cat([x, torch.zeros_like(x)], 1)
    ea = torch.cat([ev, torch.zeros_like(ev)], 1)
    g1 = G.PowerMethodGNN.PowerMethodGNN(5)(va, ei, ea, torch.zeros(3, dtype=dt), None)[2]
    g2 = G.PowerMethodGNN.PowerMethodGNN(5)(va, ei, ea, torch.zeros(3, dtype=dt), None)[2]
    assert torch.equal(g1, g2)
    bk = x.double()
    for _ in range(5):
        bk = A @ bk
        bk = bk / bk.norm()
    lam = ((bk * (A @ bk)).sum() / (bk * bk).sum()).item()
    assert abs(g1[2].item() - lam) <= TOL[dt] * abs(lam)


def test_vcycle_multi_rhs_and_large_grid(G, dev):
    """Config-5 extension: 8 right-hand sides at once == the reference cycle run per column
    (oracle), and a grid far beyond the reference's dense-P limit still contracts the residual."""
    V = G.VCycle
    N, k = 12, 8
    n = N * N
    torch.manual_seed(24601)
    ei, ev = G.UtilsGNN.laplacianfun_torch(N, device=dev)
    A = torch.sparse_coo_tensor(ei, ev.flatten(), dtype=torch.float)
    b, x = torch.rand(n, k), torch.rand(n, k)
    xo = V.runVCycle(A, b.to(dev), x.to(dev), 3, 3, 5, True).cpu()
    split = torch.zeros(n)
    split[0::2] = 1
    oi, ov = port.laplacian_2d(N)
    for c in range(k):
        ref = port.two_grid_vcycle(oi, ov, b[:, c:c + 1], x[:, c:c + 1], split)
        assert relerr(xo[:, c:c + 1], ref) <= 1e-5
    N = 384
    n = N * N
    ei, ev = G.UtilsGNN.laplacianfun_torch(N, device=dev)
    A = torch.sparse_coo_tensor(ei, ev.flatten(), dtype=torch.float)
    b = torch.rand(n, k, device=dev)
    x = torch.zeros(n, k, device=dev)
    norms = [torch.norm(V.runResidual(A, b, x)).item()]
    for _ in range(3):
        x = V.runVCycle(A, b, x, 3, 3, 5, True)
        norms.append(torch.norm(V.runResidual(A, b, x)).item())
    assert all(b_ < a_ for a_, b_ in zip(norms, norms[1:])), norms


@pytest.mark.parametrize("N,align", [(80, 256), (20, 16), (420, 256)])
@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_fused_halo_step_two_ranks_emulated_on_one_gpu(G, dev, dt, N, align, monkeypatch):
    """The multi-GPU data path (glab_halo_push, the fused glab_jacobi_halo / glab_cheby_*_halo
    kernels with in-kernel acquire + communication CTA) exercised on ONE GPU: two row blocks of
    the operator live in the same process and their kernels run one after the other on one
    stream, each storing its boundary rows into the other block's halo tail exactly as it would
    through an NVLink peer mapping.  Results must equal the unpartitioned sweeps bit for bit.
    N = 420: blocks of 88 K rows, so rank 1's plan is MIXED -- the tile that reads the halo tail
    (column offsets beyond int16) streams int32 indices, every other tile 16-bit ones."""
    from glab_b200 import dist as gd
    from glab_b200._lib import HaloStep, PushDesc
    rt = G.runtime
    if N == 80:      # keep one case on int32 halo kernels (default: 16-bit indices in the qualifying tiles)
        monkeypatch.setenv("GLAB_IDX16_HALO", "0")
    world, sweeps = 2, 4
    n = N * N
    ei, ev = G.generators.heat_fem_2d((N + 1, N + 1), (1.0, 2.0), torch.float64, dev)
    ev = ev.to(dt).contiguous()
    torch.manual_seed(24601)
    b = torch.rand(n, 1, dtype=dt, device=dev)
    x0 = torch.rand(n, 1, dtype=dt, device=dev)
    diag = G.generators.diagonal_of(ei, ev, n).reshape(-1).contiguous()
    w = torch.tensor([0.7], dtype=dt, device=dev)
    plan = G.get_plan(ei, n)
    vals = rt.get_vals(plan, ev)
    xa, xb = x0.clone(), torch.empty_like(x0)
    for _ in range(sweeps):
        rt.jacobi(plan, vals, diag, b, xa, xb, w)
        xa, xb = xb, xa
    ref = xa

    part = gd.RowPartition(n, world, align=align)    # align=16: blocks too small for an interior range
    blocks = []
    for r in range(world):
        r0, r1 = part.bounds(r)
        mine = (ei[0] >= r0) & (ei[0] < r1)
        blocks.append((ei[0][mine] - r0, ei[1][mine], ev[mine].contiguous()))
    halos = gd.HaloPlan.build_all(part, [blk[1] for blk in blocks])
    ops = []
    for r in range(world):
        rows, gcols, v = blocks[r]
        h = halos[r]
        lei = torch.stack([rows, h.local_columns(gcols)]).contiguous()
        p = G.Plan.from_coo(lei, h.n_local, h.n_local + h.n_halo)
        lo, hi = h.interior_rows(lei[0], lei[1])
        vec = [torch.zeros(h.n_local + h.n_halo, 1, dtype=dt, device=dev) for _ in range(2)]
        flags = torch.zeros(64, dtype=torch.int32, device=dev)     # 16-byte spaced words
        ops.append(dict(plan=p, vals=rt.get_vals(p, v), halo=h, vec=vec, flags=flags, lo=lo, hi=hi, keep=lei))
    if N == 420:
        assert ops[0]["plan"].index16_tiles == ops[0]["plan"].tiles
        assert 0 < ops[1]["plan"].index16_tiles < ops[1]["plan"].tiles
    for r in range(world):
        r0, r1 = part.bounds(r)
        ops[r]["vec"][0][:r1 - r0].copy_(x0[r0:r1])

    def word(t, i):
        return t.data_ptr() + 16 * i

    # per (rank, vector) flag words: arrival from the other rank = word v, pushed = word 2 + v, done = word 7
    def push_descs(r, v):
        h = ops[r]["halo"]
        descs = (PushDesc * max(len(h.peers_send), 1))()
        for i, q in enumerate(h.peers_send):
            idx = h.send_rows[q]
            first = int(idx[0].item())
            descs[i].send_idx = idx.data_ptr()
            descs[i].first_row = first if torch.equal(idx.long(), torch.arange(first, first + idx.numel(), device=dev)) else -1
            descs[i].count = idx.numel()
            descs[i].dst = ops[q]["vec"][v].data_ptr()
            descs[i].dst_offset = ops[q]["halo"].n_local + ops[q]["halo"].recv_offsets[r]
            descs[i].flag = word(ops[q]["flags"], v)
        return descs

    keep = []
    # initial publication of vector 0 with the stand-alone push kernel
    for r in range(world):
        d = push_descs(r, 0)
        keep.append(d)
        rt._call("halo_push", dt, dev, rt.ptr(ops[r]["vec"][0]), 1, len(ops[r]["halo"].peers_send), d,
                 ctypes.c_void_p(word(ops[r]["flags"], 2)), rt.stream_ptr())
    cur = 0
    for _ in range(sweeps):
        nxt = 1 - cur
        for r in range(world):
            o = ops[r]
            h = o["halo"]
            r0, r1 = part.bounds(r)
            st = HaloStep()
            st.interior_begin, st.interior_end = o["lo"], o["hi"]
            fl = (ctypes.c_void_p * 1)(word(o["flags"], cur))
            st.n_wait = len(h.peers_recv)
            st.wait_flags = fl
            st.wait_target = word(o["flags"], 2 + cur)
            d = push_descs(r, nxt)
            st.n_push = len(h.peers_send)
            st.push = d
            st.pushed_counter = word(o["flags"], 2 + nxt)
            st.push_src = o["vec"][nxt].data_ptr()
            st.done_counter = word(o["flags"], 7)
            keep += [fl, d, st]
            rt.jacobi(o["plan"], o["vals"], diag[r0:r1].contiguous(), b[r0:r1].contiguous(), o["vec"][cur],
                      o["vec"][nxt], w, halo=st)
        cur = nxt
    torch.cuda.synchronize()
    for r in range(world):
        r0, r1 = part.bounds(r)
        assert torch.equal(ops[r]["vec"][cur][:r1 - r0], ref[r0:r1]), r
    # both arrival counters saw 1 stand-alone push + one push per sweep of each vector
    f0 = ops[0]["flags"].view(-1, 4)[:, 0].tolist()
    assert f0[0] + f0[1] == 1 + sweeps and f0[7] == 0


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_scatter_seam_four_way(G, dev, dt):
    """The reference's seam as a function: torch_scatter.scatter(..., reduce=sum|max|min|mean)
    (4-way aggregation of TrainableJacobiGNN.py:65-68) against the oracle's stand-in."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "shim"))
    import torch_scatter as ts
    g = torch.Generator().manual_seed(5)
    n, z = 97, 1200
    idx = torch.randint(0, n, (z,), generator=g)
    idx[idx == 11] = 12                        # an empty segment
    src = (torch.rand(z, 2, generator=g, dtype=torch.float64) - 0.5).to(dt)
    for red in ("sum", "max", "min", "mean"):
        ref = ts.scatter(src, idx, dim=0, dim_size=n, reduce=red)
        out = G.runtime.scatter(src.to(dev), idx.to(dev), dim=0, dim_size=n, reduce=red).cpu()
        if red == "mean":
            assert relerr(out, ref) <= TOL[dt]
        else:
            assert same(out, ref), red
        assert torch.all(out[11] == 0)
    out1 = G.runtime.scatter(src[:, 0].contiguous().to(dev), idx.to(dev), dim=0, dim_size=n, reduce="max").cpu()
    assert same(out1, ts.scatter(src[:, 0].contiguous(), idx, dim=0, dim_size=n, reduce="max"))


def test_c_host_example(G, dev, tmp_path):
    """A plain-C host (no Python, no torch) linked against the C ABI: examples/jacobi_c_host.c
    reproduces the reference's Jacobi sweeps bit for bit."""
    import os
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not on PATH")
    exe = str(tmp_path / "jacobi_c_host")
    libdir = os.path.dirname(G.LIB_PATH)
    cmd = ["nvcc", "-O2", "-o", exe, os.path.join(root, "examples", "jacobi_c_host.c"), "-I" + os.path.join(root, "include"),
           "-L" + libdir, "-lglab_b200", "-Xlinker", "-rpath=" + libdir]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    run = subprocess.run([exe, "257", "5"], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0 and "max|gpu-cpu|=0" in run.stdout, run.stdout + run.stderr


def test_generic_metalayer_with_user_callbacks(G, dev):
    """A composition the package does NOT recognise (user-written callbacks, here the reference's
    residual block spelled out) runs through the generic GPU block: gathers + user callbacks +
    the CUDA segment-sum seam.  Same result as the fused GNNResidual and as the oracle."""
    agg = G.MatVecGNN.edge_to_vertex_aggregation

    class Edge(torch.nn.Module):
        def forward(self, vi, vj, e, g, batch):
            return torch.cat([e, e * vj[:, 1].view(-1, 1)], 1)

    class Vertex(torch.nn.Module):
        def forward(self, v, ei, e, g, batch):
            cbar = agg(ei, e[:, 1].contiguous(), v.shape[0]).view(-1, 1)
            return torch.cat([v, v[:, 0].view(-1, 1) - cbar], 1)

    torch.manual_seed(2)
    ei, ev = lap(13, torch.float32)
    b, x = torch.rand(169, 1), torch.rand(169, 1)
    layer = G.MetaLayer(Edge(), Vertex())
    v, e, _ = layer(torch.cat([b, x], 1).to(dev), ei.to(dev), ev.to(dev), None, None)
    ref = port.residual(torch.cat([b, x], 1), ei, ev)
    assert v.is_cuda and same(v.cpu()[:, 2:3], ref)
    assert same(G.GNNResidual.GNNResidual()(torch.cat([b, x], 1).to(dev), ei.to(dev), ev.to(dev)).cpu(), ref)


def test_config4_full_size_masks_bit_exact(G, dev):
    """BASELINE config 4 at FULL size (anisotropic periodic FEM operator, 4096^2 = 16.7 M rows,
    134 M off-diagonal edges, fp32).  The operator is a constant stencil and SOC / SA / direct
    interpolation are row-local, so EVERY row must reproduce -- bit for bit, NaN pattern included
    -- the row of the same kind (first / interior / last grid line in x and y, parity of x) that
    the CPU oracle computes on a 24 x 24 torus."""
    dt = torch.float32
    Ns, N = 24, 4096

    def setup(M, device):
        ei, ev = G.generators.constant_diffusion_fem(1.0, 0.01, M, dtype=dt, device=device)
        n = M * M
        diag = G.generators.diagonal_of(ei, ev, n)
        keep = ei[0] != ei[1]
        return n, ei[:, keep].contiguous(), ev[keep].contiguous(), diag

    # oracle on the small torus
    ns, eo_s, ao_s, diag_s = setup(Ns, "cpu")
    S_s = port.soc_classic(0.25, torch.zeros(ns, 1, dtype=dt), eo_s, ao_s)
    sa_s = port.soc_sa(diag_s, eo_s, ao_s)[:, 1]
    split_s = torch.zeros(ns, 1, dtype=dt)
    split_s[0::2] = 1
    w_s = port.direct_interp(torch.hstack([diag_s, split_s]), eo_s, torch.hstack([ao_s, (S_s.reshape(-1, 1) > 0).to(dt)]))
    # full size on the GPU through the drop-in layers
    n, eo, ao, diag = setup(N, dev)
    assert eo.shape[1] == 8 * n
    S = G.SOCClassicGNN.SOCClassicGNN(0.25)(torch.zeros(n, 1, dtype=dt, device=dev), eo, ao)
    sa = G.SOCSAGNN.SOCSAGNN()(diag, eo, ao)[1][:, 1]
    split = torch.zeros(n, 1, dtype=dt, device=dev)
    split[0::2] = 1
    w = G.DirectInterpGNN.DirectInterpGNN()(torch.hstack([diag, split]), eo, torch.hstack([ao, (S.reshape(-1, 1) > 0).to(dt)]))
    assert int((S > 0).sum().item()) == 6 * n                       # 6 of the 8 neighbours are strong
    # Row (gx, gy) of the big torus has the same neighbour order (ascending wrapped column index) and
    # the same C/F flags as the row of the same kind on the small one: first / interior / last in
    # each direction, with the parity of x (N and Ns are both even).
    idx = torch.arange(n, device=dev)
    gx, gy = idx % N, idx // N
    sx = torch.where(gx == 0, 0, torch.where(gx == N - 1, Ns - 1, 6 + gx % 2))
    sy = torch.where(gy == 0, 0, torch.where(gy == N - 1, Ns - 1, 5))
    small = sy * Ns + sx
    for name, t, t_s in (("S", S, S_s), ("sa", sa, sa_s), ("w", w, w_s)):
        got = t.view(n, 8)
        want = t_s.view(ns, 8).to(dev)[small]
        same_bits = (got == want) | (torch.isnan(got) & torch.isnan(want))
        assert bool(same_bits.all()), name
        del want, same_bits
    assert bool(torch.isnan(w_s).any()) and bool((S_s > 0).any()) and not bool((S_s > 0).all())


def _grid_stencil_apply(x, c0, ce, cn, cc):
    """9-point constant stencil with zero (eliminated Dirichlet) boundary on an [N, N] grid,
    written with shifted slices -- independent of every CSR code path."""
    y = c0 * x
    y[:, 1:] += ce * x[:, :-1]
    y[:, :-1] += ce * x[:, 1:]
    y[1:, :] += cn * x[:-1, :]
    y[:-1, :] += cn * x[1:, :]
    y[1:, 1:] += cc * x[:-1, :-1]
    y[1:, :-1] += cc * x[:-1, 1:]
    y[:-1, 1:] += cc * x[1:, :-1]
    y[:-1, :-1] += cc * x[1:, 1:]
    return y


def _grid_power_method(b0_grid, iters, c):
    b = b0_grid.double()
    for _ in range(iters):
        b = _grid_stencil_apply(b, *c)
        nrm = b.norm()
        b = b / nrm
    ab = _grid_stencil_apply(b, *c)
    return ((b * ab).sum() / (b * b).sum()).item(), nrm.item(), b


def test_config3_full_size_power_method(G, dev):
    """BASELINE config 3 at FULL size on one GPU: PowerMethodGNN(100) on the 8192 x 8192 heat-equation
    FEM operator (67.1 M rows, 603.9 M nnz, fp32) against an independent fp64 power iteration that
    applies the closed-form stencil (centre 8/3, the eight neighbours -1/3) with shifted
    grid slices.  That reference is first checked against the CPU oracle at a size the oracle holds."""
    c = (8.0 / 3.0, -1.0 / 3.0, -1.0 / 3.0, -1.0 / 3.0)
    # the grid reference == the oracle (fp64, small)
    Ns = 48
    ei, ev = G.generators.heat_fem_2d((Ns + 1, Ns + 1), (1.0, 1.0), torch.float64)
    torch.manual_seed(24601)
    b0 = torch.rand(Ns * Ns, 1, dtype=torch.float64)
    g_ref = port.power_method(30, torch.cat([b0, torch.zeros_like(b0)], 1), ei, torch.cat([ev, torch.zeros_like(ev)], 1),
                              torch.zeros(3, dtype=torch.float64))[2]
    lam_s, nrm_s, _ = _grid_power_method(b0.view(Ns, Ns), 30, c)
    assert abs(lam_s - g_ref[2].item()) <= 1e-12 * abs(lam_s) and abs(nrm_s - g_ref[0].item()) <= 1e-12 * nrm_s
    # full size
    N, iters = 8192, 100
    n = N * N
    ei, ev = G.generators.heat_fem_2d((N + 1, N + 1), (1.0, 1.0), torch.float32, dev)
    assert ei.shape[1] == (3 * N - 2) ** 2
    b0 = torch.rand(n, 1, generator=torch.Generator().manual_seed(24601)).to(dev)
    va = torch.cat([b0, torch.zeros_like(b0)], 1)
    ea = torch.cat([ev, torch.zeros_like(ev)], 1)
    v, e, g = G.PowerMethodGNN.PowerMethodGNN(iters)(va, ei, ea, torch.zeros(3, device=dev), None)
    assert v.shape == (n, 2) and e.shape == (ei.shape[1], 2) and g.shape == (3,)
    del e, ea
    lam_ref, nrm_ref, b_ref = _grid_power_method(b0.view(N, N), iters, c)
    assert abs(g[2].item() - lam_ref) <= 1e-5 * abs(lam_ref), (g[2].item(), lam_ref)
    assert abs(g[0].item() - nrm_ref) <= 1e-5 * abs(nrm_ref), (g[0].item(), nrm_ref)
    assert lam_ref < 4.0                                        # unconverged estimate below lambda_max -> 4
    # the iterate itself (normalised vector after the last step)
    assert relerr(v[:, 0].view(N, N), b_ref) <= 1e-5


def test_config5_full_size_vcycle_multi_rhs(G, dev):
    """BASELINE config 5 at FULL size on one GPU: two-grid V-cycle (VCycle.py:193-237) on the
    8192 x 8192 Laplacian (67.1 M rows) with 8 right-hand-side columns, fp32.  Size-independent
    properties: every column's residual norm decreases cycle by cycle, and column 3 of the batched
    run equals the same cycle run with that column alone, bit for bit (the k = 8 kernels and the
    k = 1 kernels accumulate in the same order)."""
    V = G.VCycle
    N, k = 8192, 8
    n = N * N
    ei, ev = G.UtilsGNN.laplacianfun_torch(N, device=dev)
    A = torch.sparse_coo_tensor(ei, ev.flatten().float(), (n, n))
    del ei, ev
    gen = torch.Generator().manual_seed(24601)
    b = torch.rand(n, k, generator=gen).to(dev)
    x = torch.zeros(n, k, device=dev)
    x1 = torch.zeros(n, 1, device=dev)
    b1 = b[:, 3:4].contiguous()
    norms = [torch.norm(V.runResidual(A, b, x), dim=0)]
    for _ in range(2):
        x = V.runVCycle(A, b, x, 3, 3, 5, True)
        x1 = V.runVCycle(A, b1, x1, 3, 3, 5, True)
        norms.append(torch.norm(V.runResidual(A, b, x), dim=0))
        assert torch.equal(x[:, 3:4], x1)
    for a_, b_ in zip(norms, norms[1:]):
        assert bool((b_ < a_).all()), norms
    tg = V._two_grid(A, None)
    assert tg.P.shape == (n, n // 2) and tg.Ac.shape == (n // 2, n // 2)


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_index16_and_index32_plans_agree(G, dev, dt, monkeypatch):
    """Banded operators get 16-bit row-relative column indices in their plan (2 B instead of 4 B of
    index traffic per nonzero in the pipeline kernels).  Every fused step must give bit-identical
    results with and without them (GLAB_IDX16=0 at plan creation), non-banded operators must stay
    on int32, and the band limit |col - row| <= 32767 is respected exactly."""
    rt = G.runtime
    torch.manual_seed(7)
    for kind, N in (("laplace", 70), ("heat", 45)):
        if kind == "laplace":
            ei, ev = G.UtilsGNN.laplacianfun_torch(N, device=dev)
        else:
            ei, ev = G.generators.heat_fem_2d((N + 1, N + 1), (1.0, 2.0), torch.float64, dev)
        ev = ev.to(dt).contiguous()
        n = int(ei.max().item()) + 1
        monkeypatch.setenv("GLAB_IDX16", "1")
        p16 = rt.Plan.from_coo(ei, n)
        monkeypatch.setenv("GLAB_IDX16", "0")
        p32 = rt.Plan.from_coo(ei, n)
        monkeypatch.delenv("GLAB_IDX16")
        assert p16.index_bytes == 2 and p32.index_bytes == 4
        vals = ev.view(-1)
        diag = G.generators.diagonal_of(ei, ev, n).reshape(-1).contiguous()
        w = torch.tensor([0.7], dtype=dt, device=dev)
        sc = torch.tensor([0.3, -0.25, 0.11], dtype=dt, device=dev)
        for k in (1, 2, 8):
            x, b = torch.rand(n, k, dtype=dt, device=dev), torch.rand(n, k, dtype=dt, device=dev)
            outs = []
            for plan in (p16, p32):
                y = rt.spmm(plan, vals, x)
                r = rt.residual(plan, vals, x, b)
                xj = rt.jacobi(plan, vals, diag, b, x, torch.empty_like(x), w)
                xo, rr, pp = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
                rt.cheby_first(plan, vals, b, x, xo, rr, pp, sc[0:1])
                p2 = torch.empty_like(x)
                rt.cheby_next(plan, vals, pp, p2, rr, xo, sc[0:1], sc[1:2], sc[2:3])
                half = (n // 2) // 256 * 256          # row ranges (16-byte aligned starts stay on the pipeline)
                y2 = torch.full_like(y, float("nan"))
                rt.spmm(plan, vals, x, y2, rows=(0, half))
                rt.spmm(plan, vals, x, y2, rows=(half, n))
                outs.append((y, r, xj, xo, rr, p2, y2))
            for a_, b_ in zip(*outs):
                assert torch.equal(a_, b_)
            assert torch.equal(outs[0][0], outs[0][6])
    # MIXED plans: the periodic operator's wrap-around rows (first / last grid line) do not fit int16
    Np = 256
    n = Np * Np
    ei, ev = G.generators.constant_diffusion_fem(1.0, 0.01, Np, dtype=dt, device=dev)
    ev = ev.contiguous()
    monkeypatch.setenv("GLAB_IDX16", "2")
    pm = rt.Plan.from_coo(ei, n)
    monkeypatch.setenv("GLAB_IDX16", "1")
    p1 = rt.Plan.from_coo(ei, n)
    monkeypatch.setenv("GLAB_IDX16", "0")
    p0 = rt.Plan.from_coo(ei, n)
    monkeypatch.delenv("GLAB_IDX16")
    assert pm.tiles == 256 and pm.index16_tiles == 254 and pm.index_bytes == 4
    assert p1.index16_tiles == 0 and p0.index16_tiles == 0
    vals = ev.view(-1)
    diag = G.generators.diagonal_of(ei, ev, n).reshape(-1).contiguous()
    for k in (1, 4):
        x, b = torch.rand(n, k, dtype=dt, device=dev), torch.rand(n, k, dtype=dt, device=dev)
        res = []
        for plan in (pm, p0):
            xj = rt.jacobi(plan, vals, diag, b, x, torch.empty_like(x), w)
            y = torch.full_like(x, float("nan"))
            rt.spmm(plan, vals, x, y, rows=(0, 1024))
            rt.spmm(plan, vals, x, y, rows=(1024, n))
            y3 = torch.full_like(x, float("nan"))
            rt.spmm(plan, vals, x, y3, rows=(0, 1000))          # unaligned split: falls back to int32 / generic kernel
            rt.spmm(plan, vals, x, y3, rows=(1000, n))
            res.append((xj, y, y3))
        for a_, b_ in zip(*res):
            assert torch.equal(a_, b_)
        assert torch.equal(res[0][1], res[0][2])
    # power-method scalars through the reducing epilogues
    ei, ev = G.UtilsGNN.laplacianfun_torch(33, device=dev)
    ev = ev.to(dt)
    x = torch.rand(33 * 33, 1, dtype=dt, device=dev)
    va, ea = torch.cat([x, torch.zeros_like(x)], 1), torch.cat([ev, torch.zeros_like(ev)], 1)
    g16 = G.PowerMethodGNN.PowerMethodGNN(7)(va, ei, ea, torch.zeros(3, dtype=dt), None)[2]
    monkeypatch.setenv("GLAB_IDX16", "0")
    rt.clear_caches()
    g32 = G.PowerMethodGNN.PowerMethodGNN(7)(va, ei.clone(), ea, torch.zeros(3, dtype=dt), None)[2]
    monkeypatch.delenv("GLAB_IDX16")
    rt.clear_caches()
    assert torch.equal(g16, g32)
    # band limit and non-banded operators
    n = 70000
    d = torch.arange(n, device=dev)
    for off, want in ((32767, 2), (32768, 4)):
        rows = torch.cat([d, d[:n - off], d[off:]])
        cols = torch.cat([d, d[:n - off] + off, d[off:] - off])
        order = torch.argsort(rows * n + cols)
        ei = torch.stack([rows[order], cols[order]])
        plan = rt.Plan.from_coo(ei, n)
        assert plan.index_bytes == want, (off, plan.index_bytes)
        v = torch.rand(ei.shape[1], dtype=dt, device=dev)
        x = torch.rand(n, 1, dtype=dt, device=dev)
        y = rt.spmm(plan, v, x)
        A = torch.sparse_coo_tensor(ei, v.double(), (n, n)).to_sparse_csr()
        assert relerr(y, A @ x.double()) <= TOL[dt]


def _ms_operator(G, dev, dt, case):
    if case == "lap5":
        N = 300
        ei, ev = G.generators.laplacian_2d(N, torch.float64, dev)
    elif case == "lap5_big":
        N = 2048
        ei, ev = G.generators.laplacian_2d(N, torch.float64, dev)
    elif case == "heat9":
        N = 200
        ei, ev = G.generators.heat_fem_2d((N + 1, N + 1), (1.0, 2.0), torch.float64, dev)
    elif case == "periodic":
        N = 128
        ei, ev = G.generators.constant_diffusion_fem(1.0, 0.01, N, torch.float64, dev)
    elif case == "tiny":
        ei, ev = G.generators.laplacian_2d(9, torch.float64, dev)
    else:   # random sparsity: no band structure -> one full completion check per sweep
        n, deg = 50000, 7
        g = torch.Generator().manual_seed(7)
        rows = torch.arange(n).repeat_interleave(deg)
        cols = torch.randint(0, n, (n * deg,), generator=g)
        cols[::deg] = torch.arange(n)                      # a diagonal entry per row
        ei = torch.stack([rows, cols]).to(dev)
        ev = (torch.rand(n * deg, 1, generator=g, dtype=torch.float64) - 0.5).to(dev)
        ev[::deg] = 8.0
    return ei.contiguous(), ev.to(dt).contiguous()


@pytest.mark.parametrize("strict", [False, True])
@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
@pytest.mark.parametrize("case", ["lap5", "heat9", "periodic", "random", "tiny", "lap5_big"])
def test_multi_sweep_jacobi_bit_exact(G, dev, dt, case, strict, monkeypatch):
    """glab_jacobi_sweeps_* (all sweeps in one persistent launch, per-tile completion counters
    instead of a grid barrier) == the same number of single-sweep launches, bit for bit: banded
    operators (dependency band of a few tiles), a periodic and a random one (full completion check
    per sweep), one tile only, and an operator with many tiles per CTA; k = 1 and k = 8; repeated
    launches on the same plan (the completion counters keep counting across launches).  Both hand-off
    protocols between CTAs: the default one (L2 read-back + relaxed counter + acquiring load) and the
    formally fenced one (GLAB_MS_STRICT=1: MEMBAR-based release / acquire fences)."""
    rt = G.runtime
    monkeypatch.setenv("GLAB_MS_STRICT", "1" if strict else "0")
    assert os.environ.get("GLAB_MS") == "2", "run the GPU suite with GLAB_MS=2 (tests/conftest.py sets it)"
    ei, ev = _ms_operator(G, dev, dt, case)
    n = int(ei[0].max().item()) + 1
    plan = G.Plan.from_coo(ei, n)
    vals = rt.get_vals(plan, ev)
    diag = G.generators.diagonal_of(ei, ev, n).reshape(-1).contiguous()
    w = torch.tensor([0.7], dtype=dt, device=dev)
    for k in ((1,) if case == "lap5_big" else (1, 8)):
        torch.manual_seed(24601)
        b = torch.rand(n, k, dtype=dt, device=dev)
        x0 = torch.rand(n, k, dtype=dt, device=dev)
        for sweeps in ((10,) if case == "lap5_big" else (1, 2, 3, 6)):
            xa, xb = x0.clone(), torch.empty_like(x0)
            for _ in range(sweeps):
                rt.jacobi(plan, vals, diag, b, xa, xb, w)
                xa, xb = xb, xa
            ya, yb = x0.clone(), torch.full_like(x0, float("nan"))
            res = rt.jacobi_sweeps(plan, vals, diag, b, ya, yb, w, sweeps)
            assert res is (yb if sweeps % 2 else ya)
            assert torch.equal(res, xa), (case, k, sweeps, (res - xa).abs().max().item())
    # the layer API takes the same path
    if case == "lap5":
        b1, x1 = torch.rand(n, 1, dtype=dt, device=dev), torch.rand(n, 1, dtype=dt, device=dev)
        va = torch.cat([diag.view(-1, 1), b1, x1], 1)
        ea = torch.cat([ev, torch.zeros_like(ev)], 1)
        out = G.JacobiGNN.JacobiGNN()(5, va, ei, ea, torch.tensor([0.7], dtype=dt))
        xa, xb = x1.clone(), torch.empty_like(x1)
        for _ in range(5):
            rt.jacobi(plan, vals, diag, b1, xa, xb, w)
            xa, xb = xb, xa
        assert torch.equal(out, xa)


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_fused_four_way_aggregation_and_batched_graphs(G, dev, dt):
    """glab_segment_agg4_*: [min | mean | sum | max] in one pass, against the oracle's torch_scatter
    stand-in -- (a) edge -> vertex aggregation on a batch of small graphs (block-diagonal union, 5
    feature columns, unsorted edges, NaN and an isolated vertex), bit for bit incl. the sequential
    sums; (b) vertex -> graph and edge -> graph aggregation with the reference's `batch` vector
    (TrainableJacobiGNN.py:53-70, LearnDiffusionCoeffs.py:312-342): segments of ~1000 entries, one
    warp each -- min / max bit-exact, sum / mean within tolerance; (c) the plan is cached per index."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "shim"))
    import torch_scatter as ts
    rt = G.runtime
    g = torch.Generator().manual_seed(11)
    # a batch of 6 small graphs (band matrices of different sizes), PyG-style block-diagonal edge list
    sizes = [700, 1300, 950, 1, 1200, 1024]
    rows, cols, batch = [], [], []
    off = 0
    for gi, m in enumerate(sizes):
        i = torch.arange(m)
        for d in (-2, -1, 0, 1, 2):
            j = i + d
            ok = (j >= 0) & (j < m)
            rows.append(i[ok] + off)
            cols.append(j[ok] + off)
        batch.append(torch.full((m,), gi))
        off += m
    row, col, batch = torch.cat(rows), torch.cat(cols), torch.cat(batch)
    perm = torch.randperm(row.numel(), generator=g)
    row, col = row[perm], col[perm]                       # arbitrary edge order
    keep = row != 5                                       # vertex 5 has no edges at all
    row, col = row[keep], col[keep]
    n, z, F = off, row.numel(), 5
    src = (torch.rand(z, F, generator=g, dtype=torch.float64) - 0.5).to(dt)
    src[17, 2] = float("nan")

    def ref4(s, idx, dim_size):
        return torch.cat([ts.scatter(s, idx, dim=0, dim_size=dim_size, reduce=r) for r in ("min", "mean", "sum", "max")], 1)

    # (a) edge -> vertex
    want = ref4(src, row, n)
    got = rt.aggregate4(src.to(dev), row.to(dev), n).cpu()
    assert got.shape == (n, 4 * F)
    assert same(got[:, :F], want[:, :F]) and same(got[:, 2 * F:], want[:, 2 * F:])        # min, sum, max: bit-exact
    fin = ~torch.isnan(want[:, F:2 * F])
    assert relerr(got[:, F:2 * F][fin], want[:, F:2 * F][fin]) <= TOL[dt]                 # mean: one division
    assert torch.all(got[5] == 0)
    # (b) vertex -> graph and edge -> graph (long segments)
    vattr = (torch.rand(n, 3, generator=g, dtype=torch.float64) - 0.5).to(dt)
    want = ref4(vattr, batch, len(sizes))
    got = rt.aggregate4(vattr.to(dev), batch.to(dev), len(sizes)).cpu()
    assert same(got[:, :3], want[:, :3]) and same(got[:, 9:], want[:, 9:])                 # min / max
    assert relerr(got[:, 3:9], want[:, 3:9]) <= 10 * TOL[dt]                               # mean, sum (other order)
    ebatch = batch[row]
    src2 = torch.nan_to_num(src)
    want = ref4(src2, ebatch, len(sizes))
    got = rt.aggregate4(src2.to(dev), ebatch.to(dev), len(sizes)).cpu()
    assert same(got[:, :F], want[:, :F]) and same(got[:, 3 * F:], want[:, 3 * F:])
    assert relerr(got[:, F:3 * F], want[:, F:3 * F]) <= 10 * TOL[dt]
    # scatter() goes through the same cached plan for every reduce and any number of columns
    row_d, src_d = row.to(dev), src2.to(dev)
    for red in ("sum", "max", "min", "mean"):
        o = rt.scatter(src_d, row_d, dim=0, dim_size=n, reduce=red).cpu()
        w = ts.scatter(src2, row, dim=0, dim_size=n, reduce=red)
        assert (relerr(o, w) <= TOL[dt]) if red == "mean" else same(o, w), red
    # (c) one plan per index tensor
    p1 = rt._index_plan(row_d, n)
    assert rt._index_plan(row_d, n) is p1


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
@pytest.mark.parametrize("n", [0, 1, 255, 1024, 1025, 5000, 70001])
def test_pack_unpack_layout_glue(G, dev, dt, n):
    """rt.pack == torch.cat(parts, 1) and rt.unpack == dense column slices, bit for bit: ragged tails,
    widths 1..8, partial spans, 16-byte-misaligned parts (views at odd offsets), wide blocks that take the
    narrow-tile kernel."""
    rt = G.runtime
    torch.manual_seed(n + 17)
    for widths in ([1, 1, 1], [1, 2], [3, 1, 4], [8, 8, 8], [1], [5, 7, 2, 1], [16, 24, 8]):
        parts = [torch.rand(n, w, dtype=dt, device=dev) for w in widths]
        cat = torch.cat(parts, 1)
        assert torch.equal(rt.pack(parts), cat)
        spans, o = [], 0
        for w in widths:
            spans.append((o, w))
            o += w
        for got, want in zip(rt.unpack(cat, spans), parts):
            assert torch.equal(got.reshape(n, want.shape[1]), want)
        # a subset of the column blocks, into preallocated destinations
        if len(spans) > 1:
            dst = torch.full((n, spans[-1][1]), -1.0, dtype=dt, device=dev)
            got = rt.unpack(cat, [spans[0], spans[-1]], outs=[None, dst])
            assert torch.equal(got[0].reshape(n, parts[0].shape[1]), parts[0]) and torch.equal(dst, parts[-1])
    if n > 0:
        # misaligned dense parts: views starting one element into a larger buffer
        big = [torch.rand(n * 2 + 3, dtype=dt, device=dev) for _ in range(3)]
        parts = [b_[1:1 + n].view(n, 1) for b_ in big]
        cat = torch.cat(parts, 1)
        assert torch.equal(rt.pack(parts), cat)
        outs = [torch.zeros(n + 1, dtype=dt, device=dev)[1:].view(n, 1) for _ in range(3)]
        rt.unpack(cat, [(0, 1), (1, 1), (2, 1)], outs=outs)
        for got, want in zip(outs, parts):
            assert torch.equal(got, want)


def test_config2_full_size_layer_pass_vs_independent_fp64_grid_stencil(G, dev):
    """BASELINE config 2 at FULL size through the drop-in layers (JacobiGNN(10) + ChebyRelaxGNN(4) on the
    4096^2 Laplacian) against an INDEPENDENT formulation: the same 10 sweeps and the degree-4 Chebyshev
    recurrence evaluated in fp64 with shifted [N, N] grid slices -- no CSR, no plan, no kernel of this
    package on the checking side (the check bench.py's parity block runs; <= 1e-5 relative)."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from bench_support import SingleGpuSmoother
    prob = SingleGpuSmoother(G, 4096, dev)
    res = prob.parity()
    assert res["ok"], res
    assert res["jacobi10_rel_err"] <= 1e-5 and res["chebyshev4_x_rel_err"] <= 1e-5, res
    assert res["rows_checked"] == 4096 * 4096
    del prob
    torch.cuda.empty_cache()
