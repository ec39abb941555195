"""Generates tests/golden/reference_layers.pt by running the UNMODIFIED reference layer files
(/root/reference/pytorch, loaded through oracle/ref_loader.py) on seeded inputs.

Run here (the build container) only:  python tests/golden/make_golden.py
The fixture travels to the GPU box; /root/reference does not.
"""
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import ref_loader  # noqa: E402


def case(R, N, dt, seed):
    torch.manual_seed(seed)
    n = N * N
    ei, ev64 = R.UtilsGNN.laplacianfun_torch(N)
    ev = ev64.to(dt)
    x = torch.rand(n, 1, dtype=dt)
    b = torch.rand(n, 1, dtype=dt)
    x3 = torch.rand(n, 2, dtype=dt)
    batch = torch.zeros(n)
    out = {"N": N, "dtype": dt, "edge_index": ei, "edge_val64": ev64, "x": x, "b": b, "x2": x3}

    mv = R.MetaLayer(R.MatVecGNN.EdgeUpdate(), R.MatVecGNN.VertexUpdate(R.MatVecGNN.edge_to_vertex_aggregation))
    out["matvec"] = mv(x, ei, ev, None, batch=batch)[:2]
    out["matvec2"] = mv(x3, ei, ev, None, batch=batch)[:2]
    out["residual"] = R.GNNResidual.GNNResidual()(torch.cat([b, x], 1), ei, ev)

    va = torch.cat([-4 * torch.ones(n, 1, dtype=dt), b, x], 1)
    ea = torch.cat([ev, torch.zeros_like(ev)], 1)
    gw = torch.tensor(0.7).reshape(-1)
    J = R.JacobiGNN.JacobiGNN()
    out["jacobi10"] = J(10, va, ei, ea, gw)
    out["jacobi_iterate"] = J.iterate(va, ei, ea, gw)

    gc = torch.tensor([-3.46, -4.0])
    out["cheby"] = {deg: R.ChebyGNN.ChebyRelaxGNN(deg)(torch.cat([b, x], 1), ei, ev, gc) for deg in (1, 2, 3, 4, 8)}

    vp = torch.cat([x, torch.zeros_like(x)], 1)
    out["power10"] = R.PowerMethodGNN.PowerMethodGNN(10)(vp, ei, ea, torch.zeros(3, dtype=dt), batch)

    eo, ao = R.UtilsGNN.remove_diag_entries(ei, ev)
    out["off_index"], out["off_val"] = eo, ao
    out["theta"] = 0.25
    S = R.SOCClassicGNN.SOCClassicGNN(0.25)(torch.zeros(n, 1, dtype=dt), eo, ao)
    out["soc_classic"] = S
    dv = -4 * torch.ones(n, 1, dtype=dt)
    out["soc_sa"] = R.MetaLayer(R.SOCSAGNN.EdgeUpdate())(dv, eo, ao, batch=batch)[1]
    split = torch.zeros(n, 1, dtype=dt)
    split[0::2] = 1
    out["splitting"] = split
    ed = torch.hstack([ao, (S.reshape(-1, 1) > 0)])
    out["direct_interp"] = R.DirectInterpGNN.DirectInterpGNN()(torch.hstack([dv, split]), eo, ed, None)
    mw = R.MatrixWeightedNorm
    gnn = R.MetaLayer(mw.EdgeUpdate(), mw.VertexUpdate(mw.EdgeToVertexAggregation),
                      mw.GlobalUpdate(mw.VertexToGlobalAggregation))
    # W = -A is SPD for the negative Laplacian
    out["mwnorm"] = gnn(x, ei, -ev, None, batch)[2]
    return out


def vcycle_case(R, N, seed, cycles=4):
    V = R.VCycle
    torch.manual_seed(seed)
    n = N * N
    ei, ev = R.UtilsGNN.laplacianfun_torch(N)
    V.N = N
    A = torch.sparse_coo_tensor(ei, ev.flatten(), dtype=torch.float)
    x = torch.rand(n, 1)
    b = torch.rand(n, 1)
    xs, res = [], []
    xr = x.clone()
    for _ in range(cycles):
        xr = V.runVCycle(A, b, xr, 3, 3, 5, True)   # CLJP stand-in: split[0::2] = 1
        xs.append(xr.clone())
        res.append(torch.norm(V.runResidual(A, b, xr)).item())
    return {"N": N, "x0": x, "b": b, "xs": xs, "residual_norms": res}


def known_answers(R):
    """The reference's own __main__ self-checks."""
    ka = {}
    ei = torch.tensor([[0, 1], [1, 0], [1, 2], [2, 1], [0, 0], [1, 1], [2, 2]], dtype=torch.long).T
    A_ij = torch.tensor([[1.], [1.], [2.], [3.], [10.], [10.], [10.]])
    x = torch.tensor([[1.], [10.], [100.]])
    x2 = torch.tensor([[1., 7.], [10., 70.], [100., 700.]])
    mv = R.MetaLayer(R.MatVecGNN.EdgeUpdate(), R.MatVecGNN.VertexUpdate(R.MatVecGNN.edge_to_vertex_aggregation))
    ka["mv3"] = {"edge_index": ei, "A_ij": A_ij, "x": x, "x2": x2,
                 "y": mv(x, ei, A_ij, None, batch=torch.zeros(3))[0][:, 1:],
                 "y2": mv(x2, ei, A_ij, None, batch=torch.zeros(3))[0][:, -2:]}
    A = torch.tensor([[1., 2., 0], [-2., 1., 2.], [1., 3., 1.]])
    torch.manual_seed(7)
    b = torch.rand(3, 1)
    eij = torch.tensor([[i, j] for i in range(3) for j in range(3)]).T
    ea = torch.tensor([[A[i, j].item(), 0.] for i in range(3) for j in range(3)])
    va = torch.cat([b, torch.zeros(3, 1)], 1)
    ka["power3"] = {"edge_index": eij, "edge_attr": ea, "vertex_attr": va,
                    "out": R.PowerMethodGNN.PowerMethodGNN(10)(va, eij, ea, torch.zeros(3), torch.zeros(3))}
    return ka


def main():
    R = ref_loader.load()
    fx = {"layers": [case(R, 5, torch.float32, 24601), case(R, 5, torch.float64, 24601),
                     case(R, 8, torch.float32, 11), case(R, 8, torch.float64, 11)],
          "vcycle": [vcycle_case(R, 5, 0), vcycle_case(R, 8, 3)],
          "known": known_answers(R),
          "torch_version": torch.__version__}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_layers.pt")
    torch.save(fx, path)
    print("wrote", path, os.path.getsize(path), "bytes")
    print("vcycle residuals:", fx["vcycle"][0]["residual_norms"])
    print("mv3 y:", fx["known"]["mv3"]["y"].flatten().tolist(), "power3 lambda:", fx["known"]["power3"]["out"][2])


if __name__ == "__main__":
    main()
