"""Generates tests/golden/reference_setup.pt: the AMG-setup intermediates of the UNMODIFIED
reference two-grid cycle (/root/reference/pytorch/VCycle.py, loaded through oracle/ref_loader.py):
the strength flags of runSOC (:72-92), the prolongator of runDirectInterp (:94-137, dense
construction, CLJP stand-in split[0::2] = 1) and the Galerkin operator P^T A P (:209).

Run here (the build container) only:  python tests/golden/make_golden_setup.py
"""
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import ref_loader  # noqa: E402


def setup_case(R, N):
    V = R.VCycle
    V.N = N                                  # the script reads a module-level N (VCycle.py:117,165)
    ei, ev = R.UtilsGNN.laplacianfun_torch(N)
    A = torch.sparse_coo_tensor(ei, ev.flatten(), dtype=torch.float)
    S = V.runSOC(A)
    P = V.runDirectInterp(A, S, N).coalesce()
    Ac = (P.t() @ (A @ P)).coalesce()        # VCycle.py:209
    return {"N": N, "S": S, "P_indices": P.indices(), "P_values": P.values(), "P_shape": tuple(P.shape),
            "Ac_indices": Ac.indices(), "Ac_values": Ac.values(), "Ac_shape": tuple(Ac.shape)}


def main():
    R = ref_loader.load()
    fx = {"cases": [setup_case(R, N) for N in (5, 8, 13)], "torch_version": torch.__version__}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_setup.pt")
    torch.save(fx, path)
    print("wrote", path, os.path.getsize(path), "bytes")
    for c in fx["cases"]:
        print(c["N"], "P", c["P_shape"], c["P_values"].numel(), "Ac", c["Ac_shape"], c["Ac_values"].numel())


if __name__ == "__main__":
    main()
