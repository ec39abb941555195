"""Drop-in for the reference's pytorch/UtilsGNN.py (layout definition of the hot path)."""
import torch

from . import generators


def laplacianfun_torch(N, device="cpu"):
    """UtilsGNN.py:53-67 -- 2-D 5-point (negative) Laplacian as (edge_index int64 [2,z] sorted
    row-major incl. the diagonal, edge_val float64 [z,1]).  `device` is an extension: the
    reference always builds on the CPU (through scipy); here any device works."""
    return generators.laplacian_2d(N, torch.float64, device)


def remove_diag_entries(edge_index, edge_val):
    """UtilsGNN.py:69-72 -- drop self loops, keep the order of the remaining edges."""
    keep = edge_index[0] != edge_index[1]
    return edge_index[:, keep], edge_val[keep]


def coo_to_gnn_input(A):
    """UtilsGNN.py:74-78 -- torch sparse COO -> (edgeij_pair, edge_attr [z,1])."""
    A = A.coalesce()
    return A.indices(), A.values().reshape(-1, 1)
