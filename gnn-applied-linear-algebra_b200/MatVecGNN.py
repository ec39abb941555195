"""Drop-in for pytorch/MatVecGNN.py: y = A x (multi-column x supported) as one GN block.

Reference usage (MatVecGNN.py:146-150):
    gnn_matvec = MetaLayer(EdgeUpdate(), VertexUpdate(edge_to_vertex_aggregation))
    vertex_attr, edge_attr, _ = gnn_matvec(x, edgeij_pair, A_ij, None, batch=batch)
With this package's MetaLayer that composition is ONE launch of glab_spmm (plus one
glab_edge_messages launch for the returned c_ij column).
"""
import torch

from . import _runtime as rt
from ._io import Placement, float_dtype
from .metalayer import MetaLayer


def edge_to_vertex_aggregation(edgeij_pair, c_ij, n_vertices):
    """cbar_i = sum_j c_ij  (MatVecGNN.py:43-62): torch_scatter.scatter(reduce="sum") as a CUDA
    segment sum over the cached CSR plan, sequential in edge order."""
    io = Placement(c_ij, edgeij_pair)
    plan = rt.get_plan(edgeij_pair, n_vertices)
    src = io.up(c_ij)
    src2 = src.view(-1, 1) if src.dim() == 1 else src
    if plan.identity and src2.is_contiguous() and src2.data_ptr() % 16 == 0:
        slots = src2
    else:
        slots = torch.stack([rt.get_vals(plan, src2, j) for j in range(src2.shape[1])], 1).contiguous()
    out = rt.segment_sum(plan, slots)
    return io.down(out.view(-1) if src.dim() == 1 else out)


class EdgeUpdate(torch.nn.Module):
    """return [A_ij, c_ij] where c_ij = A_ij x_j  (MatVecGNN.py:64-84)."""

    def forward(self, vattr_i, vattr_j, edge_attr, g, batch):
        return torch.cat([edge_attr, edge_attr * vattr_j], 1)

    @staticmethod
    def _glab_fused_block(layer, x, edgeij_pair, edge_attr, u, batch):
        nm = layer.node_model
        if (not isinstance(nm, VertexUpdate) or layer.global_model is not None
                or nm.edge_aggregation_function is not edge_to_vertex_aggregation
                or x.dim() != 2 or x.shape[1] not in rt.SUPPORTED_K
                or edge_attr is None or edge_attr.dim() != 2 or edge_attr.shape[1] != 1):
            return None
        io = Placement(x, edgeij_pair, edge_attr)
        dt = float_dtype(x, edge_attr)
        plan = rt.get_plan(edgeij_pair, x.shape[0])
        vals = rt.get_vals(plan, edge_attr, 0, dt)
        xd = rt.dense(io.up(x, dt))
        y = rt.spmm(plan, vals, xd)
        e_out = rt.with_messages(plan, vals, xd)
        return io.down(rt.pack([xd, y])), io.down(e_out), u


class VertexUpdate(torch.nn.Module):
    """return [x_i, y_i], y_i = sum_j c_ij  (MatVecGNN.py:86-114)."""

    def __init__(self, edge_aggregation_function):
        super().__init__()
        self.edge_aggregation_function = edge_aggregation_function

    def forward(self, vertex_attr, edgeij_pair, edge_attr, g, batch):
        c_ij = edge_attr[:, 1:1 + vertex_attr.shape[1]]
        y = self.edge_aggregation_function(edgeij_pair, c_ij, vertex_attr.shape[0])
        return torch.cat([vertex_attr, y], 1)


class MatVecGNN(MetaLayer):
    """Convenience: the composition of MatVecGNN.py:146 as a ready-made module."""

    def __init__(self):
        super().__init__(EdgeUpdate(), VertexUpdate(edge_to_vertex_aggregation))
