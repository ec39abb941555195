"""Drop-in for pytorch/GNNResidual.py: r = b - A x."""
import torch

from . import _runtime as rt
from ._io import Placement, float_dtype


class GNNResidual(torch.nn.Module):
    """GNNResidual.py:121-132.  forward(vertex_attr=[b,x], edgeij_pair, edge_attr=[A_ij])
    -> r [n,1].  Extension: vertex_attr = [b (k cols) | x (k cols)] gives r [n,k]."""

    def forward(self, vertex_attr, edgeij_pair, edge_attr, batch=None):
        from .dist import is_partitioned
        if is_partitioned(edgeij_pair):      # this rank's row block of a row-partitioned operator
            io = Placement(vertex_attr, edge_attr)
            dt = float_dtype(vertex_attr, edge_attr)
            k = vertex_attr.shape[1] // 2
            op = edgeij_pair.operator(edge_attr, k, dt)
            va = io.up(vertex_attr, dt)
            ent = op.entry()
            b, _ = rt.unpack(va, [(0, k), (k, k)], outs=[None, op.local(ent)])
            op.publish(ent)
            return io.down(op.spmv(ent, torch.empty_like(b), b=b))
        io = Placement(vertex_attr, edgeij_pair, edge_attr)
        dt = float_dtype(vertex_attr, edge_attr)
        n, F = vertex_attr.shape
        k = F // 2
        plan = rt.get_plan(edgeij_pair, n)
        vals = rt.get_vals(plan, edge_attr, 0, dt)
        va = io.up(vertex_attr, dt)
        b, x = rt.unpack(va, [(0, k), (k, k)])
        return io.down(rt.residual(plan, vals, x, b))
