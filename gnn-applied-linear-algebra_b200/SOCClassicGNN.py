"""Drop-in for pytorch/SOCClassicGNN.py: classical strength of connection."""
import torch

from . import _runtime as rt
from ._io import Placement, float_dtype


class SOCClassicGNN(torch.nn.Module):
    """SOCClassicGNN.py:131-147.  forward(vertex_attr placeholder [n,1], edgeij_pair,
    edge_attr=[A_ij] off-diagonal) -> S_ij [z] = relu(-A_ij / max_k(-A_ik) - theta), in the
    caller's edge order.  Both reference GN blocks (row max, then the edge update) are one
    launch of glab_soc_classic with bit-exact element-wise math."""

    def __init__(self, theta):
        super().__init__()
        self.theta = theta

    def forward(self, vertex_attr, edgeij_pair, edge_attr, batch=None):
        io = Placement(edge_attr, edgeij_pair, vertex_attr)
        dt = float_dtype(edge_attr)
        plan = rt.get_plan(edgeij_pair, vertex_attr.shape[0])
        vals = rt.get_vals(plan, edge_attr, 0, dt)
        return io.down(rt.soc_classic(plan, vals, self.theta))
