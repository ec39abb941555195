"""Drop-in for pytorch/DirectInterpGNN.py: direct-interpolation weights."""
import torch

from . import _runtime as rt
from ._io import Placement, float_dtype


class DirectInterpGNN(torch.nn.Module):
    """DirectInterpGNN.py:155-174.  vertex_attr=[A_ii, C_i] (C_i = 1 for coarse points),
    off-diagonal edge_attr=[A_ij, S_ij in {0,1}]; returns w_ij [z] for EVERY edge, in the
    caller's edge order.  Both reference GN blocks (two row sums, then the per-edge scaling) are
    one launch of glab_direct_interp.  IEEE behaviour is kept: a C row without strong C
    neighbours gives 0*inf = NaN exactly like the reference (its MATLAB twin zeroes them,
    test_direct_interpolation.m:130-132; the Python does not)."""

    def forward(self, vertex_attr, edgeij_pair, edge_attr, g=None, batch=None):
        io = Placement(vertex_attr, edgeij_pair, edge_attr)
        dt = float_dtype(vertex_attr, edge_attr)
        plan = rt.get_plan(edgeij_pair, vertex_attr.shape[0])
        vals = rt.get_vals(plan, edge_attr, 0, dt)
        S = rt.get_vals(plan, edge_attr, 1, dt)
        va = io.up(vertex_attr, dt)
        diag, cflag = rt.unpack(va, [(0, 1), (1, 1)])
        w = rt.direct_interp(plan, vals, S, diag.view(-1), cflag.view(-1))
        return io.down(w)
