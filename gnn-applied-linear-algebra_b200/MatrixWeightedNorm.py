"""Drop-in for pytorch/MatrixWeightedNorm.py: sqrt(x^T W x) as one fused SpMV + dot."""
import torch

from . import _runtime as rt
from ._io import Placement, float_dtype


class MatrixWeightedNorm(torch.nn.Module):
    """MatrixWeightedNorm.py:49-161 composes MetaLayer(EdgeUpdate, VertexUpdate, GlobalUpdate);
    forward(x [n,1], edgeij_pair, W_ij [z,1]) -> 0-d tensor sqrt(sum_i x_i (W x)_i).
    One launch of glab_xtax (fp64 accumulation, deterministic reduction)."""

    def forward(self, x, edgeij_pair, edge_attr, g=None, batch=None):
        io = Placement(x, edgeij_pair, edge_attr)
        dt = float_dtype(x, edge_attr)
        plan = rt.get_plan(edgeij_pair, x.shape[0])
        vals = rt.get_vals(plan, edge_attr, 0, dt)
        xv = rt.column(io.up(x, dt), 0)
        sums = torch.zeros(2, dtype=torch.float64, device=io.device)
        rt.xtax(plan, vals, xv, sums)
        return io.down(torch.sqrt(sums[0]).to(dt))
