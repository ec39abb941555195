"""Drop-in for pytorch/SOCSAGNN.py: smoothed-aggregation strength measure."""
import torch

from . import _runtime as rt
from ._io import Placement, float_dtype
from .metalayer import MetaLayer


class EdgeUpdate(torch.nn.Module):
    """return [A_ij, S_ij], S_ij = (A_ij*A_ij)/(A_ii*A_jj)  (SOCSAGNN.py:49-71).
    Used as MetaLayer(EdgeUpdate()) (:91), which this package runs as one glab_soc_sa launch."""

    def forward(self, vattr_i, vattr_j, edge_attr, g, batch):
        return torch.cat([edge_attr, (edge_attr * edge_attr) / (vattr_i * vattr_j)], 1)

    @staticmethod
    def _glab_fused_block(layer, x, edgeij_pair, edge_attr, u, batch):
        if (layer.node_model is not None or layer.global_model is not None or x.dim() != 2
                or x.shape[1] != 1 or edge_attr is None or edge_attr.dim() != 2
                or edge_attr.shape[1] != 1):
            return None
        io = Placement(x, edgeij_pair, edge_attr)
        dt = float_dtype(x, edge_attr)
        plan = rt.get_plan(edgeij_pair, x.shape[0])
        vals = rt.get_vals(plan, edge_attr, 0, dt)
        diag = rt.column(io.up(x, dt), 0)
        S = rt.soc_sa(plan, vals, diag)
        e_out = torch.stack([io.up(edge_attr, dt)[:, 0], S], 1)
        return x, io.down(e_out), u


class SOCSAGNN(MetaLayer):
    """Convenience: MetaLayer(EdgeUpdate()) of SOCSAGNN.py:91."""

    def __init__(self):
        super().__init__(EdgeUpdate())
