"""MetaLayer with the graph-network block ordering the reference relies on
(torch_geometric.nn.MetaLayer; spec twin matlab/gnn.m:47-54): edge update -> vertex update ->
global update, a stage skipped when its model is None.

Compositions built from this package's own callbacks are recognised and executed as ONE fused
CUDA kernel (no [nnz, F] gathers are ever materialised).  Any other composition runs the
generic block on the GPU: the two gathers x[row], x[col] feed the user's callbacks, whose
edge->vertex aggregation should be this package's `edge_to_vertex_aggregation` (a CUDA
segment-sum).  There is no CPU path.
"""
import torch

from . import _runtime as rt


class MetaLayer(torch.nn.Module):
    def __init__(self, edge_model=None, node_model=None, global_model=None):
        super().__init__()
        self.edge_model = edge_model
        self.node_model = node_model
        self.global_model = global_model

    def forward(self, x, edge_index, edge_attr=None, u=None, batch=None):
        fused = getattr(self.edge_model, "_glab_fused_block", None)
        if fused is not None:
            out = fused(self, x, edge_index, edge_attr, u, batch)
            if out is not None:
                return out
        device = rt.compute_device(x, edge_index, edge_attr)
        host = not x.is_cuda
        x = rt.to_device(x, device)
        edge_index = rt.to_device(edge_index, device)
        edge_attr = rt.to_device(edge_attr, device)
        row, col = edge_index[0], edge_index[1]
        if self.edge_model is not None:
            edge_attr = self.edge_model(x[row], x[col], edge_attr, u, None)
        if self.node_model is not None:
            x = self.node_model(x, edge_index, edge_attr, u, batch)
        if self.global_model is not None:
            u = self.global_model(x, edge_index, edge_attr, u, batch)
        if host:
            x = x.cpu()
            edge_attr = None if edge_attr is None else edge_attr.cpu()
        return x, edge_attr, u
