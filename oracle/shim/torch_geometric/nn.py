"""TEST INFRASTRUCTURE ONLY -- stand-in for torch_geometric.nn.

MetaLayer restates the published graph-network block ordering that the reference
relies on (and documents in its MATLAB twin, matlab/gnn.m:47-54,118-166):
edge update -> vertex update (which calls the edge->vertex aggregation) -> global
update, each stage skipped when its model is None; returns (x, edge_attr, u).
"""
import torch


class MetaLayer(torch.nn.Module):
    def __init__(self, edge_model=None, node_model=None, global_model=None):
        super().__init__()
        self.edge_model = edge_model
        self.node_model = node_model
        self.global_model = global_model

    def forward(self, x, edge_index, edge_attr=None, u=None, batch=None):
        row = edge_index[0]
        col = edge_index[1]
        if self.edge_model is not None:
            edge_attr = self.edge_model(x[row], x[col], edge_attr, u,
                                        batch if batch is None else batch[row])
        if self.node_model is not None:
            x = self.node_model(x, edge_index, edge_attr, u, batch)
        if self.global_model is not None:
            u = self.global_model(x, edge_index, edge_attr, u, batch)
        return x, edge_attr, u


class MessagePassing(torch.nn.Module):  # imported (unused) by UtilsGNN.py:41
    pass
