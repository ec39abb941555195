"""TEST INFRASTRUCTURE ONLY -- stand-in for torch_geometric.utils (only
remove_self_loops is actually called: UtilsGNN.py:71)."""
import torch


def remove_self_loops(edge_index, edge_attr=None):
    mask = edge_index[0] != edge_index[1]
    edge_index = edge_index[:, mask]
    if edge_attr is None:
        return edge_index, None
    return edge_index, edge_attr[mask]


def add_self_loops(*a, **k):  # imported, never called on the hot path
    raise NotImplementedError


def degree(*a, **k):  # imported, never called on the hot path
    raise NotImplementedError
