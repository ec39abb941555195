"""TEST INFRASTRUCTURE ONLY -- stand-in for the un-vendored `torch_scatter` package.

The reference (pytorch/requirements.txt:2) calls exactly one entry point,
``torch_scatter.scatter(src, index, dim=0, dim_size=n, reduce=...)`` (call sites:
MatVecGNN.py:60, GNNResidual.py:60, JacobiGNN.py:66, ChebyGNN.py:87,
PowerMethodGNN.py:80, SOCClassicGNN.py:69, DirectInterpGNN.py:89,92,
MatrixWeightedNorm.py:86).  The published semantics restated here:

* ``sum``  == ``zeros(dim_size).scatter_add_(dim, broadcast(index), src)`` (this is
  literally what torch_scatter.scatter_sum does),
* ``max``/``min`` == segment extreme, **0 for empty segments**,
* ``mean`` == sum / max(count, 1).

Only used by oracle/ref_loader.py so the unmodified reference files import in a
container that has no torch_scatter wheel.
"""
import torch


def _broadcast(index, src, dim):
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(dim):
            index = index.unsqueeze(0)
    while index.dim() < src.dim():
        index = index.unsqueeze(-1)
    return index.expand(src.size())


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    assert out is None
    index = _broadcast(index, src, dim)
    size = list(src.size())
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    size[dim] = dim_size
    if reduce in ("sum", "add"):
        return torch.zeros(size, dtype=src.dtype, device=src.device).scatter_add_(dim, index, src)
    if reduce == "mean":
        tot = torch.zeros(size, dtype=src.dtype, device=src.device).scatter_add_(dim, index, src)
        cnt = torch.zeros(size, dtype=src.dtype, device=src.device).scatter_add_(
            dim, index, torch.ones_like(src))
        return tot / cnt.clamp(min=1)
    if reduce in ("max", "min"):
        res = torch.zeros(size, dtype=src.dtype, device=src.device)
        res.scatter_reduce_(dim, index, src, reduce="amax" if reduce == "max" else "amin",
                            include_self=False)
        return res  # untouched (empty) segments keep the 0 they were created with
    raise ValueError(reduce)
