"""glab_b200 -- B200-native (sm_100a) drop-in for the edge-wise message-passing hot path of
sandialabs/gnn-applied-linear-algebra's PyTorch layers.

The sub-modules keep the reference's file and class names, forward() signatures and tensor
layouts (vertex_attr [n,Fv], edgeij_pair int64 [2,nnz], edge_attr [nnz,Fe], g):

    MatVecGNN, GNNResidual, JacobiGNN, ChebyGNN, PowerMethodGNN, SOCClassicGNN, SOCSAGNN,
    DirectInterpGNN, MatrixWeightedNorm, VCycle, UtilsGNN        (+ MetaLayer, generators, dist)

Each layer step is one launch of a hand-written CUDA kernel in libglab_b200.so, reached
through the C ABI of include/glab.h.  There is no CPU / PyTorch fallback: importing this
package without the built library raises ImportError, and calling a layer without a CUDA
device raises GlabError.
"""
from ._lib import GlabError, LIB_PATH, lib  # noqa: F401  (raises if the .so is missing)
from . import _runtime as runtime  # noqa: F401
from ._runtime import Plan, clear_caches, get_plan  # noqa: F401
from . import generators  # noqa: F401
from .metalayer import MetaLayer  # noqa: F401
from . import (ChebyGNN, DirectInterpGNN, GNNResidual, JacobiGNN, MatVecGNN,  # noqa: F401
               MatrixWeightedNorm, PowerMethodGNN, SOCClassicGNN, SOCSAGNN, UtilsGNN, VCycle)

__version__ = "0.1.0"
