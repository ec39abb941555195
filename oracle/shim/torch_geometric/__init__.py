"""TEST INFRASTRUCTURE ONLY -- stand-in for the un-vendored `torch_geometric` package
(pytorch/requirements.txt:4).  See nn.py / utils.py / data.py."""
from . import nn, utils, data  # noqa: F401
