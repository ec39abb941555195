"""TEST INFRASTRUCTURE ONLY -- stand-in for the un-vendored `pyamg` package
(pytorch/requirements.txt:7).  C/F splitting parity is UNPINNED: see classical/split.py."""
from . import classical  # noqa: F401
