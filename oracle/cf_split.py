"""ORACLE (test infrastructure, never on the product path): CPU restatement of the PMIS
coarse/fine splitting that glab_cf_split_pmis_* runs on the device.

PARITY UNPINNED against the reference: the reference calls pyamg.classical.split.CLJP
(VCycle.py:114, DirectInterpGNN.py:194), an un-pinned third-party dependency
(pytorch/requirements.txt:7) that is not installed here, uses random weights, and whose output no
reference test fixes.  The splitting is an INPUT of the hot path (vertex_attr[:,1] of
DirectInterpGNN); this file only pins the device implementation to a plain numpy statement of the
same integer algorithm (include/glab.h, "Coarse/fine splitting"), bit for bit.
"""
import numpy as np


def mix32(x):
    """murmur3 finaliser on uint32 (a bijection, so vertex keys never tie)."""
    x = np.asarray(x, dtype=np.uint32).copy()
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x85EBCA6B)
    x ^= x >> np.uint32(13)
    x *= np.uint32(0xC2B2AE35)
    x ^= x >> np.uint32(16)
    return x


def pmis(n, row, col, strong, seed=0):
    """row/col: off-diagonal edges (i depends on j), strong: bool per edge (S_ij > 0, VCycle.py:90).
    Returns (cflag uint8 [n] with 1 = coarse, number of rounds)."""
    row = np.asarray(row, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    keep = np.asarray(strong, dtype=bool) & (row != col)
    rs, cs = row[keep], col[keep]
    lam = np.bincount(cs, minlength=n).astype(np.uint64)
    with np.errstate(over="ignore"):
        h = mix32(np.arange(n, dtype=np.uint32) + np.uint32(seed & 0xFFFFFFFF))
    key = ((lam + np.uint64(1)) << np.uint64(32)) | h.astype(np.uint64)
    state = np.zeros(n, dtype=np.int8)          # 0 undecided, 1 coarse, 2 fine
    rounds = 0
    while (state == 0).any():
        rounds += 1
        both = (state[rs] == 0) & (state[cs] == 0)
        mx = np.zeros(n, dtype=np.uint64)
        np.maximum.at(mx, rs[both], key[cs[both]])
        np.maximum.at(mx, cs[both], key[rs[both]])
        state[(state == 0) & (key > mx)] = 1
        dep = (state[rs] == 0) & (state[cs] == 1)
        state[rs[dep]] = 2
    return (state == 1).astype(np.uint8), rounds
