"""TEST INFRASTRUCTURE ONLY -- deterministic stand-in for pyamg.classical.split.CLJP.

pyamg is not installed and no reference test pins a CLJP output, so C/F-splitting
parity is UNPINNED.  The splitting is an *input* of the hot path (vertex_attr[:,1]
of DirectInterpGNN), so we use the reference's own deterministic alternative,
``C(1:2:end) = 1`` (matlab/test_direct_interpolation.m:64-65, test_vcycle.m:66-67).
"""
import numpy as np


def CLJP(S, color=False):
    n = S.shape[0]
    split = np.zeros(n, dtype="intc")
    split[0::2] = 1
    return split
