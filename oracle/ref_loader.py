"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference layer files as an oracle.

Only `tests/`, `tests/golden/make_golden.py`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module; the
product package (gnn-applied-linear-algebra_b200/) never does.

`/root/reference` exists only in the build container (not on the GPU box), so
`available()` gates every use; what travels to the GPU box is `oracle/port.py`
(the restatement, pinned to this loader's output by tests/test_oracle_pinning.py and by
the committed fixtures under tests/golden/).

What it does (SURVEY.md section 8c "harness fixes needed (not edits)"):
  * puts oracle/shim (stand-ins for torch_scatter / torch_geometric / pyamg) and
    <reference>/pytorch on sys.path,
  * `ChebyGNN.GNNResidual = GNNResidual.GNNResidual` before importing JacobiGNN
    (JacobiGNN.py:49 imports the class from the wrong module),
  * registers `<m>_Meta` aliases for SOCClassicGNN, DirectInterpGNN, JacobiGNN, ChebyGNN
    (VCycle.py:48-51, DirectInterpGNN.py:180 import module names that do not exist),
  * imports VCycle.py with stdout swallowed (it runs its N=5 demo at import, :239-277),
  * restores sys.path / sys.modules afterwards so the reference's top-level module
    names never shadow anything else.
"""
import contextlib
import importlib
import io
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "shim")
_NAMES = ["UtilsGNN", "MatVecGNN", "GNNResidual", "ChebyGNN", "JacobiGNN", "PowerMethodGNN",
          "SOCClassicGNN", "SOCSAGNN", "DirectInterpGNN", "MatrixWeightedNorm", "VCycle"]
_SHIM_PKGS = ["torch_scatter", "torch_geometric", "pyamg"]
_cache = None


def reference_root():
    return os.environ.get("GLAB_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(reference_root(), "pytorch", "MatVecGNN.py"))


def load():
    """Return a namespace whose attributes are the reference modules (UtilsGNN, MatVecGNN,
    GNNResidual, ChebyGNN, JacobiGNN, PowerMethodGNN, SOCClassicGNN, SOCSAGNN,
    DirectInterpGNN, MatrixWeightedNorm, VCycle) plus `MetaLayer` and `scatter`."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError("reference sources not present at %s" % reference_root())
    import torch
    ref_py = os.path.join(reference_root(), "pytorch")
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in
                  _NAMES + [n + "_Meta" for n in _NAMES] + _SHIM_PKGS}
    rng_state = torch.get_rng_state()
    mods = {}
    try:
        sys.path.insert(0, ref_py)
        sys.path.insert(0, _SHIM)
        for k in list(sys.modules):
            if k.split(".")[0] in _SHIM_PKGS or k in _NAMES:
                del sys.modules[k]
        sys.dont_write_bytecode, old_dwb = True, sys.dont_write_bytecode
        try:
            for name in ["UtilsGNN", "MatVecGNN", "GNNResidual", "ChebyGNN"]:
                mods[name] = importlib.import_module(name)
            mods["ChebyGNN"].GNNResidual = mods["GNNResidual"].GNNResidual
            for name in ["JacobiGNN", "PowerMethodGNN", "SOCClassicGNN", "SOCSAGNN",
                         "DirectInterpGNN", "MatrixWeightedNorm"]:
                mods[name] = importlib.import_module(name)
            for name in ["SOCClassicGNN", "DirectInterpGNN", "JacobiGNN", "ChebyGNN"]:
                sys.modules[name + "_Meta"] = mods[name]
            with contextlib.redirect_stdout(io.StringIO()):
                mods["VCycle"] = importlib.import_module("VCycle")
            import torch_scatter
            from torch_geometric.nn import MetaLayer
        finally:
            sys.dont_write_bytecode = old_dwb
    finally:
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k.split(".")[0] in _SHIM_PKGS or k in _NAMES or k.endswith("_Meta"):
                if k in _NAMES or k.split(".")[0] in _SHIM_PKGS or k[:-5] in _NAMES:
                    del sys.modules[k]
        for k, v in saved_mods.items():
            if v is not None:
                sys.modules[k] = v
        torch.set_rng_state(rng_state)
    ns = types.SimpleNamespace(**mods)
    ns.MetaLayer = MetaLayer
    ns.scatter = torch_scatter.scatter
    _cache = ns
    return ns
