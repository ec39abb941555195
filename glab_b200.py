"""Import alias: `import glab_b200` loads the package that lives in the directory
``gnn-applied-linear-algebra_b200/`` (whose name, fixed by the repo layout, is not a valid
Python identifier).  Sub-modules are importable as ``glab_b200.JacobiGNN`` etc."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gnn-applied-linear-algebra_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
