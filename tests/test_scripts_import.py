"""CPU: every measurement script under scripts/ (and bench.py / bench_support.py) imports cleanly
against the current package -- they only run on a GPU box, so a stale attribute or signature would
otherwise surface there, with GPU minutes already spent."""
import glob
import importlib.util
import os

from conftest import ROOT


def test_scripts_import_without_running():
    paths = sorted(glob.glob(os.path.join(ROOT, "scripts", "*.py"))) + [os.path.join(ROOT, "bench.py"),
                                                                        os.path.join(ROOT, "bench_support.py")]
    assert len(paths) >= 8
    for path in paths:
        if '__name__ == "__main__"' not in open(path).read() and not path.endswith("bench_support.py"):
            continue                            # ad-hoc tuner without an import guard (scripts/tune_pipe.py)
        spec = importlib.util.spec_from_file_location("_script_" + os.path.basename(path)[:-3], path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)            # all of them guard execution with __name__ == "__main__"
        assert hasattr(mod, "main") or path.endswith("bench_support.py"), path
