"""CPU: the C-ABI shared library loads and exports every symbol include/glab.h declares; the
host package binds them all; argument errors come back as codes (no compute without a GPU)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "glab.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(glab_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(G):
    names = _declared()
    assert len(names) >= 45
    lib = ctypes.CDLL(G.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libglab_b200.so does not export %s" % n


def test_python_binds_every_symbol(G):
    bound = set(G._lib.exported_symbols())
    assert set(_declared()) <= bound, sorted(set(_declared()) - bound)


def test_version_and_error_strings(G):
    assert G.lib.glab_version() == 200
    assert b"invalid argument" in G.lib.glab_error_string(-1)
    assert G.lib.glab_error_string(0) == b"ok"
    # argument validation happens before any CUDA call
    assert G.lib.glab_plan_info(None, None, None, None, None, None) == -1
    assert G.lib.glab_spmm_f32(None, None, None, 1, None, 0, 0, None) == -1
    assert G.lib.glab_halo_wait(1, None, None, None) == -1


def test_sass_is_sm100a(G):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", G.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback(G):
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(G.GlabError):
        G.JacobiGNN.JacobiGNN()(1, torch.zeros(4, 3), torch.zeros(2, 4, dtype=torch.long), torch.zeros(4, 2),
                                torch.tensor([0.7]))


def test_plan_cache_validity_rules(G):
    """The plan / value caches key on (storage pointer, shape, strides, dtype, device, version) and
    are valid only while the tensor that created the entry is alive and unmodified."""
    import torch
    C = G.runtime._Cache(capacity=2)
    t = torch.arange(10)
    assert C.get(t) is None
    C.put(t, "plan-A", extra=(5,))
    assert C.get(t, extra=(5,)) == "plan-A" and C.get(t, extra=(6,)) is None
    view = t[:]                      # another tensor object on the same storage: valid while `t` lives
    assert C.get(view, extra=(5,)) == "plan-A"
    t.add_(1)                        # in-place modification bumps _version -> miss
    assert C.get(t, extra=(5,)) is None
    u = torch.arange(10)
    C.put(u, "plan-B")
    del u                            # creator gone: its address may be recycled -> entry must not be served
    import gc
    gc.collect()
    w = torch.arange(10)
    assert C.get(w) is None
    for i in range(4):               # LRU capacity
        C.put(torch.zeros(3 + i), i)
    assert len(C.d) <= 2


def test_ctypes_signatures_match_header_arity(G):
    """Every prototype of include/glab.h is bound in _lib.py with the same number of arguments and a
    compatible return type (int / int64_t / const char*): catches ABI drift between the header, the
    library and the host package without running anything on a GPU."""
    import ctypes as C
    src = open(os.path.join(ROOT, "include", "glab.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = re.findall(r"\b(int64_t|int|const char\*)\s+(glab_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S)
    assert len(protos) >= 90
    ret = {"int": C.c_int, "int64_t": C.c_int64, "const char*": C.c_char_p}
    for rtype, name, params in protos:
        params = params.strip()
        arity = 0 if params in ("", "void") else params.count(",") + 1
        fn = getattr(G.lib, name)
        assert fn.argtypes is not None, name
        assert len(fn.argtypes) == arity, (name, len(fn.argtypes), arity)
        assert fn.restype is ret[rtype], (name, fn.restype, rtype)


def test_header_is_plain_c(tmp_path):
    """include/glab.h must compile as C99 on its own (the boundary is a C ABI, not C++)."""
    import subprocess
    src = tmp_path / "use_glab.c"
    src.write_text('#include "glab.h"\nint main(void) { return glab_version() == GLAB_VERSION ? 0 : 1; }\n')
    out = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I",
                          os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
