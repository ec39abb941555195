"""ctypes binding of libglab_b200.so (the C ABI declared in include/glab.h).

There is NO fallback: if the shared library is missing or a symbol is absent, importing the
package raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` (or
``make -C gnn-applied-linear-algebra_b200/csrc``).
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# GLAB_LIB_PATH: load another build of the same library (kernel experiments); never a different implementation
LIB_PATH = os.environ.get("GLAB_LIB_PATH") or os.path.join(_HERE, "libglab_b200.so")


class GlabError(RuntimeError):
    pass


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "libglab_b200.so not found at %s -- the CUDA extension is mandatory (no CPU or "
            "PyTorch fallback exists).  Run __graft_entry__.build() or "
            "`make -C gnn-applied-linear-algebra_b200/csrc`." % LIB_PATH)
    return ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)


lib = _load()

P = c_void_p  # every device pointer / opaque handle crosses the ABI as void*
_I64, _I32, _INT = c_int64, c_int32, c_int


def _sig(name, restype, *argtypes):
    fn = getattr(lib, name)  # AttributeError here == symbol missing == hard failure
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


_sig("glab_version", c_int)
_sig("glab_halo_fits", c_int, P, c_int, c_int)
_sig("glab_error_string", c_char_p, c_int)
_sig("glab_plan_create", c_int, _I64, _I64, _I64, P, P, P, POINTER(P))
_sig("glab_plan_create_csr", c_int, _I64, _I64, _I64, P, P, P, POINTER(P))
_sig("glab_plan_destroy", c_int, P)
_sig("glab_plan_info", c_int, P, POINTER(_I64), POINTER(_I64), POINTER(_I64), POINTER(_I32), POINTER(_I32))
_sig("glab_plan_csr", c_int, P, POINTER(P), POINTER(P), POINTER(P))
_sig("glab_plan_index_width", c_int, P, POINTER(_I32))
_sig("glab_plan_index16_tiles", c_int, P, POINTER(_I64), POINTER(_I64))
_sig("glab_plan_adopt_vals_f32", c_int, P, P, POINTER(P), P)
_sig("glab_plan_adopt_vals_f64", c_int, P, P, POINTER(P), P)
_sig("glab_plan_l2_persist", c_int, P, c_int, P)
_sig("glab_reduce_workspace_bytes", c_int64)
_sig("glab_ipc_handle_bytes", c_int)
_sig("glab_ipc_alloc", c_int, _I64, POINTER(P), P)
_sig("glab_ipc_open", c_int, P, POINTER(P))
_sig("glab_ipc_close", c_int, P)
_sig("glab_ipc_free", c_int, P)


class PushDesc(ctypes.Structure):
    _fields_ = [("send_idx", c_void_p), ("first_row", c_int64), ("count", c_int64), ("dst", c_void_p),
                ("dst_offset", c_int64), ("flag", c_void_p)]


MAX_PEERS = 8


class PeerReduce(ctypes.Structure):
    """glab_peer_reduce (include/glab.h)."""
    _fields_ = [("world", c_int32), ("rank", c_int32), ("mail_local", c_void_p), ("flag_local", c_void_p),
                ("mail_peer", c_void_p * MAX_PEERS), ("flag_peer", c_void_p * MAX_PEERS),
                ("parity_counter", c_void_p)]


class HaloStep(ctypes.Structure):
    """glab_halo_step (include/glab.h)."""
    _fields_ = [("interior_begin", c_int64), ("interior_end", c_int64), ("n_wait", c_int32),
                ("wait_flags", POINTER(c_void_p)), ("wait_target", c_void_p), ("n_push", c_int32),
                ("push", POINTER(PushDesc)), ("pushed_counter", c_void_p), ("push_src", c_void_p),
                ("done_counter", c_void_p), ("status", c_void_p), ("timeout_ms", c_int64),
                ("reduce", POINTER(PeerReduce))]


_sig("glab_halo_wait", c_int, c_int, POINTER(c_void_p), P, P)
_sig("glab_spgemm_products", c_int, P, P, P, POINTER(_I64), POINTER(_I64), P)
_sig("glab_spgemm_workspace_bytes", _I64, _I64, _I64, _I64, c_int)
_sig("glab_cf_split_workspace_bytes", _I64, _I64)
_sig("glab_interp_workspace_bytes", _I64, _I64)

for _suf, _ct in (("f32", c_float), ("f64", c_double)):
    _sig("glab_gather_vals_" + _suf, c_int, P, P, _I64, _I64, P, P)
    _sig("glab_scatter_edges_" + _suf, c_int, P, P, P, _I64, _I64, P)
    _sig("glab_spmm_" + _suf, c_int, P, P, P, _INT, P, _I64, _I64, P)
    _sig("glab_spmm_add_" + _suf, c_int, P, P, P, P, _INT, P, _I64, _I64, P)
    _sig("glab_residual_" + _suf, c_int, P, P, P, P, _INT, P, _I64, _I64, P)
    _sig("glab_jacobi_" + _suf, c_int, P, P, P, P, P, P, P, _INT, _I64, _I64, P)
    _sig("glab_cheby_first_" + _suf, c_int, P, P, P, P, P, P, P, P, _INT, _I64, _I64, P)
    _sig("glab_cheby_next_" + _suf, c_int, P, P, P, P, P, P, P, P, P, _INT, _I64, _I64, P)
    _sig("glab_power_step_" + _suf, c_int, P, P, P, P, P, P, P, _I64, _I64, P)
    _sig("glab_rayleigh_" + _suf, c_int, P, P, P, P, P, P, P, P, _I64, _I64, P)
    _sig("glab_xtax_" + _suf, c_int, P, P, P, P, P, _I64, _I64, P)
    _sig("glab_edge_messages_" + _suf, c_int, P, P, P, _INT, P, _I64, _I64, P)
    _sig("glab_edge_attr_" + _suf, c_int, P, P, P, _INT, P, _I64, P)
    _H = POINTER(HaloStep)
    _sig("glab_spmm_halo_" + _suf, c_int, P, P, P, _INT, P, _H, P)
    _sig("glab_residual_halo_" + _suf, c_int, P, P, P, P, _INT, P, _H, P)
    _sig("glab_jacobi_halo_" + _suf, c_int, P, P, P, P, P, P, P, _INT, _H, P)
    _sig("glab_jacobi_sweeps_" + _suf, c_int, P, P, P, P, P, P, P, _INT, _INT, P)
    _sig("glab_jacobi_sweeps_halo_" + _suf, c_int, P, P, P, P, P, P, P, _INT, _INT, _H, _H, P)
    _sig("glab_cheby_first_halo_" + _suf, c_int, P, P, P, P, P, P, P, P, _INT, _H, P)
    _sig("glab_cheby_next_halo_" + _suf, c_int, P, P, P, P, P, P, P, P, P, _INT, _H, P)
    _sig("glab_power_step_halo_" + _suf, c_int, P, P, P, P, P, P, P, _H, P)
    _sig("glab_rayleigh_halo_" + _suf, c_int, P, P, P, P, P, P, P, P, _H, P)
    _sig("glab_pack_" + _suf, c_int, _I64, _I64, _INT, POINTER(c_void_p), POINTER(c_int32), POINTER(c_int32), P, P)
    _sig("glab_unpack_" + _suf, c_int, _I64, _I64, _INT, POINTER(c_void_p), POINTER(c_int32), POINTER(c_int32), P, P)
    _sig("glab_segment_sum_" + _suf, c_int, P, P, _INT, P, P)
    _sig("glab_segment_max_" + _suf, c_int, P, P, P, P)
    _sig("glab_segment_min_" + _suf, c_int, P, P, P, P)
    _sig("glab_segment_mean_" + _suf, c_int, P, P, _INT, P, P)
    _sig("glab_segment_agg4_" + _suf, c_int, P, P, _INT, P, P)
    _sig("glab_soc_classic_" + _suf, c_int, P, P, _ct, P, P, P)
    _sig("glab_soc_sa_" + _suf, c_int, P, P, P, P, P)
    _sig("glab_direct_interp_" + _suf, c_int, P, P, P, P, P, P, P)
    _sig("glab_halo_push_" + _suf, c_int, P, _INT, _INT, POINTER(PushDesc), P, P)
    _sig("glab_interp_count_" + _suf, c_int, P, P, P, _INT, P, _I64, P, P, POINTER(_I64), POINTER(_I64), P)
    _sig("glab_interp_fill_" + _suf, c_int, P, P, P, _INT, P, P, P, P, P, P)
    _sig("glab_spgemm_symbolic_" + _suf, c_int, P, P, P, P, P, _I64, _I64, _I64, POINTER(_I64), P)
    _sig("glab_spgemm_numeric_" + _suf, c_int, P, P, P, P, P, _I64, _I64, _I64, _I64, P, P, P, P)
    _sig("glab_cf_split_pmis_" + _suf, c_int, P, P, c_uint32, P, _I64, P, POINTER(_I32), P)


def check(rc, what=""):
    if rc != 0:
        msg = lib.glab_error_string(int(rc)).decode()
        raise GlabError("%s failed: %s (code %d)" % (what or "glab call", msg, rc))


def exported_symbols():
    """Names of every entry point this module bound (used by the CPU-side ABI test)."""
    return sorted(n for n in dir(lib) if n.startswith("glab_"))
