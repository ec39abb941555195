"""Host-side runtime above the C ABI: plans (int64 COO -> int32 CSR), value arrays, scalar
staging and one thin Python wrapper per C entry point.  torch is used for device memory,
streams and (elsewhere) torch.distributed -- plumbing only; every numerical operation on the
hot path is a kernel of libglab_b200.so.  Nothing here falls back to PyTorch math or the CPU.
"""
import collections
import ctypes
import weakref

import torch

from ._lib import GlabError, P, check, lib

SUPPORTED_K = (1, 2, 4, 8)
_SUF = {torch.float32: "f32", torch.float64: "f64"}


def _require_cuda():
    if not torch.cuda.is_available():
        raise GlabError("glab_b200 needs a CUDA device (B200 / sm_100a); there is no CPU path")


def suffix(dtype):
    try:
        return _SUF[dtype]
    except KeyError:
        raise GlabError("unsupported dtype %s (float32 / float64 only)" % dtype)


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def compute_device(*tensors):
    """Device the step runs on: the first CUDA tensor's device, else the current CUDA device
    (host tensors are uploaded, results are returned to the host -- the reference's users hold
    CPU tensors)."""
    _require_cuda()
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device
    return torch.device("cuda", torch.cuda.current_device())


def to_device(t, device, dtype=None):
    if t is None:
        return None
    if t.device != device:
        t = t.to(device, non_blocking=True)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t


def dense(t):
    """Contiguous, 16-byte aligned view/copy (vector loads in the kernels need it)."""
    t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone(memory_format=torch.contiguous_format)
    return t


def column(t, j):
    """Column j of a row-major [n, F] tensor as a dense [n] tensor."""
    return dense(t[:, j])


class _DevArray:
    """Exposes a raw device pointer owned by a plan to torch (zero-copy) via the CUDA array
    interface; keeps the owner alive."""

    def __init__(self, pointer, n, typestr, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (pointer, False),
                                         "version": 2}


class Plan:
    """Opaque int32 CSR structure of one operator on one GPU (glab_plan)."""

    def __init__(self, handle, device):
        self._h = handle
        self.device = device
        n_rows, n_cols, nnz = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        mx, ident = ctypes.c_int32(), ctypes.c_int32()
        check(lib.glab_plan_info(self._h, ctypes.byref(n_rows), ctypes.byref(n_cols),
                                 ctypes.byref(nnz), ctypes.byref(mx), ctypes.byref(ident)),
              "glab_plan_info")
        self.n_rows, self.n_cols, self.nnz = n_rows.value, n_cols.value, nnz.value
        self.max_row_nnz, self.identity = mx.value, bool(ident.value)
        iw = ctypes.c_int32()
        check(lib.glab_plan_index_width(self._h, ctypes.byref(iw)), "glab_plan_index_width")
        self.index_bytes = iw.value      # 2: every tile streams 16-bit row-relative column indices, 4: not all
        t16, tt = ctypes.c_int64(), ctypes.c_int64()
        check(lib.glab_plan_index16_tiles(self._h, ctypes.byref(t16), ctypes.byref(tt)), "glab_plan_index16_tiles")
        self.index16_tiles, self.tiles = t16.value, tt.value

    @property
    def handle(self):
        return self._h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib.glab_plan_destroy(h)
            except Exception:
                pass

    @classmethod
    def from_coo(cls, edge_index, n_rows, n_cols=None):
        """edge_index: int64 [2, z]; row 0 = aggregation target i, row 1 = source j
        (the reference's edgeij_pair, UtilsGNN.py:74-78)."""
        device = compute_device(edge_index)
        if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.shape[0] != 2:
            raise GlabError("edgeij_pair must be an int64 [2, nnz] tensor")
        ei = to_device(edge_index, device).contiguous()
        n_cols = n_rows if n_cols is None else n_cols
        z = ei.shape[1]
        h = P()
        with torch.cuda.device(device):
            check(lib.glab_plan_create(n_rows, n_cols, z, ptr(ei[0]), ptr(ei[1]), stream_ptr(),
                                       ctypes.byref(h)), "glab_plan_create")
        return cls(h, device)

    @classmethod
    def from_csr(cls, rowptr, colidx, n_rows, n_cols):
        device = compute_device(rowptr)
        rp = to_device(rowptr, device, torch.int32).contiguous()
        ci = to_device(colidx, device, torch.int32).contiguous()
        h = P()
        with torch.cuda.device(device):
            check(lib.glab_plan_create_csr(n_rows, n_cols, ci.numel(), ptr(rp), ptr(ci), stream_ptr(),
                                           ctypes.byref(h)), "glab_plan_create_csr")
        return cls(h, device)

    def csr(self):
        """(rowptr, colidx, perm-or-None) as zero-copy int32 torch tensors."""
        rp, ci, pm = P(), P(), P()
        check(lib.glab_plan_csr(self._h, ctypes.byref(rp), ctypes.byref(ci), ctypes.byref(pm)),
              "glab_plan_csr")

        def wrap(p, n):
            if not p.value or n == 0:
                return torch.empty(0, dtype=torch.int32, device=self.device)
            with torch.cuda.device(self.device):
                return torch.as_tensor(_DevArray(p.value, n, "<i4", self), device=self.device)

        return wrap(rp, self.n_rows + 1), wrap(ci, self.nnz), (wrap(pm, self.nnz) if pm.value else None)


def adopt_vals(plan, vals):
    """Copy the CSR-ordered values into plan-owned storage right behind colidx (so that ONE L2
    access-policy window can cover the whole operator) and return a zero-copy tensor over it."""
    out = P()
    with torch.cuda.device(plan.device):
        check(getattr(lib, "glab_plan_adopt_vals_" + suffix(vals.dtype))(plan.handle, ptr(vals), ctypes.byref(out),
                                                                        stream_ptr()), "glab_plan_adopt_vals")
        typestr = "<f4" if vals.dtype == torch.float32 else "<f8"
        if plan.nnz == 0:
            return vals
        return torch.as_tensor(_DevArray(out.value, plan.nnz, typestr, plan), device=plan.device)


def l2_persist(plan, enable=True):
    """Mark the plan's [colidx | adopted values] as persisting in L2 for kernels launched on (or
    captured from) the current stream."""
    with torch.cuda.device(plan.device):
        check(lib.glab_plan_l2_persist(plan.handle, 1 if enable else 0, stream_ptr()), "glab_plan_l2_persist")


# --------------------------------------------------------------------------- caches
class _Cache:
    """Small LRU keyed on (storage pointer, shape, strides, version, extra).  An entry is valid
    only while the tensor object that created it is alive (an alive tensor pins its storage, so
    the address cannot have been recycled for different data) and unmodified (_version)."""

    def __init__(self, capacity=8):
        self.capacity = capacity
        self.d = collections.OrderedDict()

    @staticmethod
    def key(t, extra=()):
        return (t.data_ptr(), tuple(t.shape), tuple(t.stride()), t.dtype, str(t.device), t._version) + tuple(extra)

    def get(self, t, extra=()):
        k = self.key(t, extra)
        hit = self.d.get(k)
        if hit is None:
            return None
        ref, value = hit
        if ref() is None:
            del self.d[k]
            return None
        self.d.move_to_end(k)
        return value

    def put(self, t, value, extra=()):
        k = self.key(t, extra)
        self.d[k] = (weakref.ref(t), value)
        self.d.move_to_end(k)
        while len(self.d) > self.capacity:
            self.d.popitem(last=False)

    def clear(self):
        self.d.clear()


_plan_cache = _Cache(8)
_vals_cache = _Cache(8)


def clear_caches():
    _plan_cache.clear()
    _vals_cache.clear()


def get_plan(edge_index, n_rows, n_cols=None):
    """Cached plan for the caller's edgeij_pair tensor (device or host)."""
    extra = (n_rows, n_cols)
    plan = _plan_cache.get(edge_index, extra)
    if plan is None:
        plan = Plan.from_coo(edge_index, n_rows, n_cols)
        _plan_cache.put(edge_index, plan, extra)
    return plan


def get_vals(plan, edge_attr, col=0, dtype=None):
    """A_ij in CSR slot order as a dense device array of `dtype` (default: edge_attr's).
    Zero-copy when the caller's edge order is already the CSR order and the column is dense."""
    dtype = dtype or edge_attr.dtype
    extra = (id(plan), col, dtype)
    hit = _vals_cache.get(edge_attr, extra)
    if hit is not None and hit[0] is plan:
        return hit[1]
    ea = to_device(edge_attr, plan.device)
    if ea.dim() == 1:
        ea = ea.view(-1, 1)
    if ea.shape[0] != plan.nnz:
        raise GlabError("edge_attr has %d rows, plan has %d edges" % (ea.shape[0], plan.nnz))
    if ea.dtype != dtype:
        ea = ea[:, col:col + 1].to(dtype)
        col = 0
    if plan.identity and ea.shape[1] == 1 and ea.is_contiguous() and ea.data_ptr() % 16 == 0:
        v = ea.view(-1)
    else:
        ea = ea if ea.stride(1) == 1 else ea.contiguous()
        v = torch.empty(plan.nnz, dtype=dtype, device=plan.device)
        with torch.cuda.device(plan.device):
            check(getattr(lib, "glab_gather_vals_" + suffix(dtype))(
                plan.handle, ptr(ea), ea.stride(0), col, ptr(v), stream_ptr()), "glab_gather_vals")
    _vals_cache.put(edge_attr, (plan, v), extra)
    return v


def slot_order(plan, per_edge, dtype=None):
    """Any per-edge 1-D array (S_ij ...) -> CSR slot order, dense."""
    return get_vals(plan, per_edge.view(-1, 1) if per_edge.dim() == 1 else per_edge, 0, dtype)


_scalar_cache = {}


def scalar(value, device, dtype):
    """A 1-element device tensor holding `value` (python number or 0-d/1-element tensor),
    rounded to `dtype` the way torch rounds a 0-d operand of a tensor op.  Device tensors stay on
    the device (no sync); host values are uploaded once per distinct (value, dtype, device) and
    cached -- a layer's omega is the same number call after call, and an upload from pageable
    memory per call costs more host time than the kernels it parameterises."""
    if isinstance(value, torch.Tensor):
        if value.is_cuda:
            return value.reshape(-1)[:1].to(device=device, dtype=dtype, non_blocking=True)
        value = value.reshape(-1)[0].item()
    key = (float(value), str(device), dtype)
    hit = _scalar_cache.get(key)
    if hit is None:
        if len(_scalar_cache) > 256:
            _scalar_cache.clear()
        hit = torch.tensor([value], dtype=dtype).to(device)
        _scalar_cache[key] = hit
    return hit


_workspaces = {}


def reduce_workspace(device):
    """Zero-initialised scratch for the deterministic grid reductions (one per device+stream)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None:
        ws = torch.zeros(int(lib.glab_reduce_workspace_bytes()) // 8 + 1, dtype=torch.float64, device=device)
        _workspaces[key] = ws
    return ws


# --------------------------------------------------------------------------- op wrappers
def _k_of(x):
    k = 1 if x.dim() == 1 else x.shape[1]
    if k not in SUPPORTED_K:
        raise GlabError("number of right-hand-side columns must be one of %s, got %d" % (SUPPORTED_K, k))
    return k


def _rows(plan, rows):
    return (0, plan.n_rows) if rows is None else (int(rows[0]), int(rows[1]))


launch_count = 0  # number of libglab kernels launched through the wrappers below (bench.py reads it)


def _call(name, dtype, device, *args):
    global launch_count
    launch_count += 1
    with torch.cuda.device(device):
        check(getattr(lib, "glab_%s_%s" % (name, suffix(dtype)))(*args), "glab_" + name)


def spmm(plan, vals, x, out=None, rows=None, halo=None):
    k = _k_of(x)
    if out is None:
        out = torch.empty((plan.n_rows,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    if halo is not None:
        _call("spmm_halo", x.dtype, plan.device, plan.handle, ptr(vals), ptr(x), k, ptr(out), halo, stream_ptr())
        return out
    rb, re = _rows(plan, rows)
    _call("spmm", x.dtype, plan.device, plan.handle, ptr(vals), ptr(x), k, ptr(out), rb, re, stream_ptr())
    return out


def residual(plan, vals, x, b, out=None, rows=None, halo=None):
    k = _k_of(x)
    if out is None:
        out = torch.empty_like(b)
    if halo is not None:
        _call("residual_halo", x.dtype, plan.device, plan.handle, ptr(vals), ptr(x), ptr(b), k, ptr(out), halo,
              stream_ptr())
        return out
    rb, re = _rows(plan, rows)
    _call("residual", x.dtype, plan.device, plan.handle, ptr(vals), ptr(x), ptr(b), k, ptr(out), rb, re,
          stream_ptr())
    return out


def spmm_add(plan, vals, x, b, out=None, rows=None):
    k = _k_of(x)
    if out is None:
        out = torch.empty_like(b)
    rb, re = _rows(plan, rows)
    _call("spmm_add", x.dtype, plan.device, plan.handle, ptr(vals), ptr(x), ptr(b), k, ptr(out), rb, re,
          stream_ptr())
    return out


def jacobi(plan, vals, diag, b, x_in, x_out, omega_dev, rows=None, halo=None):
    k = _k_of(x_in)
    if halo is not None:
        _call("jacobi_halo", x_in.dtype, plan.device, plan.handle, ptr(vals), ptr(diag), ptr(b), ptr(x_in),
              ptr(x_out), ptr(omega_dev), k, halo, stream_ptr())
        return x_out
    rb, re = _rows(plan, rows)
    _call("jacobi", x_in.dtype, plan.device, plan.handle, ptr(vals), ptr(diag), ptr(b), ptr(x_in),
          ptr(x_out), ptr(omega_dev), k, rb, re, stream_ptr())
    return x_out


def jacobi_sweeps(plan, vals, diag, b, xa, xb, omega_dev, n_sweeps, halo=None):
    """n_sweeps weighted-Jacobi sweeps ping-ponging xa -> xb -> xa ... in ONE launch of the multi-sweep
    kernel (glab_jacobi_sweeps_*; the library falls back to one fused launch per sweep for operators
    that kernel cannot take).  Returns the buffer that holds the result.  halo = (step gathering xa and
    producing xb, step gathering xb and producing xa) on a row-partitioned operator."""
    k = _k_of(xa)
    if n_sweeps <= 0:
        return xa
    if halo is not None:
        _call("jacobi_sweeps_halo", xa.dtype, plan.device, plan.handle, ptr(vals), ptr(diag), ptr(b), ptr(xa), ptr(xb),
              ptr(omega_dev), k, int(n_sweeps), halo[0], halo[1], stream_ptr())
    else:
        _call("jacobi_sweeps", xa.dtype, plan.device, plan.handle, ptr(vals), ptr(diag), ptr(b), ptr(xa), ptr(xb),
              ptr(omega_dev), k, int(n_sweeps), stream_ptr())
    return xb if n_sweeps % 2 else xa


def cheby_first(plan, vals, b, x_in, x_out, r, p, alpha_dev, rows=None, halo=None):
    k = _k_of(x_in)
    if halo is not None:
        _call("cheby_first_halo", x_in.dtype, plan.device, plan.handle, ptr(vals), ptr(b), ptr(x_in), ptr(x_out),
              ptr(r), ptr(p), ptr(alpha_dev), k, halo, stream_ptr())
        return
    rb, re = _rows(plan, rows)
    _call("cheby_first", x_in.dtype, plan.device, plan.handle, ptr(vals), ptr(b), ptr(x_in), ptr(x_out),
          ptr(r), ptr(p), ptr(alpha_dev), k, rb, re, stream_ptr())


def cheby_next(plan, vals, p_in, p_out, r, x, alpha_old_dev, alpha_dev, beta_dev, rows=None, halo=None):
    k = _k_of(p_in)
    if halo is not None:
        _call("cheby_next_halo", p_in.dtype, plan.device, plan.handle, ptr(vals), ptr(p_in), ptr(p_out), ptr(r),
              ptr(x), ptr(alpha_old_dev), ptr(alpha_dev), ptr(beta_dev), k, halo, stream_ptr())
        return
    rb, re = _rows(plan, rows)
    _call("cheby_next", p_in.dtype, plan.device, plan.handle, ptr(vals), ptr(p_in), ptr(p_out), ptr(r),
          ptr(x), ptr(alpha_old_dev), ptr(alpha_dev), ptr(beta_dev), k, rb, re, stream_ptr())


def power_step(plan, vals, b_in, y, sumsq_in, sumsq_out, rows=None, halo=None):
    ws = reduce_workspace(plan.device)
    if halo is not None:
        _call("power_step_halo", b_in.dtype, plan.device, plan.handle, ptr(vals), ptr(b_in), ptr(y),
              ptr(sumsq_in), ptr(sumsq_out), ptr(ws), halo, stream_ptr())
        return
    rb, re = _rows(plan, rows)
    _call("power_step", b_in.dtype, plan.device, plan.handle, ptr(vals), ptr(b_in), ptr(y), ptr(sumsq_in),
          ptr(sumsq_out), ptr(ws), rb, re, stream_ptr())


def rayleigh(plan, vals, b_in, b_out, y_out, sumsq_in, sums_out, rows=None, halo=None):
    ws = reduce_workspace(plan.device)
    if halo is not None:
        _call("rayleigh_halo", b_in.dtype, plan.device, plan.handle, ptr(vals), ptr(b_in), ptr(b_out), ptr(y_out),
              ptr(sumsq_in), ptr(sums_out), ptr(ws), halo, stream_ptr())
        return
    rb, re = _rows(plan, rows)
    _call("rayleigh", b_in.dtype, plan.device, plan.handle, ptr(vals), ptr(b_in), ptr(b_out), ptr(y_out),
          ptr(sumsq_in), ptr(sums_out), ptr(ws), rb, re, stream_ptr())


def xtax(plan, vals, x, sums_out, rows=None):
    rb, re = _rows(plan, rows)
    ws = reduce_workspace(plan.device)
    _call("xtax", x.dtype, plan.device, plan.handle, ptr(vals), ptr(x), ptr(sums_out), ptr(ws), rb, re,
          stream_ptr())


def edge_messages(plan, vals, x, out_edges, col):
    """out_edges[:, col:col+k] = A_ij * x_j in the caller's edge order (out_edges is [z, ld])."""
    k = _k_of(x)
    _call("edge_messages", x.dtype, plan.device, plan.handle, ptr(vals), ptr(x), k, ptr(out_edges),
          out_edges.stride(0), col, stream_ptr())


def with_messages(plan, vals, x, A_col=None):
    """The reference's returned edge_attr = cat([A_ij, c_ij], 1) (e.g. MatVecGNN.py:84), in the
    caller's edge order, produced by one kernel (vals already holds A_ij in slot order)."""
    k = _k_of(x)
    out = torch.empty((plan.nnz, 1 + k), dtype=x.dtype, device=plan.device)
    _call("edge_attr", x.dtype, plan.device, plan.handle, ptr(vals), ptr(x), k, ptr(out), out.stride(0),
          stream_ptr())
    return out


def pack(parts):
    """torch.cat(parts, 1) for dense [n, w_j] device tensors, as one coalesced kernel."""
    n = parts[0].shape[0]
    dt, dev = parts[0].dtype, parts[0].device
    if n == 0:
        return torch.cat([p_.reshape(0, p_.shape[1] if p_.dim() > 1 else 1) for p_ in parts], 1)
    parts = [p_.view(n, -1) for p_ in parts]
    widths = [p_.shape[1] for p_ in parts]
    ld = sum(widths)
    out = torch.empty((n, ld), dtype=dt, device=dev)
    if len(parts) > 8 or ld > 64:
        return torch.cat(parts, 1)
    ptrs = (ctypes.c_void_p * len(parts))(*[p_.data_ptr() for p_ in parts])
    w = (ctypes.c_int32 * len(parts))(*widths)
    offs, acc = [], 0
    for x_ in widths:
        offs.append(acc)
        acc += x_
    o = (ctypes.c_int32 * len(parts))(*offs)
    _call("pack", dt, dev, n, ld, len(parts), ptrs, w, o, ptr(out), stream_ptr())
    return out


def unpack(src, spans, outs=None):
    """Dense copies of column blocks of an interleaved [n, F] device tensor: spans = [(offset,
    width), ...] -> list of [n, width] tensors (one coalesced kernel instead of strided copies).
    outs: optional list with a preallocated dense [n, width] destination (or None) per span."""
    n, ld = src.shape
    given = list(outs) if outs is not None else [None] * len(spans)
    if n == 0:
        return [torch.empty((0, w), dtype=src.dtype, device=src.device) if t is None else t
                for (_, w), t in zip(spans, given)]
    if len(spans) > 8 or ld > 64:
        res = []
        for (o, w), t in zip(spans, given):
            if t is None:
                res.append(dense(src[:, o:o + w]))
            else:
                t.view(n, w).copy_(src[:, o:o + w])
                res.append(t)
        return res
    src = src.contiguous()
    outs = [torch.empty((n, w), dtype=src.dtype, device=src.device) if t is None else t.view(n, w)
            for (_, w), t in zip(spans, given)]
    ptrs = (ctypes.c_void_p * len(spans))(*[t.data_ptr() for t in outs])
    w = (ctypes.c_int32 * len(spans))(*[w_ for _, w_ in spans])
    o = (ctypes.c_int32 * len(spans))(*[o_ for o_, _ in spans])
    _call("unpack", src.dtype, src.device, n, ld, len(spans), ptrs, w, o, ptr(src), stream_ptr())
    return outs


def segment_sum(plan, src_slots, out=None):
    k = _k_of(src_slots)
    if out is None:
        out = torch.empty((plan.n_rows,) + tuple(src_slots.shape[1:]), dtype=src_slots.dtype,
                          device=plan.device)
    _call("segment_sum", src_slots.dtype, plan.device, plan.handle, ptr(src_slots), k, ptr(out), stream_ptr())
    return out


def segment_max(plan, src_slots, out=None):
    if out is None:
        out = torch.empty(plan.n_rows, dtype=src_slots.dtype, device=plan.device)
    _call("segment_max", src_slots.dtype, plan.device, plan.handle, ptr(src_slots), ptr(out), stream_ptr())
    return out


def segment_min(plan, src_slots, out=None):
    if out is None:
        out = torch.empty(plan.n_rows, dtype=src_slots.dtype, device=plan.device)
    _call("segment_min", src_slots.dtype, plan.device, plan.handle, ptr(src_slots), ptr(out), stream_ptr())
    return out


def segment_mean(plan, src_slots, out=None):
    k = _k_of(src_slots)
    if out is None:
        out = torch.empty((plan.n_rows,) + tuple(src_slots.shape[1:]), dtype=src_slots.dtype,
                          device=plan.device)
    _call("segment_mean", src_slots.dtype, plan.device, plan.handle, ptr(src_slots), k, ptr(out), stream_ptr())
    return out


_scatter_plans = _Cache(16)


def _index_plan(index, n):
    """Plan that groups by `index` (edgeij_pair[0] for edge -> vertex aggregation, the `batch` vector
    of a batch of graphs for vertex / edge -> graph aggregation), cached per index tensor."""
    plan = _scatter_plans.get(index, (n,))
    if plan is None:
        device = compute_device(index)
        idx = to_device(index, device).reshape(-1).to(torch.int64)
        ei = torch.stack([idx, torch.zeros_like(idx)]).contiguous()
        plan = Plan.from_coo(ei, n, 1)
        _scatter_plans.put(index, plan, (n,))
    return plan


def _slots(plan, src2):
    """[z, F] rows of src in CSR slot order (zero-copy when the index was already sorted)."""
    if plan.identity:
        return src2.contiguous()
    perm = plan.csr()[2]
    return src2.index_select(0, perm.long()).contiguous()


def aggregate4(src, index, dim_size=None):
    """cat([scatter(min), scatter(mean), scatter(sum), scatter(max)], 1) of the reference's 4-way
    aggregations (TrainableJacobiGNN.py:53-70, LearnDiffusionCoeffs.py:291-342) in ONE kernel pass:
    src [z, F] (F <= 64), index [z] -> [n, 4 F].  `index` may be edgeij_pair[0] or a `batch` vector."""
    device = compute_device(src, index)
    host = not src.is_cuda
    n = int(dim_size) if dim_size is not None else (int(index.max().item()) + 1 if index.numel() else 0)
    plan = _index_plan(index, n)
    s2 = to_device(src, device)
    s2 = s2.view(-1, 1) if s2.dim() == 1 else s2
    F = s2.shape[1]
    slots = _slots(plan, s2)
    out = torch.empty((n, 4 * F), dtype=s2.dtype, device=device)
    _call("segment_agg4", s2.dtype, device, plan.handle, ptr(slots), F, ptr(out), stream_ptr())
    return out.cpu() if host else out


def scatter(src, index, dim=0, dim_size=None, reduce="sum"):
    """torch_scatter.scatter(src, index, dim=0, dim_size=n, reduce=sum|max|min|mean) on the GPU:
    the reference's seam as a drop-in function (index = edgeij_pair[0], or a `batch` vector).  The
    plan is cached per index tensor; any number of feature columns in one launch."""
    if dim != 0:
        raise GlabError("only dim=0 (aggregation over edges) is supported")
    if reduce not in ("sum", "add", "mean", "max", "min"):
        raise GlabError("unknown reduce %r" % (reduce,))
    device = compute_device(src, index)
    host = not src.is_cuda
    n = int(dim_size) if dim_size is not None else (int(index.max().item()) + 1 if index.numel() else 0)
    plan = _index_plan(index, n)
    s2 = to_device(src, device)
    s2 = s2.view(-1, 1) if s2.dim() == 1 else s2
    F = s2.shape[1]
    slots = _slots(plan, s2)
    if reduce in ("sum", "add", "mean") and F in SUPPORTED_K and plan.max_row_nnz <= 64:
        out = (segment_sum if reduce != "mean" else segment_mean)(plan, slots)
    else:
        all4 = torch.empty((n, 4 * F), dtype=s2.dtype, device=device)
        _call("segment_agg4", s2.dtype, device, plan.handle, ptr(slots), F, ptr(all4), stream_ptr())
        block = {"min": 0, "mean": 1, "sum": 2, "add": 2, "max": 3}[reduce]
        out = all4[:, block * F:(block + 1) * F].contiguous()
    out = out.view(-1) if src.dim() == 1 else out
    return out.cpu() if host else out


def soc_classic(plan, vals, theta, rowmax=None):
    S = torch.empty(plan.nnz, dtype=vals.dtype, device=plan.device)
    _call("soc_classic", vals.dtype, plan.device, plan.handle, ptr(vals), float(theta), ptr(S), ptr(rowmax),
          stream_ptr())
    return S


def soc_sa(plan, vals, diag):
    S = torch.empty(plan.nnz, dtype=vals.dtype, device=plan.device)
    _call("soc_sa", vals.dtype, plan.device, plan.handle, ptr(vals), ptr(diag), ptr(S), stream_ptr())
    return S


def direct_interp(plan, vals, S_slots, diag, cflag):
    w = torch.empty(plan.nnz, dtype=vals.dtype, device=plan.device)
    _call("direct_interp", vals.dtype, plan.device, plan.handle, ptr(vals), ptr(S_slots), ptr(diag),
          ptr(cflag), ptr(w), stream_ptr())
    return w


# --------------------------------------------------------------------------- AMG setup (device)
def interp_assemble(plan_off, w_slots, cflag, mode=0):
    """Sparse prolongator P = [I + W](:, C) from the DirectInterpGNN weights (slot order) and the
    C/F flags: returns (edge_index int64 [2, nnz_P] sorted by (row, col), values [nnz_P],
    n_coarse).  mode 0 = Python reference (VCycle.py:126-137), 1 = coarse rows are identity rows
    (matlab/test_direct_interpolation.m:130-132)."""
    dev, n, dt = plan_off.device, plan_off.n_rows, w_slots.dtype
    cflag = dense(to_device(cflag, dev, dt).reshape(-1))
    prow = torch.empty(n + 1, dtype=torch.int32, device=dev)
    cid = torch.empty(n + 1, dtype=torch.int32, device=dev)
    nnz_p, n_coarse = ctypes.c_int64(), ctypes.c_int64()
    with torch.cuda.device(dev):
        ws_bytes = int(lib.glab_interp_workspace_bytes(n))
    if ws_bytes < 0:
        check(ws_bytes, "glab_interp_workspace_bytes")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _call("interp_count", dt, dev, plan_off.handle, ptr(w_slots), ptr(cflag), int(mode), ptr(ws), ws_bytes, ptr(prow),
          ptr(cid), ctypes.byref(nnz_p), ctypes.byref(n_coarse), stream_ptr())
    ei = torch.empty((2, nnz_p.value), dtype=torch.int64, device=dev)
    vals = torch.empty(nnz_p.value, dtype=dt, device=dev)
    if nnz_p.value:
        _call("interp_fill", dt, dev, plan_off.handle, ptr(w_slots), ptr(cflag), int(mode), ptr(prow), ptr(cid),
              ptr(ei[0]), ptr(ei[1]), ptr(vals), stream_ptr())
    return ei, vals, int(n_coarse.value)


def spgemm(plan_x, vals_x, plan_y, vals_y):
    """Z = X * Y on the device (row-local accumulation, or expand - sort - compress for dense-ish
    rows): returns (edge_index int64 [2, nnz_Z] sorted by (row, col) without duplicates, values
    [nnz_Z]) -- the reference's COO layout.  `spgemm.last` describes the most recent call (setup
    diagnostics: product count, workspace size, path taken)."""
    global launch_count
    dev, dt = plan_x.device, vals_x.dtype
    if vals_y.dtype != dt:
        raise GlabError("spgemm: operand dtypes differ (%s, %s)" % (dt, vals_y.dtype))
    scratch = torch.zeros(2, dtype=torch.int64, device=dev)
    n_prod, max_row = ctypes.c_int64(), ctypes.c_int64()
    with torch.cuda.device(dev):
        launch_count += 1
        check(lib.glab_spgemm_products(plan_x.handle, plan_y.handle, ptr(scratch), ctypes.byref(n_prod),
                                       ctypes.byref(max_row), stream_ptr()), "glab_spgemm_products")
        ws_bytes = int(lib.glab_spgemm_workspace_bytes(plan_x.n_rows, n_prod.value, max_row.value,
                                                       vals_x.element_size()))
    if ws_bytes < 0:
        check(ws_bytes, "glab_spgemm_workspace_bytes")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    nnz = ctypes.c_int64()
    _call("spgemm_symbolic", dt, dev, plan_x.handle, ptr(vals_x), plan_y.handle, ptr(vals_y), ptr(ws), ws_bytes,
          n_prod.value, max_row.value, ctypes.byref(nnz), stream_ptr())
    ei = torch.empty((2, nnz.value), dtype=torch.int64, device=dev)
    vals = torch.empty(nnz.value, dtype=dt, device=dev)
    if nnz.value:
        _call("spgemm_numeric", dt, dev, plan_x.handle, ptr(vals_x), plan_y.handle, ptr(vals_y), ptr(ws), ws_bytes,
              n_prod.value, max_row.value, nnz.value, ptr(ei[0]), ptr(ei[1]), ptr(vals), stream_ptr())
    spgemm.last = {"products": n_prod.value, "max_row_products": max_row.value, "workspace_bytes": ws_bytes,
                   "path": {1: "row-local", 2: "esc"}.get(int(ws[:32].view(torch.int64)[3].item()), "empty")}
    return ei, vals


def cf_split_pmis(plan_off, S_slots, seed=0):
    """PMIS coarse/fine splitting on the strength graph (strong <=> S > 0): returns (cflag [n] in
    S's dtype with 1 = coarse, number of rounds).  Bit-exact against oracle/cf_split.py."""
    dev, n, dt = plan_off.device, plan_off.n_rows, S_slots.dtype
    ws_bytes = int(lib.glab_cf_split_workspace_bytes(n))
    if ws_bytes < 0:
        check(ws_bytes, "glab_cf_split_workspace_bytes")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    cflag = torch.empty(n, dtype=dt, device=dev)
    rounds = ctypes.c_int32()
    _call("cf_split_pmis", dt, dev, plan_off.handle, ptr(S_slots), ctypes.c_uint32(int(seed) & 0xFFFFFFFF), ptr(ws),
          ws_bytes, ptr(cflag), ctypes.byref(rounds), stream_ptr())
    return cflag, int(rounds.value)
