"""Row-block partitioning + halo exchange for the multi-GPU path (one process per GPU).

The reference is single-process CPU code; this module is the new piece that BASELINE.json's
north_star asks for: "large structured operators are row-block partitioned across the GPUs of
one box; each smoother/matvec step exchanges halo rows; the power method's norms use an
allreduce".

Layout per rank r (P ranks):
  * rows [off[r], off[r+1]) of the operator, as a local CSR plan with n_rows = n_local and
    n_cols = n_local + n_halo: local columns keep their order (col - off[r]); columns owned by
    other ranks are renumbered n_local + (position in the sorted list of needed halo columns),
    so the halo tail is grouped by owner and ordered by global index;
  * every gathered vector is [n_local + n_halo, k]; the halo tail is refreshed before each
    SpMV-bearing step.
Two exchange engines share the same plan:
  * `exchange()`  -- torch.distributed isend/irecv of packed rows (gloo on CPU for the tests,
    NCCL on GPUs): the portable baseline;
  * `PeerHalo`    -- CUDA-IPC peer memory over NVLink: this rank's boundary rows are STORED
    directly into the neighbour's halo tail and the neighbour's arrival counter is release-
    incremented; the consumer acquires its counters on the device.  Either fused into the step
    kernel (glab_*_halo_*: one launch per sweep, communication CTA) or as stand-alone kernels
    (glab_halo_push_* / glab_halo_wait).  No NCCL call on the data path.
The partition logic (this file's torch index arithmetic) is covered on CPU by
tests/test_dist_cpu.py with world_size-2 gloo.
"""
import ctypes

import os

import torch
import torch.distributed as dist

from . import _runtime as rt
from ._lib import GlabError, check, lib

MAX_PEERS = 8      # GLAB_MAX_PEERS of include/glab.h


class RowPartition:
    """Contiguous row blocks: rank r owns rows [offsets[r], offsets[r+1])."""

    def __init__(self, n, world, align=1, offsets=None):
        if offsets is None:
            base = (n // world) // align * align
            offsets = [min(r * base, n) for r in range(world)] + [n]
        self.n, self.world = n, world
        self.offsets = torch.tensor([int(o) for o in offsets], dtype=torch.int64)

    def bounds(self, rank):
        return int(self.offsets[rank]), int(self.offsets[rank + 1])

    def owner(self, cols):
        """Owning rank of each global column index."""
        offs = self.offsets.to(cols.device)
        return torch.bucketize(cols, offs[1:-1], right=True)


class HaloPlan:
    """What this rank needs from / must send to every peer, derived from the global column
    indices of its row block (works for any matrix, not only stencils)."""

    def __init__(self, part, rank, halo_cols, recv_counts, send_rows):
        self.part, self.rank = part, rank
        self.r0, self.r1 = part.bounds(rank)
        self.n_local = self.r1 - self.r0
        self.halo_cols = halo_cols                      # sorted global ids, int64
        self.n_halo = int(halo_cols.numel())
        self.recv_counts = recv_counts                  # python list per peer
        self.recv_offsets = [0] * part.world
        acc = 0
        for q in range(part.world):
            self.recv_offsets[q] = acc
            acc += recv_counts[q]
        self.send_rows = send_rows                      # per peer: LOCAL row ids (int32 tensors)
        self.peers_recv = [q for q in range(part.world) if recv_counts[q] > 0]
        self.peers_send = [q for q in range(part.world) if send_rows[q].numel() > 0]
        # Every pair of ranks that exchanges rows in EITHER direction signals in BOTH: a rank that only
        # receives from q still bumps q's arrival counter (an empty push), which is the acknowledgement
        # the push protocol needs before q may overwrite that halo tail again (structurally
        # non-symmetric operators: q's rows reach us, ours never reach q).
        self.peers = sorted(set(self.peers_recv) | set(self.peers_send))

    @classmethod
    def build(cls, part, rank, global_cols, group=None):
        """global_cols: int64 tensor of the GLOBAL column index of every local edge."""
        r0, r1 = part.bounds(rank)
        dev = global_cols.device
        outside = (global_cols < r0) | (global_cols >= r1)
        halo_cols = torch.unique(global_cols[outside])             # sorted
        owners = part.owner(halo_cols)
        recv_counts = torch.bincount(owners, minlength=part.world).tolist()
        # tell every owner which of its rows we need
        need = [halo_cols[owners == q].cpu() for q in range(part.world)]
        if part.world == 1:
            gathered = [need]
        else:
            gathered = [None] * part.world
            dist.all_gather_object(gathered, need, group=group)
        send_rows = []
        for q in range(part.world):
            wanted = gathered[q][rank] if q != rank else torch.empty(0, dtype=torch.int64)
            send_rows.append((wanted - r0).to(torch.int32).to(dev))
        return cls(part, rank, halo_cols, recv_counts, send_rows)

    @classmethod
    def build_all(cls, part, global_cols_per_rank):
        """Single-process construction of every rank's plan (emulation / tests: all row blocks
        live in one process, so the request lists are exchanged by plain assignment)."""
        need_all, halo_all, counts_all = [], [], []
        for rank, gcols in enumerate(global_cols_per_rank):
            r0, r1 = part.bounds(rank)
            outside = (gcols < r0) | (gcols >= r1)
            halo_cols = torch.unique(gcols[outside])
            owners = part.owner(halo_cols)
            counts_all.append(torch.bincount(owners, minlength=part.world).tolist())
            need_all.append([halo_cols[owners == q].cpu() for q in range(part.world)])
            halo_all.append(halo_cols)
        plans = []
        for rank, gcols in enumerate(global_cols_per_rank):
            r0, _ = part.bounds(rank)
            send_rows = []
            for q in range(part.world):
                wanted = need_all[q][rank] if q != rank else torch.empty(0, dtype=torch.int64)
                send_rows.append((wanted - r0).to(torch.int32).to(gcols.device))
            plans.append(cls(part, rank, halo_all[rank], counts_all[rank], send_rows))
        return plans

    def local_columns(self, global_cols):
        """Global -> local column numbering (local block first, halo tail after)."""
        r0, r1 = self.r0, self.r1
        inside = (global_cols >= r0) & (global_cols < r1)
        pos = torch.searchsorted(self.halo_cols, global_cols.clamp(min=0))
        return torch.where(inside, global_cols - r0, pos + self.n_local)

    def interior_rows(self, local_rows, local_cols, align=256, n_rows=None, mark_send_rows=True):
        """[lo, hi): a contiguous range of rows none of which reads the halo tail or is sent to a
        neighbour (aligned to the kernel's 256-row tiles).  Rows outside it are the boundary rows:
        they wait for the exchange and are finished before the push."""
        n_rows = self.n_local if n_rows is None else n_rows
        touches = torch.zeros(n_rows, dtype=torch.bool, device=local_rows.device)
        touches[local_rows[local_cols >= self.n_local]] = True
        if mark_send_rows:
            for q in self.peers_send:   # rows we ship to neighbours must be finished before the push
                touches[self.send_rows[q].long().to(local_rows.device)] = True
        idx = torch.nonzero(touches).reshape(-1)
        if idx.numel() == 0:
            return 0, n_rows
        mid = n_rows // 2
        lead = idx[idx < mid]
        trail = idx[idx >= mid]
        lo = int(lead.max()) + 1 if lead.numel() else 0
        hi = int(trail.min()) if trail.numel() else n_rows
        lo = min((lo + align - 1) // align * align, n_rows)
        hi = max(hi // align * align, lo)
        return lo, hi

    # ------------------------------------------------------------------ portable exchange
    def exchange(self, x_ext, group=None):
        """Refresh the halo tail of x_ext ([n_local + n_halo, k]) with isend/irecv."""
        if self.part.world == 1:
            return
        ops, keep = [], []
        for q in self.peers_recv:
            a = self.n_local + self.recv_offsets[q]
            buf = x_ext[a:a + self.recv_counts[q]]
            ops.append(dist.P2POp(dist.irecv, buf, q, group=group))
        for q in self.peers_send:
            buf = x_ext.index_select(0, self.send_rows[q].long())
            keep.append(buf)
            ops.append(dist.P2POp(dist.isend, buf, q, group=group))
        if not ops:
            return
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def partition_coo(edge_index, edge_val, part, rank, group=None, col_part=None):
    """Rows of a GLOBAL COO that belong to `rank` (row partition `part`), with the columns
    renumbered for the local plan against the partition of the GATHERED vector (`col_part`,
    default: the same partition -- square operators; a prolongator gathers coarse vectors and
    produces fine ones, so its two partitions differ).
    Returns (local_edge_index [2, z_loc], local_vals [z_loc, F], HaloPlan of the gathered vector)."""
    r0, r1 = part.bounds(rank)
    mine = (edge_index[0] >= r0) & (edge_index[0] < r1)
    rows = edge_index[0][mine] - r0
    gcols = edge_index[1][mine]
    halo = HaloPlan.build(part if col_part is None else col_part, rank, gcols, group)
    cols = halo.local_columns(gcols)
    return torch.stack([rows, cols]), edge_val[mine], halo


def extend(x_local, n_halo):
    """[n_local, k] -> [n_local + n_halo, k] with a zero halo tail."""
    tail = torch.zeros((n_halo,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    return torch.cat([x_local, tail], 0).contiguous()


# ---------------------------------------------------------------------------- NVLink peer path
class PeerBuffer:
    """A device buffer other ranks can map (CUDA IPC).  `local` is a torch view of our copy,
    `peer_ptr[q]` the address of rank q's copy in OUR address space."""

    def __init__(self, numel, dtype, device, group=None):
        self.dtype, self.device, self.numel = dtype, device, numel
        esz = torch.empty(0, dtype=dtype).element_size()
        nbytes = max(numel * esz, 16)
        hb = lib.glab_ipc_handle_bytes()
        handle = (ctypes.c_ubyte * hb)()
        p = ctypes.c_void_p()
        with torch.cuda.device(device):
            check(lib.glab_ipc_alloc(nbytes, ctypes.byref(p), handle), "glab_ipc_alloc")
        self._ptr = p.value
        typestr = {torch.float32: "<f4", torch.float64: "<f8", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]
        self.local = torch.as_tensor(rt._DevArray(self._ptr, max(numel, 1), typestr, self), device=device)[:numel]
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, bytes(handle), group=group)
        self.peer_ptr = {}
        self._opened = []
        for q in range(world):
            if q == rank:
                self.peer_ptr[q] = self._ptr
                continue
            hq = (ctypes.c_ubyte * hb).from_buffer_copy(handles[q])
            pq = ctypes.c_void_p()
            with torch.cuda.device(device):
                check(lib.glab_ipc_open(hq, ctypes.byref(pq)), "glab_ipc_open")
            self.peer_ptr[q] = pq.value
            self._opened.append(pq.value)

    def close(self):
        for p in self._opened:
            lib.glab_ipc_close(ctypes.c_void_p(p))
        self._opened = []
        if self._ptr:
            lib.glab_ipc_free(ctypes.c_void_p(self._ptr))
            self._ptr = None


class PeerHalo:
    """Push-style halo exchange over peer memory for a set of named vectors.

    Every exchanged vector lives in a PeerBuffer of (n_local + n_halo) * k elements.  push(name)
    is ONE kernel launch that packs this rank's boundary rows of vector `name` straight into
    each neighbour's halo tail of the SAME-named vector and increments that neighbour's arrival
    counter; wait(name) is ONE single-CTA kernel that orders the consumer behind the arrival of
    all its neighbours' pushes.  Counters only count up (no host-side epoch), so a captured CUDA
    graph of sweeps can be replayed."""

    def __init__(self, halo, k, dtype, device, names, group=None):
        from ._lib import PushDesc
        self.halo, self.k, self.dtype, self.device, self.group = halo, k, dtype, device, group
        self.rank = halo.rank
        ext = halo.n_local + halo.n_halo
        self.bufs = {nm: PeerBuffer(ext * k, dtype, device, group) for nm in names}
        self.views = {nm: b.local.view(ext, k) for nm, b in self.bufs.items()}
        world = halo.part.world
        # words (16-byte spaced): arrival counter [name][peer] (bumped remotely by that peer) and
        # pushed counter [name] (how often THIS rank has pushed the vector; the wait target)
        self.flag_index = {nm: i for i, nm in enumerate(names)}
        self.flags = PeerBuffer((2 * len(names) * world + 4) * 4, torch.int32, device, group)
        self._steps = {}
        self._reduce = None
        offs = [None] * world
        if world > 1:
            dist.all_gather_object(offs, halo.recv_offsets, group=group)
        else:
            offs = [halo.recv_offsets]
        peer_n_local = [halo.part.bounds(q)[1] - halo.part.bounds(q)[0] for q in range(world)]
        self._suf = rt.suffix(dtype)
        self._push_desc, self._wait_args = {}, {}
        nw = len(names) * world
        # contiguous send blocks (every stencil slab) are copied without index loads
        self._first_row = {}
        if len(halo.peers) > MAX_PEERS:
            raise GlabError("row block exchanges halo rows with %d ranks; the peer-memory engine handles %d "
                            "(use engine='torch')" % (len(halo.peers), MAX_PEERS))
        for q in halo.peers:
            idx = halo.send_rows[q].long()
            if idx.numel() == 0:
                self._first_row[q] = 0                  # empty push: only the arrival counter is bumped
                continue
            first = int(idx[0].item())
            contiguous = bool(torch.equal(idx, torch.arange(first, first + idx.numel(), device=idx.device)))
            self._first_row[q] = first if contiguous else -1
        for nm in names:
            descs = (PushDesc * max(len(halo.peers), 1))()
            for i, q in enumerate(halo.peers):
                idx = halo.send_rows[q]
                descs[i].send_idx = idx.data_ptr() if idx.numel() else None
                descs[i].first_row = self._first_row[q]
                descs[i].count = idx.numel()
                descs[i].dst = self.bufs[nm].peer_ptr[q]
                descs[i].dst_offset = peer_n_local[q] + offs[q][self.rank]
                descs[i].flag = self.flags.peer_ptr[q] + (self.flag_index[nm] * world + self.rank) * 16
            self._push_desc[nm] = descs
            nf = len(halo.peers)
            fl = (ctypes.c_void_p * max(nf, 1))()
            base = self.flags.peer_ptr[self.rank]
            for i, q in enumerate(halo.peers):
                fl[i] = base + (self.flag_index[nm] * world + q) * 16
            pushed = base + (nw + self.flag_index[nm]) * 16
            self._wait_args[nm] = (nf, fl, ctypes.c_void_p(pushed))
        if world > 1:
            dist.barrier(group=group)

    def reduce_desc(self):
        """glab_peer_reduce of this rank (mailboxes + arrival counters in peer memory, created on first
        use -- collective): the reducing step kernels sum their two partial sums over the ranks
        themselves, no NCCL call per iteration."""
        if self._reduce is None:
            from ._lib import PeerReduce, MAX_PEERS as MP
            world = self.halo.part.world
            if world > MP:
                raise GlabError("in-kernel reduction handles %d ranks, world is %d" % (MP, world))
            mail = PeerBuffer(2 * MP * 2, torch.float64, self.device, self.group)
            flag = PeerBuffer((MP + 1) * 4, torch.int32, self.device, self.group)    # [MP] = publish count
            d = PeerReduce()
            d.world, d.rank = world, self.rank
            d.mail_local, d.flag_local = mail.peer_ptr[self.rank], flag.peer_ptr[self.rank]
            for q in range(world):
                d.mail_peer[q] = mail.peer_ptr[q]
                d.flag_peer[q] = flag.peer_ptr[q]
            d.parity_counter = flag.peer_ptr[self.rank] + MP * 16
            self._reduce = (d, mail, flag)
            if world > 1:
                dist.barrier(group=self.group)
        return self._reduce[0]

    def step(self, name_in, name_out, interior, reduce=False):
        """glab_halo_step for a FUSED step that gathers vector `name_in` and produces `name_out`
        (None: nothing to push): wait on the arrivals of name_in, push name_out from the kernel.
        reduce: attach the in-kernel sum over the ranks (power_step / rayleigh)."""
        from ._lib import HaloStep
        key = (name_in, name_out, interior, reduce)
        hit = self._steps.get(key)
        if hit is not None:
            return hit
        world = self.halo.part.world
        st = HaloStep()
        st.interior_begin, st.interior_end = interior
        nf, fl, pushed_in = self._wait_args[name_in]
        st.n_wait = nf
        st.wait_flags = fl
        st.wait_target = pushed_in
        if name_out is not None:
            st.n_push = len(self.halo.peers)
            st.push = self._push_desc[name_out]
            st.pushed_counter = self._wait_args[name_out][2]
            st.push_src = self.views[name_out].data_ptr()
        else:
            st.n_push = 0
            st.push = None
            st.pushed_counter = None
            st.push_src = None
        st.done_counter = self._spare_word(0)      # two counters, 16 bytes apart (words 0 and 1)
        st.status = self._spare_word(2)
        st.timeout_ms = 0                           # library default (GLAB_SPIN_TIMEOUT_MS)
        st.reduce = ctypes.pointer(self.reduce_desc()) if reduce else None
        self._steps[key] = st
        return st

    def _spare_word(self, i):
        """16-byte spaced scratch words behind the counters of the flag buffer: 0, 1 = boundary-tile
        counters of the fused steps, 2 = status word of the bounded in-kernel waits."""
        world = self.halo.part.world
        return self.flags.peer_ptr[self.rank] + (2 * len(self.flag_index) * world + i) * 16

    def status(self):
        """GLAB_STATUS_* bits recorded by in-kernel waits that gave up (0 = none).  Synchronises."""
        world = self.halo.part.world
        torch.cuda.synchronize(self.device)
        return int(self.flags.local[(2 * len(self.flag_index) * world + 2) * 4].item())

    def check(self):
        st = self.status()
        if st:
            names = [n for bit, n in ((1, "a neighbour's halo rows did not arrive"),
                                      (2, "this GPU's boundary tiles did not finish"),
                                      (4, "a tile of the previous sweep did not finish")) if st & bit]
            raise GlabError("fused halo step timed out on rank %d: %s (status %d); results of that step are "
                            "undefined" % (self.rank, "; ".join(names), st))

    def push(self, name):
        """After the kernel that produced vector `name`: send boundary rows to every neighbour."""
        n = len(self.halo.peers)
        fn = getattr(lib, "glab_halo_push_" + self._suf)
        rt.launch_count += 1
        with torch.cuda.device(self.device):
            check(fn(rt.ptr(self.views[name]), self.k, n, self._push_desc[name], self._wait_args[name][2],
                     rt.stream_ptr()), "glab_halo_push")

    def wait(self, name):
        """Before the kernel that gathers vector `name`'s halo tail."""
        nf, fl, pushed = self._wait_args[name]
        if nf == 0:
            return
        rt.launch_count += 1
        with torch.cuda.device(self.device):
            check(lib.glab_halo_wait(nf, fl, pushed, rt.stream_ptr()), "glab_halo_wait")

    def close(self):
        torch.cuda.synchronize(self.device)
        if dist.is_initialized() and self.halo.part.world > 1:
            dist.barrier(group=self.group)
        for b in self.bufs.values():
            b.close()
        self.flags.close()
        if self._reduce is not None:
            self._reduce[1].close()
            self._reduce[2].close()
            self._reduce = None
            self._steps = {}


class DistOperator:
    """One rank's row block of an operator + the halo machinery, with the fused layer steps of
    the single-GPU path applied block-wise:

        interior rows (no halo column)   -> kernel launched immediately
        wait for the neighbours' pushes  -> glab_halo_wait
        boundary rows                    -> kernel(s) on the two edge ranges
        push the new boundary values     -> glab_halo_push

    engine = "peer" (NVLink peer memory, fused one-kernel steps; default on GPUs), "peer-split"
    (peer memory, separate wait / boundary / push kernels) or "torch" (isend/irecv through
    torch.distributed: NCCL on GPUs; the same HaloPlan logic is what tests/test_dist_cpu.py
    exercises with gloo)."""

    ENTRY = ("v0", "v1")      # vectors a caller loads / publishes; never overwritten by the steps below

    def __init__(self, local_edge_index, local_vals, halo, k=1, engine="peer", group=None, n_rows=None,
                 names=("v0", "v1", "va", "vb")):
        """n_rows: number of local ROWS when it differs from the local length of the gathered
        vector (rectangular operators: restriction / prolongation); such operators are applied
        with apply_rect() (stand-alone wait, whole-block kernel), not with the fused steps."""
        self.halo, self.k, self.group, self.engine = halo, k, group, engine
        self.device = local_vals.device
        self.dtype = local_vals.dtype
        n_loc, n_ext = halo.n_local, halo.n_local + halo.n_halo
        self.square = n_rows is None or n_rows == n_loc
        self.n_rows = n_loc if n_rows is None else n_rows
        self.plan = rt.Plan.from_coo(local_edge_index.contiguous(), self.n_rows, n_ext)
        self.vals = rt.get_vals(self.plan, local_vals.view(-1, 1))
        self._keep = (local_edge_index, local_vals)
        self.lo, self.hi = halo.interior_rows(local_edge_index[0], local_edge_index[1], n_rows=self.n_rows,
                                              mark_send_rows=self.square)
        self.names = list(names)
        if engine in ("peer", "peer-split"):
            self.peer = PeerHalo(halo, k, self.dtype, self.device, self.names, group)
            self.vec = self.peer.views
        else:
            self.peer = None
            self.vec = {nm: torch.zeros(n_ext, k, dtype=self.dtype, device=self.device) for nm in self.names}
        self.n_local, self.n_ext = n_loc, n_ext
        self.side = torch.cuda.Stream(self.device) if (self.peer is not None and local_vals.is_cuda) else None
        import os
        self.multi_sweep = os.environ.get("GLAB_DIST_MS", "1") != "0"
        self._entry = 0
        # engine "peer": the one-kernel steps need a 256-row tile of the operator in two shared-memory stages;
        # operators with very wide rows (deep coarse levels) take the separate wait / boundary / push kernels,
        # which speak the same counters, so neighbours may differ in their choice
        esz = torch.empty(0, dtype=self.dtype).element_size()
        self.fused = engine == "peer" and self.square and bool(lib.glab_halo_fits(self.plan.handle, k, esz))

    def entry(self):
        """Name of the vector the next layer call loads its input into: "v0" and "v1" alternate.
        A step that only READS a gathered vector (residual, Rayleigh quotient, the message column)
        pushes nothing back, so the neighbours get no signal that this rank is done with that
        vector's halo tail.  With alternating entry vectors a neighbour overwrites a tail only two
        layer calls later, and it cannot get there without having received this rank's publish of the
        call in between -- which is stream-ordered behind every kernel of the earlier call."""
        self._entry ^= 1
        return self.ENTRY[self._entry]

    # -- halo plumbing -----------------------------------------------------------------------
    def publish(self, name):
        """Make the local rows of vector `name` visible in the neighbours' halo tails."""
        if self.halo.part.world == 1:
            return
        if self.peer is not None:
            self.peer.push(name)
        else:
            self.halo.exchange(self.vec[name], self.group)

    def acquire(self, name):
        if self.peer is not None and self.halo.part.world > 1:
            self.peer.wait(name)

    def _ranges(self):
        """(interior range, [boundary ranges])"""
        out = []
        if self.lo > 0:
            out.append((0, self.lo))
        if self.hi < self.n_local:
            out.append((self.hi, self.n_local))
        return (self.lo, self.hi), out

    def run_step(self, name_in, launch, name_out=None):
        """launch(rows=None, halo=None) issues the fused kernel; the gathered vector is `name_in`
        (already published by the producer of its values).  If `name_out` is given, the vector of
        that name -- which the step writes -- is published to the neighbours.

        engine "peer"       ONE kernel: interior tiles first, boundary tiles after an in-kernel
                            acquire of the neighbours' arrival counters, halo push by the grid's
                            last CTA (glab_*_halo_*).
        engine "peer-split" separate kernels with fork/join over two streams: interior rows on the
                            caller's stream; wait kernel, boundary rows and push kernel on a side
                            stream (kept as the comparison point and for operators that do not fit
                            the pipeline).
        engine "torch"      whole block, then isend/irecv."""
        interior, boundary = self._ranges()
        if self.halo.part.world == 1 or self.peer is None:
            launch(rows=(0, self.n_local))
            if name_out is not None:
                self.publish(name_out)
            return
        if self.fused:
            launch(halo=self.peer.step(name_in, name_out, interior))
            return
        main = torch.cuda.current_stream(self.device)
        fork = torch.cuda.Event()
        join = torch.cuda.Event()
        fork.record(main)
        with torch.cuda.stream(self.side):
            self.side.wait_event(fork)
            self.acquire(name_in)
            for rng in boundary:
                launch(rows=rng)
            if name_out is not None:
                self.publish(name_out)
            join.record(self.side)
        if interior[1] > interior[0]:
            launch(rows=interior)
        main.wait_event(join)

    # -- fused layer steps -------------------------------------------------------------------
    def load(self, name, x_local):
        self.vec[name][:self.n_local].copy_(x_local.view(self.n_local, self.k))
        self.publish(name)

    def local(self, name):
        """The locally owned rows of gathered vector `name` (a view into the peer buffer)."""
        return self.vec[name][:self.n_local]

    def apply_rect(self, name_in, out, add_to=None):
        """out = A x (or add_to + A x) for a gathered vector that was publish()-ed by its producer;
        works for rectangular row blocks.  Stand-alone wait + one whole-block kernel."""
        self.acquire(name_in)
        xin = self.vec[name_in]
        if add_to is None:
            rt.spmm(self.plan, self.vals, xin, out)
        else:
            rt.spmm_add(self.plan, self.vals, xin, add_to, out)
        return out

    def jacobi(self, n_iters, diag, b, omega_dev, start="v0"):
        """n_iters sweeps starting from vector `start` (already load()-ed / published), ping-ponging
        through "va"/"vb" (never overwriting `start` if it is "v0"); returns the name of the
        vector holding the result."""
        cur = start
        todo = n_iters
        if self.fused and self.halo.part.world > 1 and self.multi_sweep and n_iters > 1:
            # One launch for all sweeps (glab_jacobi_sweeps_halo_*).  The multi-sweep kernel overwrites both
            # of its buffers, so a start vector that must stay intact ("v0") takes one ordinary sweep first.
            if cur in self.ENTRY:
                xin, xout, first = self.vec[cur], self.vec["va"], cur
                self.run_step(first, lambda **kw: rt.jacobi(self.plan, self.vals, diag, b, xin, xout, omega_dev, **kw),
                              "va")
                cur, todo = "va", todo - 1
            if todo > 1:
                oth = "vb" if cur == "va" else "va"
                interior = self._ranges()[0]
                steps = (self.peer.step(cur, oth, interior), self.peer.step(oth, cur, interior))
                res = rt.jacobi_sweeps(self.plan, self.vals, diag, b, self.vec[cur], self.vec[oth], omega_dev, todo,
                                       halo=steps)
                return oth if res is self.vec[oth] else cur
        for _ in range(todo):
            nxt = "va" if cur != "va" else "vb"
            xin, xout = self.vec[cur], self.vec[nxt]
            self.run_step(cur, lambda **kw: rt.jacobi(self.plan, self.vals, diag, b, xin, xout, omega_dev, **kw),
                          nxt)
            cur = nxt
        return cur

    def spmv(self, name_in, out, b=None):
        xin = self.vec[name_in]
        if b is None:
            self.run_step(name_in, lambda **kw: rt.spmm(self.plan, self.vals, xin, out, **kw))
        else:
            self.run_step(name_in, lambda **kw: rt.residual(self.plan, self.vals, xin, b, out, **kw))
        return out

    def chebyshev(self, deg, b, table, start="va", x=None, r=None):
        """Chebyshev relaxation of degree `deg` from vector `start` ("va" or "vb").  x lives in its
        own local buffer after iteration 1 (only p is gathered), so the named peer vectors
        ping-pong p.  Returns (x, r, name of p)."""
        n, k = self.n_local, self.k
        # p ping-pongs between two named vectors, never `start` if that is "v0" (kept intact so
        # that the same start vector can be reused by the next call)
        other = "va" if start != "va" else "vb"
        second = "vb" if start in self.ENTRY else start
        xin = self.vec[start]
        x = torch.empty(n, k, dtype=self.dtype, device=self.device) if x is None else x
        r = torch.empty_like(x) if r is None else r
        pv = self.vec[other]
        self.run_step(start, lambda **kw: rt.cheby_first(self.plan, self.vals, b, xin, x, r, pv, table[0, 1:2], **kw),
                      other)
        cur, nxt = other, second
        self.last_gathered = start
        for it in range(1, deg):
            pin, pout = self.vec[cur], self.vec[nxt]
            self.run_step(cur, lambda **kw: rt.cheby_next(self.plan, self.vals, pin, pout, r, x,
                                                          table[it, 0:1], table[it, 1:2], table[it, 2:3], **kw), nxt)
            self.last_gathered = cur
            cur, nxt = nxt, cur
        return x, r, cur

    def power_method(self, num_iter, start="v0"):
        """Power iteration + Rayleigh quotient on the partitioned operator.  Engine "peer": the squared
        norm of every iterate is summed over the ranks INSIDE the step kernels (glab_peer_reduce:
        mailboxes in peer memory, rank-ordered sum, identical on every rank) -- one NCCL all-reduce in
        total, for the three reported numbers.  Other engines: one all-reduce (2 fp64 scalars) per
        iteration.  Returns (lambda, n, n_A) as a device fp64 tensor [3] and the last b, y blocks."""
        if self.halo.part.world > 1 and self.fused and os.environ.get("GLAB_DIST_FUSED_REDUCE", "1") != "0":
            return self._power_method_fused(num_iter, start)
        n = self.n_local
        cur = start
        sums = torch.zeros(2 * (num_iter + 2), dtype=torch.float64, device=self.device)
        prev = None
        multi = self.halo.part.world > 1
        for it in range(num_iter):
            nxt = "va" if cur != "va" else "vb"
            ss = sums[2 * it:2 * it + 2]
            bin_, yout = self.vec[cur], self.vec[nxt]
            part = torch.zeros(2, dtype=torch.float64, device=self.device)
            if multi and self.fused:
                rt.power_step(self.plan, self.vals, bin_, yout, prev, part,
                              halo=self.peer.step(cur, nxt, self._ranges()[0]))
            else:
                self._reduced_step(cur, lambda rows, acc: rt.power_step(self.plan, self.vals, bin_, yout, prev, acc,
                                                                        rows), part)
                self.publish(nxt)
            if multi:
                dist.all_reduce(part, group=self.group)
            ss.copy_(part)
            cur = nxt
            prev = ss
        bout = torch.empty(n, self.k, dtype=self.dtype, device=self.device)
        yout = torch.empty_like(bout)
        part = torch.zeros(2, dtype=torch.float64, device=self.device)
        bin_ = self.vec[cur]
        if multi and self.fused:
            rt.rayleigh(self.plan, self.vals, bin_, bout, yout, prev, part,
                        halo=self.peer.step(cur, None, self._ranges()[0]))
        else:
            self._reduced_step(cur, lambda rows, acc: rt.rayleigh(self.plan, self.vals, bin_, bout, yout, prev, acc,
                                                                  rows), part)
        if multi:
            dist.all_reduce(part, group=self.group)
        norm = torch.sqrt(prev[0]) if prev is not None else torch.ones((), dtype=torch.float64, device=self.device)
        return torch.stack([part[0] / part[1], norm, part[0]]), bout, yout

    def _power_method_fused(self, num_iter, start):
        n, rng = self.n_local, self._ranges()[0]
        cur = start
        part = torch.zeros(2 * (num_iter + 1), dtype=torch.float64, device=self.device)   # rank-local sums
        prev = None
        for it in range(num_iter):
            nxt = "va" if cur != "va" else "vb"
            rt.power_step(self.plan, self.vals, self.vec[cur], self.vec[nxt], prev, part[2 * it:2 * it + 2],
                          halo=self.peer.step(cur, nxt, rng, reduce=True))
            cur, prev = nxt, part[2 * it:2 * it + 2]
        bout = torch.empty(n, self.k, dtype=self.dtype, device=self.device)
        yout = torch.empty_like(bout)
        last = part[2 * num_iter:2 * num_iter + 2]
        rt.rayleigh(self.plan, self.vals, self.vec[cur], bout, yout, prev, last,
                    halo=self.peer.step(cur, None, rng, reduce=True))
        tot = torch.stack([last[0], last[1], prev[0] if prev is not None else last[1] * 0 + 1.0 / self.halo.part.world])
        dist.all_reduce(tot, group=self.group)
        return torch.stack([tot[0] / tot[1], torch.sqrt(tot[2]), tot[0]]), bout, yout

    def _reduced_step(self, name_in, launch, total):
        """Reducing kernels write their partial sums per launch; ranges are accumulated."""
        interior, boundary = self._ranges()
        if self.halo.part.world == 1 or self.peer is None:
            launch((0, self.n_local), total)
            return
        tmp = torch.zeros(2 * (1 + len(boundary)), dtype=torch.float64, device=self.device)
        launch(interior, tmp[0:2])
        self.acquire(name_in)
        for i, rng in enumerate(boundary):
            launch(rng, tmp[2 * i + 2:2 * i + 4])
        total.copy_(tmp.view(-1, 2).sum(0))

    def check(self):
        """Raise if an in-kernel wait of a fused step gave up (bounded spins); synchronises."""
        if self.peer is not None:
            self.peer.check()

    def close(self):
        if self.peer is not None:
            self.peer.close()


# ---------------------------------------------------------------------------- layer-level handle
class PartitionedGraph:
    """The row-partitioned counterpart of `edgeij_pair` for the drop-in layers.

    One process per GPU; every rank holds the edges of ITS contiguous row block with GLOBAL indices,
    exactly the rows [offsets[rank], offsets[rank+1]) of the reference's `edgeij_pair`.  Passing this
    handle where a layer expects `edgeij_pair`

        pg = glab_b200.dist.PartitionedGraph(edgeij_pair_of_my_rows, n_global)
        x_local = JacobiGNN()(n_iters, vertex_attr_of_my_rows, pg, edge_attr_of_my_rows, g)

    runs the same fused layer steps on the row block, with the halo rows of every gathered vector
    pushed into the neighbours' memory from inside the kernels (NVLink peer memory; engine "torch"
    falls back to isend/irecv).  vertex_attr / edge_attr / results are this rank's rows / edges; the
    column layouts and return values are those of the single-GPU layers.  Construction and the first
    use with a new edge_attr are collective (all ranks, same order)."""

    def __init__(self, edge_index, n_global, part=None, rank=None, world=None, group=None, engine=None,
                 align=256):
        if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.shape[0] != 2:
            raise GlabError("edgeij_pair must be an int64 [2, nnz] tensor")
        inited = dist.is_initialized()
        self.group = group
        self.rank = (dist.get_rank(group) if inited else 0) if rank is None else rank
        self.world = (dist.get_world_size(group) if inited else 1) if world is None else world
        self.part = RowPartition(n_global, self.world, align=align) if part is None else part
        self.n_global = n_global
        self.r0, self.r1 = self.part.bounds(self.rank)
        self.device = rt.compute_device(edge_index)
        ei = rt.to_device(edge_index, self.device)
        if ei.numel() and (int(ei[0].min()) < self.r0 or int(ei[0].max()) >= self.r1):
            raise GlabError("rank %d owns rows [%d, %d); edgeij_pair holds other rows" % (self.rank, self.r0, self.r1))
        self.halo = HaloPlan.build(self.part, self.rank, ei[1], group)
        self.local_edge_index = torch.stack([ei[0] - self.r0, self.halo.local_columns(ei[1])]).contiguous()
        self.n_local = self.halo.n_local
        self.nnz_local = int(ei.shape[1])
        if engine is None:
            import os
            engine = os.environ.get("GLAB_DIST_ENGINE", "peer" if self.device.type == "cuda" else "torch")
        self.engine = engine
        self._ops = rt._Cache(4)

    def operator(self, edge_attr, k, dtype, col=0):
        """The DistOperator for these values (cached per edge_attr tensor, k and dtype)."""
        extra = (k, dtype, col)
        hit = self._ops.get(edge_attr, extra)
        if hit is not None:
            return hit
        ea = rt.to_device(edge_attr, self.device)
        ea = ea.view(-1, 1) if ea.dim() == 1 else ea
        vals = ea[:, col].to(dtype).contiguous()
        op = DistOperator(self.local_edge_index, vals, self.halo, k=k, engine=self.engine, group=self.group)
        self._ops.put(edge_attr, op, extra)
        return op


def is_partitioned(edgeij_pair):
    return isinstance(edgeij_pair, PartitionedGraph)
