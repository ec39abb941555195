"""Row-partitioned multilevel cycle (BASELINE config 5: "VCycle multilevel (Jacobi smoother +
interpolation) on 67M-row Laplacian, 8-column SpMM RHS batch, 8xB200 with halo exchange").

The hierarchy is the single-GPU one of multilevel.py (built REPLICATED on every rank: it fits one
B200 up to the 67 M-row Laplacian; setup, once per operator).  The cycle is DISTRIBUTED on the large
levels and replicated on the small ones:

  levels with >= `replicate_below` rows   rows of A_l / P_l / P_l^T partitioned in contiguous blocks; a
                                          coarse point lives with the root of its aggregate, so every
                                          level's partition follows the finest one.  Smoothing and
                                          residual are the fused one-kernel-per-step halo steps
                                          (multi-sweep Jacobi where it pays), restriction and
                                          prolongation gather across the partition boundary;
  the first level below that size         its right-hand side is all-gathered (NCCL; a few MB) and every
                                          rank runs the rest of the cycle redundantly on the whole
                                          vectors with the single-GPU kernels -- coarse levels are
                                          latency-bound, and 7 more exchanges per visit would cost more
                                          than the redundant arithmetic;  the prolongation back needs no
                                          exchange (every rank holds the whole coarse correction).

Every kernel performs the same row-local arithmetic in the same order as on one GPU, so the result is
bit-identical to VCycle.runVCycleML (bench.py's parity block and tests/dist_gpu_check.py check it).
"""
import torch
import torch.distributed as dist

from . import _runtime as rt
from .dist import DistOperator, RowPartition, partition_coo
from .multilevel import Hierarchy


class _DLevel:
    pass


class DistMultilevel:
    def __init__(self, A, k, rank, world, engine="peer", group=None, n_pre=3, n_post=3, gamma=2,
                 replicate_below=1 << 20, **options):
        import time
        self.rank, self.world, self.k, self.group = rank, world, k, group
        self.n_pre, self.n_post, self.gamma = n_pre, n_post, gamma
        t0 = time.perf_counter()
        self.h = Hierarchy(A, **options)                    # replicated setup
        torch.cuda.synchronize()
        self.setup_times = {"replicated_hierarchy": time.perf_counter() - t0}
        t1 = time.perf_counter()
        lv = self.h.levels
        dev, dt = lv[0].device, lv[0].dtype
        self.dtype, self.device = dt, dev
        # number of partitioned levels: those with at least replicate_below rows (never the coarsest)
        self.n_part = 0
        while self.n_part < len(lv) - 1 and lv[self.n_part].n >= replicate_below:
            self.n_part += 1
        if self.n_part == 0 and len(lv) > 1:
            self.n_part = 1                                 # always distribute the finest level
        parts = [RowPartition(lv[0].n, world, align=256)]
        for l in range(self.n_part):
            lev = lv[l]
            # coarse point = aggregate id; roots are numbered in index order, so the coarse rows owned by a
            # rank are those whose root lies in its fine block (left-over singleton aggregates: the last rank)
            nroot = torch.cumsum(lev.root.to(torch.int64), 0)
            offs = [0 if int(o) == 0 else int(nroot[int(o) - 1].item()) for o in parts[l].offsets.tolist()]
            offs[-1] = lev.n_agg
            parts.append(RowPartition(lev.n_agg, world, offsets=offs))
        self.parts = parts
        self.lev = []
        for l in range(self.n_part):
            src, D = lv[l], _DLevel()
            f0, f1 = parts[l].bounds(rank)
            c0, c1 = parts[l + 1].bounds(rank)
            ai, av, ah = partition_coo(src.edge_index, src.edge_val, parts[l], rank, group)
            D.A = DistOperator(ai, av.contiguous(), ah, k=k, engine=engine, group=group)
            D.diag = src.diag[f0:f1].clone()          # own allocation: the kernels need 16-byte aligned streams
            D.w = src.w
            D.nf, D.nc = f1 - f0, c1 - c0
            pidx, pval = src.P_index, src.P_vals.reshape(-1, 1)
            tidx = torch.stack([pidx[1], pidx[0]])
            ti, tv, th = partition_coo(tidx, pval, parts[l + 1], rank, group, col_part=parts[l])
            D.PT = DistOperator(ti, tv.contiguous(), th, k=k, engine=engine, group=group, n_rows=D.nc, names=("g",))
            mine = (pidx[0] >= f0) & (pidx[0] < f1)
            if l + 1 < self.n_part:
                pi, pv, ph = partition_coo(pidx, pval, parts[l], rank, group, col_part=parts[l + 1])
                D.P = DistOperator(pi, pv.contiguous(), ph, k=k, engine=engine, group=group, n_rows=D.nf, names=("g",))
            else:
                # the level below is replicated: prolongation gathers from the WHOLE coarse vector, no exchange
                gi = torch.stack([pidx[0][mine] - f0, pidx[1][mine]]).contiguous()
                D.P = None
                D.P_plan = rt.Plan.from_coo(gi, D.nf, src.n_agg)
                D.P_vals = rt.get_vals(D.P_plan, pval[mine].contiguous())
                D._keep = gi
            D.b = torch.empty(D.nf, k, dtype=dt, device=dev)
            D.cur = "v0"
            self.lev.append(D)
        # replicated tail
        self.rep_bufs = self.h._bufs(k)
        nrep = lv[self.n_part].n
        cl = [parts[self.n_part].bounds(q) for q in range(world)]
        self.rep_counts = [b_ - a_ for a_, b_ in cl]
        self.rep_pad = max(self.rep_counts) if self.rep_counts else 0
        self.rep_gather = torch.empty(world * max(self.rep_pad, 1), k, dtype=dt, device=dev)
        self.rep_send = torch.zeros(max(self.rep_pad, 1), k, dtype=dt, device=dev)
        self.nrep = nrep
        self.zero = [torch.zeros(D.nf, k, dtype=dt, device=dev) for D in self.lev]
        torch.cuda.synchronize()
        self.setup_times["partition_levels"] = time.perf_counter() - t1
        if world > 1:
            dist.barrier(group=group)

    # ------------------------------------------------------------------ data in / out
    def load_x(self, x_local):
        self.lev[0].A.load("v0", x_local.to(self.dtype))
        self.lev[0].cur = "v0"

    def x_local(self):
        return self.lev[0].A.local(self.lev[0].cur)

    def residual_local(self, b_local, out=None):
        D = self.lev[0]
        out = torch.empty(D.nf, self.k, dtype=self.dtype, device=self.device) if out is None else out
        return D.A.spmv(D.cur, out, b=b_local)

    # ------------------------------------------------------------------ the cycle
    def _replicated(self):
        """Right-hand side slabs of the first replicated level -> every rank; run the rest of the cycle."""
        l0 = self.n_part
        bufs = self.rep_bufs
        if self.world > 1:
            dist.all_gather_into_tensor(self.rep_gather, self.rep_send, group=self.group)
            off = 0
            for q, cnt in enumerate(self.rep_counts):
                bufs[l0]["b"][off:off + cnt].copy_(self.rep_gather[q * self.rep_pad:q * self.rep_pad + cnt])
                off += cnt
        else:
            bufs[l0]["b"].copy_(self.rep_send[:self.nrep])
        return self.h.run_from(l0, self.k, self.n_pre, self.n_post, self.gamma, first_zero=True)

    def _visit(self, l, b_local, zero_guess):
        D = self.lev[l]
        A = D.A
        if zero_guess:
            A.load("v0", self.zero[l])
            D.cur = "v0"
        cur = A.jacobi(self.n_pre, D.diag, b_local, D.w, D.cur)
        A.spmv(cur, D.PT.local("g"), b=b_local)                       # r = b - A x into the restriction's input
        D.PT.publish("g")
        if l + 1 < self.n_part:
            nxt = self.lev[l + 1]
            D.PT.apply_rect("g", nxt.b)                               # b_{l+1} = P^T r
            for g in range(self.gamma):
                self._visit(l + 1, nxt.b, zero_guess=(g == 0))
            D.P.local("g").copy_(nxt.A.local(nxt.cur))
            D.P.publish("g")
            xl = A.local(cur)
            D.P.apply_rect("g", xl, add_to=xl)                        # x += P x_c
        else:
            D.PT.apply_rect("g", self.rep_send[:D.nc])
            xc = self._replicated()
            xl = A.local(cur)
            rt.spmm_add(D.P_plan, D.P_vals, xc, xl, xl)
        A.publish(cur)
        D.cur = A.jacobi(self.n_post, D.diag, b_local, D.w, cur)

    def cycle(self, b_local):
        self._visit(0, b_local, zero_guess=False)
        return self.x_local()

    def info(self):
        out = self.h.info(self.n_pre, self.n_post, self.gamma)
        out["partitioned_levels"] = self.n_part
        out["replicated_from_rows"] = self.h.levels[self.n_part].n
        return out

    def check(self):
        for D in self.lev:
            for o in (D.A, D.PT, D.P):
                if o is not None:
                    o.check()

    def close(self):
        for D in self.lev:
            for o in (D.A, D.PT, D.P):
                if o is not None:
                    o.close()
