"""Row-partitioned two-grid V-cycle (BASELINE config 5: Jacobi smoother + direct interpolation,
k right-hand-side columns, multi-GPU with halo exchange).

Same cycle as VCycle.runVCycle (reference VCycle.py:175-237): 3 weighted-Jacobi sweeps, residual,
restriction P^T r, Chebyshev degree-4 coarse "solve" on the Galerkin operator, correction
x + P xc, 3 sweeps.  The AMG setup (classical SOC, direct interpolation, sparse P, P^T A P) is
REPLICATED: every rank runs the single-GPU setup on the whole operator (it fits one B200 up to
the 67 M-row Laplacian) and then keeps only its row blocks of A, P, P^T and A_c.  The solve is
DISTRIBUTED: fine vectors are partitioned like the rows of A, coarse vectors like the coarse
points (a coarse point lives with its fine point, so both partitions are contiguous blocks);
A and A_c steps are the fused one-kernel-per-sweep halo steps, restriction / prolongation gather
across the partition boundary with a stand-alone wait + kernel + push.
Every kernel performs the same row-local arithmetic in the same order as on one GPU, so the
result is bit-identical to VCycle.runVCycle (tests/dist_gpu_check.py).
"""
import torch
import torch.distributed as dist

from . import _runtime as rt
from . import VCycle as V
from .ChebyGNN import _recurrence
from .dist import DistOperator, RowPartition, partition_coo


class DistTwoGrid:
    def __init__(self, edge_index, edge_val, k, rank, world, engine="peer", splitting=None, group=None,
                 n_pre=3, n_post=3):
        import time
        dev = edge_val.device
        n = int(edge_index[0].max().item()) + 1
        t_ = [time.perf_counter()]

        def lap(name):
            torch.cuda.synchronize()
            t_.append(time.perf_counter())
            self.setup_times[name] = t_[-1] - t_[-2]

        self.setup_times = {}
        self.rank, self.world, self.k, self.group = rank, world, k, group
        self.n_pre, self.n_post = n_pre, n_post
        # ---- replicated setup (single-GPU code path, whole operator)
        A = torch.sparse_coo_tensor(edge_index, edge_val.flatten(), (n, n), dtype=torch.float)
        self._A = A
        op = V._operator(A)
        tg = V._two_grid(A, splitting)
        cop = V._operator(tg.Ac)
        lap("replicated_setup_soc_interp_galerkin")
        split = V.default_splitting(n, dev) if splitting is None else splitting.to(dev).reshape(-1)
        coarse = split > 0
        new_id = torch.cumsum(coarse.to(torch.int64), 0)          # number of coarse points with index <= i
        nc = int(new_id[-1].item())
        # ---- partitions: fine rows in 256-aligned blocks; a coarse point lives with its fine point
        self.fine = RowPartition(n, world, align=256)
        offs_c = [0 if int(o) == 0 else int(new_id[int(o) - 1].item()) for o in self.fine.offsets.tolist()]
        self.coarse = RowPartition(nc, world, offsets=offs_c)
        f0, f1 = self.fine.bounds(rank)
        c0, c1 = self.coarse.bounds(rank)
        self.nf, self.ncl = f1 - f0, c1 - c0
        dt = op.edge_attr.dtype
        self.dtype = dt
        # ---- row blocks
        ai, av, ah = partition_coo(op.edge_index, op.edge_attr, self.fine, rank, group)
        lap("partition_A")
        self.A = DistOperator(ai, av.contiguous(), ah, k=k, engine=engine, group=group)
        lap("operator_A")
        ci, cv, ch = partition_coo(cop.edge_index, cop.edge_attr, self.coarse, rank, group)
        self.Ac = DistOperator(ci, cv.contiguous(), ch, k=k, engine=engine, group=group)
        lap("partition_and_operator_Ac")
        P = tg.P
        pidx, pval = P.indices(), P.values().reshape(-1, 1)
        pi, pv, ph = partition_coo(pidx, pval, self.fine, rank, group, col_part=self.coarse)
        self.P = DistOperator(pi, pv.contiguous(), ph, k=k, engine=engine, group=group, n_rows=self.nf, names=("g",))
        tidx = torch.stack([pidx[1], pidx[0]])
        ti, tv, th = partition_coo(tidx, pval, self.coarse, rank, group, col_part=self.fine)
        self.PT = DistOperator(ti, tv.contiguous(), th, k=k, engine=engine, group=group, n_rows=self.ncl, names=("g",))
        lap("partition_and_operators_P_PT")
        self.nnz = {"A": int(op.edge_index.shape[1]), "P": int(pidx.shape[1]), "Ac": int(cop.edge_index.shape[1])}
        # ---- per-rank data of the cycle
        self.diag = op.diag.to(dt).reshape(-1)[f0:f1].clone()
        self.w = torch.tensor(0.7).reshape(-1).to(device=dev, dtype=dt)                  # VCycle.py:195
        rows, _ = _recurrence(V.cheb_deg, torch.tensor([-3.4, -4.0]))                    # VCycle.py:221-222
        self.table = torch.stack([torch.stack(r_) for r_ in rows]).to(device=dev, dtype=dt).contiguous()
        self.rc = torch.empty(self.ncl, k, dtype=dt, device=dev)
        self.xc = torch.empty(self.ncl, k, dtype=dt, device=dev)
        self.rr = torch.empty(self.ncl, k, dtype=dt, device=dev)
        self.Ac.load("v0", torch.zeros(self.ncl, k, dtype=dt, device=dev))              # xc = 0 (VCycle.py:218)
        self.cur = "v0"
        del tg, cop
        if world > 1:
            dist.barrier(group=group)

    def load_x(self, x_local):
        self.A.load("v0", x_local.to(self.dtype))
        self.cur = "v0"

    def x_local(self):
        return self.A.local(self.cur)

    def residual_local(self, b_local, out=None):
        out = torch.empty(self.nf, self.k, dtype=self.dtype, device=b_local.device) if out is None else out
        return self.A.spmv(self.cur, out, b=b_local)

    def cycle(self, b_local):
        A, Ac, P, PT = self.A, self.Ac, self.P, self.PT
        cur = A.jacobi(self.n_pre, self.diag, b_local, self.w, self.cur)                 # :194-196
        # r = b - A x straight into the restriction's gathered vector, then ship its boundary rows
        A.spmv(cur, PT.local("g"), b=b_local)                                            # :212
        PT.publish("g")
        PT.apply_rect("g", self.rc)                                                      # :215  rc = P^T r
        _, _, _ = Ac.chebyshev(V.cheb_deg, self.rc, self.table, "v0", self.xc, self.rr)  # :221-223
        P.local("g").copy_(self.xc)
        P.publish("g")
        xl = A.local(cur)
        P.apply_rect("g", xl, add_to=xl)                                                 # :226  x += P xc
        A.publish(cur)
        self.cur = A.jacobi(self.n_post, self.diag, b_local, self.w, cur)                # :229-231
        return self.x_local()

    def check(self):
        """Raise if a bounded in-kernel wait of any fused step gave up (synchronises)."""
        for o in (self.A, self.Ac, self.P, self.PT):
            o.check()

    def close(self):
        for o in (self.A, self.Ac, self.P, self.PT):
            o.close()
