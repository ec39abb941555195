"""Multi-GPU parity check, launched with torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_gpu_check.py

Every rank also runs the UNPARTITIONED operator on its own GPU and compares its slab of the
partitioned result with it: Jacobi / Chebyshev / SpMV must agree bit for bit (row-local
arithmetic is identical), the power-method Rayleigh quotient to 1e-6 / 1e-12.
Both exchange engines are exercised: NVLink peer memory (push/wait kernels) and NCCL send/recv.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import glab_b200 as G
    from glab_b200 import dist as gd
    rt = G.runtime
    ok = True
    # N = 512: row blocks of > 32767 rows, so the plans are MIXED (tiles that read the halo tail
    # stream int32 column indices, interior tiles 16-bit row-relative ones)
    for dt, N, gen in ((torch.float32, 96, "lap"), (torch.float64, 64, "heat"), (torch.float32, 512, "lap")):
        # the first case keeps the fused halo kernels on int32 column indices (default: 16-bit where a tile qualifies)
        os.environ["GLAB_IDX16_HALO"] = "0" if N == 96 else "1"
        n = N * N
        if gen == "lap":
            ei, ev = G.generators.laplacian_2d(N, torch.float64, dev)
        else:
            ei, ev = G.generators.heat_fem_2d((N + 1, N + 1), (1.0, 1.0), torch.float64, dev)
        ev = ev.to(dt)
        torch.manual_seed(24601)
        b = torch.rand(n, 1, dtype=dt, device=dev)
        x0 = torch.rand(n, 1, dtype=dt, device=dev)
        diag = G.generators.diagonal_of(ei, ev, n).reshape(-1).contiguous()
        w = torch.tensor([0.7], dtype=dt, device=dev)
        # ---- single-GPU reference on every rank
        plan = G.get_plan(ei, n)
        vals = rt.get_vals(plan, ev)
        xa, xb = x0.clone(), torch.empty_like(x0)
        for _ in range(5):
            rt.jacobi(plan, vals, diag, b, xa, xb, w)
            xa, xb = xb, xa
        jac_ref = xa.clone()
        rows_, _ = G.ChebyGNN._recurrence(4, torch.tensor([-3.4, -4.0]))
        table = torch.stack([torch.stack(r) for r in rows_]).to(device=dev, dtype=dt).contiguous()
        cv, _, _ = G.ChebyGNN.ChebyRelaxGNN(4)(torch.cat([b, jac_ref], 1), ei, ev, torch.tensor([-3.4, -4.0]))
        va = torch.cat([x0, torch.zeros_like(x0)], 1)
        ea = torch.cat([ev, torch.zeros_like(ev)], 1)
        g_ref = G.PowerMethodGNN.PowerMethodGNN(20)(va, ei, ea, torch.zeros(3, dtype=dt), None)[2]
        for engine in ("peer", "peer-split", "torch"):
            part = gd.RowPartition(n, world, align=256)
            r0, r1 = part.bounds(rank)
            lei, lev, halo = gd.partition_coo(ei, ev, part, rank)
            op = gd.DistOperator(lei, lev.contiguous(), halo, k=1, engine=engine)
            t16 = (op.plan.index16_tiles, op.plan.tiles)
            op.load("v0", x0[r0:r1])
            cur = op.jacobi(5, diag[r0:r1].contiguous(), b[r0:r1].contiguous(), w, "v0")
            jac = op.vec[cur][:halo.n_local]
            e1 = torch.equal(jac, jac_ref[r0:r1])
            x, r, pname = op.chebyshev(4, b[r0:r1].contiguous(), table, cur)
            e2 = torch.equal(x, cv[r0:r1, 1:2]) and torch.equal(r, cv[r0:r1, 2:3]) and \
                torch.equal(op.vec[pname][:halo.n_local], cv[r0:r1, 3:4])
            op.load("v0", x0[r0:r1])
            lam, bout, yout = op.power_method(20, "v0")
            tol = 1e-5 if dt == torch.float32 else 1e-12
            e3 = abs(lam[0].item() - g_ref[2].item()) <= tol * abs(g_ref[2].item())
            # again (odd number of publishes in between: the mailbox parity carries over), and every rank must
            # hold the same bits (rank-ordered in-kernel sum / NCCL sum)
            op.load("v0", x0[r0:r1])
            lam2, _, _ = op.power_method(7, "v0")
            op.load("v0", x0[r0:r1])
            lam3, _, _ = op.power_method(20, "v0")
            allr = [torch.zeros_like(lam3) for _ in range(world)]
            dist.all_gather(allr, lam3)
            e3 = e3 and torch.equal(lam3, lam) and all(torch.equal(a, lam3) for a in allr) and \
                bool(torch.isfinite(lam2).all())
            op.check()
            torch.cuda.synchronize()
            print("rank %d %s N=%d engine=%s halo=%d interior=[%d,%d) idx16 tiles %d/%d: jacobi %s cheby %s power %s "
                  "(%.9g vs %.9g)" % (rank, str(dt)[6:], N, engine, halo.n_halo, op.lo, op.hi, t16[0], t16[1], e1, e2, e3,
                                      lam[0].item(), g_ref[2].item()), flush=True)
            ok = ok and e1 and e2 and e3
            op.close()
            dist.barrier()
    os.environ.pop("GLAB_IDX16_HALO", None)
    # ---- a structurally NON-symmetric operator: row i couples to i and to i + N only (an upwind stencil), so
    # rank q reads rows of rank q + 1 but rank q + 1 reads nothing of rank q -- one-directional neighbours: the
    # receiver must still advance its own push counters and the sender gets no data back
    for engine in ("peer", "peer-split", "torch"):
        N = 128
        n = N * N
        dt = torch.float32
        i_ = torch.arange(n, device=dev)
        up = i_[: n - N]
        ei = torch.cat([torch.stack([i_, i_]), torch.stack([up, up + N])], 1)
        order = torch.argsort(ei[0] * n + ei[1])
        ei = ei[:, order].contiguous()
        ev = torch.where(ei[0] == ei[1], torch.tensor(4.0, device=dev), torch.tensor(-1.0, device=dev)).view(-1, 1)
        torch.manual_seed(7)
        b = torch.rand(n, 1, dtype=dt, device=dev)
        x0 = torch.rand(n, 1, dtype=dt, device=dev)
        diag = torch.full((n,), 4.0, dtype=dt, device=dev)
        w = torch.tensor([0.9], dtype=dt, device=dev)
        plan = G.get_plan(ei, n)
        vals = rt.get_vals(plan, ev)
        xa, xb = x0.clone(), torch.empty_like(x0)
        for _ in range(6):
            rt.jacobi(plan, vals, diag, b, xa, xb, w)
            xa, xb = xb, xa
        part = gd.RowPartition(n, world, align=256)
        r0, r1 = part.bounds(rank)
        lei, lev, halo = gd.partition_coo(ei, ev, part, rank)
        op = gd.DistOperator(lei, lev.contiguous(), halo, k=1, engine=engine)
        good = True
        for rep in range(2):                      # twice: counters carry over
            op.load(op.entry(), x0[r0:r1])
            cur = op.jacobi(6, diag[r0:r1].clone(), b[r0:r1].clone(), w, op.ENTRY[op._entry])
            good = good and torch.equal(op.vec[cur][:halo.n_local], xa[r0:r1])
        op.check()
        torch.cuda.synchronize()
        print("rank %d non-symmetric upwind operator engine=%s send->%s recv<-%s: %s" % (
            rank, engine, halo.peers_send, halo.peers_recv, good), flush=True)
        ok = ok and good
        op.close()
        dist.barrier()
    # ---- the drop-in layers on a dist.PartitionedGraph (edgeij_pair of this rank's rows) vs the same layer
    # calls on the whole operator on one GPU: Jacobi / Chebyshev / residual bit for bit, power method to tolerance
    for dt, N in ((torch.float32, 200), (torch.float64, 72)):
        n = N * N
        ei, ev = G.generators.laplacian_2d(N, torch.float64, dev)
        ev = ev.to(dt)
        torch.manual_seed(24601)
        b = torch.rand(n, 1, dtype=dt, device=dev)
        x0 = torch.rand(n, 1, dtype=dt, device=dev)
        diag = G.generators.diagonal_of(ei, ev, n)
        ea2 = torch.cat([ev, torch.zeros_like(ev)], 1)
        gw = torch.tensor([0.7], dtype=dt)
        gc = torch.tensor([-3.4, -4.0])
        jac, cheb, res, pw = G.JacobiGNN.JacobiGNN(), G.ChebyGNN.ChebyRelaxGNN(3), G.GNNResidual.GNNResidual(), \
            G.PowerMethodGNN.PowerMethodGNN(15)
        x_ref = jac(7, torch.cat([diag, b, x0], 1), ei, ea2, gw)
        v_ref, e_ref, g_ref2 = cheb(torch.cat([b, x_ref], 1), ei, ev, gc)
        r_ref = res(torch.cat([b, x_ref], 1), ei, ev)
        pv_ref, pe_ref, pg_ref = pw(torch.cat([x0, torch.zeros_like(x0)], 1), ei, ea2, torch.zeros(3, dtype=dt), None)
        part = gd.RowPartition(n, world, align=256)
        r0, r1 = part.bounds(rank)
        mine = (ei[0] >= r0) & (ei[0] < r1)
        pg = gd.PartitionedGraph(ei[:, mine].contiguous(), n, part, rank, world)
        ea2_l, ev_l = ea2[mine].contiguous(), ev[mine].contiguous()
        x_l = jac(7, torch.cat([diag, b, x0], 1)[r0:r1].contiguous(), pg, ea2_l, gw)
        v_l, e_l, g_l = cheb(torch.cat([b[r0:r1], x_l], 1), pg, ea2_l, gc)
        r_l = res(torch.cat([b[r0:r1], x_l], 1), pg, ea2_l)
        pv_l, pe_l, pg_l = pw(torch.cat([x0, torch.zeros_like(x0)], 1)[r0:r1].contiguous(), pg, ea2_l,
                              torch.zeros(3, dtype=dt), None)
        tol = 1e-5 if dt == torch.float32 else 1e-12
        checks = {"jacobi": torch.equal(x_l, x_ref[r0:r1]), "cheby_v": torch.equal(v_l, v_ref[r0:r1]),
                  "cheby_e": torch.equal(e_l, e_ref[mine]), "cheby_g": torch.equal(g_l, g_ref2),
                  "residual": torch.equal(r_l, r_ref[r0:r1]),
                  "power_lambda": abs(pg_l[2].item() - pg_ref[2].item()) <= tol * abs(pg_ref[2].item()),
                  "power_vec": ((pv_l[:, 0] - pv_ref[r0:r1, 0]).norm() <= 10 * tol * pv_ref[r0:r1, 0].norm().clamp_min(1e-30)).item(),
                  "power_msgs": ((pe_l[:, 1] - pe_ref[mine][:, 1]).norm() <= 10 * tol * pe_ref[mine][:, 1].norm().clamp_min(1e-30)).item()}
        for op_ in list(pg._ops.d.values()):
            op_[1].check()
        torch.cuda.synchronize()
        print("rank %d layers on PartitionedGraph %s N=%d: %s" % (rank, str(dt)[6:], N, checks), flush=True)
        if not checks["residual"]:
            bad = torch.nonzero((r_l != r_ref[r0:r1]).reshape(-1)).reshape(-1)
            print("rank %d residual mismatch: %d rows, first %s last %s of %d local rows, max abs %g" % (
                rank, bad.numel(), bad[:5].tolist(), bad[-5:].tolist(), r1 - r0,
                (r_l - r_ref[r0:r1]).abs().max().item()), flush=True)
        ok = ok and all(checks.values())
        for op_ in list(pg._ops.d.values()):
            op_[1].close()
        dist.barrier()
    # ---- row-partitioned two-grid V-cycle (config 5) vs the single-GPU cycle: bit-identical
    from glab_b200.dist_vcycle import DistTwoGrid
    V = G.VCycle
    for k, N in ((1, 48), (8, 64)):
        n = N * N
        ei, ev = G.UtilsGNN.laplacianfun_torch(N, device=dev)
        A = torch.sparse_coo_tensor(ei, ev.flatten(), dtype=torch.float)
        torch.manual_seed(24601)
        b = torch.rand(n, k, device=dev)
        x0 = torch.rand(n, k, device=dev)
        x_ref = x0
        refs = []
        for _ in range(2):
            x_ref = V.runVCycle(A, b, x_ref, 3, 3, 5, True)
            refs.append(x_ref)
        for engine in ("peer", "torch"):
            tg = DistTwoGrid(ei, ev, k, rank, world, engine=engine)
            f0, f1 = tg.fine.bounds(rank)
            tg.load_x(x0[f0:f1])
            bl = b[f0:f1].contiguous()
            good = True
            for c in range(2):
                xl = tg.cycle(bl)
                good = good and torch.equal(xl, refs[c][f0:f1])
            rl = tg.residual_local(bl)
            r_ref = V.runResidual(A, b, refs[-1])
            good = good and torch.equal(rl, r_ref[f0:f1])
            torch.cuda.synchronize()
            print("rank %d two-grid V-cycle N=%d k=%d engine=%s coarse rows %d: %s" % (rank, N, k, engine, tg.ncl, good),
                  flush=True)
            ok = ok and good
            tg.close()
            dist.barrier()
    # ---- row-partitioned MULTILEVEL cycle vs VCycle.runVCycleML on one GPU: bit-identical
    from glab_b200.dist_multilevel import DistMultilevel
    for k, N, below in ((1, 96, 1500), (8, 128, 4000), (8, 128, 1000), (2, 256, 1000)):
        n = N * N
        ei, ev = G.UtilsGNN.laplacianfun_torch(N, device=dev)
        A = torch.sparse_coo_tensor(ei, ev.flatten().float(), (n, n))
        torch.manual_seed(24601)
        b = torch.rand(n, k, device=dev)
        x_ref = torch.zeros(n, k, device=dev)
        refs = []
        for _ in range(3):
            x_ref = V.runVCycleML(A, b, x_ref, 3, 3, coarsest_n=64)
            refs.append(x_ref)
        ml = DistMultilevel(A, k, rank, world, replicate_below=below, coarsest_n=64)
        f0, f1 = ml.parts[0].bounds(rank)
        ml.load_x(torch.zeros(f1 - f0, k, device=dev))
        bl = b[f0:f1].contiguous()
        good = True
        for c in range(3):
            xl = ml.cycle(bl)
            good = good and torch.equal(xl, refs[c][f0:f1])
        ml.check()
        torch.cuda.synchronize()
        print("rank %d multilevel cycle N=%d k=%d levels %d (partitioned %d, replicated from %d rows): %s" % (
            rank, N, k, len(ml.h.levels), ml.n_part, ml.h.levels[ml.n_part].n, good), flush=True)
        ok = ok and good
        ml.close()
        dist.barrier()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_GPU_CHECK", "PASS" if flag.item() == 1 else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
