#!/usr/bin/env python
"""Debug helper (torchrun, 2+ ranks): GNNResidual on a PartitionedGraph vs the unpartitioned layer,
optionally preceded by the Jacobi / Chebyshev calls of tests/dist_gpu_check.py."""
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
    mode = sys.argv[1] if len(sys.argv) > 1 else "res"
    dt, N = torch.float32, 200
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
    jac, cheb, res = G.JacobiGNN.JacobiGNN(), G.ChebyGNN.ChebyRelaxGNN(3), G.GNNResidual.GNNResidual()
    part = gd.RowPartition(n, world, align=256)
    r0, r1 = part.bounds(rank)
    mine = (ei[0] >= r0) & (ei[0] < r1)
    pg = gd.PartitionedGraph(ei[:, mine].contiguous(), n, part, rank, world)
    ea2_l = ea2[mine].contiguous()
    x_in = x0
    if "jac" in mode:
        x_ref = jac(7, torch.cat([diag, b, x0], 1), ei, ea2, gw)
        x_l = jac(7, torch.cat([diag, b, x0], 1)[r0:r1].contiguous(), pg, ea2_l, gw)
        print("rank %d jacobi %s" % (rank, torch.equal(x_l, x_ref[r0:r1])), flush=True)
        x_in = x_ref
    if "cheb" in mode:
        v_ref, _, _ = cheb(torch.cat([b, x_in], 1), ei, ev, gc)
        v_l, _, _ = cheb(torch.cat([b, x_in], 1)[r0:r1].contiguous(), pg, ea2_l, gc)
        print("rank %d cheb %s" % (rank, torch.equal(v_l, v_ref[r0:r1])), flush=True)
    for rep in range(3):
        r_ref = res(torch.cat([b, x_in], 1), ei, ev)
        r_l = res(torch.cat([b, x_in], 1)[r0:r1].contiguous(), pg, ea2_l)
        torch.cuda.synchronize()
        bad = torch.nonzero((r_l != r_ref[r0:r1]).reshape(-1)).reshape(-1)
        print("rank %d mode %s rep %d residual equal %s bad rows %d %s" % (rank, mode, rep, bad.numel() == 0, bad.numel(),
                                                                          bad[:3].tolist()), flush=True)
        x_in = x_in + 0.125          # new values every repetition
        dist.barrier()
    for op_ in list(pg._ops.d.values()):
        op_[1].close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
