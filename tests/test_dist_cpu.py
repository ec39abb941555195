"""CPU, world_size-2 (and 3) gloo: the host-side logic of the multi-GPU path -- row partition,
halo plan derived from the CSR column indices, column renumbering, send/recv lists and the
portable exchange -- checked against the unpartitioned oracle.  The local arithmetic is done
with the oracle's scatter_sum here (no GPU), so what is under test is exactly the index logic
that the GPU path (DistOperator / PeerHalo) shares."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem(kind):
    from oracle import port
    torch.manual_seed(5)
    if kind == "laplacian":
        N = 12
        ei, ev = port.laplacian_2d(N)
        return N * N, ei, ev
    if kind == "upwind":
        # structurally NON-symmetric: row i couples to i and i + N only, so rank q reads rows of rank q + 1 but
        # sends nothing back -- one-directional neighbours (peers = union of send and receive sides)
        N = 12
        n = N * N
        i_ = torch.arange(n)
        up = i_[: n - N]
        ei = torch.cat([torch.stack([i_, i_]), torch.stack([up, up + N])], 1)
        order = torch.argsort(ei[0] * n + ei[1])
        ei = ei[:, order].contiguous()
        ev = torch.where(ei[0] == ei[1], torch.tensor(4.0, dtype=torch.float64),
                         torch.tensor(-1.0, dtype=torch.float64)).view(-1, 1)
        return n, ei, ev
    # random sparse operator: every rank talks to every other rank
    n, z = 157, 1500
    g = torch.Generator().manual_seed(3)
    rows = torch.randint(0, n, (z,), generator=g)
    cols = torch.randint(0, n, (z,), generator=g)
    order = torch.argsort(rows, stable=True)
    ei = torch.stack([rows[order], cols[order]])
    ev = torch.rand(z, 1, generator=g, dtype=torch.float64)
    return n, ei, ev


def _worker(rank, world, port_no, kind, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import glab_b200  # noqa: F401  (loads the .so; no GPU needed for the partition logic)
        from glab_b200 import dist as gd
        from oracle import port
        n, ei, ev = _problem(kind)
        x = torch.rand(n, 2, generator=torch.Generator().manual_seed(9), dtype=torch.float64)
        y_ref = port.scatter_sum(ev * x[ei[1]], ei[0], n)               # unpartitioned oracle
        part = gd.RowPartition(n, world, align=4)
        r0, r1 = part.bounds(rank)
        lei, lev, halo = gd.partition_coo(ei, ev, part, rank)
        assert halo.n_local == r1 - r0 and lei[0].max() < halo.n_local
        assert lei[1].max() < halo.n_local + halo.n_halo
        owners = part.owner(halo.halo_cols)
        assert torch.all(owners != rank) and torch.all(halo.halo_cols[1:] > halo.halo_cols[:-1])
        # exchange and local product
        x_ext = gd.extend(x[r0:r1].clone(), halo.n_halo)
        halo.exchange(x_ext)
        assert torch.equal(x_ext[halo.n_local:], x[halo.halo_cols])     # tail == owners' rows
        y_loc = port.scatter_sum(lev * x_ext[lei[1]], lei[0], halo.n_local)
        assert torch.equal(y_loc, y_ref[r0:r1])                         # bit-exact (same edge order)
        # interior range: none of its rows reads the halo tail
        lo, hi = halo.interior_rows(lei[0], lei[1], align=4)
        inner = (lei[0] >= lo) & (lei[0] < hi)
        assert torch.all(lei[1][inner] < halo.n_local)
        # a second exchange after a local update (what a Jacobi sweep does)
        x_ext[:halo.n_local] *= 2
        halo.exchange(x_ext)
        assert torch.equal(x_ext[halo.n_local:], 2 * x[halo.halo_cols])
        # rectangular operator (restriction / prolongation): rows and gathered vector live in
        # DIFFERENT partitions
        m = 61
        g2 = torch.Generator().manual_seed(11)
        rr = torch.randint(0, n, (900,), generator=g2)
        cc = torch.randint(0, m, (900,), generator=g2)
        order2 = torch.argsort(rr, stable=True)
        ri = torch.stack([rr[order2], cc[order2]])
        rv = torch.rand(900, 1, generator=g2, dtype=torch.float64)
        xc = torch.rand(m, 2, generator=g2, dtype=torch.float64)
        cpart = gd.RowPartition(m, world, offsets=[0] + [min(m, 7 + (m // world) * (q + 1)) for q in range(world - 1)] + [m])
        li, lv, lh = gd.partition_coo(ri, rv, part, rank, col_part=cpart)
        c0, c1 = cpart.bounds(rank)
        xe = gd.extend(xc[c0:c1].clone(), lh.n_halo)
        lh.exchange(xe)
        yr = port.scatter_sum(lv * xe[li[1]], li[0], r1 - r0)
        assert torch.equal(yr, port.scatter_sum(rv * xc[ri[1]], ri[0], n)[r0:r1])
        # allreduce of partial sums (power-method norms)
        ss = (y_loc ** 2).sum(0)
        dist.all_reduce(ss)
        assert torch.allclose(ss, (y_ref ** 2).sum(0), rtol=1e-13)
        if kind == "upwind":
            # receive only from the next rank, send only to the previous one; both are peers
            assert halo.peers_recv == ([rank + 1] if rank + 1 < world else [])
            assert halo.peers_send == ([rank - 1] if rank > 0 else [])
            assert halo.peers == sorted(set(halo.peers_recv) | set(halo.peers_send))
        q.put((rank, "ok", halo.n_halo, halo.peers_recv))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "fail: %r\n%s" % (e, traceback.format_exc()), 0, []))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,world", [("laplacian", 2), ("laplacian", 3), ("random", 2), ("random", 3),
                                        ("upwind", 2), ("upwind", 3)])
def test_partition_and_halo_exchange(kind, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port_no, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, status, n_halo, peers in sorted(results):
        assert status == "ok", "rank %d: %s" % (rank, status)
        if kind == "laplacian":
            # a slab of grid lines has at most two neighbours and one grid line of halo per side
            assert len(peers) <= 2 and n_halo in (12, 24)


def test_row_partition_owner():
    sys.path.insert(0, ROOT)
    from glab_b200 import dist as gd
    part = gd.RowPartition(1000, 3, align=256)
    assert part.offsets.tolist() == [0, 256, 512, 1000]
    cols = torch.tensor([0, 255, 256, 511, 512, 999])
    assert part.owner(cols).tolist() == [0, 0, 1, 1, 2, 2]
