#!/usr/bin/env python
"""Diagnostic for DESIGN.md section 8 item 1: why a fused halo sweep (glab_jacobi_halo_*) on half of
an operator is not faster with 2-byte column indices although the plain kernel is.

Runs on ONE GPU: the operator is split into two row blocks that live in the same process (the
emulation of tests/test_parity_gpu.py::test_fused_halo_step_two_ranks_emulated_on_one_gpu); the two
blocks' halo kernels run back to back on one stream and push into each other's halo tails, so a
"sweep" is two launches.  Printed per configuration (GLAB_IDX16 = 0 / 2 / 3): time per halo launch,
time per plain launch on the same row block (vectors in torch memory and in CUDA-IPC buffers), and
the bytes each moves.  Use it under ncu to capture the HALO instantiation of k_row_pipe, which the
real multi-rank run cannot be (in-kernel flag waits):

    python scripts/diag_halo_sweep.py [--grid 4096] [--sweeps 50]
    ncu --set full --clock-control none -k regex:k_row_pipe --launch-skip 40 -c 4 -o gpurun_out/prof_halo \\
        python scripts/diag_halo_sweep.py --sweeps 30 --modes 3
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import glab_b200 as G  # noqa: E402
from glab_b200 import dist as gd  # noqa: E402
from glab_b200._lib import HaloStep, PushDesc  # noqa: E402

rt = G.runtime


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def run_mode(mode, N, sweeps, dev):
    # mode: "0" = int32 plans, "2h0" = 16-bit plans but the fused halo kernels stay on int32, "2" = both on 16-bit
    os.environ["GLAB_IDX16"] = "0" if mode == "0" else "2"
    os.environ["GLAB_IDX16_HALO"] = "0" if mode == "2h0" else "1"
    rt.clear_caches()
    dt, world = torch.float32, 2
    n = N * N
    ei, ev = G.generators.laplacian_2d(N, dt, dev)
    ev = ev.contiguous()
    torch.manual_seed(24601)
    b = torch.rand(n, 1, device=dev)
    x0 = torch.rand(n, 1, device=dev)
    diag = torch.full((n,), -4.0, device=dev)
    w = torch.tensor([0.7], device=dev)
    part = gd.RowPartition(n, world, align=256)
    blocks = []
    for r in range(world):
        r0, r1 = part.bounds(r)
        mine = (ei[0] >= r0) & (ei[0] < r1)
        blocks.append((ei[0][mine] - r0, ei[1][mine], ev[mine].contiguous()))
    halos = gd.HaloPlan.build_all(part, [blk[1] for blk in blocks])
    ops = []
    for r in range(world):
        rows, gcols, v = blocks[r]
        h = halos[r]
        lei = torch.stack([rows, h.local_columns(gcols)]).contiguous()
        p = G.Plan.from_coo(lei, h.n_local, h.n_local + h.n_halo)
        lo, hi = h.interior_rows(lei[0], lei[1])
        ext = h.n_local + h.n_halo
        bufs = [gd.PeerBuffer(ext, dt, dev) for _ in range(2)]           # CUDA-IPC capable allocations
        vec = [bf.local.view(ext, 1) for bf in bufs]
        for t in vec:
            t.zero_()
        flags = torch.zeros(64, dtype=torch.int32, device=dev)
        r0, r1 = part.bounds(r)
        ops.append(dict(plan=p, vals=rt.get_vals(p, v), halo=h, vec=vec, bufs=bufs, flags=flags, lo=lo, hi=hi, keep=lei,
                        diag=diag[r0:r1].contiguous(), b=b[r0:r1].contiguous()))
        vec[0][:r1 - r0].copy_(x0[r0:r1])

    def word(t, i):
        return t.data_ptr() + 16 * i

    def push_descs(r, v):
        h = ops[r]["halo"]
        descs = (PushDesc * max(len(h.peers_send), 1))()
        for i, q in enumerate(h.peers_send):
            idx = h.send_rows[q]
            first = int(idx[0].item())
            descs[i].send_idx = idx.data_ptr()
            descs[i].first_row = first
            descs[i].count = idx.numel()
            descs[i].dst = ops[q]["vec"][v].data_ptr()
            descs[i].dst_offset = ops[q]["halo"].n_local + ops[q]["halo"].recv_offsets[r]
            descs[i].flag = word(ops[q]["flags"], v)
        return descs

    keep = []
    for r in range(world):
        d = push_descs(r, 0)
        keep.append(d)
        rt._call("halo_push", dt, dev, rt.ptr(ops[r]["vec"][0]), 1, len(ops[r]["halo"].peers_send), d,
                 ctypes.c_void_p(word(ops[r]["flags"], 2)), rt.stream_ptr())
    steps = {}
    for cur in (0, 1):
        nxt = 1 - cur
        for r in range(world):
            o = ops[r]
            h = o["halo"]
            st = HaloStep()
            st.interior_begin, st.interior_end = o["lo"], o["hi"]
            fl = (ctypes.c_void_p * 1)(word(o["flags"], cur))
            st.n_wait = len(h.peers_recv)
            st.wait_flags = fl
            st.wait_target = word(o["flags"], 2 + cur)
            d = push_descs(r, nxt)
            st.n_push = len(h.peers_send)
            st.push = d
            st.pushed_counter = word(o["flags"], 2 + nxt)
            st.push_src = o["vec"][nxt].data_ptr()
            st.done_counter = word(o["flags"], 7)
            keep += [fl, d, st]
            steps[(cur, r)] = st
    state = {"cur": 0}

    def halo_sweep():
        cur = state["cur"]
        nxt = 1 - cur
        for r in range(world):
            o = ops[r]
            rt.jacobi(o["plan"], o["vals"], o["diag"], o["b"], o["vec"][cur], o["vec"][nxt], w, halo=steps[(cur, r)])
        state["cur"] = nxt

    t_halo = timed(halo_sweep, sweeps) / world
    # plain kernel on rank 1's block (mixed plan: its first tiles read the halo tail): vectors in the
    # IPC buffers, then in torch memory
    o = ops[1]
    t_plain_ipc = timed(lambda: rt.jacobi(o["plan"], o["vals"], o["diag"], o["b"], o["vec"][0], o["vec"][1], w), sweeps)
    xa, xb = o["vec"][0].clone(), torch.empty_like(o["vec"][1])
    t_plain = timed(lambda: rt.jacobi(o["plan"], o["vals"], o["diag"], o["b"], xa, xb, w), sweeps)
    nl, z = o["halo"].n_local, o["plan"].nnz
    alg = z * 8 + 4 * (nl + 1) + 4 * nl * 4
    res = {"GLAB_IDX16": mode, "grid": N, "rows_per_block": nl, "nnz_per_block": z,
           "index16_tiles": [[q["plan"].index16_tiles, q["plan"].tiles] for q in ops],
           "halo_kernel_ms": t_halo, "plain_kernel_ipc_vectors_ms": t_plain_ipc, "plain_kernel_torch_vectors_ms": t_plain,
           "algorithmic_MB": alg / 1e6, "halo_GBps_algorithmic": alg / t_halo / 1e6,
           "plain_GBps_algorithmic": alg / t_plain / 1e6}
    for q in ops:
        for bf in q["bufs"]:
            bf.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=4096)
    ap.add_argument("--sweeps", type=int, default=50)
    ap.add_argument("--modes", default="0,2h0,2")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    for mode in args.modes.split(","):
        print(json.dumps(run_mode(mode, args.grid, args.sweeps, dev)), flush=True)
    os.environ.pop("GLAB_IDX16", None)
    os.environ.pop("GLAB_IDX16_HALO", None)


if __name__ == "__main__":
    main()
