#!/usr/bin/env python
"""Times jacobi / cheby_next for a given k and dtype under the current GLAB_CTAS / GLAB_STAGES
environment (the pipeline's tuning knobs are read once per process)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import glab_b200 as G  # noqa: E402

rt = G.runtime
k = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dt = torch.float64 if (len(sys.argv) > 2 and sys.argv[2] == "f64") else torch.float32
N = 4096
dev = torch.device("cuda:0")
ei, ev = G.generators.laplacian_2d(N, torch.float64, dev)
ev = ev.to(dt).contiguous()
n = N * N
plan = G.Plan.from_coo(ei, n)
vals = rt.get_vals(plan, ev)
s = ev.element_size()
z = plan.nnz
x = torch.rand(n, k, dtype=dt, device=dev)
b = torch.rand(n, k, dtype=dt, device=dev)
y = torch.empty_like(x)
r = torch.rand(n, k, dtype=dt, device=dev)
y2 = torch.empty_like(x)
diag = torch.full((n,), -4.0, dtype=dt, device=dev)
w = torch.tensor([0.7], dtype=dt, device=dev)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    c.record()
    torch.cuda.synchronize()
    return a.elapsed_time(c) / reps


base = z * (4 + s) + 4 * (n + 1)
tj = timeit(lambda: rt.jacobi(plan, vals, diag, b, x, y, w))
tc = timeit(lambda: rt.cheby_next(plan, vals, x, y, r, y2, w, w, w))
ts = timeit(lambda: rt.spmm(plan, vals, x, y))
print("CTAS=%s STAGES=%s k=%d %s: spmm %.4f ms (%.0f GB/s)  jacobi %.4f ms (%.0f GB/s)  cheby_next %.4f ms (%.0f GB/s)" % (
    os.environ.get("GLAB_CTAS", "auto"), os.environ.get("GLAB_STAGES", "auto"), k, str(dt)[6:],
    ts, (base + 2 * n * k * s) / ts / 1e6, tj, (base + (3 * k + 1) * n * s) / tj / 1e6, tc, (base + 6 * n * k * s) / tc / 1e6))
