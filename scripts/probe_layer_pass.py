#!/usr/bin/env python
"""Where does a drop-in layer pass (JacobiGNN(10) + ChebyRelaxGNN(4) through the layer API, device-resident
vertex_attr) spend its device time?  Lists every kernel of one pass with torch.profiler (CUPTI) and the
event-timed pieces.  Single GPU."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import glab_b200 as G
    from bench_support import SingleGpuSmoother
    dev = torch.device("cuda", 0)
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    prob = SingleGpuSmoother(G, N, dev)
    va = prob.va_host.to(dev)
    for _ in range(3):
        prob.layer_pass(va)
    torch.cuda.synchronize()

    def timed(fn, reps=5):
        out = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            r = fn()
            b.record()
            torch.cuda.synchronize()
            out.append(a.elapsed_time(b))
        return sorted(out)[len(out) // 2], r

    res = {}
    res["layer_pass_ms"], _ = timed(lambda: prob.layer_pass(va))
    res["jacobi_layer_ms"], x1 = timed(lambda: prob.jac(prob.N_JACOBI, va, prob.ei, prob.ea2, prob.gw))
    res["pack_ms"], vb = timed(lambda: prob.rt.pack([va[:, 1:2].contiguous(), x1]))
    res["cheby_layer_ms"], out = timed(lambda: prob.cheb(vb, prob.ei, prob.ev, prob.gc))
    res["slice_copy_ms"], _ = timed(lambda: out[0][:, 1:2].contiguous())
    res["step_kernels_ms"], _ = timed(lambda: prob.step_kernels())
    print(json.dumps(res))
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        prob.layer_pass(va)
        torch.cuda.synchronize()
    rows = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            rows.append((e.time_range.start, e.name[:90], e.device_time))
    rows.sort()
    t0 = rows[0][0] if rows else 0
    for st, nm, us in rows:
        print("%9.1f us  %8.1f us  %s" % (st - t0, us, nm))


if __name__ == "__main__":
    main()
