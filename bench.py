#!/usr/bin/env python
"""bench.py -- fused SpMV-layer throughput (nnz/s) and HBM-roofline fraction on B200.

Workload (BASELINE.json configs[1]): JacobiGNN (10 sweeps, w = 0.7) + ChebyRelaxGNN (degree 4,
c = -3.4, d = -4) smoothing on the 4096 x 4096 5-point Laplacian (16.7 M rows, 83.9 M nnz), fp32.
One "step" = one smoothing pass = 14 fused SpMV-bearing layer launches (10 glab_jacobi +
1 glab_cheby_first + 3 glab_cheby_next).  metric value = 14 * nnz / step time.

  value     kernels only, operator + vectors resident in HBM (C-ABI calls on torch's stream)
  e2e       through the drop-in layer API: per step the vectors are copied from PINNED HOST
            memory to the GPU, JacobiGNN.forward + ChebyRelaxGNN.forward run (including the
            returned edge_attr message column), and the result x is read back to the host
            (three staging buffers, so consecutive steps overlap their PCIe copies with compute).
            The operator (edge list + CSR plan) is step-invariant and stays resident, like
            model weights.
  roofline  dominant kernel = glab_jacobi: algorithmic bytes z(4+s)+4(n+1)+4ns per launch over
            its CUDA-event time, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline / --impl reference
            the reference's own layer code (oracle/ref_loader.py when /root/reference exists,
            else the bit-exact restatement oracle/port.py) on the host cores, on a bounded sample.

N > 1 (torchrun): the same operator row-block partitioned across ranks (strong scaling), halo
rows pushed over NVLink peer memory each sweep.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

GRID = {"L4096": 4096, "L8192": 8192, "L2048": 2048, "L1448": 1448, "L1024": 1024, "L256": 256}
N_JACOBI, CHEB_DEG, OMEGA, CHEB_C, CHEB_D = 10, 4, 0.7, -3.4, -4.0
LAUNCHES_PER_STEP = N_JACOBI + CHEB_DEG


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def committed_traffic():
    """dram bytes per glab_jacobi launch from the committed ncu --set full capture, or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(path):
        return json.load(open(path))
    return None


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions: an NVML polling thread
    (every ~4 ms; samples taken between region_begin() / region_end() count as "under load"),
    with the recipe's `nvidia-smi -lms` loop as the fallback when NVML cannot be opened."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"),
               (0x4, "sw_power_cap"), (0x80, "hw_power_brake_slowdown"))

    def __init__(self, cuda_index):
        self.cuda_index = cuda_index
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        self.smi_index = cuda_index
        if vis:
            try:
                self.smi_index = int(vis.split(",")[cuda_index])
            except (ValueError, IndexError):
                pass
        self.proc = self.path = self.thread = self.nvml = None
        self.samples = []          # (sm_mhz, reasons bitmask, in timed region)
        self.sm_max = None
        self.in_region = False
        self.stop_flag = False

    # ---- NVML thread
    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        self.nvml = pynvml
        try:
            uuid = str(torch.cuda.get_device_properties(self.cuda_index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            return pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        except Exception:
            return pynvml.nvmlDeviceGetHandleByIndex(self.smi_index)

    def _poll(self, handle):
        nv = self.nvml
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
                self.samples.append((float(mhz), int(get_reasons(handle)), self.in_region))
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        try:
            handle = self._nvml_handle()
            self.sm_max = float(self.nvml.nvmlDeviceGetMaxClockInfo(handle, self.nvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, args=(handle,), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.smi_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            time.sleep(0.5)        # nvidia-smi needs ~0.3 s to start
        except Exception:
            self.proc = None

    def region_begin(self):
        self.in_region = True

    def region_end(self):
        self.in_region = False

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "samples_under_load": 0}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            loaded = [s_ for s_ in self.samples if s_[2]]
            use = sorted(m for m, _, _ in (loaded or self.samples))
            mask = 0
            for _, r, _ in (loaded or self.samples):
                mask |= r
            if use:
                out.update(sm_mhz=use[len(use) // 2], sm_min_mhz=use[0], sm_max_mhz=self.sm_max,
                           samples=len(self.samples), samples_under_load=len(loaded),
                           reasons=sorted(name for bit, name in self.REASONS if mask & bit), source="nvml")
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, loaded = [], [], set(), []
        try:
            for line in open(self.path):
                f = [t.strip() for t in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                    if len(f) > 9 and float(f[9]) >= 50:
                        loaded.append(float(f[1]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            use = sorted(loaded) if loaded else sorted(sm)
            out.update(sm_mhz=use[len(use) // 2], sm_max_mhz=max(mx), samples=len(sm),
                       samples_under_load=len(loaded), reasons=sorted(reasons), source="nvidia-smi")
        return out


def bind_to_gpu_numa(cuda_index):
    """Pin this process (and therefore the pinned host buffers it allocates from now on) to the NUMA node
    its GPU hangs off, when the platform exposes it.  Returns what was found for the JSON line."""
    info = {"numa_node": None, "bound": False}
    try:
        props = torch.cuda.get_device_properties(cuda_index)
        bdf = "%04x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
        info["pci"] = bdf
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        info["numa_node"] = node
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        info["numa_nodes_visible"] = len(nodes)
        if node >= 0 and len(nodes) > 1:
            cpus = set()
            for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
            allowed = os.sched_getaffinity(0)
            use = cpus & allowed
            if use:
                os.sched_setaffinity(0, use)
                info["bound"] = True
                info["cpus"] = len(use)
    except Exception as exc:          # no sysfs / no permission: not fatal, just reported
        info["error"] = str(exc)[:80]
    return info


def e2e_phases(prob, reps=3):
    """Where an end-to-end step spends its time, measured on serialised steps AFTER the timed region:
    H2D of the step's inputs, the layer calls (device time and host time to issue them), D2H."""
    import statistics
    if not hasattr(prob, "layer_pass"):
        return None
    dev = prob.dev
    va_dev = torch.empty_like(prob.va_host, device=dev)
    out_host = torch.empty(prob.va_host.shape[0], 1).pin_memory()
    rows = []
    for _ in range(reps + 1):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        torch.cuda.synchronize()
        ev[0].record()
        va_dev.copy_(prob.va_host, non_blocking=True)
        ev[1].record()
        t0 = time.perf_counter()
        res = prob.layer_pass(va_dev)
        t_issue = time.perf_counter() - t0
        ev[2].record()
        out_host.copy_(res, non_blocking=True)
        ev[3].record()
        torch.cuda.synchronize()
        rows.append((ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), t_issue * 1e3, ev[2].elapsed_time(ev[3])))
    rows = rows[1:]
    med = [statistics.median(r[i] for r in rows) for i in range(4)]
    h2d_gbs = prob.h2d_bytes / (med[0] * 1e-3) / 1e9 if med[0] > 0 else None
    return {"h2d_ms": med[0], "layers_device_ms": med[1], "layers_host_issue_ms": med[2], "d2h_ms": med[3],
            "h2d_GBps": h2d_gbs, "serialised_sum_ms": med[0] + med[1] + med[3],
            "note": "medians of %d serialised steps on this rank; the timed e2e loop overlaps H2D / compute / D2H of "
                    "neighbouring steps, so its step time is bounded below by the largest phase" % reps}


def e2e_timeline(prob, steps=8, keep=4):
    """Timeline of the OVERLAPPED e2e loop (after the timed region): per step the start / end of its H2D,
    layer calls and D2H on the device clock and the host's issue window, ms since the first traced step."""
    if not hasattr(prob, "step_e2e"):
        return None
    torch.cuda.synchronize()
    base = torch.cuda.Event(enable_timing=True)
    prob.trace = []
    base.record()
    t_base = time.perf_counter()
    for _ in range(steps):
        prob.step_e2e()
    torch.cuda.synchronize()
    rows = []
    for ev, h0, h1 in prob.trace[-keep:]:
        t = [round(base.elapsed_time(e), 3) for e in ev]
        rows.append({"h2d": t[0:2], "layers": t[2:4], "d2h": t[4:6],
                     "host_issue": [round((h0 - t_base) * 1e3, 3), round((h1 - t_base) * 1e3, 3)]})
    prob.trace = None
    return {"steps_traced": steps, "last_steps_ms": rows}


def jacobi_bytes(n, z, s=4, k=1):
    return z * (4 + s) + 4 * (n + 1) + (3 * k + 1) * n * s


def cheby_first_bytes(n, z, s=4, k=1):
    return z * (4 + s) + 4 * (n + 1) + 5 * n * k * s


def cheby_next_bytes(n, z, s=4, k=1):
    return z * (4 + s) + 4 * (n + 1) + 6 * n * k * s


# ------------------------------------------------------------------------------- reference arm
def run_cpu_reference(workload, steps, warmup, sample_jacobi=2, sample_cheb=2):
    """Times the reference's own CPU implementation of the path on a bounded sample of the
    workload: `sample_jacobi` Jacobi sweeps + Chebyshev degree `sample_cheb` on the full operator."""
    from oracle import port, ref_loader
    N = GRID[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(24601)
    n = N * N
    ei, ev = port.laplacian_2d(N)
    ev = ev.float()
    z = ei.shape[1]
    b, x = torch.rand(n, 1), torch.rand(n, 1)
    va = torch.cat([-4 * torch.ones(n, 1), b, x], 1)
    ea = torch.cat([ev, torch.zeros_like(ev)], 1)
    gw = torch.tensor(OMEGA).reshape(-1)
    gc = torch.tensor([CHEB_C, CHEB_D])
    if ref_loader.available():
        R = ref_loader.load()
        kind = "reference"
        jac = R.JacobiGNN.JacobiGNN()
        cheb = R.ChebyGNN.ChebyRelaxGNN(sample_cheb)

        def one():
            x1 = jac(sample_jacobi, va, ei, ea, gw)
            return cheb(torch.cat([b, x1], 1), ei, ev, gc)[0]
    else:
        kind = "port"

        def one():
            x1 = port.jacobi(sample_jacobi, va, ei, ea, gw)
            return port.chebyshev(sample_cheb, torch.cat([b, x1], 1), ei, ev, gc)[0]

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    layers = sample_jacobi + sample_cheb
    return {"value": layers * z / dt, "ms_per_step": dt * 1e3, "cores": cores, "kind": kind, "n": n, "nnz": z,
            "sample": "%d of %d Jacobi sweeps + Chebyshev degree %d of %d on the full %s operator "
                      "(%d SpMV-bearing GN blocks per step)" % (sample_jacobi, N_JACOBI, sample_cheb, CHEB_DEG,
                                                               workload, layers)}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = max(1, min(args.steps, 3))
    warm = max(0, min(args.warmup, 1))
    # the reference arm runs the WHOLE pass (10 sweeps + degree 4 = 14 GN blocks, ~6 s per step on 16 cores)
    r = run_cpu_reference(args.workload, steps, warm, sample_jacobi=N_JACOBI, sample_cheb=CHEB_DEG)
    line = {"impl": "reference", "metric": "fused SpMV-layer nnz/s", "value": r["value"], "unit": "nnz/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args.workload, r["n"], r["nnz"], args.gpus),
            "cpu_baseline": {"value": r["value"], "unit": "nnz/s", "cores": r["cores"], "kind": r["kind"],
                             "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def workload_config(workload, n, z, gpus):
    N = GRID[workload]
    return {"workload": "%s: JacobiGNN(10 sweeps, w=0.7) + ChebyRelaxGNN(deg 4, c=-3.4, d=-4) on the %dx%d "
                        "5-point Laplacian" % (workload, N, N),
            "rows": n, "nnz": z, "rhs_columns": 1,
            "fused_layer_launches_per_step": LAUNCHES_PER_STEP,
            "partition": "single GPU" if gpus == 1 else "1-D row blocks over %d GPUs, halo push over NVLink peer memory" % gpus,
            "l2": ("inputs larger than L2 (CSR %.0f MB + vectors per GPU; 126 MB L2), no flush needed"
                   if z * 8 / gpus > 1.5 * 126e6 else
                   "per-GPU working set (CSR %.0f MB + vectors) is comparable to the 126 MB L2: partly L2-resident across sweeps, as it would be in production at this size") % (z * 8 / 1e6 / gpus),
            "e2e_operator": "edge list + CSR plan resident (step-invariant); vectors H2D from pinned host and x D2H every step"}


# ------------------------------------------------------------------------------- GPU arm
def main_gpu(args):
    import glab_b200 as G
    rt = G.runtime
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun with %d ranks" % (args.gpus, args.gpus))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa(local_rank)      # before any pinned allocation
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    N = GRID[args.workload]
    n_global = N * N
    torch.manual_seed(24601)

    if world == 1:
        from bench_support import SingleGpuSmoother
        prob = SingleGpuSmoother(G, N, dev)
    else:
        from bench_support import PartitionedSmoother
        engine = os.environ.get("GLAB_DIST_ENGINE", "peer")
        import torch.distributed as dist
        try:
            prob = PartitionedSmoother(G, N, dev, rank, world, engine=engine,
                                       use_graph=os.environ.get("GLAB_DIST_GRAPH", "1") != "0")
            ok = 1
        except G.GlabError as exc:      # e.g. CUDA IPC / peer access not permitted on this box
            print("rank %d: peer-memory engine unavailable (%s)" % (rank, exc), file=sys.stderr)
            ok = 0
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if flag.item() == 0:            # every rank falls back together: halo rows via NCCL send/recv
            prob = PartitionedSmoother(G, N, dev, rank, world, engine="torch", use_graph=False)
    z_global = prob.nnz_global

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- kernels only (value) + roofline of the dominant kernel
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()      # before the warm-up; region_begin/end mark the two timed regions
    for _ in range(args.warmup):
        prob.step_kernels()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    jac_events = []
    launches0 = rt.launch_count
    barrier()
    sampler.region_begin()
    ev0.record()
    for _ in range(args.steps):
        jac_events.append(prob.step_kernels(time_jacobi=True))
    ev1.record()
    barrier()
    sampler.region_end()
    launches = rt.launch_count - launches0
    ms_total = ev0.elapsed_time(ev1)
    jac_ms = sum(a.elapsed_time(b) for a, b in jac_events) / (len(jac_events) * N_JACOBI)

    # ---------------- end to end through the layer API with pinned host buffers
    for _ in range(min(args.warmup, 3)):
        prob.step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.region_begin()
    t_wall0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        prob.step_e2e()
    e1.record()
    barrier()
    e2e_wall = time.perf_counter() - t_wall0
    sampler.region_end()
    e2e_ms = max(e0.elapsed_time(e1), e2e_wall * 1e3)  # D2H is synchronous: host clock bounds it too
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms_total, jac_ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, jac_ms, e2e_ms = t.tolist()

    # ---------------- parity of the measured path (outside the timed regions), phases of an e2e step
    parity = prob.parity()
    phases = e2e_phases(prob)
    if phases is not None:
        phases["overlapped_timeline"] = e2e_timeline(prob)
    if world > 1 and getattr(prob, "op", None) is not None:
        prob.op.check()                      # bounded in-kernel waits: raises if one gave up

    ms_step = ms_total / args.steps
    value = LAUNCHES_PER_STEP * z_global / (ms_step * 1e-3)
    e2e_value = LAUNCHES_PER_STEP * z_global / (e2e_ms / args.steps * 1e-3)
    peak, peak_src = peaks()
    # per-GPU roofline of the dominant kernel (its own rows / nnz)
    bytes_jac = jacobi_bytes(prob.n_local, prob.nnz_local)
    achieved = bytes_jac / (jac_ms * 1e-3) / 1e9
    traffic = committed_traffic()
    idx_bytes = getattr(prob, "index_bytes", 4)
    moved_jac = int(bytes_jac - (4 - idx_bytes) * prob.nnz_local)   # what the kernel actually streams
    line = {
        "metric": "fused SpMV-layer nnz/s", "value": value, "unit": "nnz/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, n_global, z_global, world),
        "roofline": {"bound": "hbm", "kernel": "glab_jacobi_f32 (k_row_pipe<float,1,5,EpiJacobi>: TMA-fed persistent pipeline)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "frac_of_nominal_8TBps": achieved / 8000.0,
                     "peak_source": peak_src, "bytes_per_launch": bytes_jac, "ms_per_launch": jac_ms,
                     "index_bytes_streamed": idx_bytes, "bytes_moved_per_launch": moved_jac,
                     "moved_gbs": moved_jac / (jac_ms * 1e-3) / 1e9, "moved_frac": moved_jac / (jac_ms * 1e-3) / 1e9 / peak,
                     "note": ("algorithmic bytes use SURVEY 8d's 4-byte column indices; with index_bytes_streamed = 2 the "
                              "kernel streams 16-bit row-relative indices, so it moves fewer bytes than the algorithmic "
                              "count and `frac` can exceed 1; `moved_frac` is the fraction of peak actually moved "
                              "(row blocks of a partitioned operator: fractional = share of 256-row tiles on 16-bit "
                              "indices; the tiles that read the halo tail stay on int32)"),
                     "gnnz_per_s": prob.nnz_local / (jac_ms * 1e-3) / 1e9,
                     "traffic": None if not traffic else traffic.get(
                         "jacobi_dram_bytes_per_launch_idx16" if idx_bytes == 2 else "jacobi_dram_bytes_per_launch")
                     if world == 1 else None,
                     "traffic_source": (None if not traffic else traffic.get(
                         "source_idx16" if idx_bytes == 2 else "source")) if world == 1 else None},
        "e2e": {"value": e2e_value, "unit": "nnz/s", "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": prob.h2d_bytes * world, "d2h_bytes_per_step": prob.d2h_bytes * world,
                "api": ("JacobiGNN.forward(10, vertex_attr, edgeij_pair, ...) + ChebyRelaxGNN(4).forward(...) on device "
                        "copies of pinned host vectors" if world == 1 else
                        "JacobiGNN.forward(10, vertex_attr_slab, dist.PartitionedGraph, ...) + ChebyRelaxGNN(4).forward(...) "
                        "on each rank's slab, device copies of pinned host vectors"),
                "uploaded_per_step": "vertex_attr = [A_ii, b, x] (3 vectors) of every rank's rows",
                "pipelining": "3 staging buffers: the H2D of the next steps and the D2H of the previous one overlap step i's compute",
                "phases": phases, "numa": numa},
        "parity": parity,
        "gpu_launches": launches,
        "clocks": clocks,
        "setup": prob.setup_info,
    }
    if parity is not None and not parity.get("ok", False):
        line["parity_failed"] = True
    # ---------------- the other BASELINE configs as sub-results (each guarded: a failure is reported, not fatal)
    extras = [e for e in (args.extras or "").split(",") if e and e != "none"]
    if "all" in extras:
        extras = ["L8192_layers", "config3_power", "config4_amg", "config5_vcycle"]
    if extras:
        if hasattr(prob, "close"):
            prob.close()
        del prob
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        # The headline numbers above are complete.  The sub-results must never cost the driver its JSON
        # line: if they overrun their budget (a stuck collective cannot be interrupted from Python), a
        # watchdog prints the line with what has finished and ends every rank.
        done_extras = {}
        line["extra"] = done_extras
        budget = float(os.environ.get("GLAB_BENCH_EXTRAS_BUDGET_S", "240"))

        def give_up():
            import faulthandler
            faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
            if rank == 0:
                done_extras["_watchdog"] = {"error": "sub-results exceeded %.0f s; the remaining ones were abandoned" % budget}
                try:
                    sys.stdout.write(json.dumps(line) + "\n")
                    sys.stdout.flush()
                finally:
                    os._exit(0)
            time.sleep(2.0)          # let rank 0 print before the launcher sees ranks leave
            os._exit(0)

        import threading
        dog = threading.Timer(budget, give_up)
        dog.daemon = True
        dog.start()
        run_extras(G, dev, rank, world, peak, extras, not args.no_cpu_baseline, done_extras)
        dog.cancel()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = run_cpu_reference(args.workload, 1, 0)
        line["cpu_baseline"] = {"value": r["value"], "unit": "nnz/s", "cores": r["cores"], "kind": r["kind"],
                                "sample": r["sample"], "ms_per_step": r["ms_per_step"]}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_extras(G, dev, rank, world, peak, names, cpu_baseline, out=None):
    """bench_extra.* one by one; every rank takes part in the partitioned ones.  The single-GPU-only
    sub-results (the 67 M-row layer table and the AMG setup kernels) run on rank 0's GPU at N = 1 only."""
    import bench_extra as X
    out = {} if out is None else out
    grid_big = int(os.environ.get("GLAB_BENCH_BIG", "8192"))       # shrink for smoke runs of the bench itself
    grid_amg = int(os.environ.get("GLAB_BENCH_AMG", "4096"))
    for name in names:
        t0 = time.perf_counter()
        try:
            if name == "L8192_layers":
                res = X.l8192_layers(G, dev, peak, N=grid_big) if world == 1 else None
            elif name == "config3_power":
                res = X.config3_power(G, dev, rank, world, peak, N=grid_big, cpu_baseline=cpu_baseline)
                # config 3 in fp32 AND fp64 -- the fp64 run as a compact sub-result
                r64 = X.config3_power(G, dev, rank, world, peak, N=grid_big, cpu_baseline=False, dtype=torch.float64)
                res["fp64"] = {k_: r64[k_] for k_ in ("ms_per_iteration", "value_nnz_per_s", "lambda", "norm", "roofline",
                                                      "parity") if k_ in r64}
            elif name == "config4_amg":
                res = X.config4_amg(G, dev, peak, N=grid_amg, cpu_baseline=cpu_baseline) if world == 1 else None
            elif name == "config5_vcycle":
                res = X.config5_vcycle(G, dev, rank, world, peak, N=grid_big, cpu_baseline=cpu_baseline)
            else:
                res = {"error": "unknown extra %r" % name}
        except Exception as exc:       # noqa: BLE001 -- reported in the JSON line
            import traceback
            sys.stderr.write("[bench extras] rank %d: %s failed\n%s\n" % (rank, name, traceback.format_exc()))
            sys.stderr.flush()
            res = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300]),
                   "where": traceback.format_exc().strip().splitlines()[-3:]}
            torch.cuda.synchronize()
        if res is not None:
            res["wall_s"] = time.perf_counter() - t0
            out[name] = res
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="glab", choices=["glab", "reference"])
    ap.add_argument("--workload", default=os.environ.get("GLAB_BENCH_WORKLOAD", "L4096"), choices=sorted(GRID))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extras", default=os.environ.get("GLAB_BENCH_EXTRAS", "all"),
                    help="comma list of L8192_layers,config3_power,config4_amg,config5_vcycle | all | none")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "glab":
        args.warmup = 3
    if args.impl == "reference":
        return main_reference(args)
    return main_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
