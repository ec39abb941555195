#!/usr/bin/env python
"""Turn an `ncu --set full` report into the text summary committed under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_jacobi_full.txt [--traffic-json profiles/traffic.json]

Reads the report with `ncu -i ... --page raw --csv` (works without a GPU); a first argument ending in
`.csv` is taken as that export itself (reports are ~30 MB each: on the GPU box they are written to /tmp
and only the CSV pages travel back)."""
import csv
import io
import json
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    tj = sys.argv[sys.argv.index("--traffic-json") + 1] if "--traffic-json" in sys.argv else None
    if rep.endswith(".csv"):
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    lines = ["# ncu --set full --clock-control none summary of %s" % rep,
             "# %d launches captured; values per launch" % len(data), ""]
    for r in data:
        lines.append("kernel: " + r[ki][:150])
    lines.append("")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            lines.append("%-86s %-16s %s" % (w, units[i], "  ".join(r[i] for r in data)))
    i_r, i_w, i_t = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")

    def mb(v, u):
        v = float(v.replace(",", ""))
        return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[u]

    tot = [mb(r[i_r], units[i_r]) + mb(r[i_w], units[i_w]) for r in data]
    avg = sum(tot) / len(tot)
    lines += ["", "dram traffic per launch (read+write): avg %.1f MB  (min %.1f, max %.1f)" % (avg / 1e6, min(tot) / 1e6, max(tot) / 1e6)]
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    if tj:
        json.dump({"jacobi_dram_bytes_per_launch": avg, "kernel": data[0][ki][:120],
                   "source": "ncu --set full --clock-control none, %s (%d launches)" % (out, len(data))},
                  open(tj, "w"), indent=1)


if __name__ == "__main__":
    main()
