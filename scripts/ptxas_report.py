#!/usr/bin/env python
"""Summarise `-Xptxas -v` logs: registers / spill bytes / stack per kernel.

    python scripts/ptxas_report.py gnn-applied-linear-algebra_b200/csrc/build/*.ptxas.log [--spills] [--grep PATTERN]
"""
import re
import subprocess
import sys


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def parse(path):
    rows, cur = [], None
    for line in open(path, errors="replace"):
        m = re.search(r"Function properties for (\S+)", line)
        if m:
            cur = {"name": m.group(1), "spill_st": 0, "spill_ld": 0, "stack": 0, "regs": None}
            rows.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            cur["stack"], cur["spill_st"], cur["spill_ld"] = map(int, m.groups())
        m = re.search(r"Used (\d+) registers", line)
        if m:
            cur["regs"] = int(m.group(1))
    return rows


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    only_spills = "--spills" in sys.argv
    pat = None
    if "--grep" in sys.argv:
        pat = sys.argv[sys.argv.index("--grep") + 1]
        args = [a for a in args if a != pat]
    rows = []
    for p in args:
        rows += parse(p)
    names = demangle([r["name"] for r in rows])
    n_spill = 0
    for r in rows:
        nm = names.get(r["name"], r["name"])
        nm = re.sub(r"\(.*", "", nm).replace("glab::", "")
        if r["spill_st"] or r["spill_ld"]:
            n_spill += 1
        if only_spills and not (r["spill_st"] or r["spill_ld"]):
            continue
        if pat and not re.search(pat, nm):
            continue
        print("%3s regs  spill %4d/%4d  stack %4d  %s" % (r["regs"], r["spill_st"], r["spill_ld"], r["stack"], nm[:150]))
    print("kernels: %d, with spills: %d" % (len(rows), n_spill))


if __name__ == "__main__":
    main()
