#!/usr/bin/env python
"""profiles/r2_counters.json from the committed ncu summaries (tools/ncu_summary.py kernel -> *.json): the figures
bench.py quotes in its roofline object (per-launch DRAM traffic of the dominant kernel, issue-slot utilisation, lanes per
instruction, warp occupancy), with their source named.

    python tools/ncu_counters.py profiles/r2p_k_trace.json [profiles/r2ac_c4flat_k_trace.json [profiles/r2ad_c4tl_k_trace2.json]]
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def num(v):
    m = re.match(r"\s*([0-9.eE+-]+)\s*(\S*)", str(v))
    x, unit = float(m.group(1)), m.group(2)
    return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)


def summarise(path, prefix):
    rows = [r for r in json.load(open(path)) if "k_trace" in r["kernel"]]
    n = len(rows)
    out = {
        f"{prefix}_dram_bytes_per_launch": sum(num(r["dram__bytes_read.sum"]) + num(r["dram__bytes_write.sum"]) for r in rows) / n,
        f"{prefix}_issue_active": sum(num(r["smsp__issue_active.avg.pct_of_peak_sustained_active"]) for r in rows) / n / 100.0,
        f"{prefix}_lanes_per_inst": sum(num(r["smsp__thread_inst_executed_per_inst_executed.ratio"]) for r in rows) / n,
        f"{prefix}_warps_active": sum(num(r["sm__warps_active.avg.pct_of_peak_sustained_active"]) for r in rows) / n / 100.0,
        f"{prefix}_launches_profiled": n,
        f"{prefix}_source": os.path.relpath(path, ROOT) + " (ncu --set full --clock-control none, " + rows[0]["kernel"] + ")",
    }
    return out


if __name__ == "__main__":
    out = summarise(sys.argv[1], "k_trace")
    out["counters_source"] = out["k_trace_source"]
    if len(sys.argv) > 2:
        out.update(summarise(sys.argv[2], "c4_k_trace"))      # config 4 flattened: k_trace on the 1 GB tree
    if len(sys.argv) > 3:
        out.update(summarise(sys.argv[3], "c4tl_k_trace"))    # config 4 two-level (default): k_trace2
    dst = os.path.join(ROOT, "profiles", "r2_counters.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(open(dst).read())
