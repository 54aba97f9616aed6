#!/usr/bin/env python
"""Summarise ncu output into small tracked files under profiles/.

  launches: python tools/ncu_summary.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
            (from `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...`)
  kernel  : python tools/ncu_summary.py kernel gpurun_out/prof_trace_r1.ncu-rep profiles/r1_k_trace.md
            (from `ncu --set full --clock-control none --import-source on -k regex:... -o ...`)
"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import OrderedDict, defaultdict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__thread_inst_executed_per_inst_executed.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_bytes.sum",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    # the L1 data pipe: wavefronts of the load instructions + sector fills after misses (profiles/r2_sweeps.md section 14)
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
]


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)
    name = name.replace("crb::<unnamed>::", "").replace("void ", "")
    return name.strip()


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1e6 if r[ui] == "ns" else (v / 1e3 if r[ui] == "us" else v)  # -> ms
        tot[short(r[ki])] += v
        cnt[short(r[ki])] += 1
    total = sum(tot.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\nPer-launch times are cold-cache and serialised under the profiler: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total ms | share | mean ms |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write(f"| `{k}` | {cnt[k]} | {v:.3f} | {v / total:.1%} | {v / cnt[k]:.4f} |\n")
        f.write(f"\ntotal {total:.3f} ms over {sum(cnt.values())} launches\n")
    print(open(dst).read())


def kernel(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = OrderedDict()
        d["kernel"] = short(r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                d[k] = f"{r[hdr.index(k)]} {units[hdr.index(k)]}".strip()
        out.append(d)
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n")
        for i, d in enumerate(out):
            f.write(f"## launch {i}: `{d['kernel']}`\n\n| metric | value |\n|---|---|\n")
            for k, v in d.items():
                if k != "kernel":
                    f.write(f"| {k} | {v} |\n")
            f.write("\n")
    json.dump(out, open(dst.replace(".md", ".json"), "w"), indent=1)
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
