#!/usr/bin/env python
"""Summarise an Nsight Compute report (read here, no GPU): tools/ncu_summary.py report.ncu-rep out.json "what".
Keeps, per profiled launch, the metrics DESIGN.md / bench.py cite (duration, DRAM bytes and throughput, L1 / LSU
wavefronts, issue, occupancy, pipe utilisation, registers, stall reasons per issue)."""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes.sum.per_second", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def main():
    rep, out, what = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, r):
            if h in ("Kernel Name", "ID") or h in KEEP:
                try:
                    d[h] = [float(v.replace(",", "")), u] if h not in ("Kernel Name",) else v
                except ValueError:
                    d[h] = v
        launches.append(d)
    json.dump({"what": what, "launches": launches}, open(out, "w"), indent=1)
    for d in launches:
        print(d.get("Kernel Name"), d.get("gpu__time_duration.sum"), d.get("dram__bytes_write.sum"), d.get("dram__bytes_read.sum"))


if __name__ == "__main__":
    main()
