#!/usr/bin/env python3
"""Summarise one `ncu --set full` capture (.ncu-rep) into the small JSON files kept under profiles/.
   python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x.json      (needs the `ncu` CLI; no GPU)"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__occupancy_limit_registers": "occupancy_limit_registers_blocks",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active": "pipe_fmaheavy_pct",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active": "pipe_fmaheavy_cycles_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "pipe_alu_cycles_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_lanes_per_instruction",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "launch__local_size": "local_bytes_per_thread",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_throttle",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_instruction",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "stall_branch_resolving",
}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = {"_source": rep, "kernels": []}
    for vals in rows[2:]:
        k = {}
        for h, u, v in zip(hdr, units, vals):
            if h == "Kernel Name":
                k["kernel"] = v
            elif h in WANT:
                try:
                    k[WANT[h]] = {"value": float(v.replace(",", "")), "unit": u}
                except ValueError:
                    k[WANT[h]] = {"value": v, "unit": u}
        out["kernels"].append(k)
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
