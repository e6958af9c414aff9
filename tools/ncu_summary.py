#!/usr/bin/env python3
"""Condenses an .ncu-rep (ncu --set full) into the small CSV kept under profiles/: launch configuration, duration, DRAM
bytes, pipe utilisation and the warp-stall breakdown of every captured kernel.

    python tools/ncu_summary.py gpurun_out/k1.ncu-rep > profiles/rNN_k1_ncu_raw_subset.csv
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = csv.writer(sys.stdout)
    out.writerow(["kernel", "metric", "value", "unit"])
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = d["Kernel Name"]
        for k in KEEP:
            if k in d and d[k] != "":
                out.writerow([name, k, d[k], u.get(k, "")])
        for k in hdr:
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                try:
                    v = float(d[k])
                except ValueError:
                    continue
                if v >= 0.1:
                    out.writerow([name, k, f"{v:.3f}", "warps/issue"])


if __name__ == "__main__":
    main()
