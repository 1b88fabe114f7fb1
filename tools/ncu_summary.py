#!/usr/bin/env python
"""Print the handful of `ncu --set full` raw metrics we read per kernel: `python tools/ncu_summary.py report.ncu-rep [extra-substring ...]`."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum ",
    "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum ", "lts__throughput.avg.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct", "sm__cycles_elapsed.avg ",
    "sm__inst_executed_pipe_tmem", "sm__inst_executed_pipe_tma", "smsp__inst_executed_op_ldsm", "sm__pipe_shared_cycles_active", "sm__mio",
]


def main():
    rep, extra = sys.argv[1], sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("====", d.get("Kernel Name", "?")[:100], "id", d.get("ID"))
        for h, u, v in zip(hdr, units, r):
            hh = h + " "
            if any(k in hh for k in KEYS + extra):
                if "stalled" in h and "ratio" in h:
                    try:
                        if float(v) < 0.15:
                            continue
                    except ValueError:
                        pass
                print(f"  {h:95s} {u:12s} {v}")


if __name__ == "__main__":
    main()
