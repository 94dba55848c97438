#!/usr/bin/env python
"""Turn an .ncu-rep into the tall CSV summaries kept under profiles/: one line per (kernel, metric)
for the metrics the roofline claims rest on.   usage: ncu_summary.py <report.ncu-rep> <out.csv>"""
import csv
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tmem", "sm__inst_executed_pipe_uniform", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__icc_request_hit_rate", "smsp__pcsamp_warps_issue_stalled", "launch__grid_size", "launch__block_size",
        "launch__cluster_dim_x", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu",
        "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_lsu", "smsp__cycles_active.avg", "gpc__cycles_elapsed.max",
        "dram__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.avg.per_second")


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["launch", "kernel", "metric", "unit", "value"])
        for i, r in enumerate(rows[2:]):
            kern = r[name_col].split("(")[0]
            for h, u, v in zip(hdr, units, r):
                if v != "" and any(k in h for k in KEEP):
                    w.writerow([i, kern, h, u, v])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
