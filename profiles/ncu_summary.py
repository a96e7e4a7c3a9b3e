#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, on CPU): key metrics of each captured kernel and,
optionally, the hottest source lines.  Usage: python profiles/ncu_summary.py REPORT [--source N]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg",
    "sm__cycles_active.avg", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_op_read.sum",
    "lts__t_sectors_op_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
    "sm__cycles_active.avg.pct_of_peak_sustained_elapsed",
]


def raw(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return hdr, units, data


def main():
    report = sys.argv[1]
    hdr, units, data = raw(report)
    name_i = hdr.index("Kernel Name")
    for row in data:
        print("==", row[name_i])
        for i, h in enumerate(hdr):
            short = h.split(".TriageCompute.")[-1] if ".Triage" in h else h
            if short in KEYS or ("issue_stalled" in short and short.endswith("per_warp_active.pct")):
                print(f"  {short:90s} {row[i]:>18s} {units[i]}")
    if "--source" in sys.argv:
        n = int(sys.argv[sys.argv.index("--source") + 1])
        out = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if rows:
            h = rows[0]
            print(h)
            try:
                si = h.index("Warp Stall Sampling (All Samples)")
            except ValueError:
                si = None
            body = [r for r in rows[1:] if len(r) == len(h)]
            if si is not None:
                body.sort(key=lambda r: -float(r[si] or 0))
            for r in body[:n]:
                print(r)


if __name__ == "__main__":
    main()
