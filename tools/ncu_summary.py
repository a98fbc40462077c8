#!/usr/bin/env python
"""Condense an `ncu --set full` report into the per-launch counters that DESIGN.md cites:
  python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.csv"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct", "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic"]


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    keys = [k for k in KEYS if k in ix]
    w = csv.writer(sys.stdout)
    w.writerow(["Kernel Name"] + keys)
    w.writerow([""] + [units[ix[k]] for k in keys])
    for d in data:
        w.writerow([d[ix["Kernel Name"]]] + [d[ix[k]] for k in keys])


if __name__ == "__main__":
    main()
