#!/usr/bin/env python3
"""Condense an `ncu --set full` report into the small metric,value,unit CSV kept under profiles/.

usage: tools/ncu_summary.py gpurun_out/<name>.ncu-rep profiles/<name>.csv [kernel-substring]
Reads the report with `ncu -i ... --page raw --csv` (no GPU needed) and keeps the pipe, issue,
stall, occupancy, launch-shape and DRAM metrics of the first launch whose name contains the substring.
"""
import csv
import io
import subprocess
import sys

KEEP_EXACT = {
    "Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
    "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__registers_per_thread", "sm__cycles_active.avg",
    "sm__cycles_elapsed.max", "sm__inst_executed.avg.per_cycle_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__cycles_active.avg",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
}
KEEP_PIPES = ("alu", "fma", "lsu", "uniform", "xu")


def keep(name):
    if name in KEEP_EXACT:
        return True
    if name.startswith("smsp__average_warps_issue_stalled_") and name.endswith("_per_issue_active.ratio"):
        return True
    for p in KEEP_PIPES:
        if name in (f"sm__inst_executed_pipe_{p}.avg.pct_of_peak_sustained_active",
                    f"sm__pipe_{p}_cycles_active.avg.pct_of_peak_sustained_active"):
            return True
    return False


def main():
    rep, out = sys.argv[1], sys.argv[2]
    want = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = hdr.index("Kernel Name")
    row = next(r for r in rows[2:] if want in r[col])
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh, quoting=csv.QUOTE_ALL)
        fh.write("metric,value,unit\n")
        for name, unit, val in sorted(zip(hdr, units, row)):
            if keep(name):
                w.writerow([name, val, unit])
    print(out)


if __name__ == "__main__":
    main()
