#!/usr/bin/env python
"""Key metrics of one `ncu --set full` report as a small CSV (the r*_ncu_full_*.csv files in this directory).

    python profiles/ncu_summary.py gpurun_out/r2y_full_front2.ncu-rep profiles/r2_ncu_full_front2_fp16x2.csv
"""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__block_size",
        "launch__grid_size", "launch__registers_per_thread", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]


def main():
    src, dst = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as fh:
        for k in KEYS:
            if k in col:
                v = vals[col[k]].replace(",", ";") if k == "Kernel Name" else vals[col[k]]
                fh.write("%s,%s%s\n" % (k, v, ("," + units[col[k]]) if k != "Kernel Name" else ""))


if __name__ == "__main__":
    main()
