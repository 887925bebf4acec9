#!/usr/bin/env python
"""Key metrics of every kernel in an .ncu-rep (read on the CPU box with `ncu -i`):

    python scripts/ncu_summary.py gpurun_out/prof_gemm.ncu-rep > profiles/r01_gemm_ncu.txt
"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}")
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        print(name[:150])
        for w in WANT:
            if w in col:
                print(f"    {w:85s} {r[col[w]]:>16s} {units[col[w]]}")


if __name__ == "__main__":
    main(sys.argv[1])
