#!/usr/bin/env python
"""Per-instruction warp-stall samples of every kernel in an .ncu-rep captured with --import-source on: the SASS
instructions that collected the most samples (read on the CPU box with `ncu -i`).

    python scripts/ncu_stalls.py gpurun_out/x.ncu-rep [top_n] >> profiles/rNN_x_ncu.txt
"""
import csv
import subprocess
import sys


def main(path, top=16):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    blocks, cur = [], None
    for r in csv.reader(out.splitlines()):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    for b in blocks:
        h = b["rows"][0]
        isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        data = [(int(r[isamp] or 0), r[isrc].strip(), int(r[iex] or 0), i)
                for i, r in enumerate(b["rows"][1:]) if len(r) > isamp]
        tot = sum(d[0] for d in data) or 1
        print(f"# warp-stall samples by SASS instruction: {b['name'][:120]}")
        print(f"#   {tot} samples over {len(data)} instructions, {sum(d[2] for d in data)} warp instructions executed")
        for s, src, ex, i in sorted(data, reverse=True)[:top]:
            print(f"    {s:6d} {100 * s / tot:5.1f}%  sass#{i:<5d} executed {ex:9d}  {src[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 16)
