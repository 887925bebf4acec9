#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time share of ONE step
(the launches between the last two adamw_kernel launches).

    python scripts/summarize_launches.py gpurun_out/launches_itm.csv [n_steps] > profiles/r01_launches_itm.txt
"""
import collections
import csv
import re
import sys


def short(n):
    m = re.search(r"(\w+)(<[^>]*>)?\(", n)
    return (m.group(1) + (m.group(2) or "")) if m else n[:60]


def main(path, n_steps=1):
    with open(path) as f:
        rows = list(csv.DictReader([l for l in f if not l.startswith("==")]))
    names = [short(r["Kernel Name"]) for r in rows]
    marks = [i for i, n in enumerate(names) if n.startswith("adamw_kernel")]
    if len(marks) >= n_steps + 1:
        lo, hi = marks[-1 - n_steps] + 1, marks[-1] + 1
    else:
        lo, hi = 0, len(rows)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r, n in zip(rows[lo:hi], names[lo:hi]):
        k = (n, r["Grid Size"], r["Block Size"])
        agg[k][0] += 1
        agg[k][1] += float(r["Metric Value"]) / 1e3
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: launches [{lo}, {hi}) = {n_steps} step(s), {hi - lo} launches, {tot:.1f} us summed kernel time "
          f"(ncu: serialised, cold cache -- compare shares, not absolutes)")
    print(f"{'us':>10} {'n':>5} {'share':>7}  kernel grid block")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:10.1f} {v[0]:5d} {100 * v[1] / tot:6.1f}%  {k[0]} {k[1]} {k[2]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
