#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libuc2_b200.so: the Blackwell-native mnemonics (tcgen05.mma -> UTC*MMA,
tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UTMAPF, tcgen05.commit -> UTCBAR) next to the legacy tensor
path (mma.sync -> HMMA).  Runs on the CPU box:

    python scripts/sass_histogram.py > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "uc2_b200", "libuc2_b200.so")
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "HMMA", "MUFU", "LDGSTS"]


def demangle(name):
    for tool in ("cu++filt", "c++filt"):
        try:
            d = subprocess.run([tool, name], capture_output=True, text=True).stdout.strip()
        except OSError:
            continue
        if d and d != name:
            d = re.sub(r"^void ", "", d)
            if d.endswith(")"):                      # drop the trailing parameter list, keep template arguments
                depth = 0
                for i in range(len(d) - 1, -1, -1):
                    depth += d[i] == ")"
                    depth -= d[i] == "("
                    if depth == 0:
                        d = d[:i]
                        break
            d = d.replace("uc2::(anonymous namespace)::", "").replace("uc2::<unnamed>::", "").replace("uc2::", "")
            return d.replace("<unnamed>::", "")
    k = re.search(r"\d+([A-Za-z_0-9]+?_kernel)(I\w+E)?", name)
    return (k.group(1) + (k.group(2) or "")) if k else name


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fn = None
    hist = collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = demangle(m.group(1))
            hist.setdefault(fn, collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and fn:
            hist[fn][m.group(1)] += 1
    rev = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    print(f"# cuobjdump -sass uc2_b200/libuc2_b200.so (built from commit {rev} + working tree), opcode counts per kernel")
    print(f"# {'kernel':78s} " + " ".join(f"{w:>8s}" for w in WATCH) + "    total")
    tot = collections.Counter()
    for fn, c in sorted(hist.items()):
        if not c:
            continue
        tot.update(c)
        print(f"{fn[:80]:80s} " + " ".join(f"{c.get(w, 0):8d}" for w in WATCH) + f" {sum(c.values()):8d}")
    print(f"{'ALL KERNELS':80s} " + " ".join(f"{tot.get(w, 0):8d}" for w in WATCH) + f" {sum(tot.values()):8d}")


if __name__ == "__main__":
    main()
