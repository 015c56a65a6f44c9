#!/usr/bin/env python
"""Per-kernel resource table from the ptxas logs the Makefile leaves next to the objects (csrc/*.ptxas.log are build
artefacts and not tracked; this summary is): registers, spills, stack, shared memory.

  python scripts/ptxas_summary.py > profiles/r02_ptxas_summary.txt
"""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def demangle(names):
    try:
        out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        return [o.split("(")[0] for o in out]
    except Exception:
        return names


def main():
    rows = []
    for f in sorted(glob.glob(os.path.join(ROOT, "gadget_leicester_b200", "csrc", "*.ptxas.log"))):
        txt = open(f).read().splitlines()
        for k, ln in enumerate(txt):
            m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", ln)
            if not m:
                continue
            name = m.group(1)
            blk = " ".join(txt[k:k + 4])
            sp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", blk)
            us = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", blk)
            sm = re.search(r"(\d+) bytes smem", blk)
            rows.append((os.path.basename(f)[:-10], name, int(us.group(1)) if us else -1, int(sp.group(1)) if sp else 0,
                         int(sp.group(2)) if sp else 0, int(sp.group(3)) if sp else 0, int(sm.group(1)) if sm else 0))
    names = demangle([r[1] for r in rows])
    print("# ptxas -v, nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false (gadget_leicester_b200/csrc/Makefile)")
    print("# %-12s %-58s %5s %6s %7s %7s %6s" % ("file", "kernel", "regs", "stack", "spill-st", "spill-ld", "smem"))
    for r, nm in zip(rows, names):
        print("  %-12s %-58s %5d %6d %7d %7d %6d" % (r[0], nm[:58], r[2], r[3], r[4], r[5], r[6]))


if __name__ == "__main__":
    main()
