#!/usr/bin/env python3
"""Per-kernel resource usage and the SASS evidence for the fp64 tensor path, from the built library.

    python tools/resource_table.py > profiles/rNN_kernel_resources.txt

Needs only the CUDA toolkit (cuobjdump, cu++filt); no GPU.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "matrix_solver_b200", "lib", "libms_b200.so")
MNEMONICS = ("DMMA", "UTMALDG", "SYNCS", "DFMA", "DADD", "DMUL", "MUFU", "LDS", "STS", "LDG", "STG", "CCTL",
             "BAR", "ATOM", "RED", "STL", "LDL")


def demangle(names):
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout
    return [re.sub(r"\(.*", "", ln.strip()).replace("void ", "") for ln in out.strip().split("\n")]


def main():
    if not os.path.exists(LIB):
        sys.exit("build the library first: make lib")
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    usage = {}
    fn = None
    for ln in res.split("\n"):
        m = re.match(r"\s*Function (\S+):", ln)
        if m:
            fn = m.group(1)
            continue
        m = re.match(r"\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", ln)
        if m and fn:
            usage[fn] = tuple(int(x) for x in m.groups())
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.defaultdict(collections.Counter)
    fn = None
    for ln in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            fn = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m and fn:
            op = m.group(1)
            for mn in MNEMONICS:
                if op.startswith(mn):
                    counts[fn][mn] += 1
    names = sorted(usage)
    pretty = dict(zip(names, demangle(names)))
    print("kernel resources of matrix_solver_b200/lib/libms_b200.so (sm_100a; cuobjdump -res-usage / -sass)")
    print("static shared memory only: the tile, TRSM and sweep kernels add dynamic shared memory at launch\n")
    print(f"{'kernel':58s} {'regs':>5s} {'stack':>6s} {'smem':>7s} {'local':>6s}  SASS mnemonic counts")
    for n in sorted(names, key=lambda k: pretty[k]):
        r, st, sh, lo = usage[n]
        c = " ".join(f"{k}={v}" for k, v in sorted(counts[n].items(), key=lambda kv: -kv[1]) if v)
        print(f"{pretty[n][:58]:58s} {r:5d} {st:6d} {sh:7d} {lo:6d}  {c}")


if __name__ == "__main__":
    main()
