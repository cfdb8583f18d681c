#!/usr/bin/env python3
"""Per-kernel counts of the SASS instructions that show how the library moves data: TMA tensor loads
(UTMALDG), L2 tensor prefetch (UTMAPF), bulk stores (UBLKCP), mbarrier transactions (SYNCS.*),
16-byte shared-memory vectors (LDS.128 / STS.128), release/acquire accesses of the progress counters,
block barriers, system-scope atomics of the peer-memory transport.

    python tools/sass_evidence.py [library.so] > profiles/rN_sass_evidence.txt

Runs `cuobjdump -sass` on the in-tree library (no GPU needed).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hpcclassmultigridproject_b200", "libmgb200.so")

KEEP = ("UTMALDG", "UTMAPF", "UBLKCP", "UTMASTG", "SYNCS", "LDS", "STS", "LDG", "STG", "BAR", "MEMBAR", "ATOMG", "ATOMS",
        "REDG", "RED", "DFMA", "DMUL", "DADD", "MUFU.RCP64H", "SHFL", "ELECT", "NANOSLEEP", "FENCE", "CCTL", "ERRBAR",
        "LD.E", "ST.E", "WARPSYNC", "BSSY", "BSYNC")


def demangle(names):
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True)
    if out.returncode != 0:
        return names
    return out.stdout.splitlines()


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else LIB
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            kernels[cur]["__total__"] += 1
            op = m.group(1)
            if op.startswith(KEEP):
                kernels[cur][op] += 1
    names = list(kernels)
    pretty = demangle(names)
    print("# SASS evidence: cuobjdump -sass %s" % os.path.relpath(lib, ROOT))
    print("# per kernel: total instructions, then the counts of the data-movement / synchronisation / FP64 instructions.")
    print("# k_syst_pass<arith (0 fast, 1 exact), PRE role present, epilogue (0 none, 1 injection, 2 sum of squares)>.")
    for name, nice in sorted(zip(names, pretty), key=lambda t: t[1]):
        c = kernels[name]
        total = c.pop("__total__", 0)
        print(nice[:150])
        print("    instructions %d: %s" % (total, ", ".join("%s %d" % kv for kv in sorted(c.items()))))


if __name__ == "__main__":
    main()
