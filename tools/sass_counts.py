#!/usr/bin/env python
"""SASS evidence (build container, no GPU): per kernel of libphynet_b200.so, the count of the mnemonics that prove the
Blackwell-native paths.  `python tools/sass_counts.py > profiles/sass_r2.txt`"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "phynetpy_b200", "libphynet_b200.so")
KEYS = ["UBLKCP", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK.TRANS64.TRYWAIT", "LDG.E.ENL2.256", "STG.E.ENL2.256", "LDS.128",
        "LDS", "STS", "ATOMS", "LD.E", "ST.E", "DFMA", "DMUL", "ATOMG", "MEMBAR", "CCTL.IVALL", "MATCH.ANY", "VOTE", "ST.E.STRONG.SYS",
        "LD.E.STRONG.SYS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            op = m.group(1)
            funcs[cur]["instructions"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    funcs[cur][k] += 1
    names = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
    print("# SASS evidence, round 2 (tools/sass_counts.py: cuobjdump -sass phynetpy_b200/libphynet_b200.so; built by phynetpy_b200/build.py:")
    print("# nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo).  Counts of the mnemonics that prove the Blackwell-native paths, per kernel:")
    print("# UBLKCP = cp.async.bulk (TMA 1-D bulk copy global->shared); SYNCS.* = mbarrier; LDG/STG.E.ENL2.256 = 256-bit global access;")
    print("# LDS.128 = ld.shared.v2.f64; LD.E / ST.E = generic accesses (0 in the shared-memory variants of the K-m kernels); ATOMS = shared atomics")
    print("# (the K-m duplicate table); ST/LD.E.STRONG.SYS = the epoch flags of the in-kernel cross-GPU barrier; MATCH.ANY = __match_any_sync")
    print()
    for (mangled, c), name in zip(funcs.items(), names):
        print(name[:170])
        print("    instructions=%d  " % c["instructions"] + "  ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k]))


if __name__ == "__main__":
    sys.exit(main())
