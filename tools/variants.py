#!/usr/bin/env python
"""Build tuning variants of libphynet_b200.so in the build container (nvcc cross-compiles; no GPU minutes),
each loadable through PNB_LIB=<path>.  `python tools/variants.py name:-DPNB_PPT=4,-DPNB_TPB=64 ...`
writes phynetpy_b200/variants/lib_<name>.so and prints the pruning kernel's register / spill line."""
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phynetpy_b200 import build as B  # noqa: E402

OUT = os.path.join(ROOT, "phynetpy_b200", "variants")


def one(spec):
    name, _, defs = spec.partition(":")
    defs = [d for d in defs.split(",") if d]
    out = os.path.join(OUT, "lib_%s.so" % name)
    cmd = [B._nvcc()] + B.NVCC_FLAGS + defs + ["-Xptxas", "-v"] + [os.path.join(B.CSRC, s) for s in B.SOURCES] + ["-o", out]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        return "%s: BUILD FAILED\n%s" % (name, p.stderr[-2000:])
    lines = p.stderr.splitlines()
    info = []
    for i, l in enumerate(lines):
        m = re.search(r"pnb_prune_kernel_tILb(\d)ELb(\d)", l)
        if m and "Compiling" in l:
            info.append("fused=%s rescale=%s: %s | %s" % (m.group(1), m.group(2), lines[i + 2].strip(), lines[i + 3].strip().replace("ptxas info    : ", "")))
    return "%s %s\n  " % (name, " ".join(defs)) + "\n  ".join(info)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    with ThreadPoolExecutor(4) as ex:
        for r in ex.map(one, sys.argv[1:]):
            print(r, flush=True)
