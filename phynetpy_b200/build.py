"""Build recipe for libphynet_b200.so (sm_100a only, in-tree).

``python -m phynetpy_b200.build`` or ``__graft_entry__.build()``.  nvcc
cross-compiles without a GPU; the .so is git-ignored but travels with the tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libphynet_b200.so")
SOURCES = ["pnb_runtime.cu", "pnb_compile.cu", "pnb_eval.cu", "pnb_batch.cu", "pnb_compress.cu", "pnb_msnc.cu", "pnb_candidates.cu", "pnb_simulate.cu"]
HEADERS = ["pnb_device.cuh", "pnb_host.hpp", "pnb_msnc_core.hpp", os.path.join("..", "..", "include", "phynet_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-fvisibility=hidden",
    "-shared",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=...)")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, out: str = LIB, defs=None) -> str:
    """Compile every source into one shared library.  ``out`` / ``defs`` build a tuning variant next to the
    product library (``tools/variants.py``; loaded through ``PNB_LIB=<path>``)."""
    if out == LIB and not force and not needs_build():
        return LIB
    if defs is None:
        defs = os.environ.get("PNB_NVCC_DEFS", "").split()  # tuning knobs, e.g. "-DPNB_PPT=1 -DPNB_MIN_BLOCKS=8"
    cmd = [_nvcc()] + NVCC_FLAGS + list(defs) + (["-Xptxas", "-v"] if verbose else []) + \
        [os.path.join(CSRC, s) for s in SOURCES] + ["-o", out + ".tmp"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    os.replace(out + ".tmp", out)
    if verbose:
        sys.stderr.write(proc.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
