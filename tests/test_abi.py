"""CPU: the C-ABI shared library loads and exports every symbol include/phynet_b200.h
declares.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

from phynetpy_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        from phynetpy_b200.build import build
        build()
    return _lib.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "phynet_b200.h")).read()
    return sorted(set(re.findall(r"PNB_API\s+(?:const\s+char\s*\*|int)\s*(pnb_\w+)\s*\(", text)))


def test_header_symbols_all_exported(lib):
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "symbol %s declared in the header is not exported" % name
    assert sorted(_lib.ABI_SYMBOLS) == declared


def test_abi_version_and_error_string(lib):
    assert lib.pnb_abi_version() == 1
    assert isinstance(lib.pnb_last_error(), bytes)


def test_no_cpu_fallback_without_device(lib):
    """Without a CUDA device the product path must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    rc = lib.pnb_ctx_create(0, ctypes.byref(h))
    assert rc == _lib.PNB_ERR_CUDA
    assert b"no CPU fallback" in lib.pnb_last_error()
    from phynetpy_b200.infer import FelsensteinCalculator, JC69, GeneTree, PhyNetB200Error
    fc = FelsensteinCalculator({"A": "ACGT", "B": "ACGA"}, device_compress=False)
    with pytest.raises(PhyNetB200Error):
        fc.log_likelihood(GeneTree.from_newick("(A:0.1,B:0.1);"), JC69())


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under phynetpy_b200/ may reference it."""
    pkg = os.path.join(ROOT, "phynetpy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower() or f == "build.py", "%s mentions the oracle" % f
