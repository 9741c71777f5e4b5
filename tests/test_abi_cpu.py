"""CPU-side checks of the C-ABI library: it loads, exports every symbol the header declares,
and fails loudly (never silently falls back) when no GPU is present."""
import ctypes
import os
import re

import pytest

from mcphylo_jl_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_every_declared_symbol():
    with open(os.path.join(ROOT, "include", "mcphylo_b200.h")) as fh:
        header = fh.read()
    declared = set(re.findall(r"\b(mcp_[a-z_]+)\s*\(", header))
    assert declared == set(capi.SYMBOLS)
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert capi.load().mcp_abi_version() == 2


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.McpError) as ei:
        capi.Context(0)
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle():
    """The product package must never route through oracle/."""
    pkg = os.path.join(ROOT, "mcphylo.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                with open(os.path.join(dirpath, f)) as fh:
                    src = fh.read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, re.M), f
                assert "liboracle" not in src, f
