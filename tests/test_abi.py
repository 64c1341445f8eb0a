"""The C-ABI library loads on a CPU-only box and exports every symbol include/gemma_b200.h
declares; without a GPU the context constructor fails loudly (no CPU fallback)."""
import os
import re

import pytest

import gemma_b200
from gemma_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    h = open(os.path.join(ROOT, "include", "gemma_b200.h")).read()
    return sorted(set(re.findall(r"GB200_API[^;(]*?(gb200_[a-z0-9_]+)\s*\(", h)))


def test_header_and_binding_agree():
    assert _declared() == sorted(api.SIGNATURES)


def test_library_exports_every_declared_symbol():
    lib = gemma_b200.load_library()
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.gb200_abi_version() == 1


def test_sumstat_layout_is_reference_sumstat():
    # SUMSTAT src/param.h:54-66: 8 doubles in this order
    assert api.SUMSTAT_DTYPE.itemsize == 64
    assert api.SUMSTAT_DTYPE.names == ("beta", "se", "lambda_remle", "lambda_mle", "p_wald", "p_lrt", "p_score",
                                       "logl_H1")


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(gemma_b200.GB200Error):
        gemma_b200.Context(0)


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under gemma_b200/ may import, link or load it."""
    pkg = os.path.join(ROOT, "gemma_b200")
    pat = re.compile(r"(import\s+oracle|from\s+oracle|from\s+\.\.?oracle|libgemma_oracle|oracle/)")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), "%s references the oracle" % f
