"""The C-ABI library loads and exports every symbol include/bs2e.h declares;
without a GPU the compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import bs2e

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "bs2e.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(bs2e_[A-Za-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 30
    raw = C.CDLL(bs2e.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/bs2e.h but not exported"
    assert sorted(bs2e.ABI) == names, "python binding table out of sync with the header"


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    grid = bs2e.generate_grid(4, 2, 1, 1.0, 5.0)
    with pytest.raises(bs2e.Bs2eError):
        bs2e.Context(4, grid, 2, 7)
    msg = bs2e.lib().bs2e_last_error().decode()
    assert "bs2e_ctx_create" in msg


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "b-spline-two-e_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".hpp", ".f90")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower(), f"{f} mentions the oracle"
                assert "hostcheck" not in txt.replace("tests/hostcheck", "").lower() or f in ("core.h", "slater_core.h", "geom_host.h"), f
