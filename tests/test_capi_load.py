"""CPU: the C-ABI library builds for sm_100a, loads without a GPU, exports every symbol include/sarasmg.h declares,
and fails loudly (no CPU fallback) when asked to create a solver without a CUDA device."""
import os
import re

import numpy as np
import pytest

from saras_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sarasmg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sarasmg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    lib = capi.load()
    decl = _declared_symbols()
    assert len(decl) >= 30
    missing = [s for s in decl if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(capi.SYMBOLS) == decl, "capi.SYMBOLS out of sync with include/sarasmg.h"
    assert lib.sarasmg_version() >= 100


def test_no_cpu_fallback():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    z = np.zeros(3)
    coefs = {k: z for k in ("hx", "hy", "hz", "hx2", "hz2", "hzhx", "hxhy", "hyhz", "hxhyhz")}
    with pytest.raises(capi.SarasMgError) as e:
        capi.Solver(dim=3, vc_depth=2, vc_count=1, pre_smooth=1, post_smooth=1, gs_smooth=0, res_type=2, mg_tolerance=1e-6,
                    x_per=0, y_per=0, z_per=0, n0=(17, 17, 17), coefs=coefs, metrics={})
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_touch_oracle():
    """nothing under saras_b200/ may import, link or load anything under oracle/"""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "saras_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"mg_oracle|mgoracle|sarasref|oracle/", text):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
