"""CPU test of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/pansvr_b200.h declares; compute entry points fail loudly without a GPU (no fallback)."""
import os
import re

import numpy as np
import pytest
import torch

from pansvr_b200 import build, ksw, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "pansvr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pansvr_\w+|ksw_extd2_sse)\s*\(", text)) - {"pansvr_ksw_ctx"})


def test_library_builds_and_exports_header_symbols():
    build.build()
    lib = ksw.load_library()
    names = declared_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pansvr_b200.h but not exported"
    assert set(names) == set(ksw.EXPORTS)


def test_band_cells_is_host_side_and_matches_definition():
    assert ksw.band_cells(150, 1100, 100) == 25100
    assert ksw.band_cells(250, 280, 500) == 70000
    assert ksw.band_cells(0, 10, 5) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(RuntimeError, match="pansvr_ksw_create failed"):
        ksw.KswContext(0)
