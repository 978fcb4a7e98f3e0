"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/andi_b200.h
declares, computes the anchor threshold on the host like the reference, and fails loudly (no
fallback) when asked to compute without a CUDA device."""
import re
from pathlib import Path

import pytest

import oracle
from andi_b200 import native

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    L = native.load()
    header = (ROOT / "include" / "andi_b200.h").read_text()
    declared = set(re.findall(r"\b(andi_[a-z_0-9]+)\s*\(", header))
    assert declared == set(native.ABI_SYMBOLS), declared ^ set(native.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name


def test_threshold_is_the_reference_formula():
    # src/sequence.c:296-373; pinned through the oracle (itself pinned to the reference)
    for p in (0.025, 0.1, 0.001):
        for gc in (0.3, 0.41, 0.5, 0.62):
            for l in (101, 2001, 200001, 4200001, 10000001, 240000001):
                assert native.threshold(p, gc, l) == oracle.lib().orc_min_anchor_length(p, gc, l)
    assert native.threshold(0.025, 0.5, 200001) == 12  # config 1


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(native.AndiError):
        native.Context(0)


def test_product_never_imports_the_oracle():
    for path in (ROOT / "andi_b200").rglob("*"):
        if path.suffix in (".py", ".cu", ".cuh", ".c", ".h", ".cpp"):
            text = path.read_text()
            assert "import oracle" not in text and "from oracle" not in text, path
            assert "andi_oracle" not in text and "orc_" not in text, path


def test_library_exports_the_reference_named_symbols():
    """include/andi_compat.h: the reference's own C surface (src/esa.h:61-64, src/process.c:141,
    src/process.h:11, src/dist_hack.h:34) -- what a relinked reference program binds."""
    L = native.load()
    header = (ROOT / "include" / "andi_compat.h").read_text()
    declared = set(re.findall(r"^\w[\w \*]*?\b(\w+)\s*\([^;{]*\);", header, flags=re.M))
    assert {"esa_init", "esa_free", "get_match", "get_match_cached", "dist_anchor", "distMatrix", "distMatrixLM",
            "calculate_distances", "andi_compat_set_model"} <= declared, declared
    for name in declared:
        assert hasattr(L, name), name
