"""CPU: the C-ABI library loads and exports every symbol include/qa_b200.h declares (no compute calls)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "qa_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qa_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    names = _declared()
    assert "qa_quant_recon" in names and "qa_tile_stats" in names and "qa_greedy_assign" in names
    assert len(names) >= 14


def test_library_exports_every_declared_symbol():
    from quantization_analysis_b200 import _lib
    _lib.ensure_built()
    L = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in _declared():
        assert hasattr(L, name), name
    assert set(_declared()) == set(_lib.EXPORTS)
    assert L.qa_version() >= 100


def test_no_cpu_fallback_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import numpy as np
    from quantization_analysis_b200 import _lib, quantization_formats as qf
    with pytest.raises(_lib.QaError):
        qf.quantize_weight_values(np.ones(16, np.float32), "bfp8")
