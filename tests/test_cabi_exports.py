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


def test_batch_descriptor_layout_and_host_helpers():
    """qa_batch_desc as bound by ctypes has the header's layout (3 pointers + 5 int64 = 64 bytes, no padding), and the host-side
    helpers of the descriptor-array entry points answer without a GPU."""
    from quantization_analysis_b200 import _lib
    assert ctypes.sizeof(_lib.BatchDesc) == 64
    assert [f[0] for f in _lib.BatchDesc._fields_] == ["x", "table", "init", "rows", "cols", "ld", "item_begin", "block_begin"]
    text = (ROOT / "include" / "qa_b200.h").read_text()
    body = re.search(r"typedef struct qa_batch_desc \{(.*?)\} qa_batch_desc;", text, flags=re.S).group(1)
    assert re.findall(r"\b(x|table|init|rows|cols|ld|item_begin|block_begin)\b", body) == \
        ["x", "table", "init", "rows", "cols", "ld", "item_begin", "block_begin"]
    L = _lib.lib()
    assert L.qa_tile_stats_items(2048, 7168) == 64 * 14            # 32-row stripes x 512-column chunks
    assert L.qa_tile_stats_items(45, 77) == 2 and L.qa_tile_stats_items(0, 5) == 0
    assert L.qa_greedy_init_bytes(1) > 256


def test_synthetic_fp8_checkpoint_round_trip():
    """synthetic.fp8_checkpoint_cpu: block scale grid = ceil(shape / block), inverse scale = block amax / 448, and the dequantized
    tensor (torch float8 -> float32, times the block scale: hf_model_utils.py:199-215) is within half an e4m3 step of the source."""
    import torch
    from quantization_analysis_b200 import synthetic
    w, inv = synthetic.fp8_checkpoint_cpu((300, 520), 7, block=(128, 128))
    assert w.dtype == torch.uint8 and tuple(w.shape) == (300, 520) and tuple(inv.shape) == (3, 5) and inv.dtype == torch.float32
    g = torch.Generator(device="cpu")
    g.manual_seed(7)
    x = torch.randn((300, 520), generator=g, dtype=torch.float32) * 0.02
    scale = inv.repeat_interleave(128, 0).repeat_interleave(128, 1)[:300, :520]
    deq = w.view(torch.float8_e4m3fn).float() * scale
    assert torch.isfinite(deq).all()
    assert (deq.abs() <= 448 * scale * (1 + 1e-6)).all()
    # e4m3: 3 mantissa bits -> relative half-step 2^-4 for normals; absolute half-step 2^-10 * scale in the subnormal range
    assert ((deq - x).abs() <= torch.maximum(x.abs() * 2.0 ** -4, scale * 2.0 ** -10) * (1 + 1e-5)).all()
