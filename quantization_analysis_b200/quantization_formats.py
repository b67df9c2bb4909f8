"""Drop-in for the reference's ``quantization_formats`` module (quantization_formats.py:8,29-45,84-194),
backed by the sm_100a kernels behind include/qa_b200.h.  No NumPy arithmetic happens here.

Inputs may be NumPy float32 arrays (returns float32 NumPy arrays, like the reference) or torch
tensors on the GPU (returns a bf16 torch tensor on the GPU: every reconstruction is bf16-exact).
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine

SUPPORTED_FORMATS = ["mxfp4", "nvfp4", "bf16", "bfp8", "bfp4", "bfp2", "fp0"]   # reference order (:8)
_MANT_TO_FMT = {7: "bfp8", 3: "bfp4", 1: "bfp2"}


def _is_torch(x) -> bool:
    return isinstance(x, torch.Tensor)


def _run(x, fmt: str):
    if _is_torch(x):
        if x.numel() == 0:
            return x.to(torch.bfloat16)
        p = engine.prepare_rows(x)
        return engine.quant_recon(p, [fmt])[fmt].reshape(p.shape)
    x = np.asarray(x, dtype=np.float32)
    if x.size == 0:
        return x.astype(np.float32)
    p = engine.prepare_rows(x)
    y = engine.quant_recon(p, [fmt])[fmt]
    out = engine.result_to_numpy(p, y)
    if x.ndim == 0:
        return np.array(out.reshape(()), dtype=np.float32)
    return out


def quantize_dequantize_bf16(x):
    """fp32 -> bf16 (round to nearest even on the bit pattern) -> fp32 (:44-45)."""
    return _run(x, "bf16")


def fp32_to_bf16_round_to_nearest_even(x) -> np.ndarray:
    """uint16 bf16 patterns (:29-35)."""
    y = quantize_dequantize_bf16(np.asarray(x, dtype=np.float32))
    return (np.ascontiguousarray(y).view(np.uint32) >> np.uint32(16)).astype(np.uint16)


def bf16_to_fp32(bf16) -> np.ndarray:
    """Pure bit re-interpretation (:38-41); no arithmetic."""
    b = np.asarray(bf16, dtype=np.uint16)
    return (b.astype(np.uint32) << np.uint32(16)).view(np.float32)


def quantize_dequantize_bfp_ttnn(x, mant_bits: int):
    """TTNN-style BFP quantize->dequantize, shared exponent per 16-element row group (:84-164)."""
    fmt = _MANT_TO_FMT.get(int(mant_bits))
    if fmt is None:
        raise ValueError(f"Unsupported mant_bits for the device path: {mant_bits} (supported: 7, 3, 1)")
    return _run(x, fmt)


def quantize_fp0(x):
    """All zeros (:167-168)."""
    if _is_torch(x):
        return torch.zeros(x.shape, dtype=torch.bfloat16, device=x.device)
    return np.zeros(np.shape(x), dtype=np.float32)


def quantize_weight_values(x, fmt: str):
    """Format dispatch (:171-194)."""
    fmt = fmt.lower()
    if fmt == "bf16":
        return quantize_dequantize_bf16(x)
    if fmt == "bfp8":
        return quantize_dequantize_bfp_ttnn(x, mant_bits=7)
    if fmt == "bfp4":
        return quantize_dequantize_bfp_ttnn(x, mant_bits=3)
    if fmt == "bfp2":
        return quantize_dequantize_bfp_ttnn(x, mant_bits=1)
    if fmt == "fp0":
        return quantize_fp0(x)
    if fmt in ("mxfp4", "nvfp4"):
        return _scalar_proxy(x, fmt)
    raise ValueError(f"Unsupported weight format: {fmt}")


def _scalar_proxy(x, fmt: str):
    """mxfp4 / nvfp4 scalar proxies (:171-183): elementwise on the device (qa_scalar_proxy), float32 out."""
    which = 0 if fmt == "mxfp4" else 1
    if _is_torch(x):
        xd = x.contiguous()
        if xd.dtype not in (torch.bfloat16, torch.float32):
            xd = xd.to(torch.float32)
        out = torch.empty(xd.shape, dtype=torch.float32, device=xd.device)
        if xd.numel():
            engine.scalar_proxy(xd, which, out)
        return out
    xn = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    if xn.size == 0:
        return xn.astype(np.float32)
    xd = torch.from_numpy(xn.reshape(-1)).to(engine._require_cuda())
    out = torch.empty_like(xd)
    engine.scalar_proxy(xd, which, out)
    return out.cpu().numpy().reshape(xn.shape)
