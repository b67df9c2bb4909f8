"""Quantize x.T and transpose back: shared exponents run along axis 0
(reference: compression_algorithms/transpose.py:13-33).  On the device this is the column-group kernel
``qa_quant_recon_cols`` (a thread owns 16 consecutive rows of one column; loads and stores stay coalesced across the
warp), so no transposed copy of the tensor is ever made."""
from __future__ import annotations

import numpy as np
import torch

from .. import engine
from .base import CompressionAlgorithm, CompressionResult
from .none import quantize_all


def quantize_all_transposed(xf, formats) -> dict:
    """{fmt: reconstruction of shape xf.shape} with the BFP groups along axis 0."""
    is_t = isinstance(xf, torch.Tensor)
    ndim = xf.dim() if is_t else np.asarray(xf).ndim
    n = int(xf.numel()) if is_t else int(np.asarray(xf).size)
    if ndim < 2 or n == 0:
        return quantize_all(xf, formats)              # x.T is x: ordinary row groups
    fmts = [f.lower() for f in formats]
    elementwise = [f for f in fmts if f not in engine.FMT_INDEX]       # fp0 / mxfp4 / nvfp4 do not depend on the grouping
    out = quantize_all(xf, elementwise) if elementwise else {}
    recon, shape = engine.quant_recon_cols(xf, [f for f in fmts if f in engine.FMT_INDEX])
    for f, y in recon.items():
        out[f] = y.reshape(shape) if is_t else y.to(torch.float32).cpu().numpy().reshape(shape)
    return out


class TransposeCompression(CompressionAlgorithm):
    name = "transpose"

    def run(self, xf, formats, quantizer=None, cache=None):
        if quantizer is not None and getattr(quantizer, "backend", "emulation") != "emulation":
            raise RuntimeError("the ttnn backend is not available in this build (emulation only)")
        is_t = isinstance(xf, torch.Tensor)
        results, missing, cached = [], [], {}
        for fmt in formats:
            y = cache.load_array(self.name, fmt) if cache is not None else None
            if y is not None and tuple(y.shape) == tuple(xf.shape):
                cached[fmt] = y
            else:
                missing.append(fmt)
        fresh = quantize_all_transposed(xf, missing) if missing else {}
        for fmt in formats:
            y = cached.get(fmt)
            if y is None:
                y = fresh[fmt.lower()]
                if cache is not None and not is_t:
                    cache.save_array(self.name, fmt, y)
            results.append(CompressionResult(fmt=fmt.upper(), compression=self.name, y=y))
        return results
