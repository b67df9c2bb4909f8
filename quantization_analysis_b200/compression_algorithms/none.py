"""Baseline: quantize the whole tensor per format (reference: compression_algorithms/none.py:13-31).
All formats that are not served from the on-disk cache come out of ONE fused device pass."""
from __future__ import annotations

import numpy as np
import torch

from .. import engine
from .base import CompressionAlgorithm, CompressionResult


def quantize_all(xf, formats) -> dict:
    """{fmt: reconstruction} for every requested format with a single read of xf."""
    is_t = isinstance(xf, torch.Tensor)
    fmts = [f.lower() for f in formats]
    proxies = ("mxfp4", "nvfp4")                      # elementwise scalar proxies: their own kernel (qa_scalar_proxy)
    for f in fmts:
        if f not in engine.FMT_INDEX and f != "fp0" and f not in proxies:
            raise ValueError(f"Unsupported weight format: {f}")
    n = int(xf.numel()) if is_t else int(np.asarray(xf).size)
    out = {}
    if n == 0:
        for f in fmts:
            out[f] = xf.to(torch.float32 if f in proxies else torch.bfloat16) if is_t else np.asarray(xf, dtype=np.float32)
        return out
    p = engine.prepare_rows(xf)
    recon = engine.quant_recon(p, [f for f in fmts if f in engine.FMT_INDEX])
    for f in fmts:
        if f in proxies:
            from ..quantization_formats import quantize_weight_values
            out[f] = quantize_weight_values(xf, f)
        elif f == "fp0":
            out[f] = (torch.zeros(p.shape, dtype=torch.bfloat16, device=p.data.device) if is_t
                      else np.zeros(p.shape, dtype=np.float32))
        elif is_t:
            out[f] = recon[f].reshape(p.shape)
        else:
            out[f] = engine.result_to_numpy(p, recon[f])
    return out


class NoneCompression(CompressionAlgorithm):
    name = "none"

    def run(self, xf, formats, quantizer=None, cache=None):
        if quantizer is not None and getattr(quantizer, "backend", "emulation") != "emulation":
            raise RuntimeError("the ttnn backend is not available in this build (emulation only)")
        cached = {}
        if cache is not None:
            for fmt in formats:
                y = cache.load_array(self.name, fmt)
                if y is not None and tuple(y.shape) == tuple(xf.shape):
                    cached[fmt] = y
        missing = [f for f in formats if f not in cached]
        fresh = quantize_all(xf, missing) if missing else {}
        results = []
        for fmt in formats:
            if fmt in cached:
                y = cached[fmt]
            else:
                y = fresh[fmt.lower()]
                if cache is not None and not isinstance(y, torch.Tensor):
                    cache.save_array(self.name, fmt, y)
            results.append(CompressionResult(fmt=fmt.upper(), compression=self.name, y=y))
        return results
