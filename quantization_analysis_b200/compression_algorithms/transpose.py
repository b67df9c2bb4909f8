"""Quantize x.T and transpose back: shared exponents run along columns
(reference: compression_algorithms/transpose.py:13-33).  The transpose is device data movement."""
from __future__ import annotations

import numpy as np
import torch

from .base import CompressionAlgorithm, CompressionResult
from .none import quantize_all


class TransposeCompression(CompressionAlgorithm):
    name = "transpose"

    def run(self, xf, formats, quantizer=None, cache=None):
        is_t = isinstance(xf, torch.Tensor)
        results, missing, cached = [], [], {}
        for fmt in formats:
            y = cache.load_array(self.name, fmt) if cache is not None else None
            if y is not None and tuple(y.shape) == tuple(xf.shape):
                cached[fmt] = y
            else:
                missing.append(fmt)
        fresh = {}
        if missing:
            if is_t:
                xt = xf.permute(*reversed(range(xf.dim()))).contiguous()
            else:
                xt = np.ascontiguousarray(np.transpose(np.asarray(xf, dtype=np.float32)))
            for f, yt in quantize_all(xt, missing).items():
                fresh[f] = yt.permute(*reversed(range(yt.dim()))) if is_t else np.transpose(yt)
        for fmt in formats:
            y = cached.get(fmt)
            if y is None:
                y = fresh[fmt.lower()]
                if cache is not None and not is_t:
                    cache.save_array(self.name, fmt, y)
            results.append(CompressionResult(fmt=fmt.upper(), compression=self.name, y=y))
        return results
