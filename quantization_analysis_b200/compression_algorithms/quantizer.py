"""Backend switch (compression_algorithms/quantizer.py:8-34).  Only ``emulation`` exists here:
it lands in the CUDA kernels; the Tenstorrent ``ttnn`` round-trip is out of scope."""
from __future__ import annotations

from ..quantization_formats import quantize_weight_values


class Quantizer:
    def __init__(self, backend: str = "emulation", ttnn=None) -> None:
        self.backend = backend
        self.ttnn = ttnn

    def quantize(self, xf, fmt: str):
        if self.backend == "ttnn":
            raise RuntimeError("The ttnn backend is not available in the B200 build (emulation only).")
        return quantize_weight_values(xf, fmt.lower())
