"""B200-native quantize-and-score path behind the reference's Python API.

Sub-modules mirror the reference's module names (``quantization_formats``,
``compression_algorithms.*``); ``install_drop_in()`` registers them under the reference's
top-level names so unmodified reference scripts (``wq``, the sweep script) import them.
"""
from __future__ import annotations

import sys

__all__ = ["install_drop_in"]


def install_drop_in() -> None:
    """Alias this package's modules to the reference's top-level module names."""
    from . import quantization_formats as _qf
    from . import compression_algorithms as _ca

    sys.modules["quantization_formats"] = _qf
    sys.modules["compression_algorithms"] = _ca
    for name in ("base", "cache", "config", "metrics", "quantizer", "tile_utils", "none",
                 "mixed_tile_greedy", "mixed_tile_threshold", "mixed_tile_random", "transpose"):
        mod = getattr(_ca, name, None)
        if mod is not None:
            sys.modules[f"compression_algorithms.{name}"] = mod
