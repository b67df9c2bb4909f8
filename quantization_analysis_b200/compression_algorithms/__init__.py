"""Algorithm registry and factory (mirrors compression_algorithms/__init__.py:11-29)."""
from __future__ import annotations

from . import base, cache, config, metrics, quantizer, tile_utils                      # noqa: F401
from . import mixed_tile_greedy, mixed_tile_random, mixed_tile_threshold, none, transpose  # noqa: F401
from .base import CompressionAlgorithm, CompressionResult
from .config import CompressionConfig, load_compression_config
from .mixed_tile_greedy import MixedTileGreedyCompression
from .mixed_tile_random import MixedTileRandomCompression
from .mixed_tile_threshold import MixedTileThresholdCompression
from .none import NoneCompression
from .transpose import TransposeCompression

ALGORITHM_REGISTRY: dict[str, type[CompressionAlgorithm]] = {
    "none": NoneCompression,
    "transpose": TransposeCompression,
    "mixed-tile-greedy": MixedTileGreedyCompression,
    "mixed-tile-threshold": MixedTileThresholdCompression,
    "mixed-tile-random": MixedTileRandomCompression,
    "mixed-tile": MixedTileGreedyCompression,
}

__all__ = ["ALGORITHM_REGISTRY", "CompressionAlgorithm", "CompressionConfig", "CompressionResult",
           "create_algorithm", "load_compression_config"]


def create_algorithm(name: str, params: dict | None = None) -> CompressionAlgorithm:
    cls = ALGORITHM_REGISTRY.get(name.strip().lower())
    if cls is None:
        raise ValueError(f"Unsupported compression algorithm '{name}'. "
                         f"Supported: {', '.join(sorted(ALGORITHM_REGISTRY))}")
    return cls.from_params(params or {})
