"""Plug-in contract shared by all compression algorithms (mirrors compression_algorithms/base.py:13-44)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Any, Iterable


@dataclass
class CompressionResult:
    fmt: str                       # upper-case format name, or "MIXED"
    compression: str               # algorithm name
    y: Any                         # reconstruction: float32 ndarray (numpy in) or bf16 CUDA tensor (torch in)
    tile_counts: dict | None = None
    tile_bytes: float | None = None
    meta: dict | None = None


class CompressionAlgorithm(ABC):
    name: str

    def __init__(self, params: dict | None = None) -> None:
        self.params = params or {}

    @classmethod
    def from_params(cls, params: dict | None = None) -> "CompressionAlgorithm":
        return cls(params=params or {})

    def expected_evals(self, formats: Iterable[str]) -> int:
        return len(list(formats))

    @abstractmethod
    def run(self, xf, formats, quantizer, cache) -> list[CompressionResult]:
        raise NotImplementedError
