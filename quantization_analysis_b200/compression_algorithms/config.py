"""Compression-config JSON loader (schema and seed rules of compression_algorithms/config.py:17-69)."""
from __future__ import annotations

import json
from dataclasses import dataclass
from pathlib import Path


@dataclass
class CompressionConfig:
    algorithm: str
    params: dict
    quantization_formats: list[str] | None
    seed: int | None
    random_seed: bool


def _parse_seed(raw, random_seed: bool):
    """int -> fixed seed; 0 or "random" -> draw a seed at run time."""
    if raw is None:
        return None, random_seed
    if isinstance(raw, str) and raw.strip().lower() == "random":
        return None, True
    try:
        value = int(raw)
    except (TypeError, ValueError) as exc:
        raise ValueError("Compression config 'seed' must be an int, 0, or 'random'") from exc
    if value == 0:
        return None, True
    return value, random_seed


def load_compression_config(path: str | None) -> CompressionConfig:
    if path is None:
        return CompressionConfig("none", {}, None, None, False)
    p = Path(path)
    if not p.exists():
        raise FileNotFoundError(f"Compression config not found: {path}")
    doc = json.loads(p.read_text(encoding="utf-8"))
    if not isinstance(doc, dict):
        raise ValueError("Compression config must be a JSON object")
    params = doc.get("params", {})
    params = {} if params is None else params
    if not isinstance(params, dict):
        raise ValueError("Compression config 'params' must be an object")
    formats = doc.get("quantization_formats")
    if formats is not None:
        if not isinstance(formats, list):
            raise ValueError("Compression config 'quantization_formats' must be a list of strings")
        formats = [str(f).strip().lower() for f in formats if str(f).strip()] or None
    seed, random_seed = _parse_seed(doc.get("seed"), bool(doc.get("random_seed", False)))
    return CompressionConfig(
        algorithm=str(doc.get("algorithm", "none")).strip().lower(),
        params=params,
        quantization_formats=formats,
        seed=seed,
        random_seed=random_seed,
    )
