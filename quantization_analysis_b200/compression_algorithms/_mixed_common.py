"""Shared host logic of the three mixed-tile algorithms: parameter parsing, device results."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from .. import engine
from .tile_utils import MIXED_TILE_FORMATS, mixed_tile_total_bytes

VALID_METRICS = {"pcc", "mae", "atol"}


def parse_formats(value) -> list[str]:
    """'bfp8,bfp4' or ['bfp8', ...] -> de-duplicated list of mixed-tile formats, order kept."""
    if value is None or value == "":
        return []
    if isinstance(value, str):
        items = value.split(",")
    elif isinstance(value, list):
        items = [str(v) for v in value]
    else:
        raise ValueError("formats must be a comma-separated string or a list of strings")
    out: list[str] = []
    for raw in items:
        name = raw.strip().lower()
        if not name:
            continue
        if name not in MIXED_TILE_FORMATS:
            raise ValueError(f"Unsupported mixed-tile format: {name}")
        if name not in out:
            out.append(name)
    return out


def filter_formats(formats, algo_name: str) -> list[str]:
    allowed = [f for f in formats if f in MIXED_TILE_FORMATS]
    if not allowed:
        raise ValueError(
            f"{algo_name} requires at least one of {', '.join(MIXED_TILE_FORMATS)} in quantization_formats")
    return allowed


@dataclass
class DeviceResult:
    """What one algorithm run leaves on the device (nothing is copied to the host unless asked)."""
    compression: str
    prepared: engine.Prepared
    assignment: torch.Tensor | None        # int8 [ntiles] (device) for mixed results
    counts: dict
    tile_bytes: float
    metrics_exact: dict                    # float64 recombination of pcc / mae / atol from the tile-stat table
    tile_formats: list[str]
    meta: dict = field(default_factory=dict)
    _y: torch.Tensor | None = None
    _metrics: dict | None = None

    @property
    def metrics(self) -> dict:
        """pcc / mae / atol of the result as the reference reports them (wq:684-687): float32, NumPy's summation orders
        (qa_tensor_scores_f32 over x and the materialised reconstruction)."""
        if self._metrics is None:
            p = self.prepared
            s = engine.tensor_scores_f32(p.data, self.y_device(), n=p.numel)[0]
            self._metrics = {"pcc": float(s[0]), "mae": float(s[1]), "atol": float(s[2])}
        return self._metrics

    def y_device(self) -> torch.Tensor:
        """bf16 reconstruction on the device (materialised on first use)."""
        if self._y is None:
            self._y = engine.apply_assignment(self.prepared, self.assignment)
        return self._y

    def assignment_numpy(self) -> np.ndarray:
        p = self.prepared
        return self.assignment.cpu().numpy().astype(np.int8).reshape(p.tiles_h, p.tiles_w)


def counts_dict(counts_dev) -> dict:
    c = counts_dev.cpu().numpy().astype(np.int64).reshape(-1)
    return {f: int(c[i]) for i, f in enumerate(MIXED_TILE_FORMATS)}


def empty_result(xf):
    """Empty input (mixed_tile_greedy.py:78-83)."""
    return np.asarray(xf, dtype=np.float32), {f: 0 for f in MIXED_TILE_FORMATS}, np.zeros((1, 1), dtype=np.int8)


def is_torch(x) -> bool:
    return isinstance(x, torch.Tensor)


def numel_of(x) -> int:
    return int(x.numel()) if is_torch(x) else int(np.asarray(x).size)


def finish(dr: DeviceResult, xf, extra_meta: dict | None = None):
    """DeviceResult -> the reference's (y, counts, assignment) triple in the caller's array type."""
    y_dev = dr.y_device()
    if is_torch(xf):
        p = dr.prepared
        y = y_dev[: p.numel].reshape(p.shape) if p.kind == "vector" else y_dev.reshape(p.shape)
    else:
        y = engine.result_to_numpy(dr.prepared, y_dev)
    return y, dr.counts, dr.assignment_numpy()


def total_bytes(counts: dict) -> float:
    return mixed_tile_total_bytes(counts)
