"""On-disk cache of quantised arrays per (algorithm, backend, format, tensor)
(layout of compression_algorithms/cache.py:17-104; disk I/O, not accelerated)."""
from __future__ import annotations

import hashlib
import re
from dataclasses import dataclass
from pathlib import Path

import numpy as np

from .tile_utils import counts_from_array, counts_to_array, format_tag


def _safe_tensor_key(tensor_name: str) -> str:
    """File-system-safe tensor key: sanitised name + sha1 prefix (hf_model_utils.py:121-126)."""
    digest = hashlib.sha1(tensor_name.encode("utf-8")).hexdigest()[:12]
    safe = re.sub(r"[^A-Za-z0-9._-]+", "_", tensor_name).strip("_") or "tensor"
    return f"{safe}--{digest}"


def _safe_float_tag(value: float) -> str:
    return f"{value:.6g}".replace("-", "m").replace(".", "p")


@dataclass
class CacheContext:
    root: Path
    tensor_name: str
    backend: str
    recompute: bool
    run_tag: str

    @property
    def safe_tensor(self) -> str:
        return _safe_tensor_key(self.tensor_name)

    def quant_path(self, compression: str, fmt: str) -> Path:
        return Path(self.root) / compression / self.backend / fmt / f"{self.safe_tensor}.npy"

    def mixed_path(self, compression, metric, threshold, cluster, k, iters, random_formats) -> Path:
        base = Path(self.root) / compression / f"run-{self.run_tag}" / f"metric-{metric}"
        thr = f"thr-{_safe_float_tag(threshold)}"
        if compression == "mixed-tile-random":
            return (base / f"iters-{iters}" / thr / f"formats-{format_tag(random_formats or [])}"
                    / self.backend / f"{self.safe_tensor}.npz")
        return base / thr / f"cluster-{cluster}" / f"k-{k}" / self.backend / f"{self.safe_tensor}.npz"

    def load_array(self, compression: str, fmt: str):
        if self.recompute:
            return None
        path = self.quant_path(compression, fmt)
        return np.load(path) if path.exists() else None

    def save_array(self, compression: str, fmt: str, y) -> None:
        path = self.quant_path(compression, fmt)
        path.parent.mkdir(parents=True, exist_ok=True)
        np.save(path, np.asarray(y))

    def load_mixed(self, path: Path):
        if self.recompute or not Path(path).exists():
            return None
        try:
            with np.load(path) as data:
                if not all(k in data for k in ("y", "counts", "assignment")):
                    return None
                return (np.asarray(data["y"], dtype=np.float32), counts_from_array(data["counts"]),
                        np.asarray(data["assignment"], dtype=np.int8))
        except Exception:
            return None

    def save_mixed(self, path: Path, y, counts: dict, assignment) -> None:
        Path(path).parent.mkdir(parents=True, exist_ok=True)
        np.savez(path, y=y, counts=counts_to_array(counts), assignment=np.asarray(assignment).astype(np.int8))
