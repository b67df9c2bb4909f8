"""Drop-in for compression_algorithms/tile_utils.py: format tables, byte model, tiling helpers.

The byte model and the reshape helpers are host bookkeeping (no hot-path arithmetic);
``tile_metrics`` runs the NumPy-faithful device scorer (qa_tile_scores_f32).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import engine

MIXED_TILE_FORMATS = ["bf16", "bfp8", "bfp4", "bfp2"]                      # tile_utils.py:8
MIXED_TILE_BYTES_PER_ELEM = {"bf16": 2.0, "bfp8": 1.088, "bfp4": 0.50097, "bfp2": 0.25097}  # :9-14
_METRICS = ("pcc", "mae", "atol")


def counts_to_array(counts: dict) -> np.ndarray:
    return np.asarray([counts.get(k, 0) for k in MIXED_TILE_FORMATS], dtype=np.int64)


def counts_from_array(values) -> dict:
    v = np.asarray(values, dtype=np.int64).reshape(-1)
    if v.size != len(MIXED_TILE_FORMATS):
        raise ValueError("Invalid mixed-tile counts payload.")
    return {k: int(v[i]) for i, k in enumerate(MIXED_TILE_FORMATS)}


def assignment_to_array(assignment) -> np.ndarray:
    return np.asarray(assignment, dtype=np.int8)


def mixed_tile_total_bytes(counts: dict, tile_hw: int = 32) -> float:
    """Sum in dict order of float(count) * tile elements * bytes/elem (tile_utils.py:32-37)."""
    per_tile = float(tile_hw * tile_hw)
    acc = 0.0
    for name, c in counts.items():
        acc += float(c) * per_tile * MIXED_TILE_BYTES_PER_ELEM.get(name, 0.0)
    return acc


def format_tag(formats) -> str:
    return "+".join(formats) if formats else "none"


def reshape_to_2d_with_padding(xf):
    """n-d -> zero-padded [rows x32, cols x32] (tile_utils.py:91-115).  Pure data movement."""
    xf = np.asarray(xf, dtype=np.float32)
    if xf.ndim == 0:
        flat, info = xf.reshape(1, 1), ("scalar", xf.shape)
    elif xf.ndim == 1:
        n = xf.shape[0]
        flat = np.zeros((int(np.ceil(n / 32.0)), 32), dtype=np.float32)
        flat.reshape(-1)[:n] = xf
        info = ("vector", n)
    else:
        flat, info = xf.reshape(int(np.prod(xf.shape[:-1])), xf.shape[-1]), ("nd", xf.shape)
    h, w = flat.shape
    hp, wp = int(np.ceil(h / 32.0)) * 32, int(np.ceil(w / 32.0)) * 32
    out = np.zeros((hp, wp), dtype=np.float32)
    out[:h, :w] = flat
    return out, info, (h, w, hp, wp)


def reconstruct_from_tiles(tiles, shape_info, pad_info, tile_hw: int = 32):
    """Inverse of the tiling (tile_utils.py:118-132)."""
    h, w, hp, wp = pad_info
    grid = np.asarray(tiles).reshape(hp // tile_hw, wp // tile_hw, tile_hw, tile_hw)
    flat = grid.transpose(0, 2, 1, 3).reshape(hp, wp)[:h, :w]
    kind = shape_info[0]
    if kind == "scalar":
        return np.array(flat[0, 0], dtype=np.float32)
    if kind == "vector":
        return flat.reshape(-1)[: shape_info[1]].astype(np.float32)
    if kind == "nd":
        return flat.reshape(shape_info[1]).astype(np.float32)
    raise ValueError("Invalid shape_info")


def tile_metrics(ref_tiles, q_tiles, metric: str) -> np.ndarray:
    """Per-tile float32 score of two [N,32,32] tile stacks (tile_utils.py:46-57) on the GPU, with NumPy's float32
    summation orders (qa_tile_scores_pair_f32): any pair of operands, like the reference."""
    if metric not in _METRICS:
        raise ValueError(f"Unsupported metric: {metric}")
    ref = np.ascontiguousarray(np.asarray(ref_tiles, dtype=np.float32))
    q = np.ascontiguousarray(np.asarray(q_tiles, dtype=np.float32))
    n = ref.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.float32)
    if ref.shape != q.shape or ref.size != n * 1024:
        raise ValueError("tile_metrics expects two [N, 32, 32] stacks of equal shape")
    return engine.tile_scores_pair(ref, q)[_METRICS.index(metric)]


def global_metric(xf, tiles, shape_info, pad_info, metric: str) -> float:
    """metric_value of the un-tiled reconstruction against xf (tile_utils.py:135-137)."""
    from .metrics import metric_value
    return metric_value(xf, reconstruct_from_tiles(tiles, shape_info, pad_info), metric)


def kmeans_1d(values, k: int, max_iters: int = 25, seed: int = 0):
    """1-D Lloyd iterations from quantile seeds (tile_utils.py:60-88; an unused helper of the reference, host-side
    bookkeeping like there).  -> (labels int32 [n], centroids float32 [k])."""
    v = np.asarray(values, dtype=np.float32).reshape(-1)
    if v.size == 0:
        return np.zeros((0,), dtype=np.int32), np.zeros((0,), dtype=np.float32)
    k = max(1, min(k, v.size))
    if k == 1:
        return np.zeros((v.size,), dtype=np.int32), np.array([float(np.mean(v))], dtype=np.float32)
    cent = np.quantile(v, np.linspace(0.0, 1.0, k, dtype=np.float32))
    rng = np.random.default_rng(seed)
    labels = np.zeros(v.size, dtype=np.int64)
    for _ in range(max_iters):
        labels = np.argmin(np.abs(v[:, None] - cent[None, :]), axis=1)
        nxt = cent.copy()
        for c in range(k):
            members = v[labels == c]
            nxt[c] = float(np.mean(members)) if members.size else float(v[rng.integers(0, v.size)])
        done = np.allclose(nxt, cent, rtol=0.0, atol=1e-6)
        cent = nxt
        if done:
            break
    return labels.astype(np.int32), cent.astype(np.float32)
