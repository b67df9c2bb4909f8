"""Threshold sweep on the device (core of scripts/sweep_mixed_tile_threshold.py:623-790).

Per tensor: one pass for the NumPy-faithful per-tile scores, one for the tile-stat table; then all
thresholds are assigned in ONE launch (`qa_threshold_assign`, index into the ascending-bytes order with the
last format forced, :145-155) and every distinct assignment is scored from the table.
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine
from .compression_algorithms.mixed_tile_threshold import formats_by_precision
from .compression_algorithms.tile_utils import MIXED_TILE_FORMATS, mixed_tile_total_bytes

_ROW = {"pcc": 0, "mae": 1, "atol": 2}


def sweep_thresholds(scores_metric: torch.Tensor, tile_formats, metric: str, steps: int, lowest: float) -> np.ndarray:
    """np.linspace(max score of the highest-precision format, lowest, steps) as float32 (sweep:659-670)."""
    order = formats_by_precision(tile_formats)
    top = float(scores_metric[engine.FMT_INDEX[order[-1]]].max().item())
    return np.linspace(top, lowest, steps, dtype=np.float32)


def sweep_tensor(x, tile_formats=MIXED_TILE_FORMATS, metric: str = "pcc", steps: int = 32, lowest: float = 0.9,
                 thresholds=None):
    """-> list of dict rows (threshold, counts, total_bytes, pcc, mae, atol) and the int8 maps [steps, ntiles]."""
    p = engine.prepare_tiles(x)
    scores = engine.tile_scores(p, tile_formats)[_ROW[metric]].contiguous()
    table = engine.tile_stats(p, MIXED_TILE_FORMATS)
    if thresholds is None:
        thresholds = sweep_thresholds(scores, tile_formats, metric, steps, lowest)
    order = formats_by_precision(tile_formats)
    maps, counts = engine.threshold_assign(scores, order, metric == "pcc", thresholds)
    counts = counts.cpu().numpy()
    sums = engine.assignment_sums_batch(table, maps).cpu().numpy()     # every threshold's map scored in one launch
    rows = []
    for i, thr in enumerate(thresholds):
        # (the reference reuses the previous row when the assignment did not change, sweep:736-742: same numbers)
        m = engine.metrics_from_sums(sums[i], p.numel)
        c = {f: int(counts[i, j]) for j, f in enumerate(MIXED_TILE_FORMATS)}
        rows.append({"threshold": float(thr), "counts": c, "total_bytes": mixed_tile_total_bytes(c), **m})
    return rows, maps


def write_sweep_csv(out_dir, rows, formats=MIXED_TILE_FORMATS):
    """``sweep_results.csv`` in the reference's layout (scripts/sweep_mixed_tile_threshold.py:792-797):
    step,threshold,size_bytes,pcc,mae,atol,<fmt>_tiles..."""
    from pathlib import Path
    out = Path(out_dir)
    out.mkdir(parents=True, exist_ok=True)
    headers = ["step", "threshold", "size_bytes", "pcc", "mae", "atol", *[f"{fmt}_tiles" for fmt in formats]]
    with (out / "sweep_results.csv").open("w", encoding="utf-8") as f:
        f.write(",".join(headers) + "\n")
        for i, r in enumerate(rows):
            vals = {"step": i, "threshold": float(r["threshold"]), "size_bytes": r["total_bytes"], "pcc": r["pcc"], "mae": r["mae"],
                    "atol": r["atol"], **{f"{fmt}_tiles": r["counts"].get(fmt, 0) for fmt in formats}}
            f.write(",".join(str(vals.get(h, "")) for h in headers) + "\n")
    return out / "sweep_results.csv"
