"""Drop-in for compression_algorithms/metrics.py (:6-39).

``pearson_corr`` / ``metric_value`` run on the GPU and return the reference's own float32 numbers: `qa_tensor_scores_f32`
evaluates metrics.py:6-27 with NumPy's pairwise float32 sums and OpenBLAS's sdot accumulation order, so
the values agree with the reference bit for bit (its float32 result is ~1e-5 away from the exact correlation at 1e7
elements; ``pearson_corr_exact`` gives the float64 recombination for callers that want the mathematically exact value).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import engine


def _dev(a) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        t = a.detach()
        if t.dtype not in (torch.bfloat16, torch.float32):
            t = t.float()
        return t.to(engine._require_cuda()).contiguous().reshape(-1)
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32)).reshape(-1)).to(engine._require_cuda())


def _scores(a, b) -> np.ndarray | None:
    ta, tb = _dev(a), _dev(b)
    if ta.numel() != tb.numel():
        raise ValueError("operands could not be broadcast together: sizes differ")
    if ta.numel() == 0:
        return None
    return engine.tensor_scores_f32(ta, tb)[0]


def pearson_corr(a, b) -> float:
    s = _scores(a, b)
    return 1.0 if s is None else float(s[0])


def pearson_corr_exact(a, b) -> float:
    """float64 recombination of the same formula (qa_pair_sums): the value the float32 evaluation approximates."""
    s, n = engine.pair_sums(a, b)
    return 1.0 if n == 0 else engine.metrics_from_sums(s, n)["pcc"]


def metric_value(a, b, metric: str) -> float:
    if metric not in ("pcc", "mae", "atol"):
        raise ValueError(f"Unsupported metric: {metric}")
    s = _scores(a, b)
    if s is None:
        if metric == "pcc":
            return 1.0
        if metric == "mae":
            return float("nan")
        raise ValueError("zero-size array to reduction operation maximum which has no identity")
    return float(s[("pcc", "mae", "atol").index(metric)])


def metric_is_good(value: float, metric: str, threshold: float) -> bool:
    if metric == "pcc":
        return bool(value >= threshold)
    return bool(value <= threshold)


def metric_better(a: float, b: float, metric: str) -> bool:
    if metric == "pcc":
        return bool(a > b)
    return bool(a < b)
