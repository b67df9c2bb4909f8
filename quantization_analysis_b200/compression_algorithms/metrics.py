"""Drop-in for compression_algorithms/metrics.py (:6-39).

``pearson_corr`` / ``metric_value`` run on the GPU: the pair (a, b) is reduced to the seven
sums {sx, sx2, sy, sy2, sxy, s|a-b|, max|a-b|} by `qa_pair_sums` (float64, fixed reduction tree)
and recombined in float64.  That is the exact value of the formula the reference evaluates in float32 (the
reference's own result is only ~1e-5 accurate at 1e7 elements).
"""
from __future__ import annotations

from .. import engine


def _pair_sums(a, b):
    return engine.pair_sums(a, b)


def pearson_corr(a, b) -> float:
    s, n = _pair_sums(a, b)
    if n == 0:
        return 1.0
    return engine.metrics_from_sums(s, n)["pcc"]


def metric_value(a, b, metric: str) -> float:
    if metric not in ("pcc", "mae", "atol"):
        raise ValueError(f"Unsupported metric: {metric}")
    s, n = _pair_sums(a, b)
    if n == 0:
        return 1.0 if metric == "pcc" else float("nan")
    return engine.metrics_from_sums(s, n)[metric]


def metric_is_good(value: float, metric: str, threshold: float) -> bool:
    if metric == "pcc":
        return bool(value >= threshold)
    return bool(value <= threshold)


def metric_better(a: float, b: float, metric: str) -> bool:
    if metric == "pcc":
        return bool(a > b)
    return bool(a < b)
