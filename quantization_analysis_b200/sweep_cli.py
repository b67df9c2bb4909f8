"""Threshold-sweep CLI (scripts/sweep_mixed_tile_threshold.py:36-102, 581-841 of the reference): same arguments, same
``details/<tensor>/sweep_config.json`` + ``sweep_results.csv``; the plots (matplotlib) are out of scope.

    python -m quantization_analysis_b200.sweep_cli <repo_or_url> <tensor_name> [--no-regex] [--metric pcc|mae|atol]
           [--lowest-metric-val V] [--steps N] [--formats bf16,bfp8,bfp4,bfp2] [--out-dir DIR]
"""
from __future__ import annotations

import argparse
import fnmatch
import re
import sys
import time
from pathlib import Path

import numpy as np

from . import sweep, tensor_source
from .compression_algorithms.tile_utils import MIXED_TILE_FORMATS


def _parse_formats(value: str) -> list[str]:
    out = []
    for part in (p.strip().lower() for p in value.split(",")):
        if not part:
            continue
        if part not in MIXED_TILE_FORMATS:
            raise ValueError(f"Unsupported mixed-tile format: {part}")
        if part not in out:
            out.append(part)
    if not out:
        raise ValueError("No valid mixed-tile formats selected.")
    return out


def select_tensors(index, query: str, use_regex: bool) -> list[str]:
    """sweep:313-345: regex search (default), exact name, fnmatch pattern, then the wq filter."""
    names = list(index.tensor_to_file)
    weight_like = [n for n in names if "weight" in n.lower() and not n.lower().endswith("_scale_inv")]
    cand = weight_like if weight_like else names
    if use_regex:
        try:
            pat = re.compile(query)
        except re.error as exc:
            raise RuntimeError(f"Invalid regex '{query}': {exc}") from exc
        hits = [n for n in cand if pat.search(n)]
        if hits:
            return sorted(hits)
        raise RuntimeError("No tensors matched the regex query.")
    if query in cand:
        return [query]
    if any(ch in query for ch in "*?[]"):
        hits = [n for n in cand if fnmatch.fnmatch(n, query)]
        if hits:
            return sorted(hits)
    return tensor_source.filter_tensor_names(cand, query)


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(description="Sweep mixed-tile-threshold over a range of metric thresholds.")
    ap.add_argument("repo_or_url")
    ap.add_argument("tensor_name")
    ap.add_argument("--regex", action="store_true", default=True)
    ap.add_argument("--no-regex", dest="regex", action="store_false")
    ap.add_argument("--list-matches", action="store_true")
    ap.add_argument("--revision", default="main")
    ap.add_argument("--cache-dir", default="data/hf-cache")
    ap.add_argument("--backend", choices=["emulation", "ttnn"], default="emulation")
    ap.add_argument("--formats", default="bf16,bfp8,bfp4,bfp2")
    ap.add_argument("--metric", choices=["pcc", "mae", "atol"], default="pcc")
    ap.add_argument("--lowest-metric-val", type=float, default=0.9)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--out-dir", default=None)
    return ap


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    formats = _parse_formats(args.formats)
    if args.backend == "ttnn":
        raise RuntimeError("TTNN backend requires `ttnn` in the active Python environment.")
    index = tensor_source.build_tensor_index(args.repo_or_url, args.revision, args.cache_dir)
    selected = select_tensors(index, args.tensor_name, args.regex)
    if not selected:
        print("error: no tensors matched the filter query")
        return 1
    if args.list_matches:
        print(f"Matched {len(selected)} tensor(s):")
        for n in selected:
            print(f"  {n}")
        return 0
    base = Path(args.out_dir) if args.out_dir else (Path("results") / index.repo_id.replace("/", "__") / "mixed_tile_threshold_sweep"
                                                    / time.strftime("%Y%m%d-%H%M%S"))
    detail = base / "details"
    detail.mkdir(parents=True, exist_ok=True)
    for name in selected:
        xf = np.asarray(index.load_fp32(name), dtype=np.float32)
        out = detail / name.replace("/", "_").replace(".", "_")
        try:
            rows, _maps = sweep.sweep_tensor(xf, formats, args.metric, args.steps, args.lowest_metric_val)
        except sweep.SweepRangeError as exc:
            print(f"error: {exc}")
            return 1
        sweep.write_sweep_config(out, args.repo_or_url, name, args.revision, args.backend, formats, args.metric,
                                 args.lowest_metric_val, args.steps)
        sweep.write_sweep_csv(out, rows, formats)
    print(f"Wrote sweep results to {base}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
