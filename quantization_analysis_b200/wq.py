"""`wq` for the B200 build: per-tensor loop, timing, scoring and output tree of the reference CLI
(wq:549-884) on the device path.  Tensors come from the synthetic DeepSeek-R1 provider (no Hugging
Face access here); the hot region (wq:679-709: algo.run + whole-tensor scoring) runs on the GPU and the
scores are float64 recombinations of the tile-stat table (no reconstruction leaves the device unless a
writer needs it).

    python -m quantization_analysis_b200.wq [filter ...] --compression-config cfg.json [--limit N] [--out DIR]
"""
from __future__ import annotations

import argparse
import csv
import json
import re
import secrets
import sys
import time
from datetime import datetime
from pathlib import Path

import numpy as np
import torch

from . import engine, synthetic
from .compression_algorithms import create_algorithm, load_compression_config
from .compression_algorithms.tile_utils import MIXED_TILE_FORMATS
from .quantization_formats import SUPPORTED_FORMATS

FORMAT_BYTES_PER_ELEM = {"mxfp4": 0.5, "nvfp4": 0.5, "bf16": 2.0, "bfp8": 1.088, "bfp4": 0.50097, "bfp2": 0.25097,
                         "fp0": 0.0}                                           # wq:132-140
IN_SCOPE = ["bf16", "bfp8", "bfp4", "bfp2", "fp0"]


def _slug(s: str) -> str:
    return re.sub(r"[^a-zA-Z0-9._-]+", "_", s).strip("_") or "tensor"


def filter_tensor_names(names, query):
    """Substring, or dotted prefix path (hf_model_utils.py:60-77)."""
    if not query or not query.strip():
        return sorted(names)
    q = query.strip()
    if "." in q:
        qp = [p.lower() for p in q.split(".") if p]
        return sorted(n for n in names if n.lower().split(".")[: len(qp)] == qp)
    return sorted(n for n in names if q.lower() in n.lower())


def resolve_format_list(values, supported):
    """hf_model_utils.py:317-335."""
    if not values:
        return list(supported)
    out = []
    for raw in values:
        v = raw.strip().lower()
        if v == "all":
            out += [s for s in supported if s not in out]
            continue
        if v not in supported:
            raise ValueError(f"Unsupported format '{raw}'. Supported: {', '.join(supported)}, all")
        if v not in out:
            out.append(v)
    return out


def resolve_seed(config, algo_params: dict):
    """Seed precedence of wq:553-586."""
    used, source = None, None
    if config.seed is not None:
        used, source = (secrets.randbits(31), "random") if int(config.seed) == 0 else (int(config.seed), "config")
    elif config.random_seed:
        used, source = secrets.randbits(31), "random"
    if used is not None:
        algo_params["seed"] = used
    elif "seed" in algo_params:
        try:
            v = int(algo_params["seed"])
        except (TypeError, ValueError):
            return algo_params["seed"], "params"
        used, source = (secrets.randbits(31), "random") if v == 0 else (v, "params")
        algo_params["seed"] = used
    return used, source


def _mapping(assignment: np.ndarray) -> dict:
    return {"tile_hw": 32, "format_to_int": {f: i for i, f in enumerate(MIXED_TILE_FORMATS)},
            "int_to_format": MIXED_TILE_FORMATS, "assignment_shape": list(assignment.shape)}


def write_assignment(out_dir: Path, algo_dir: str, tensor_name: str, assignment: np.ndarray) -> None:
    """assignment.npy + assignment_mapping.json (wq:295-316)."""
    d = out_dir / algo_dir / _slug(tensor_name)
    d.mkdir(parents=True, exist_ok=True)
    np.save(d / "assignment.npy", assignment.astype(np.int8))
    (d / "assignment_mapping.json").write_text(json.dumps(_mapping(assignment), indent=2))


def write_random_outputs(out_dir: Path, tensor_name: str, samples, tile_formats, assignment) -> None:
    """<slug>.csv + <slug>_assignment.npy + mapping (wq:151-194)."""
    d = out_dir / "mixed_tile_random"
    d.mkdir(parents=True, exist_ok=True)
    slug = _slug(tensor_name)
    with (d / f"{slug}.csv").open("w", newline="", encoding="utf-8") as f:
        wr = csv.writer(f)
        wr.writerow(["sample_id", *[f"{fmt}_tiles" for fmt in tile_formats], "total_gb", "pcc", "mae", "atol"])
        for s in samples:
            wr.writerow([s["id"], *[s["counts"].get(fmt, 0) for fmt in tile_formats], float(s["total_bytes"]) / 1e9,
                         s["pcc"], s["mae"], s["atol"]])
    if assignment is not None:
        np.save(d / f"{slug}_assignment.npy", assignment.astype(np.int8))
        (d / f"{slug}_assignment_mapping.json").write_text(json.dumps(_mapping(assignment), indent=2))


def evaluate_tensor(name: str, x_dev: torch.Tensor, algorithms, formats, out_dir: Path | None):
    """The hot region of wq:679-709 for one tensor: returns table rows (dicts)."""
    rows = []
    p = engine.prepare_tiles(x_dev)
    table = None
    for algo in algorithms:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if algo.name in ("none", "transpose"):
            mixed = [f for f in formats if f in engine.FMT_INDEX]
            if algo.name == "none":
                if table is None:
                    table = engine.tile_stats(p, engine.MIXED_FORMATS)
                results = []
                for f in formats:
                    if f == "fp0":       # all zeros (quantize_fp0): scored against b = 0 (metrics.py:14-15 gives pcc 0)
                        sums, n = engine.pair_sums(x_dev, None)
                        results.append((f.upper(), engine.metrics_from_sums(sums, n), None))
                    elif f in mixed:
                        sums = engine.assignment_sums(table, None, engine.FMT_INDEX[f]).cpu().numpy()
                        results.append((f.upper(), engine.metrics_from_sums(sums, p.numel), None))
            else:
                res = algo.run(x_dev, formats, None, None)
                results = []
                for r in res:
                    sums, n = engine.pair_sums(x_dev, r.y)
                    results.append((r.fmt, engine.metrics_from_sums(sums, n), None))
            torch.cuda.synchronize()
            elapsed = time.perf_counter() - t0
            for fmt, m, _ in results:
                rows.append({"tensor": name, "fmt": fmt, "compression": algo.name, **m, "time_s": elapsed,
                             "gb": p.numel * FORMAT_BYTES_PER_ELEM.get(fmt.lower(), 0.0) / 1e9, "tile_counts": None})
            continue
        tile_formats = getattr(algo, "tile_formats", None) or getattr(algo, "formats", None) or \
            algo._filter_from_formats(formats)
        if table is None:
            table = engine.tile_stats(p, engine.MIXED_FORMATS)
        dr = algo.run_prepared(p, tile_formats, table=table)
        torch.cuda.synchronize()
        elapsed = time.perf_counter() - t0
        rows.append({"tensor": name, "fmt": "MIXED", "compression": algo.name, **dr.metrics, "time_s": elapsed,
                     "gb": dr.tile_bytes / 1e9, "tile_counts": dr.counts})
        if out_dir is not None:
            a = dr.assignment_numpy()
            if algo.name == "mixed-tile-random":
                write_random_outputs(out_dir, name, dr.meta["samples"], dr.tile_formats, a)
            else:
                write_assignment(out_dir, algo.name.replace("-", "_"), name, a)
    return rows


def format_rows(rows) -> list[str]:
    lines = [f"{'FMT':<6} {'COMP':<22} {'PCC':>10} {'MAE':>12} {'ATOL':>12} {'TIME(s)':>9} {'GB':>10}  TILES"]
    for r in rows:
        tc = "" if not r["tile_counts"] else " ".join(f"{k}:{v}" for k, v in r["tile_counts"].items())
        lines.append(f"{r['fmt']:<6} {r['compression']:<22} {r['pcc']:>10.5f} {r['mae']:>12.3e} {r['atol']:>12.3e} "
                     f"{r['time_s']:>9.4f} {r['gb']:>10.6f}  {tc}")
    return lines


def run(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="wq", description="Weight quantization analyzer (B200 device path, synthetic weights).")
    ap.add_argument("filter_query", nargs="*", help="Optional filter: substring, or dotted torch-style prefix path.")
    ap.add_argument("--limit", type=int, default=None)
    ap.add_argument("--backend", choices=["emulation", "ttnn"], default="emulation")
    ap.add_argument("--compression-config", type=str, default=None)
    ap.add_argument("--synthetic-seed", type=int, default=1000)
    ap.add_argument("--out", type=str, default="results")
    args = ap.parse_args(argv)
    if args.backend == "ttnn":
        print("error: the ttnn backend is not available in the B200 build", file=sys.stderr)
        return 1
    config = load_compression_config(args.compression_config)
    algo_params = dict(config.params)
    used_seed, seed_source = resolve_seed(config, algo_params)
    selected = create_algorithm(config.algorithm, algo_params)
    baseline = create_algorithm("none", {})
    algorithms = [baseline] if selected.name == "none" else [baseline, selected]
    formats = [f for f in resolve_format_list(config.quantization_formats, SUPPORTED_FORMATS) if f in IN_SCOPE]
    names = filter_tensor_names(list(synthetic.DEEPSEEK_R1_SHAPES), " ".join(args.filter_query).strip() or None)
    if args.limit is not None:
        names = names[: max(0, args.limit)]
    if not names:
        print("No tensors matched.", file=sys.stderr)
        return 1
    run_tag = datetime.now().strftime("%Y%m%d-%H%M%S")
    out_dir = Path(args.out) / "synthetic__DeepSeek-R1-shapes" / selected.name / run_tag
    out_dir.mkdir(parents=True, exist_ok=True)
    used = {"algorithm": config.algorithm, "quantization_formats": formats,
            "params": {k: v for k, v in algo_params.items() if not (k == "seed" and used_seed is not None)}}
    if used_seed is not None:
        used.update(seed=used_seed, seed_source=seed_source)
    (out_dir / "compression_config.used.json").write_text(json.dumps(used, indent=2))
    print(f"synthetic DeepSeek-R1 shapes - {len(names)} tensors\nformats: {', '.join(formats)}\n"
          f"compression: {', '.join(a.name for a in algorithms)}")
    lines = []
    for i, name in enumerate(names):
        x = synthetic.randn_bf16_cpu(synthetic.DEEPSEEK_R1_SHAPES[name], args.synthetic_seed + i).cuda()
        rows = evaluate_tensor(name, x, algorithms, formats, out_dir)
        block = [name, f"  shape={tuple(x.shape)} numel={x.numel()}"] + format_rows(rows)
        print("\n".join(block))
        lines += block
    (out_dir / "table.txt").write_text("\n".join(lines) + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(run())
