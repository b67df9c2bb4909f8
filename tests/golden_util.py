"""Helpers to read the committed golden fixtures (made by tests/golden/make_golden.py)."""
from __future__ import annotations

import json
from functools import lru_cache
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"
FORMATS = ["bf16", "bfp8", "bfp4", "bfp2", "fp0"]
MIXED = ["bf16", "bfp8", "bfp4", "bfp2"]


@lru_cache(maxsize=None)
def npz(name: str):
    return dict(np.load(GOLDEN / name))


@lru_cache(maxsize=None)
def js(name: str):
    return json.loads((GOLDEN / name).read_text())


def f32_from_bits(bits: np.ndarray, shape) -> np.ndarray:
    return np.ascontiguousarray(bits, dtype=np.uint32).view(np.float32).reshape(tuple(int(s) for s in shape))


def kat_cases():
    z = npz("kat_formats.npz")
    names = sorted(k[:-4] for k in z if k.endswith("__in"))
    return names


def kat(name: str):
    z = npz("kat_formats.npz")
    shape = z[f"{name}__shape"]
    x = f32_from_bits(z[f"{name}__in"], shape)
    outs = {f: z[f"{name}__{f}"].reshape(-1) for f in FORMATS}
    return x, outs


def algo_case_names():
    z = npz("algo_small.npz")
    return sorted(k[:-4] for k in z if k.endswith("__in"))


def algo_input(name: str) -> np.ndarray:
    z = npz("algo_small.npz")
    return f32_from_bits(z[f"{name}__in"], z[f"{name}__shape"])


def algo_runs(name: str):
    """Yield (key, meta, assignment, y_bits) for every recorded reference run on `name`."""
    z, meta = npz("algo_small.npz"), js("algo_small.json")
    for key, m in meta.items():
        if key.startswith(name + "__run"):
            yield key, m, z[f"{key}__assignment"], z[f"{key}__y"]


def bits(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32)).view(np.uint32).reshape(-1)
