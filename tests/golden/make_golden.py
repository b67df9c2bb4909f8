#!/usr/bin/env python3
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference has no tests or golden vectors of its own, so these files
*are* the pin for the oracle and, through it, for the CUDA path.  Nothing here is read
from /root/reference at test time; the .npz/.json outputs are committed.
"""
from __future__ import annotations

import hashlib
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

import quantization_formats as ref_qf  # noqa: E402  (reference)
from compression_algorithms import create_algorithm  # noqa: E402  (reference)
from compression_algorithms.metrics import pearson_corr  # noqa: E402
from compression_algorithms.quantizer import Quantizer  # noqa: E402
from compression_algorithms.tile_utils import tile_metrics, reshape_to_2d_with_padding  # noqa: E402

from quantization_analysis_b200 import synthetic  # noqa: E402

FORMATS = ["bf16", "bfp8", "bfp4", "bfp2", "fp0"]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bits(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32)).view(np.uint32)


def kat_inputs() -> dict[str, np.ndarray]:
    """Known-answer inputs as float32 arrays (mostly bf16-exact; a few full-fp32)."""
    rng = np.random.default_rng(20261018)
    f = lambda hexes: np.array([int(h, 16) for h in hexes.split()], dtype=np.uint32).view(np.float32)  # noqa: E731
    cases: dict[str, np.ndarray] = {}
    cases["appendix_c_row"] = f(
        "3f800000 3f810000 3f820000 3f000000 3f010000 bf400000 3b000000 38000000 "
        "bfff0000 3fff0000 00000000 80000000 00800000 3eaa0000 beaa0000 3c000000")
    cases["inf_nan_group"] = f("7f800000 3f800000 7fc00000 ff800000 " + "00000000 " * 12)
    cases["tiny_exponents"] = f("00800000 01000000 00c00000 00800000 " + "00000000 " * 12)
    cases["denormals"] = f("00000001 00400000 807fffff 00800000 3f800000 " + "00000000 " * 11)
    face = np.full((32, 32), 0.01171875, dtype=np.float32)
    face[:, :16] = 1.0
    face[0, 0] = 64.0
    cases["face_structure"] = face
    cases["ragged_1d_40"] = np.arange(1, 41, dtype=np.float32)
    cases["ragged_2x20"] = np.arange(1, 41, dtype=np.float32).reshape(2, 20)
    cases["scalar"] = np.array(0.3, dtype=np.float32)
    # random bf16 bit patterns with controlled exponent spread per 16-group
    for name, spread in (("rand_bf16_spread4", 4), ("rand_bf16_spread12", 12), ("rand_bf16_spread40", 40)):
        e0 = rng.integers(60, 180, size=(24, 5, 1))
        e = np.clip(e0 - rng.integers(0, spread + 1, size=(24, 5, 16)), 0, 254)
        m = rng.integers(0, 128, size=(24, 5, 16))
        s = rng.integers(0, 2, size=(24, 5, 16))
        u = ((s << 31) | (e << 23) | (m << 16)).astype(np.uint32)
        cases[name] = u.view(np.float32).reshape(24, 80)
    # all 65536 bf16 patterns, one per element (groups mix everything incl. inf/nan/denormal)
    allb = (np.arange(65536, dtype=np.uint32) << 16)
    cases["all_bf16_patterns"] = rng.permutation(allb).view(np.float32).reshape(64, 1024)
    # genuine fp32 (not bf16-exact) data: exercises truncating alignment + full-width rounding
    cases["rand_fp32"] = (rng.standard_normal((40, 70)) * 0.02).astype(np.float32)
    u = rng.integers(0, 2**32, size=(16, 96), dtype=np.uint64).astype(np.uint32)
    cases["rand_fp32_bits"] = u.view(np.float32)
    cases["ragged_3d"] = (rng.standard_normal((2, 5, 37)) * 3).astype(np.float32)
    return cases


def make_kats() -> None:
    out = {}
    with np.errstate(all="ignore"):
        for name, x in kat_inputs().items():
            out[f"{name}__in"] = bits(x)
            out[f"{name}__shape"] = np.asarray(x.shape, dtype=np.int64)
            for fmt in FORMATS:
                out[f"{name}__{fmt}"] = bits(ref_qf.quantize_weight_values(x, fmt))
    np.savez_compressed(HERE / "kat_formats.npz", **out)
    print("kat_formats.npz:", len(out), "arrays")


def algo_cases() -> dict[str, np.ndarray]:
    c = {}
    c["het_96x160"] = synthetic.heterogeneous_f32_np((96, 160), 11)
    c["het_70x45"] = synthetic.heterogeneous_f32_np((70, 45), 12)
    c["het_1d_1536"] = synthetic.heterogeneous_f32_np((1536,), 13)
    c["het_1d_1000"] = synthetic.heterogeneous_f32_np((1000,), 14)
    c["het_3d_2x40x64"] = synthetic.heterogeneous_f32_np((2, 40, 64), 15)
    c["randn_64x256"] = synthetic.randn_f32_np((64, 256), 16)
    c["het_256x512"] = synthetic.heterogeneous_f32_np((256, 512), 17)
    z = synthetic.heterogeneous_f32_np((64, 64), 18)
    z[:32, :32] = 0.0          # an all-zero tile
    z[32:, 32:] = 0.25         # a constant tile (denominator == 0 branch)
    c["zero_and_const_tiles"] = z
    return c


ALGO_RUNS = [
    ("mixed-tile-greedy", {"metric": "pcc", "threshold": 0.999, "seed": 123}),
    ("mixed-tile-greedy", {"metric": "pcc", "threshold": 0.99, "seed": 7}),
    ("mixed-tile-greedy", {"metric": "mae", "threshold": 2e-4, "seed": 5}),
    ("mixed-tile-greedy", {"metric": "atol", "threshold": 0.02, "seed": 9}),
    ("mixed-tile-greedy", {"metric": "pcc", "threshold": 0.995, "seed": 3, "formats": "bfp8,bfp4,bfp2"}),
    ("mixed-tile-threshold", {"metric": "pcc", "threshold": 0.999}),
    ("mixed-tile-threshold", {"metric": "pcc", "threshold": 0.94}),
    ("mixed-tile-threshold", {"metric": "mae", "threshold": 1e-3}),
    ("mixed-tile-threshold", {"metric": "atol", "threshold": 0.01}),
    ("mixed-tile-random", {"metric": "pcc", "threshold": 0.99, "iters": 6, "seed": 42}),
    ("mixed-tile-random", {"metric": "mae", "threshold": 1e-3, "iters": 5, "seed": 43, "formats": "bfp8,bfp4"}),
]


def run_tag(algo: str, params: dict) -> str:
    return algo + "|" + json.dumps(params, sort_keys=True)


def make_algos() -> None:
    quantizer = Quantizer(backend="emulation")
    arrays, meta = {}, {}
    for cname, x in algo_cases().items():
        arrays[f"{cname}__in"] = bits(x)
        arrays[f"{cname}__shape"] = np.asarray(x.shape, dtype=np.int64)
        # padded-tile f32 scores, all metrics x mixed formats
        padded, _s, _p = reshape_to_2d_with_padding(x)
        th, tw = padded.shape[0] // 32, padded.shape[1] // 32
        tiles = padded.reshape(th, 32, tw, 32).transpose(0, 2, 1, 3).reshape(-1, 32, 32)
        for fmt in ["bf16", "bfp8", "bfp4", "bfp2"]:
            yq = quantizer.quantize(x, fmt)
            pq, _s2, _p2 = reshape_to_2d_with_padding(yq)
            tq = pq.reshape(th, 32, tw, 32).transpose(0, 2, 1, 3).reshape(-1, 32, 32)
            for metric in ("pcc", "mae", "atol"):
                arrays[f"{cname}__tilescore__{fmt}__{metric}"] = np.asarray(
                    tile_metrics(tiles, tq, metric), dtype=np.float32)
        for ri, (algo, params) in enumerate(ALGO_RUNS):
            a = create_algorithm(algo, dict(params))
            res = a.run(xf=x, formats=FORMATS, quantizer=quantizer, cache=None)[0]
            key = f"{cname}__run{ri}"
            arrays[f"{key}__assignment"] = np.asarray(res.meta["assignment"], dtype=np.int8)
            arrays[f"{key}__y"] = bits(res.y)
            d = np.abs(x - res.y)
            m = {"algo": algo, "params": params, "counts": res.tile_counts, "tile_bytes": res.tile_bytes,
                 "pcc_f32": pearson_corr(x, res.y), "mae_f32": float(np.mean(d)), "atol_f32": float(np.max(d))}
            if "samples" in res.meta:
                m["samples"] = res.meta["samples"]
            meta[key] = m
    np.savez_compressed(HERE / "algo_small.npz", **arrays)
    (HERE / "algo_small.json").write_text(json.dumps(meta, indent=1))
    print("algo_small:", len(arrays), "arrays,", len(meta), "runs")


def make_cfg1() -> None:
    """Config 1/2 shape [1536,7168] (q_a_proj), synthetic seed 0: `none` metrics + greedy/threshold maps."""
    x = synthetic.randn_f32_np((1536, 7168), 0)
    quantizer = Quantizer(backend="emulation")
    meta = {"input_sha256": sha(x), "shape": list(x.shape), "seed": 0, "none": {}}
    arrays = {}
    for fmt in FORMATS:
        y = quantizer.quantize(x, fmt)
        d = np.abs(x - y)
        meta["none"][fmt] = {"pcc_f32": pearson_corr(x, y), "mae_f32": float(np.mean(d)),
                             "atol_f32": float(np.max(d)), "y_sha256": sha(y)}
    g = create_algorithm("mixed-tile-greedy", {"metric": "pcc", "threshold": 0.999, "seed": 123})
    res = g.run(xf=x, formats=FORMATS, quantizer=quantizer, cache=None)[0]
    arrays["greedy_pcc0999_seed123"] = np.asarray(res.meta["assignment"], dtype=np.int8)
    meta["greedy_pcc0999_seed123"] = {"counts": res.tile_counts, "tile_bytes": res.tile_bytes,
                                      "y_sha256": sha(res.y), "pcc_f32": pearson_corr(x, res.y)}
    t = create_algorithm("mixed-tile-threshold", {"metric": "pcc", "threshold": 0.9937})
    res = t.run(xf=x, formats=FORMATS, quantizer=quantizer, cache=None)[0]
    arrays["threshold_pcc09937"] = np.asarray(res.meta["assignment"], dtype=np.int8)
    meta["threshold_pcc09937"] = {"counts": res.tile_counts, "tile_bytes": res.tile_bytes, "y_sha256": sha(res.y)}
    np.savez_compressed(HERE / "cfg1_q_a_proj.npz", **arrays)
    (HERE / "cfg1_q_a_proj.json").write_text(json.dumps(meta, indent=1))
    print("cfg1:", json.dumps(meta["greedy_pcc0999_seed123"]["counts"]), json.dumps(meta["threshold_pcc09937"]["counts"]))


def make_rng() -> None:
    """NumPy Generator(PCG64) streams the assignment algorithms consume."""
    out = {}
    for seed in (1, 123, 2**31 - 1):
        r = np.random.default_rng(seed)
        out[f"s{seed}__perm_10752"] = r.permutation(10752).astype(np.int32)
        out[f"s{seed}__perm_777"] = r.permutation(777).astype(np.int32)
        out[f"s{seed}__int4_1000"] = r.integers(0, 4, size=1000, dtype=np.int64).astype(np.int8)
        out[f"s{seed}__int3_1000"] = r.integers(0, 3, size=1000, dtype=np.int64).astype(np.int8)
        out[f"s{seed}__perm_4096"] = r.permutation(4096).astype(np.int32)
        out[f"s{seed}__int2_777"] = r.integers(0, 2, size=777, dtype=np.int64).astype(np.int8)
        out[f"s{seed}__perm_2"] = r.permutation(2).astype(np.int32)
        out[f"s{seed}__perm_1"] = r.permutation(1).astype(np.int32)
        out[f"s{seed}__perm_65537"] = r.permutation(65537).astype(np.int32)
    np.savez_compressed(HERE / "numpy_rng.npz", **out)
    print("numpy_rng.npz:", len(out))


def make_scalar_proxies() -> None:
    """mxfp4 / nvfp4 scalar proxies (quantization_formats.py:171-183): every bf16 bit pattern and 20 000 random float32
    values over 32 octaves, through the reference's per-element Python loop."""
    x16 = (np.arange(65536, dtype=np.uint32) << 16).view(np.float32)
    rng = np.random.default_rng(5)
    xr = (rng.standard_normal(20000) * np.exp2(rng.integers(-20, 12, 20000))).astype(np.float32)
    out = {"rand__in": bits(xr)}
    with np.errstate(all="ignore"):
        for fmt in ("mxfp4", "nvfp4"):
            out[f"bf16__{fmt}"] = bits(ref_qf.quantize_weight_values(x16, fmt))
            out[f"rand__{fmt}"] = bits(ref_qf.quantize_weight_values(xr, fmt))
    np.savez_compressed(HERE / "scalar_proxies.npz", **out)
    print("scalar_proxies.npz:", len(out), "arrays")


def make_fp8_dequant() -> None:
    """hf_model_utils._dequantize_tensor_with_scale_inv on every e4m3fn pattern and on a ragged 300x520 tensor with
    128x128 blocks (scale shape 3x5), through torch's float8_e4m3fn."""
    import torch
    import hf_model_utils as ref_hf
    rng = np.random.default_rng(11)
    out = {}
    allb = np.arange(256, dtype=np.uint8).reshape(16, 16)
    s1 = np.array([[0.0123456]], dtype=np.float32)
    out["all__w"], out["all__s"] = allb, s1
    out["all__out"] = bits(ref_hf._dequantize_tensor_with_scale_inv(torch.from_numpy(allb).view(torch.float8_e4m3fn),
                                                                    torch.from_numpy(s1)).numpy())
    w = rng.integers(0, 256, size=(300, 520), dtype=np.uint8)
    w[(w & 0x7F) == 0x7F] = 0x3A                                  # no nan patterns in the random block
    s2 = (np.exp2(rng.uniform(-14, -8, size=(3, 5))) * rng.uniform(1, 2, size=(3, 5))).astype(np.float32)
    out["rag__w"], out["rag__s"] = w, s2
    out["rag__out"] = bits(ref_hf._dequantize_tensor_with_scale_inv(torch.from_numpy(w).view(torch.float8_e4m3fn),
                                                                    torch.from_numpy(s2)).numpy())
    np.savez_compressed(HERE / "fp8_dequant.npz", **out)
    print("fp8_dequant.npz:", len(out), "arrays")


def make_transpose() -> None:
    """compression_algorithms/transpose.py (quantize x.T, transpose back) through the reference's plug-in, incl. ragged, 3-D,
    1-D and genuinely-float32 inputs."""
    quantizer = Quantizer(backend="emulation")
    rng = np.random.default_rng(31)
    cases = {k: v for k, v in algo_cases().items() if k in ("het_96x160", "het_70x45", "het_3d_2x40x64", "het_1d_1000")}
    cases["rand_fp32_50x70"] = (rng.standard_normal((50, 70)) * 0.02).astype(np.float32)
    cases["rand_fp32_4d"] = (rng.standard_normal((3, 5, 4, 9)) * 2).astype(np.float32)
    out = {}

    class _NoCache:
        def load_array(self, *a):
            return None

        def save_array(self, *a):
            return None

    for name, x in cases.items():
        out[f"{name}__in"] = bits(x)
        out[f"{name}__shape"] = np.asarray(x.shape, dtype=np.int64)
        res = create_algorithm("transpose", {}).run(xf=x, formats=FORMATS, quantizer=quantizer, cache=_NoCache())
        for r in res:
            assert r.y.shape == x.shape
            out[f"{name}__{r.fmt.lower()}"] = bits(np.ascontiguousarray(r.y))
    np.savez_compressed(HERE / "transpose_small.npz", **out)
    print("transpose_small.npz:", len(out), "arrays")


def make_cfg2() -> None:
    """The bench workload itself (configs[1], rank 0): the reference's mixed-tile-greedy pcc >= 0.999, seed 123, on the five
    layer-0 self_attn shapes with bench.py's synthetic bf16 tensors (seed 1000 + i).  Full size: 187 M elements, minutes."""
    import time
    arrays, meta = {}, {}
    quantizer = Quantizer(backend="emulation")
    for i, name in enumerate(synthetic.ATTN_NAMES):
        shape = synthetic.DEEPSEEK_R1_SHAPES[name]
        x = synthetic.randn_bf16_cpu(shape, 1000 + i).float().numpy()
        t0 = time.time()
        g = create_algorithm("mixed-tile-greedy", {"metric": "pcc", "threshold": 0.999, "seed": 123})
        res = g.run(xf=x, formats=FORMATS, quantizer=quantizer, cache=None)[0]
        a = np.asarray(res.meta["assignment"], dtype=np.int8)
        key = name.split(".")[-2]
        arrays[key] = a
        meta[key] = {"name": name, "shape": list(shape), "seed": 1000 + i, "input_sha256": sha(x), "counts": res.tile_counts,
                     "tile_bytes": res.tile_bytes, "assignment_sha256": sha(a), "reference_seconds": round(time.time() - t0, 1)}
        print("cfg2", key, meta[key]["counts"], meta[key]["reference_seconds"], "s", flush=True)
    np.savez_compressed(HERE / "cfg2_bench_workload.npz", **arrays)
    (HERE / "cfg2_bench_workload.json").write_text(json.dumps(meta, indent=1))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "cfg2":
        make_cfg2()
        raise SystemExit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "scalar_proxies":
        make_scalar_proxies()
        raise SystemExit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "transpose":
        make_transpose()
        raise SystemExit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "fp8_dequant":
        make_fp8_dequant()
        raise SystemExit(0)
    make_fp8_dequant()
    make_transpose()
    make_scalar_proxies()
    make_kats()
    make_rng()
    make_algos()
    make_cfg1()
