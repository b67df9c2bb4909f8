"""Shared by tests/golden/make_golden_cli.py (runs the reference's programs) and the CLI parity tests (run this package's
programs, or the reference's programs over this package's drop-in modules): the tiny synthetic 'repository', the configs
and the comparison helpers."""
from __future__ import annotations

import re
import shutil
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from quantization_analysis_b200 import synthetic, tensor_source  # noqa: E402

REPO = "synthetic-tests/tiny-r1"
REVISION = "main"
FILTER = "model.layers.0"
SWEEP_TENSOR = "model.layers.0.self_attn.q_a_proj.weight"
CLI_GOLDEN = Path(__file__).resolve().parent / "golden" / "cli"


def tensors() -> dict[str, np.ndarray]:
    rng = np.random.default_rng(77)
    return {
        "model.layers.0.self_attn.kv_a_proj_with_mqa.weight": synthetic.heterogeneous_f32_np((96, 160), 21),
        "model.layers.0.mlp.down_proj.weight": synthetic.heterogeneous_f32_np((70, 45), 22),            # ragged tiles
        SWEEP_TENSOR: (rng.standard_normal((64, 96)) * 0.02).astype(np.float32),                          # NOT bf16-exact
        "model.layers.0.input_layernorm.weight": synthetic.heterogeneous_f32_np((1000,), 23),           # 1-D
    }


FORMATS5 = ["bf16", "bfp8", "bfp4", "bfp2", "fp0"]
WQ_CASES = {
    "greedy": {"algorithm": "mixed-tile-greedy", "quantization_formats": FORMATS5, "params": {"metric": "pcc", "threshold": 0.999}, "seed": 123},
    "greedy_mae": {"algorithm": "mixed-tile-greedy", "quantization_formats": FORMATS5, "params": {"metric": "mae", "threshold": 3e-4, "seed": 5}},
    "threshold": {"algorithm": "mixed-tile-threshold", "quantization_formats": FORMATS5, "params": {"metric": "pcc", "threshold": 0.94}},
    "random": {"algorithm": "mixed-tile-random", "quantization_formats": FORMATS5, "params": {"metric": "pcc", "threshold": 0.99, "iters": 6}, "seed": 42},
    "transpose": {"algorithm": "transpose", "quantization_formats": FORMATS5, "params": {}},
    "none_all_formats": {"algorithm": "none", "params": {}},                          # default: all 7 formats incl. mxfp4 / nvfp4
}
SWEEP_CASES = {
    "pcc": ["--metric", "pcc", "--lowest-metric-val", "0.9", "--steps", "8"],
    "mae": ["--metric", "mae", "--lowest-metric-val", "0.005", "--steps", "8"],
    "atol": ["--metric", "atol", "--lowest-metric-val", "0.05", "--steps", "8"],
    "pcc_3fmt": ["--metric", "pcc", "--lowest-metric-val", "0.95", "--steps", "5", "--formats", "bfp8,bfp4,bfp2"],
}


def seed_fp32_cache(cache_dir: Path) -> None:
    """Write the tensors where the reference's loader (and tensor_source) look first (hf_model_utils.py:129-132,248-250)."""
    d = Path(cache_dir) / "tensor-fp32" / tensor_source.safe_repo_revision_key(REPO, REVISION)
    d.mkdir(parents=True, exist_ok=True)
    for name, x in tensors().items():
        np.save(d / f"{tensor_source.safe_tensor_key(name)}.npy", x)


def fake_index_factory(hf, cache_dir: Path):
    """The single monkeypatch SURVEY.md section 4 describes: a ModelIndex that lists the cached tensors."""
    def fake(repo_or_url, revision="main", cache_dir=cache_dir, hf_token=None):
        assert hf._safe_repo_revision_key(REPO, REVISION) == tensor_source.safe_repo_revision_key(REPO, REVISION)
        return hf.ModelIndex(repo_id=REPO, revision=revision, cache_dir=Path(cache_dir), hf_token=None, safetensor_files=[],
                             tensor_to_file={n: "synthetic.safetensors" for n in tensors()}, weight_map=None)
    return fake


def latest_results_dir(results_root: Path, algo: str) -> Path:
    base = Path(results_root) / REPO.replace("/", "__") / algo
    runs = sorted(p for p in base.iterdir() if p.is_dir())
    return runs[-1]


def harvest(src: Path, dst: Path) -> None:
    for f in sorted(Path(src).rglob("*")):
        if f.is_file() and f.suffix in (".txt", ".json", ".csv", ".npy"):
            out = dst / f.relative_to(src)
            out.parent.mkdir(parents=True, exist_ok=True)
            shutil.copyfile(f, out)


_TIME_GB = re.compile(r"(\s)(\d+\.\d{3})(\s+\d+\.\d{3})(\s|$)")


def strip_times(table_text: str) -> str:
    """table.txt with the TIME(s) value of every row blanked (the only column that legitimately differs)."""
    out = []
    for line in table_text.splitlines():
        out.append(_TIME_GB.sub(lambda m: m.group(1) + "T" * len(m.group(2)) + m.group(3) + m.group(4), line, count=1))
    return "\n".join(out)


def compare_trees(got: Path, want: Path) -> list[str]:
    """Differences between two harvested result trees (empty list = identical up to TIME(s))."""
    diffs = []
    want_files = sorted(p.relative_to(want) for p in Path(want).rglob("*") if p.is_file())
    for rel in want_files:
        g, w = Path(got) / rel, Path(want) / rel
        if not g.exists():
            diffs.append(f"missing {rel}")
            continue
        if rel.suffix == ".npy":
            a, b = np.load(g), np.load(w)
            if a.dtype != b.dtype or a.shape != b.shape or not np.array_equal(a, b):
                diffs.append(f"array differs: {rel}")
        elif rel.name == "table.txt":
            if strip_times(g.read_text()) != strip_times(w.read_text()):
                diffs.append(f"table differs: {rel}")
        elif rel.name == "sweep_config.json" or rel.name == "compression_config.used.json":
            import json
            a, b = json.loads(g.read_text()), json.loads(w.read_text())
            if a != b:
                diffs.append(f"json differs: {rel}: {a} != {b}")
        else:
            if g.read_text() != w.read_text():
                diffs.append(f"text differs: {rel}")
    return diffs


REF_COPY = ROOT / "oracle" / "_ref"

_DROPIN_DRIVER = r'''
import contextlib, io, json, os, runpy, sys
root, ref, work, spec = sys.argv[1:5]
sys.path.insert(0, root)
sys.path.insert(0, ref)                      # hf_model_utils (reference, unmodified) + the reference programs
sys.path.insert(0, os.path.join(ref, "scripts"))
import quantization_analysis_b200 as q
q.install_drop_in()                          # quantization_formats / compression_algorithms.* -> this package
import hf_model_utils as hf
from tests import cli_util
hf.build_model_index = cli_util.fake_index_factory(hf, os.path.join(work, "data", "hf-cache"))
os.chdir(work)
import time
for script, argv in json.loads(spec):
    time.sleep(1.1)                          # the reference names result directories by the second
    sys.argv = [script] + argv
    out = io.StringIO()
    try:
        with contextlib.redirect_stdout(out):
            runpy.run_path(script, run_name="__main__")
    except SystemExit as e:
        if e.code not in (0, None):
            print(out.getvalue()[-3000:])
            raise
import compression_algorithms, quantization_formats
assert "quantization_analysis_b200" in compression_algorithms.__file__ and "quantization_analysis_b200" in quantization_formats.__file__
'''


def run_reference_programs_over_dropin(jobs, workdir: Path) -> None:
    """Run the reference's UNMODIFIED programs (oracle/_ref, made by oracle/make_ref.sh) in a fresh interpreter in which
    ``quantization_formats`` and ``compression_algorithms`` resolve to this package.  jobs: [(script path, argv), ...]."""
    import json
    import subprocess
    subprocess.run([sys.executable, "-c", _DROPIN_DRIVER, str(ROOT), str(REF_COPY), str(workdir),
                    json.dumps([[str(s), list(a)] for s, a in jobs])], check=True, cwd=str(workdir), timeout=1200)
