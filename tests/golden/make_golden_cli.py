#!/usr/bin/env python3
"""Goldens made by running the UNMODIFIED reference command-line programs offline (build container only):

    python tests/golden/make_golden_cli.py

* `wq` (reference: /root/reference/wq) for mixed-tile-greedy / -threshold / -random / transpose configs,
* `scripts/sweep_mixed_tile_threshold.py` for the pcc, mae and atol sweeps,

on three small synthetic tensors, with the one monkeypatch SURVEY.md section 4 describes (`build_model_index` returns a
ModelIndex over tensors pre-seeded in the fp32 cache, so nothing touches the network).  The files the programs wrote
(`table.txt`, `compression_config.used.json`, assignment maps, random CSVs, `sweep_results.csv`) are committed under
tests/golden/cli/; tests/cli_util.py re-creates the same inputs and runs either the reference program over this
package's drop-in modules or this package's own CLI against them.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import runpy
import shutil
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT / "tests"))
import cli_util  # noqa: E402

REF = Path("/root/reference")


def run_reference(script: Path, argv: list[str], workdir: Path) -> str:
    """Run one reference program in `workdir` with the model index patched; returns its stdout."""
    import hf_model_utils as hf
    hf.build_model_index = cli_util.fake_index_factory(hf, workdir / "data" / "hf-cache")
    old_argv, old_cwd = sys.argv, os.getcwd()
    sys.argv = [str(script)] + argv
    os.chdir(workdir)
    out = io.StringIO()
    try:
        with contextlib.redirect_stdout(out), contextlib.redirect_stderr(io.StringIO()):
            try:
                runpy.run_path(str(script), run_name="__main__")
            except SystemExit as e:
                if e.code not in (0, None):
                    raise RuntimeError(f"{script.name} {argv} exited with {e.code}: {out.getvalue()[-2000:]}")
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)
    return out.getvalue()


def main() -> None:
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(REF / "scripts"))
    dst = HERE / "cli"
    if dst.exists():
        shutil.rmtree(dst)
    dst.mkdir()
    with tempfile.TemporaryDirectory() as td:
        work = Path(td)
        cli_util.seed_fp32_cache(work / "data" / "hf-cache")
        for case, cfg in cli_util.WQ_CASES.items():
            cfg_path = work / f"{case}.json"
            cfg_path.write_text(json.dumps(cfg))
            run_reference(REF / "wq", [cli_util.REPO, cli_util.FILTER, "--compression-config", str(cfg_path), "--recompute", "--summary"], work)
            res = cli_util.latest_results_dir(work / "results", cfg["algorithm"])
            cli_util.harvest(res, dst / "wq" / case)
            shutil.rmtree(work / "results")
        for case, args in cli_util.SWEEP_CASES.items():
            run_reference(REF / "scripts" / "sweep_mixed_tile_threshold.py",
                          [cli_util.REPO, cli_util.SWEEP_TENSOR, "--no-regex", "--out-dir", str(work / "sweep" / case)] + args, work)
            cli_util.harvest(work / "sweep" / case, dst / "sweep" / case)
    n = sum(1 for _ in dst.rglob("*") if _.is_file())
    print("wrote", n, "files under", dst)


if __name__ == "__main__":
    main()
