"""Tile-stat kernel alone on an o_proj-size tensor (7168 x 16384 bf16, 235 MB > L2): CUDA-event time per launch,
algorithmic GB/s (2 B/elem read + 176 B per tile written) and the fraction of the measured copy peak.
    python profiles/stats_time.py [mode: 0 exact-abs | 2 approx-abs] [tag]"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from quantization_analysis_b200 import _lib, engine, synthetic

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 2
tag = sys.argv[2] if len(sys.argv) > 2 else ""
x = synthetic.device_randn_bf16((7168, 16384), 3, "cuda")
p = engine.prepare_tiles(x)
table = torch.zeros((_lib.NSTAT, p.ntiles), dtype=torch.float64, device="cuda")
L = _lib.lib()


def run():
    _lib.check(L.qa_tile_stats(p.data.data_ptr(), 0, p.rows, p.cols, p.cols, 0, 0xF, mode, table.data_ptr(),
                               torch.cuda.current_stream().cuda_stream), "qa_tile_stats")


for _ in range(5):
    run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
a.record()
for _ in range(reps):
    run()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
alg = 2 * p.numel + 176 * p.ntiles
peak = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
cyc = ms * 1e-3 * 1.965e9 * 148 * 128 / p.numel
print(f"stats_fast_kernel {tag} mode={mode}: {ms*1e3:.1f} us per launch, {alg/ms/1e6:.0f} GB/s algorithmic, "
      f"{alg/ms/1e6/peak:.3f} of {peak} GB/s, {cyc:.1f} lane-cycles per element at 1965 MHz")
