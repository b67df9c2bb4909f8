"""TMA-staged tile-stat kernel (QA_STATS_TMA=<CTAs per SM>) against the shipped one on an o_proj-size tensor: same table
(sum x^2 to 1e-13: the rows reach the float64 accumulators in another order), time per launch.  Run once per setting:
    QA_STATS_TMA=4 python profiles/stats_tma_probe.py"""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from quantization_analysis_b200 import _lib, engine, synthetic

x = synthetic.device_randn_bf16((7168, 16384), 3, "cuda")
p = engine.prepare_tiles(x)
L = _lib.lib()
sp = torch.cuda.current_stream().cuda_stream
table = torch.zeros((_lib.NSTAT, p.ntiles), dtype=torch.float64, device="cuda")


def run(mode=2):
    _lib.check(L.qa_tile_stats(p.data.data_ptr(), 0, p.rows, p.cols, p.cols, 0, 0xF, mode, table.data_ptr(), sp), "qa_tile_stats")


run()
torch.cuda.synchronize()
ref_path = Path("/tmp/stats_ref.pt")
tag = os.environ.get("QA_STATS_TMA", "0")
if tag == "0":
    torch.save(table.cpu(), ref_path)
elif ref_path.exists():
    ref = torch.load(ref_path)
    got = table.cpu()
    exact = [i for i in range(_lib.NSTAT) if torch.equal(got[i], ref[i])]
    close = torch.allclose(got, ref, rtol=1e-13, atol=0)
    print(f"QA_STATS_TMA={tag}: rows bit-equal to the shipped kernel's: {len(exact)} of {_lib.NSTAT}; all within 1e-13: {close}")
for _ in range(5):
    run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    run()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
alg = 2 * p.numel + 176 * p.ntiles
peak = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
print(f"QA_STATS_TMA={tag}: {ms*1e3:.1f} us per launch, {alg/ms/1e6:.0f} GB/s algorithmic = {alg/ms/1e6/peak:.3f} of {peak} GB/s")
