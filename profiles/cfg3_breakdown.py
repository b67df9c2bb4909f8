"""cfg3 (threshold sweep, 32 steps, pcc) on an o_proj-size tensor: where the time goes."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from quantization_analysis_b200 import engine, synthetic, sweep

dev = torch.device("cuda:0")


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


x = synthetic.device_randn_bf16((7168, 16384), 5, dev)
p = engine.prepare_tiles(x)
fm = list(engine.MIXED_FORMATS)
t_sc, scores = timed(lambda: engine.tile_scores(p, fm))
t_st, table = timed(lambda: engine.tile_stats(p, fm))
sc = scores[0].contiguous()
thr = sweep.sweep_thresholds(sc, fm, "pcc", 32, 0.9)
order = sweep.formats_by_precision(fm)
t_th, (maps, counts) = timed(lambda: engine.threshold_assign(sc, order, True, thr))
t_as, _ = timed(lambda: [engine.assignment_sums(table, maps[i]) for i in range(32)])
t_all, _ = timed(lambda: sweep.sweep_tensor(x, metric="pcc", steps=32, lowest=0.9), reps=2)
print(f"tile_scores {t_sc:.3f} ms | tile_stats {t_st:.3f} | threshold_assign(32) {t_th:.3f} | assignment_sums x32 {t_as:.3f} | sweep_tensor total {t_all:.2f} ms")
print(f"input {p.numel*2/1e6:.0f} MB -> sweep {p.numel*2/t_all/1e6:.1f} GB/s of bf16 weights")
