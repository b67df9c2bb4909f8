"""cfg1 (`none`, all formats) and cfg4 (mixed-tile-random, 1000 samples) on the [1536, 7168] config tensor: device time."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from quantization_analysis_b200 import engine, synthetic

dev = torch.device("cuda:0")


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


x = synthetic.device_randn_bf16((1536, 7168), 5, dev)
p = engine.prepare_tiles(x)
pr = engine.prepare_rows(x)
fm = list(engine.MIXED_FORMATS)
t_rec, _ = timed(lambda: engine.quant_recon(pr, ["bf16", "bfp8", "bfp4", "bfp2"]))
t_st, table = timed(lambda: engine.tile_stats(p, fm))
t_rs, (ch, met, cnt) = timed(lambda: engine.random_samples(table, p.numel, fm, 1000, engine.make_rng(123)), reps=3)
nb = p.numel * 2
print(f"cfg1 recon (3 materialised formats): {t_rec*1e3:.0f} us = {nb/t_rec/1e6:.0f} GB/s of bf16 weights ({4*nb/t_rec/1e6:.0f} GB/s algorithmic)")
print(f"cfg4: tile_stats {t_st*1e3:.0f} us + random_samples(1000) {t_rs:.2f} ms -> {nb/(t_st+t_rs)/1e6:.1f} GB/s of bf16 weights "
      f"({1000*p.ntiles/t_rs/1e6:.1f} G tile-draws/s)")
