"""Short deterministic target for `ncu --set full --profile-from-start off`: after a warm-up of everything, one eager step of a
GreedyBatch over the cfg2 tensor list (perm cache on: the launches are the tile-stat passes, the initial sums / delta records
and the chain), one eager step of the same list stored as fp8 + block scales (stats_f32_kernel with the fused dequantization),
and one reference-float32 scoring call on a q_a_proj-size pair (sdot_pipe_kernel)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import bench
from quantization_analysis_b200 import engine, synthetic
from quantization_analysis_b200.batch import GreedyBatch

dev = torch.device("cuda:0")
items = bench.workload(0)
if len(sys.argv) > 1 and sys.argv[1] == "largest":      # o_proj only: a capture of ~15 kernels stays small
    items = [max(items, key=lambda it: it[1][0] * it[1][1])]
shapes = [s for (_n, s, _sd) in items]
b = GreedyBatch(shapes, **bench.GREEDY, device=dev, perm_cache=True)
b.load_device([synthetic.randn_bf16_cpu(s, sd) for (_n, s, sd) in items])
f = GreedyBatch(shapes, **bench.GREEDY, device=dev, perm_cache=True, source="fp8")
f.load_device([synthetic.fp8_checkpoint_cpu(s, sd) for (_n, s, sd) in items])
x = synthetic.device_randn_bf16((1536, 7168), 3, dev)
y = engine.quant_recon(engine.prepare_rows(x), ["bfp4"])["bfp4"]
for _ in range(2):
    b.run()
    f.run()
    engine.tensor_scores_f32(x, y)
    torch.cuda.synchronize()
torch.cuda.profiler.start()
b.run()
torch.cuda.synchronize()
f.run()
torch.cuda.synchronize()
out = engine.tensor_scores_f32(x, y)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
r, rf = b.collect(), f.collect()
print("ok", r[0]["counts"], rf[0]["counts"], "min margins", [f"{v['min_margin']:.2e}" for v in r], "pcc", float(out[0, 0]))
