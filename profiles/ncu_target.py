"""Short deterministic target for `ncu --set full`: three eager steps of one GreedyBatch over the cfg2 tensor list
(perm cache on, so the launches are the tile-stat passes, the initial sums / delta records and the chain)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import bench
from quantization_analysis_b200 import synthetic
from quantization_analysis_b200.batch import GreedyBatch

dev = torch.device("cuda:0")
items = bench.workload(0)
b = GreedyBatch([s for (_n, s, _sd) in items], **bench.GREEDY, device=dev, perm_cache=True)
b.load_device([synthetic.randn_bf16_cpu(s, sd) for (_n, s, sd) in items])
for _ in range(3):
    b.run()
    torch.cuda.synchronize()
r = b.collect()
print("ok", r[0]["counts"], r[-1]["counts"], "min margins", [f"{x['min_margin']:.2e}" for x in r])
