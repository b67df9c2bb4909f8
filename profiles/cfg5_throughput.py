"""cfg5-style throughput: one GPU's share of a MoE layer (96 experts x gate/up/down = 288 tensors of 14 336 tiles each),
greedy pcc >= 0.999, device-resident, one CUDA graph per pass.  All tensors share their permutations (same tile count)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from quantization_analysis_b200 import synthetic
from quantization_analysis_b200.batch import GreedyBatch

dev = torch.device("cuda:0")
n_exp = int(sys.argv[1]) if len(sys.argv) > 1 else 96
items = synthetic.expert_tensor_list(n_exp)
shapes = [s for _n, s in items]
n_streams = int(sys.argv[2]) if len(sys.argv) > 2 else None
b = GreedyBatch(shapes, metric="pcc", threshold=0.999, seed=123, n_streams=n_streams)
base = [synthetic.device_randn_bf16(s, 11 + i, dev) for i, s in enumerate(shapes[:6])]
b.load_device([base[i % 6] if base[i % 6].shape == torch.Size(shapes[i]) else base[i % 6].reshape(shapes[i]) for i in range(len(shapes))])
b.capture()
for _ in range(2):
    b.run_graph()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    b.run_graph()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"{len(shapes)} tensors, {b.total_bytes()/1e9:.2f} GB per pass: {ms:.2f} ms -> {b.total_bytes()/ms/1e6:.0f} GB/s; launches per pass {b.launches_per_step}")
r = b.collect()
print("first tensor counts", r[0]["counts"], "pcc", r[0]["metrics"]["pcc"])
