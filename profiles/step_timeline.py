"""Per-tensor stage completion times inside one cfg2 step (eager launches, CUDA events on each tensor's streams)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from quantization_analysis_b200 import synthetic
from quantization_analysis_b200.batch import GreedyBatch

dev = torch.device("cuda:0")
names = synthetic.ATTN_NAMES
shapes = [synthetic.DEEPSEEK_R1_SHAPES[n] for n in names]
b = GreedyBatch(shapes, metric="pcc", threshold=0.999, seed=123)
b.load_device([synthetic.device_randn_bf16(s, 7 + i, dev) for i, s in enumerate(shapes)])
for _ in range(3):
    b.run()
torch.cuda.synchronize()
for rep in range(2):
    b.trace = {}
    b.run()
    rows = b.timeline()
    b.trace = None
    print(f"--- step {rep} (ms since the step's fork)")
    for n, r in zip(names, rows):
        print(n.split(".")[-2].ljust(22), " ".join(f"{k}={v:.3f}" for k, v in r.items() if k != "shape"))
