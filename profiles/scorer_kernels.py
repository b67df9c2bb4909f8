"""One qa_tensor_scores_f32 call on an o_proj-size pair (for the ncu launch list: which kernel holds the time)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from quantization_analysis_b200 import engine, synthetic

x = synthetic.device_randn_bf16((7168, 16384), 3, "cuda")
p = engine.prepare_rows(x)
y = engine.quant_recon(p, ["bfp4"])["bfp4"]
for _ in range(2):
    out = engine.tensor_scores_f32(x, y)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
out = engine.tensor_scores_f32(x, y)
b.record()
torch.cuda.synchronize()
print(f"one call: {a.elapsed_time(b):.2f} ms (CUDA events), pcc {out[0, 0]}")
