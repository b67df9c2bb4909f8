"""Whole-tensor reference-float32 scorer (qa_tensor_scores_f32) on an o_proj-size pair: time per call for 1 and 8 candidates,
and a bit check against the oracle's restated NumPy orders on a smaller tensor."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from oracle import qa_oracle as orc
from quantization_analysis_b200 import engine, synthetic

x = synthetic.device_randn_bf16((7168, 16384), 3, "cuda")
p = engine.prepare_rows(x)
rec = engine.quant_recon(p, ["bfp8", "bfp4", "bfp2"])
ys1 = rec["bfp4"]
ys8 = torch.stack([rec["bfp8"], rec["bfp4"], rec["bfp2"], rec["bfp8"], rec["bfp4"], rec["bfp2"], rec["bfp8"], rec["bfp4"]])
for name, ys in (("1 candidate", ys1), ("8 candidates", ys8)):
    engine.tensor_scores_f32(x, ys)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        out = engine.tensor_scores_f32(x, ys)
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms per call on 117.4 M elements; pcc {out[:3, 0]}")
xs = synthetic.randn_bf16_cpu((1536, 7168), 0)
ysm = engine.quant_recon(engine.prepare_rows(xs.cuda()), ["bfp4"])["bfp4"]
got = engine.tensor_scores_f32(xs.cuda(), ysm)[0]
want = orc.wq_scores(xs.float().numpy(), ysm.float().cpu().numpy().reshape(1536, 7168))
print("bit check [1536,7168] bfp4:", float(got[0]) == want["pcc"], float(got[1]) == want["mae"], float(got[2]) == want["atol"], float(got[0]))
