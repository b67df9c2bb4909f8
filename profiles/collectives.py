"""Cycles per call of the cluster collectives (qa_collective_bench): scan + flag exchange, 3-way min, full cluster sync
(CTA barrier + mbarrier exchange), __syncthreads, pair scan of 3 streams."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from quantization_analysis_b200 import _lib

L = _lib.lib()
out = torch.zeros(8, dtype=torch.float64, device="cuda:0")
for cl in (1, 2, 4, 8, 16):
    _lib.check(L.qa_collective_bench(out.data_ptr(), 500, cl, torch.cuda.current_stream().cuda_stream), "bench")
    torch.cuda.synchronize()
    v = out.cpu().numpy()
    print(f"cluster {cl:2d}: scan+flag {v[0]:7.0f} | min3 {v[1]:7.0f} | cluster sync {v[2]:7.0f} | __syncthreads {v[3]:6.0f} | pair scan x3 {v[4]:7.0f}  cycles")
