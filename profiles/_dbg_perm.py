import sys
sys.path.insert(0, '.')
import torch
from quantization_analysis_b200 import engine
for n in (114688, 17652):
    for _ in range(2):
        r = engine.make_rng(123)
        p = engine.numpy_permutation(r, n, parallel=True)
        torch.cuda.synchronize()
    print("----")
