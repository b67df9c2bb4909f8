import sys, time
sys.path.insert(0, "/root/repo")
import torch, numpy as np
from quantization_analysis_b200 import engine, synthetic, sweep
x = synthetic.device_randn_bf16((7168, 16384), 3, "cuda")
p = engine.prepare_tiles(x)
maps = torch.randint(1, 4, (32, p.ntiles), dtype=torch.int8, device="cuda")
orig = engine.candidate_chunk
for cap in (8, 16, 32):
    engine.candidate_chunk = lambda per, wanted, device, cap=cap: min(cap, wanted)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = sweep._score_maps(p, maps, list(range(32)))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"chunk {cap} rep {rep}: {dt*1e3:.1f} ms", flush=True)
ys = torch.stack([engine.apply_assignment(p, maps[i]) for i in range(32)])
for nb in (1, 8, 16, 32):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        engine.tensor_scores_f32(p.data, ys[:nb], n=p.numel)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"tensor_scores_f32 nbatch {nb} rep {rep}: {dt*1e3:.1f} ms", flush=True)
