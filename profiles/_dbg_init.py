import sys
sys.path.insert(0, '.')
import torch
from quantization_analysis_b200 import engine, synthetic
dev = torch.device("cuda:0")
for name in synthetic.ATTN_NAMES:
    shape = synthetic.DEEPSEEK_R1_SHAPES[name]
    x = synthetic.device_randn_bf16(shape, 7, dev)
    p = engine.prepare_tiles(x)
    table = engine.tile_stats(p, engine.MIXED_FORMATS, exact_abs=False)
    for _ in range(2):
        init = engine.greedy_init(table, "pcc", list(engine.MIXED_FORMATS))
    h = init.view(torch.float64)[:32].cpu().numpy()
    print(name.split('.')[-2], "init", int(h[7]), "pos", int(h[8]), "rounds", int(h[10]), "head/load/walk1/scan/walk2/min3/owner/bcast/tail:", [int(v) for v in h[16:25]])
