"""Where does a cfg2 step's time go when several tensor lists are in flight?  ms per step of (a) the tile-stat passes
alone, (b) the greedy chain group alone (tables from a previous pass), (c) the whole step, each with 1, 2 and 4 lists in
flight (one CUDA graph per list and phase, perm cache on)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import bench
from quantization_analysis_b200 import synthetic
from quantization_analysis_b200.batch import GreedyBatch

dev = torch.device("cuda:0")
items = bench.workload(0)
host = [synthetic.randn_bf16_cpu(s, sd) for (_n, s, sd) in items]
shapes = [s for (_n, s, _sd) in items]
NL = 4
batches = [GreedyBatch(shapes, **bench.GREEDY, device=dev, perm_cache=True) for _ in range(NL)]
for b in batches:
    b.load_device(host)
    b.run()
torch.cuda.synchronize()
for b in batches:
    for ph in ((True, False), (False, True), (True, True)):
        b.capture(*ph)
lanes = [torch.cuda.Stream(device=dev) for _ in range(NL)]


def timed(ph, nl, steps=12):
    def go(n):
        cur = torch.cuda.current_stream(dev)
        for ln in lanes[:nl]:
            ln.wait_stream(cur)
        for k in range(n):
            with torch.cuda.stream(lanes[k % nl]):
                batches[k % nl].run_graph(*ph)
        for ln in lanes[:nl]:
            cur.wait_stream(ln)
    go(4)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    go(steps)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


for name, ph in (("tile-stat passes only", (True, False)), ("greedy chain group only", (False, True)), ("whole step", (True, True))):
    print(f"{name:26s}: " + "   ".join(f"{nl} in flight {timed(ph, nl):.4f} ms/step" for nl in (1, 2, 4)))
