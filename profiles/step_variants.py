"""Graph-replayed step time for subsets of the cfg2 tensor list (which part of the step is contention between tensors?)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from quantization_analysis_b200 import synthetic
from quantization_analysis_b200.batch import GreedyBatch

dev = torch.device("cuda:0")
names = synthetic.ATTN_NAMES


def step_ms(sel, reps=20, **kw):
    shapes = [synthetic.DEEPSEEK_R1_SHAPES[n] for n in sel]
    b = GreedyBatch(shapes, metric="pcc", threshold=0.999, seed=123, **kw)
    b.load_device([synthetic.device_randn_bf16(s, 7 + i, dev) for i, s in enumerate(shapes)])
    out = {}
    for key, (st, asg) in {"step": (True, True), "stats": (True, False), "assign": (False, True)}.items():
        b.capture(st, asg)
        for _ in range(3):
            b.run_graph(st, asg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            b.run_graph(st, asg)
        e1.record()
        torch.cuda.synchronize()
        out[key] = e0.elapsed_time(e1) / reps
    return out


def timeline(sel):
    """One graph-replayed step with the cluster kernels' device timestamps (qa_debug_times)."""
    import ctypes
    from quantization_analysis_b200 import _lib
    L = _lib.lib()
    shapes = [synthetic.DEEPSEEK_R1_SHAPES[n] for n in sel]
    b = GreedyBatch(shapes, metric="pcc", threshold=0.999, seed=123)
    b.load_device([synthetic.device_randn_bf16(s, 7 + i, dev) for i, s in enumerate(shapes)])
    b.capture()
    for _ in range(3):
        b.run_graph()
    torch.cuda.synchronize()
    L.qa_debug_times(None, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); b.run_graph(); e1.record()
    torch.cuda.synchronize()
    out = (ctypes.c_ulonglong * 40)()
    L.qa_debug_times(out, 0)
    starts = [out[c * 8 + 2 * i] for c in range(5) for i in range(4) if out[c * 8 + 2 * i + 1] != 0]
    t0 = min(starts)
    lab = ["resolve", "init sums", "chain 0-1", "chain 2+"]
    print(f"  step {e0.elapsed_time(e1)*1e3:.0f} us; first start .. last end (us since the first cluster kernel started) per cluster-size class:")
    for c in range(4, -1, -1):
        if all(out[c * 8 + 2 * i + 1] == 0 for i in range(4)):
            continue
        print(f"    cluster {1 << c:2d}: " + " | ".join(f"{l} {(out[c*8+2*i]-t0)/1e3:6.1f}..{(out[c*8+2*i+1]-t0)/1e3:6.1f}" for i, l in enumerate(lab)
                                                 if out[c * 8 + 2 * i + 1] != 0))


o = [n for n in names if "o_proj" in n]
rest = [n for n in names if "o_proj" not in n]
for label, sel in (("o_proj only", o), ("the other four", rest), ("all five", names), ("all five", names), ("o_proj only", o)):
    r = step_ms(sel)
    print(f"{label:16s} step {r['step']:.3f} ms | stats only {r['stats']:.3f} | assign only {r['assign']:.3f}")
print("o_proj only:"); timeline(o)
print("all five:"); timeline(names)
