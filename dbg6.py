import sys; sys.path.insert(0,'.')
import torch
from quantization_analysis_b200 import engine as eng, synthetic
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/reps
for shape in [(1536,7168),(7168,16384),(18432,7168)]:
    x = synthetic.device_randn_bf16(shape, 1, "cuda")
    p = eng.prepare_rows(x)
    n = p.numel
    t = timeit(lambda: eng.quant_recon(p, ["bfp8","bfp4","bfp2"]))
    print(shape, "recon 3 fmts: %.3f ms  alg %.0f GB/s (8 B/elem)  frac %.3f" % (t, 8*n/t/1e6, 8*n/t/1e6/6547.2))
    t = timeit(lambda: eng.quant_recon(p, ["bfp8"]))
    print(shape, "recon 1 fmt : %.3f ms  alg %.0f GB/s (4 B/elem)  frac %.3f" % (t, 4*n/t/1e6, 4*n/t/1e6/6547.2))
    pt = eng.prepare_tiles(x)
    t = timeit(lambda: eng.tile_stats(pt, exact_abs=False))
    print(shape, "stats fast  : %.3f ms  alg %.0f GB/s (2.17 B/elem)  frac %.3f" % (t, 2.17*n/t/1e6, 2.17*n/t/1e6/6547.2))
    t = timeit(lambda: eng.tile_scores(pt))
    print(shape, "tile scores : %.3f ms  %.0f GB/s of bf16 input" % (t, 2*n/t/1e6))
    a = torch.randint(0, 4, (pt.ntiles,), dtype=torch.int8, device="cuda")
    t = timeit(lambda: eng.apply_assignment(pt, a))
    print(shape, "apply       : %.3f ms  alg %.0f GB/s (4 B/elem) frac %.3f" % (t, 4*n/t/1e6, 4*n/t/1e6/6547.2))
