"""Fast tile-stat kernels for inputs that are not bf16-exact, on an o_proj-size tensor (7168 x 16384): float32 source
(qa_tile_stats_f32, 4 B/elem), fp8 e4m3fn + 128 x 128 block scales (qa_tile_stats_fp8, 1 B/elem), the strict NumPy-order
kernel they replace, and the bf16 kernel for scale.  CUDA-event time per launch.
    python profiles/stats_f32_time.py [mode: 0 exact-abs | 2 approx-abs]"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from quantization_analysis_b200 import _lib, engine

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rows, cols = 7168, 16384
g = torch.Generator(device="cuda").manual_seed(3)
w8 = torch.randint(0, 256, (rows, cols), dtype=torch.uint8, device="cuda", generator=g)
w8[(w8 & 0x7F) == 0x7F] = 0x3C
sc = torch.exp(torch.empty((rows // 128, cols // 128), device="cuda").uniform_(-9.2, -5.8, generator=g))
xf, _, bad = engine.fp8_block_dequant(w8, sc, want_bf16=False)
xb = xf.to(torch.bfloat16)
ntiles = (rows // 32) * (cols // 32)
table = torch.zeros((_lib.NSTAT, ntiles), dtype=torch.float64, device="cuda")
L = _lib.lib()
peak = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
sp = torch.cuda.current_stream().cuda_stream


def timed(name, fn, in_bytes, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    alg = in_bytes + 176 * ntiles
    print(f"{name:34s} {ms*1e3:8.1f} us per launch, {alg/ms/1e6:7.0f} GB/s algorithmic = {alg/ms/1e6/peak:.3f} of {peak} GB/s, "
          f"{rows*cols/ms/1e6:6.1f} G elements/s, {ms*1e-3*1.965e9*148*128/(rows*cols):.1f} lane-cycles per element")


print(f"{rows} x {cols}, {bad} of {rows*cols} dequantized values are not bf16-exact, mode={mode}")
timed("qa_tile_stats (bf16 fast)", lambda: _lib.check(L.qa_tile_stats(xb.data_ptr(), 0, rows, cols, cols, 0, 0xF, mode, table.data_ptr(), sp), "s"), 2 * rows * cols)
timed("qa_tile_stats_f32 (float32 fast)", lambda: _lib.check(L.qa_tile_stats_f32(xf.data_ptr(), rows, cols, cols, 0xF, mode, table.data_ptr(), 0, -1, sp), "s"), 4 * rows * cols)
timed("qa_tile_stats_fp8 (fused dequant)", lambda: _lib.check(L.qa_tile_stats_fp8(w8.data_ptr(), sc.data_ptr(), rows, cols, cols, sc.shape[0], sc.shape[1], 0xF, mode, table.data_ptr(), 0, -1, None, sp), "s"), rows * cols + 4 * sc.numel())
timed("qa_tile_stats (float32 strict)", lambda: _lib.check(L.qa_tile_stats(xf.data_ptr(), 1, rows, cols, cols, 0, 0xF, 1, table.data_ptr(), sp), "s"), 4 * rows * cols, reps=3)
tmp = torch.empty_like(xf)
timed("qa_fp8_block_dequant alone", lambda: _lib.check(L.qa_fp8_block_dequant(w8.data_ptr(), sc.data_ptr(), rows, cols, sc.shape[0], sc.shape[1], tmp.data_ptr(), None, None, sp), "d"), 5 * rows * cols - 176 * ntiles)
