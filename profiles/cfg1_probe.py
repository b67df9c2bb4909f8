import sys, time
sys.path.insert(0, '/root/repo')
import torch
from quantization_analysis_b200 import engine, synthetic
dev = torch.device('cuda:0')
xs = [synthetic.device_randn_bf16((1536, 7168), 100 + i, dev) for i in range(8)]
preps = [engine.prepare_rows(x) for x in xs]
for k in range(5):
    engine.quant_recon(preps[k % 8], ["bfp8", "bfp4", "bfp2"])
torch.cuda.synchronize()
for K in (3, 20, 100):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for k in range(K):
        out = engine.quant_recon(preps[k % 8], ["bfp8", "bfp4", "bfp2"])
    b.record()
    torch.cuda.synchronize()
    print(K, 'steps: device %.4f ms/step, host %.4f ms/step' % (a.elapsed_time(b) / K, (time.perf_counter() - t0) * 1e3 / K))
