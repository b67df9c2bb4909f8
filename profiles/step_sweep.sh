#!/bin/sh
# Lists in flight x cluster cap sweep of the cfg2 step (perm cache on): prints value / ms_per_step / uncached value per point.
for inf in 2 3 4; do for cap in 0 8 4; do
  QA_BENCH_INFLIGHT=$inf QA_BENCH_CLUSTER_CAP=$cap python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('inflight $inf cap $cap: value %.0f GB/s  ms/step %.4f  uncached %.0f  latency %.3f ms  e2e %.1f' % (b['value'], b['ms_per_step'], b['value_uncached'], b['step_latency_ms'], b['e2e']['value']))"
done; done
