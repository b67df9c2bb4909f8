#!/bin/sh
# More lists in flight with smaller clusters (chain SM-time vs latency), cfg2 step with perm cache.
for pt in "6 0" "8 0" "6 8" "8 8" "8 4" "12 4" "12 2" "16 2"; do
  set -- $pt
  QA_BENCH_INFLIGHT=$1 QA_BENCH_CLUSTER_CAP=$2 python bench.py --steps 16 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('inflight $1 cap $2: value %.0f GB/s  ms/step %.4f  uncached %.0f  latency %.3f ms' % (b['value'], b['ms_per_step'], b['value_uncached'], b['step_latency_ms']))"
done
