# the driver's invocation (--steps 20 --warmup 5): lists in flight x cluster cap
for pt in "12 0" "12 4" "10 4" "12 6" "10 6" "12 8" "10 8" "20 4"; do
  set -- $pt
  QA_BENCH_INFLIGHT=$1 QA_BENCH_CLUSTER_CAP=$2 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('steps 20: inflight $1 cap $2: value %.0f GB/s  ms/step %.4f  latency %.3f ms' % (b['value'], b['ms_per_step'], b['step_latency_ms']))"
done
