for pt in "16 2" "24 2" "12 6"; do
  set -- $pt
  QA_BENCH_INFLIGHT=$1 QA_BENCH_CLUSTER_CAP=$2 python bench.py --steps 48 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('inflight $1 cap $2: value %.0f GB/s  ms/step %.4f  uncached %.0f  latency %.3f ms  maps ok %s' % (b['value'], b['ms_per_step'], b['value_uncached'], b['step_latency_ms'], b['result_check']['maps_equal_reference']))"
done
