P=quantization_analysis_b200
cp $P/libqa_b200.so /tmp/keep.so
for pt in "chain96 12" "chain112 12" "chain128 16" "chain96 16"; do
  set -- $pt
  cp $P/libqa_$1.so $P/libqa_b200.so
  QA_BENCH_INFLIGHT=$2 python bench.py --steps 32 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1 inflight $2: value %.0f GB/s  ms/step %.4f  uncached %.0f  latency %.3f ms  chain-only %.3f ms  maps ok %s' % (b['value'], b['ms_per_step'], b['value_uncached'], b['step_latency_ms'], b['roofline_by_kernel'][1]['ms_per_step'], b['result_check']['maps_equal_reference']))"
done
cp /tmp/keep.so $P/libqa_b200.so
