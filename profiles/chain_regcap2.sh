P=quantization_analysis_b200
cp $P/libqa_b200.so /tmp/keep.so
for v in base chain128; do
  [ $v = base ] || cp $P/libqa_$v.so $P/libqa_b200.so
  for inf in 8 12; do
    QA_BENCH_INFLIGHT=$inf python bench.py --steps 24 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v inflight $inf: value %.0f GB/s  ms/step %.4f  uncached %.0f  latency %.3f ms  chain-only %.3f ms' % (b['value'], b['ms_per_step'], b['value_uncached'], b['step_latency_ms'], b['roofline_by_kernel'][1]['ms_per_step']))"
  done
done
cp /tmp/keep.so $P/libqa_b200.so
