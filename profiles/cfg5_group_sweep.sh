# cfg5 (768 expert tensors) with the tile-stat pass / delta records as descriptor-array launches of G tensors (0: one launch per tensor)
for g in 0 8 16 32 64 96; do
  QA_BENCH_STATS_GROUP=$g python bench.py --config cfg5 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('stats_group $g: value %.0f GB/s  ms/step %.3f  launches/step %d  stats-only %.3f ms' % (b['value'], b['ms_per_step'], b['gpu_launches']/b['steps'], b['roofline_by_kernel'][0]['ms_per_step']))"
done
