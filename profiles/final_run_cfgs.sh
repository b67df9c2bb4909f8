# the other BASELINE configs on one GPU, with their CPU arms (lines land in gpurun_out/bench_r2_cfg*.json)
python bench.py --config cfg1 --steps 10 --warmup 3 > gpurun_out/bench_r2_cfg1.json 2> gpurun_out/bench_r2_cfg1.err
python bench.py --config cfg3 --steps 2 --warmup 2 > gpurun_out/bench_r2_cfg3.json 2> gpurun_out/bench_r2_cfg3.err
python bench.py --config cfg4 --steps 1 --warmup 1 > gpurun_out/bench_r2_cfg4.json 2> gpurun_out/bench_r2_cfg4.err
python bench.py --config cfg5 --steps 3 --warmup 3 > gpurun_out/bench_r2_cfg5_n1.json 2> gpurun_out/bench_r2_cfg5_n1.err
for c in cfg1 cfg3 cfg4 cfg5_n1; do python -c "
import json
d=json.loads(open('gpurun_out/bench_r2_$c.json').read().strip().splitlines()[-1])
print('$c', round(d['value'],3), round(d['ms_per_step'],3), (d.get('e2e') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value'))"; done
# per-instruction stall samples of the cluster kernels (o_proj-size tensor); only the CSV pages travel
python profiles/ncu_target.py largest > gpurun_out/ncu_target_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:greedy_par|greedy_init" -c 4 -o gpurun_out/prof_r2d -f python profiles/ncu_target.py largest > gpurun_out/ncu_full2.log 2>&1
ncu -i gpurun_out/prof_r2d.ncu-rep --page source --csv -k regex:greedy_par > gpurun_out/r2_source_greedy_par.csv 2>/dev/null
ncu -i gpurun_out/prof_r2d.ncu-rep --page source --csv -k regex:greedy_init > gpurun_out/r2_source_greedy_init.csv 2>/dev/null
rm -f gpurun_out/prof_r2d.ncu-rep
ls -la gpurun_out; du -sh gpurun_out
