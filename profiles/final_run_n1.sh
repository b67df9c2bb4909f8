# Round-2 numbers of record on one B200 (run from the repo root on the GPU box; everything lands in gpurun_out/).
set -x
python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; tail -3 gpurun_out/r2c_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1; tail -1 gpurun_out/r2c_smoke.log
python bench.py --steps 24 --warmup 5 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; cut -c1-300 gpurun_out/bench_r2_final.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err; cut -c1-200 gpurun_out/bench_r2_ref.json
python bench.py --config cfg2-fp8 --steps 24 --warmup 5 > gpurun_out/bench_r2_cfg2_fp8.json 2> gpurun_out/bench_r2_cfg2_fp8.err
python profiles/stats_f32_time.py 2 > gpurun_out/r2_stats_f32.txt 2>&1; python profiles/stats_f32_time.py 0 >> gpurun_out/r2_stats_f32.txt 2>&1
python profiles/scorer_time.py > gpurun_out/r2_scorer_time.txt 2>&1
QA_BENCH_INFLIGHT=2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu_plain.json 2> gpurun_out/bench_ncu_plain.err && QA_BENCH_INFLIGHT=2 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
du -sh gpurun_out
