# `ncu --set full` of the main kernels on the o_proj-size tensor (bf16 and fp8 source) + one scorer call; the raw page is
# exported on the box so that only the CSV has to travel if the report is large.
set -x
python profiles/ncu_target.py largest > gpurun_out/ncu_target_plain.log 2>&1 && ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:stats_fast|stats_f32|greedy_par|greedy_init|sdot_pipe" -o gpurun_out/prof_r2c -f python profiles/ncu_target.py largest > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i gpurun_out/prof_r2c.ncu-rep --page raw --csv > gpurun_out/r2_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r2c.ncu-rep --page source --csv -k regex:stats_fast > gpurun_out/r2_source_stats_fast.csv 2>/dev/null
ls -la gpurun_out
[ $(stat -c %s gpurun_out/prof_r2c.ncu-rep) -gt 40000000 ] && rm gpurun_out/prof_r2c.ncu-rep
./profiles/microbench/pipe_ops > gpurun_out/r2_pipe_ops.txt 2>&1; tail -6 gpurun_out/r2_pipe_ops.txt
P=quantization_analysis_b200
python profiles/stats_time.py 2 packed > gpurun_out/r2_stats_scalar.txt 2>&1
cp $P/libqa_b200.so /tmp/keep.so; cp $P/libqa_scalar.so $P/libqa_b200.so
python profiles/stats_time.py 2 scalar >> gpurun_out/r2_stats_scalar.txt 2>&1
cp /tmp/keep.so $P/libqa_b200.so
cat gpurun_out/r2_stats_scalar.txt
du -sh gpurun_out
