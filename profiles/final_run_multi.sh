# usage: sh profiles/final_run_multi.sh N      (torchrun the way the driver does; lines land in gpurun_out/)
N=$1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$RUN bench.py --gpus $N --steps 24 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err; tail -c 600 gpurun_out/bench_r2_n$N.json | head -c 300; echo
$RUN bench.py --gpus $N --steps 3 --warmup 3 --config cfg5 --no-cpu-baseline > gpurun_out/bench_r2_cfg5_n$N.json 2> gpurun_out/bench_r2_cfg5_n$N.err; head -c 300 gpurun_out/bench_r2_cfg5_n$N.json; echo
