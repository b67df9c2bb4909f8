#!/usr/bin/env python3
"""Headline benchmark: bf16 weight GB/s quantized + scored + assigned (BASELINE.json metric).

Workload at N=1 (BASELINE.json configs[1]): mixed-tile-greedy, metric pcc >= 0.999, seed 123,
candidate formats bf16/bfp8/bfp4/bfp2, over the five DeepSeek-R1 layer-0 self_attn weight
shapes (187.1 M elements, 374 MB of bf16 - larger than the 126 MB L2, so no flush is needed).
A step = one pass of the hot path over that tensor list: fused quantize+tile-stats, greedy
assignment and whole-tensor scoring, per tensor.  With N GPUs every rank runs the same shapes
for its own layer of the tensor list (weak scaling, no data-path collective; per-tensor result
rows are gathered to rank 0 at the end).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "bf16 weight GB/s quantized+scored (mixed-tile-greedy pcc>=0.999, DeepSeek-R1 layer-0 self_attn shapes)"
UNIT = "GB/s"
GREEDY = {"metric": "pcc", "threshold": 0.999, "seed": 123}
INFLIGHT = int(os.environ.get("QA_BENCH_INFLIGHT", "2"))          # tensor lists in flight for the device-resident throughput (double buffering)
TABLE_BYTES_PER_TILE = 22 * 8


def workload(rank: int):
    from quantization_analysis_b200 import synthetic
    return [(n.replace("layers.0", f"layers.{rank}"), synthetic.DEEPSEEK_R1_SHAPES[n], 1000 * (rank + 1) + i)
            for i, n in enumerate(synthetic.ATTN_NAMES)]


def config_dict(n_gpus: int) -> dict:
    return {"workload": "configs[1]: mixed-tile-greedy pcc>=0.999 seed 123 over q_a/q_b/kv_a/kv_b/o_proj "
                        "(synthetic randn*0.02 bf16), one such tensor list per GPU",
            "tensors_per_gpu": 5, "elements_per_gpu": 187105280, "formats": "bf16,bfp8,bfp4,bfp2",
            "l2": "inputs_larger_than_l2 (374 MB per step vs 126 MB L2)", "parallelism": f"tensor-list x{n_gpus}",
            "launch": "one CUDA graph per step for the device-resident value (kernels of all tensors on ~20 captured streams); "
                      "eager stream launches for e2e",
            "inflight": f"{INFLIGHT} double-buffered tensor lists (step k+1's tile-stat passes overlap step k's greedy chain); "
                        "step_latency_ms is one step alone"}


# --------------------------------------------------------------------------------------------
# clocks sampler (NVML)
# --------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int, enabled: bool = True):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        self.nv = None
        if not enabled:
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.004)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=1.0)

    def summary(self) -> dict:
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# --------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port of the reference algorithm on host cores
# --------------------------------------------------------------------------------------------
def _cpu_greedy_one(args):
    """One tensor through the oracle's restatement of MixedTileGreedyCompression (mixed_tile_greedy.py:72-352)."""
    shape, seed = args
    import numpy as np
    from oracle import qa_oracle as orc
    from quantization_analysis_b200 import synthetic
    x = synthetic.randn_f32_np(shape, seed)
    t0 = time.perf_counter()
    table = orc.tile_stat_table(x)
    a, counts = orc.greedy_assign(table, list(orc.MIXED_FORMATS), GREEDY["metric"], GREEDY["threshold"], GREEDY["seed"])
    y = orc.apply_assignment(x, a)
    orc.wq_scores(x, y)                                   # wq:684-687 scoring of the result
    return time.perf_counter() - t0, int(np.prod(shape))


CPU_SAMPLE = [((576, 7168), 1002), ((1536, 7168), 1000)]   # kv_a_proj + q_a_proj of the workload


def cpu_sample_run(procs: int):
    """Time the bounded CPU sample on `procs` host processes (the reference is single-threaded NumPy: one tensor per
    process, the two sample shapes alternating with different seeds until every core has one); returns (aggregate GB/s of
    bf16 weights, processes used, wall seconds)."""
    import multiprocessing as mp
    tasks = [(CPU_SAMPLE[i % len(CPU_SAMPLE)][0], CPU_SAMPLE[i % len(CPU_SAMPLE)][1] + 7 * (i // len(CPU_SAMPLE)))
             for i in range(max(procs, 1) if procs > 1 else len(CPU_SAMPLE))]
    t0 = time.perf_counter()
    if procs > 1:
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")      # one core per process: no oversubscription by np.dot
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        # spawn, not fork: a parent that already ran torch CPU ops has live OpenMP threads
        with mp.get_context("spawn").Pool(procs) as pool:
            res = pool.map(_cpu_greedy_one, tasks, chunksize=1)
    else:
        res = [_cpu_greedy_one(a) for a in tasks]
    wall = time.perf_counter() - t0
    elems = sum(r[1] for r in res)
    compute = max(r[0] for r in res) if procs > 1 else sum(r[0] for r in res)
    return 2.0 * elems / compute / 1e9, (procs if procs > 1 else 1), wall


CPU_SAMPLE_DESC = ("oracle port of mixed_tile_greedy (tile sums + greedy + apply + wq scoring) on kv_a_proj [576,7168] and "
                   "q_a_proj [1536,7168] tensors of the workload, one per host process on every core (the reference is "
                   "single-threaded NumPy); aggregate elements / slowest process; input generation excluded")


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs = max(1, os.cpu_count() or 1)
    for _ in range(args.warmup if args.warmup < 2 else 1):      # warm-up is page-cache / import warm only
        cpu_sample_run(procs)
    vals, t0 = [], time.perf_counter()
    for _ in range(args.steps):
        v, cores, _w = cpu_sample_run(procs)
        vals.append(v)
    wall = time.perf_counter() - t0
    value = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 sums over f32 values (NumPy)", "data": "synthetic",
            "config": config_dict(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": CPU_SAMPLE_DESC},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def maps_vs_reference_goldens(items, *result_sets):
    """Rank 0's tensors are the ones tests/golden/cfg2_bench_workload.* holds the UNMODIFIED reference's greedy maps for
    (made by tests/golden/make_golden.py cfg2): compare the maps this run produced - device-resident and end-to-end - with
    them.  True / False, or None when the fixtures are not there."""
    import numpy as np
    gold = ROOT / "tests" / "golden" / "cfg2_bench_workload.npz"
    if not gold.exists():
        return None
    maps = dict(np.load(gold))
    ok = True
    for results in result_sets:
        for (name, _shape, _seed), r in zip(items, results):
            key = name.split(".")[-2]
            ok = ok and key in maps and bool(np.array_equal(r["assignment"], maps[key]))
    return ok


# --------------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from quantization_analysis_b200 import _lib, synthetic
    from quantization_analysis_b200.batch import GreedyBatch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()
    W = max(3, args.warmup)
    K = max(1, args.steps)

    items = workload(rank)
    host = [synthetic.randn_bf16_cpu(shape, seed).pin_memory() for (_n, shape, seed) in items]
    batch = GreedyBatch([s for (_n, s, _sd) in items], **GREEDY, device=dev)
    batch.load_device(host)
    torch.cuda.synchronize()
    nbytes = batch.total_bytes()
    numel = nbytes // 2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ------------------------------------------------------
    # One CUDA graph per step (all tensors, all streams).  INFLIGHT tensor lists are double-buffered: a step's tail is
    # the latency-bound greedy chain on ~33 SMs, so the next list's tile-stat passes run underneath it (each list has its
    # own input / table / map buffers; a list's graph only starts after its own previous replay).
    batches = [batch] + [GreedyBatch([s for (_n, s, _sd) in items], **GREEDY, device=dev) for _ in range(INFLIGHT - 1)]
    for b in batches[1:]:
        b.load_device(host)
    lanes = [torch.cuda.Stream(device=dev) for _ in batches]
    for b in batches:
        b.capture()

    def run_steps(n_steps: int, lanes_used: int) -> None:
        cur = torch.cuda.current_stream(dev)
        for ln in lanes[:lanes_used]:
            ln.wait_stream(cur)
        for k in range(n_steps):
            with torch.cuda.stream(lanes[k % lanes_used]):
                batches[k % lanes_used].run_graph()
        for ln in lanes[:lanes_used]:
            cur.wait_stream(ln)

    run_steps(W, INFLIGHT)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # SM clocks / throttle reasons are sampled from here to the end of the end-to-end region (both timed regions); on rank 0
    # only: NVML queries from every rank of a node serialise in the driver and steal host time from the launch threads
    clk = ClockSampler(local, enabled=(rank == 0))
    clk.__enter__()
    barrier()
    e0.record()
    run_steps(K, INFLIGHT)
    e1.record()
    barrier()
    ms_local = e0.elapsed_time(e1)
    ms = reduce_max(ms_local)
    per_rank_ms = [ms_local / K]
    if world > 1:                                   # every rank's own device time per step (the value uses the max)
        t = torch.tensor([ms_local / K], dtype=torch.float64, device=dev)
        buf = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(buf, t)
        per_rank_ms = [float(b.item()) for b in buf]
    value = world * nbytes * K / (ms * 1e-3) / 1e9
    # latency of one step with nothing else in flight
    run_steps(2, 1)
    barrier()
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record()
    run_steps(K, 1)
    l1.record()
    barrier()
    ms_single = reduce_max(l0.elapsed_time(l1)) / K
    results = batch.collect()
    for b in batches[1:]:           # every in-flight list produced the same maps
        for r0, r1 in zip(results, b.collect()):
            assert (r0["assignment"] == r1["assignment"]).all() and r0["counts"] == r1["counts"]

    # ---- end to end: pinned host bf16 -> H2D -> path -> D2H of maps and metric rows ------------
    # Two batches alternate (enqueue_from_host / finish): the next list's H2D copies keep the PCIe link busy while the
    # previous list's chain finishes and its results travel back.  Every step's results are read on the host.
    def e2e_steps(n_steps: int):
        res = None
        for k in range(n_steps):
            b = batches[k % len(batches)]
            if k >= len(batches):
                res = b.finish()
            b.enqueue_from_host(host)
        for k in range(min(n_steps, len(batches))):
            res = batches[(n_steps - min(n_steps, len(batches)) + k) % len(batches)].finish()
        return res

    e2e_steps(2)
    barrier()
    t0 = time.perf_counter()
    res_e2e = e2e_steps(K)
    torch.cuda.synchronize()
    e2e_s = reduce_max(time.perf_counter() - t0)
    e2e_value = world * nbytes * K / e2e_s / 1e9
    clk.__exit__()
    # the ceiling of that number: a plain pinned-host -> device copy of the same bytes (the PCIe link of this GPU)
    big = max(range(len(host)), key=lambda i: host[i].numel())
    dst = torch.empty_like(host[big], device=dev)
    for _ in range(2):
        dst.copy_(host[big], non_blocking=True)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(5):
        dst.copy_(host[big], non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    h2d_copy_gbs = 5 * host[big].numel() * 2 / (c0.elapsed_time(c1) * 1e-3) / 1e9
    del dst
    if world > 1:                                   # per-tensor result rows to rank 0 (tiny)
        rows = [[r["metrics"]["pcc"], r["metrics"]["mae"], r["metrics"]["atol"]] for r in res_e2e]
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(rows, gathered, dst=0)

    # ---- per-kernel timing for the roofline (events on the launching stream, all tensors) ----
    def time_phase(stats: bool, assign: bool, reps: int = 5) -> float:
        for _ in range(2):
            batch.run_graph(stats=stats, assign=assign)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            batch.run_graph(stats=stats, assign=assign)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    n_stats_launches = sum(3 if (s["ntiles"] >= batch.PIPELINE_MIN_TILES and -(-s["rows"] // 32) >= 8) else 1 for s in batch.slots)
    ms_stats = time_phase(True, False)
    ms_assign = time_phase(False, True, reps=2)
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    ntiles = sum(s["ntiles"] for s in batch.slots)
    alg_stats = nbytes + ntiles * TABLE_BYTES_PER_TILE            # read x once + write the tile-stat table
    alg_assign = ntiles * (TABLE_BYTES_PER_TILE + 1)              # read the table once + write int8 map
    kernels = [
        {"kernel": "stats_fast_kernel", "ms_per_step": ms_stats, "launches_per_step": len(batch.slots),   # timed one launch per tensor
         "alg_bytes_per_step": alg_stats, "achieved_gbs": alg_stats / (ms_stats * 1e-3) / 1e9},
        {"kernel": "greedy_par_kernel (+ greedy_init_kernel, perm_resolve_chain_kernel, pa_* apply kernels on side streams)",
         "ms_per_step": ms_assign, "launches_per_step": batch.launches_per_step - n_stats_launches, "alg_bytes_per_step": alg_assign,
         "achieved_gbs": alg_assign / (ms_assign * 1e-3) / 1e9},
    ]
    # dram__bytes_read.sum + dram__bytes_write.sum of the o_proj tensor (117.4 M elements), from the `ncu --set full` captures
    # summarised in profiles/r1_summary.md.  stats: its three row-range launches, 235.0 MB read + 10.6 MB written before the
    # kernels end (algorithmic: 234.9 MB + 20.2 MB table).  greedy: the two chain launches 12.4 MB + init sums 0.6 MB
    # (algorithmic: 20.2 MB table once; the delta records and the table mostly hit in L2).
    ncu_traffic = {"stats_fast_kernel": 235032320 + 10577920, kernels[1]["kernel"]: 4258048 + 8105216 + 591360}
    dom = max(kernels, key=lambda k: k["ms_per_step"])
    roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": dom["achieved_gbs"] / peak, "traffic": ncu_traffic.get(dom["kernel"]),
                "traffic_note": "ncu dram bytes of the largest launch (o_proj); see profiles/r1_summary.md", "peak_source": peak_src,
                "alg_bytes_per_launch": dom["alg_bytes_per_step"] / dom["launches_per_step"],
                "avg_launch_ms": dom["ms_per_step"] / dom["launches_per_step"],
                "note": "dominant kernel by time in the step; per-kernel breakdown in roofline_by_kernel"}
    for k in kernels:
        k["frac"] = k["achieved_gbs"] / peak

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16 in; f32 group-scaled + f64 sums", "data": "synthetic", "config": config_dict(world),
                "clocks": clk.summary(),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": batch.d2h_bytes(),
                        "h2d_copy_gbs": h2d_copy_gbs,
                        "api": "GreedyBatch.enqueue_from_host(pinned bf16 host tensors) / finish() -> assignment maps + pcc/mae/atol on host, two batches alternating"},
                "gpu_launches": batch.launches_per_step * K,
                "step_latency_ms": ms_single,
                "per_rank_ms_per_step": per_rank_ms,
                "roofline": roofline, "roofline_by_kernel": kernels,
                "pct_of_8TBs": 100.0 * value / world / 8000.0,
                "result_check": {"counts_q_a_proj": results[0]["counts"], "pcc_q_a_proj": results[0]["metrics"]["pcc"],
                                 "maps_equal_reference": maps_vs_reference_goldens(items, results, res_e2e)}}
        if world == 1 and not args.no_cpu_baseline:
            # fresh interpreter (no CUDA context, no inherited thread pools), bounded by a timeout
            import subprocess
            try:
                out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1",
                                      "--warmup", "0"], capture_output=True, text=True, timeout=240, check=True)
                ref = json.loads(out.stdout.strip().splitlines()[-1])
                line["cpu_baseline"] = dict(ref["cpu_baseline"], host_cpus=os.cpu_count())
            except Exception as exc:  # report, never hang the GPU line
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                                        "sample": f"failed: {type(exc).__name__}: {exc}"[:300]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
