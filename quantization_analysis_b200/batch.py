"""Multi-tensor batching of the quantize-and-score path (the per-tensor loop of wq:655-709).

``GreedyBatch`` owns every device buffer for a list of same-run tensors and enqueues, per
tensor, the fused tile-stat pass, the greedy assignment and the whole-tensor sums on a small
pool of CUDA streams, with no host synchronisation until ``collect()``.  Tensors are independent
(the reference processes them one after another), so a rank's shard is just a sub-list.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, engine
from ._lib import METRIC_CODE, NFMT, NSTAT, STATS_FAST, STATS_FAST_APPROX_ABS, check

MIXED = engine.MIXED_FORMATS


class GreedyBatch:
    def __init__(self, shapes, metric: str = "pcc", threshold: float = 0.999, seed: int = 123,
                 tile_formats=MIXED, n_streams: int | None = None, device=None):
        self.device = device or engine._require_cuda()
        self.metric, self.threshold, self.seed = metric, float(threshold), int(seed)
        self.tile_formats = list(tile_formats)
        self.shapes = [tuple(int(v) for v in s) for s in shapes]
        n_streams = min(16, len(self.shapes)) if n_streams is None else n_streams     # one stream per tensor
        # Cluster kernels (resolve / init / chain) can only start when a whole GPC's worth of SMs is free at once, which a
        # streaming kernel that keeps refilling every SM rarely allows: they go on high-priority streams, the tile-stat
        # kernels on normal-priority ones, so the block scheduler drains SMs for a pending cluster first.
        self.streams = [torch.cuda.Stream(device=self.device, priority=-1) for _ in range(max(1, n_streams))]
        self.side_streams = [torch.cuda.Stream(device=self.device, priority=-1) for _ in range(max(1, n_streams))]
        self.side2_streams = [torch.cuda.Stream(device=self.device) for _ in range(max(1, n_streams))]
        # The tile-stat kernels are bandwidth-bound and each fills the GPU: run concurrently they only slow each other
        # down, and the largest tensor - whose init/chain is the critical path - would get its table last.  A short
        # list of large tensors therefore streams its tile-stat passes one after the other, largest first; a long list
        # of small ones (one MoE layer = 768 tensors) keeps a few streams to hide launch gaps.
        self.stats_streams = [torch.cuda.Stream(device=self.device) for _ in range(1 if len(self.shapes) <= 16 else 4)]
        self._stats_rr = 0
        self.prefetch = metric != "atol" and len(self.tile_formats) >= 2
        L = _lib.lib()
        self.slots = []
        rng0 = engine.make_rng(self.seed, self.device)
        for (r, c) in self.shapes:
            nt = (-(-r // 32)) * (-(-c // 32))
            self.slots.append({
                "rows": r, "cols": c, "ntiles": nt, "numel": r * c,
                "x": torch.empty(r * c, dtype=torch.bfloat16, device=self.device),
                "table": torch.zeros((NSTAT, nt), dtype=torch.float64, device=self.device),
                "assignment": torch.empty(nt, dtype=torch.int8, device=self.device),
                "counts": torch.zeros(NFMT, dtype=torch.int64, device=self.device),
                "state": torch.zeros(24, dtype=torch.float64, device=self.device),
                "sums": torch.zeros(8, dtype=torch.float64, device=self.device),
                "work": torch.empty(max(L.qa_greedy_work_bytes(nt), L.qa_greedy_par_work_bytes(nt)), dtype=torch.uint8,
                                    device=self.device),
                "rng": rng0.clone(),
                "pre_order": torch.empty((2, nt), dtype=torch.int32, device=self.device),
                "rngs": torch.stack([rng0, rng0, rng0]).contiguous(),      # stream states after permutations #1, #2, #3
                "init": torch.empty(L.qa_greedy_init_bytes(nt), dtype=torch.uint8, device=self.device),
                "jarr": torch.empty((3, nt), dtype=torch.int32, device=self.device),
                "apply_work": torch.empty((2, L.qa_perm_apply_work_bytes(nt)), dtype=torch.uint8, device=self.device),
                "ev": torch.cuda.Event(),
            })
        self._rng0 = rng0
        self._graphs = {}
        self.trace = None            # set to {} to record per-tensor stage events during eager run() (see timeline())
        self._order = _lib.int32_array([engine.FMT_INDEX[f] for f in self.tile_formats])
        # kernels of ours per tensor and step.  atol: tile_stats + greedy + assignment_sums.  pcc / mae: tile_stats +
        # greedy_init + greedy chain, plus the prefetched permutations (one chained resolve kernel, 7 grid kernels per apply)
        if metric == "atol":
            per_tensor = 3
        elif not self.prefetch:
            per_tensor = 3
        else:
            per_tensor = 4 + (1 + 2 * 7 if len(self.tile_formats) >= 3 else 1 + 7)
        self.launches_per_step = per_tensor * len(self.slots)

    # ---- data movement -------------------------------------------------------------------
    def load_device(self, tensors) -> None:
        """Copy bf16 tensors (host or device) into the resident input buffers."""
        for slot, t in zip(self.slots, tensors):
            slot["x"].copy_(t.reshape(-1), non_blocking=True)

    def total_bytes(self) -> int:
        return sum(2 * s["numel"] for s in self.slots)

    # ---- compute -------------------------------------------------------------------------
    def _enqueue(self, slot, stream, stats: bool = True, assign: bool = True, side=None) -> None:
        L = _lib.lib()
        sp = stream.cuda_stream
        pre = assign and self.prefetch and side is not None

        def mark(tag, on=None):
            if self.trace is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(on or stream)
                self.trace.setdefault(id(slot), {})[tag] = ev

        mark("start")
        if pre:
            # the first permutations depend only on (seed, ntiles): draw them on side streams while the tile-stat pass
            # streams the tensor.  Resolves chain through the RNG state (#1 -> #2 -> #3, one cluster each); each apply
            # only needs its own swap targets and runs as grid kernels, #2's on a second side stream next to resolve #3.
            side, side2 = side
            n, three = slot["ntiles"], len(self.tile_formats) >= 3
            side.wait_stream(stream)
            sa = side.cuda_stream
            # one launch for the chained resolves (#1: stream position only): the cluster keeps its SMs while the
            # tile-stat kernels flood the rest of the GPU
            check(L.qa_perm_resolve_chain(self._rng0.data_ptr(), n, 3 if three else 2, 0b110 if three else 0b010,
                                          slot["jarr"].data_ptr(), slot["rngs"].data_ptr(), sa), "qa_perm_resolve_chain")
            mark("resolve", side)
            if three:
                slot["ev"].record(side)
                side2.wait_event(slot["ev"])
                check(L.qa_perm_apply(slot["jarr"][1].data_ptr(), n, None, slot["pre_order"][0].data_ptr(),
                                      slot["apply_work"][0].data_ptr(), side2.cuda_stream), "qa_perm_apply")
                check(L.qa_perm_apply(slot["jarr"][2].data_ptr(), n, None, slot["pre_order"][1].data_ptr(),
                                      slot["apply_work"][1].data_ptr(), sa), "qa_perm_apply")
            else:
                check(L.qa_perm_apply(slot["jarr"][1].data_ptr(), n, None, slot["pre_order"][0].data_ptr(),
                                      slot["apply_work"][0].data_ptr(), sa), "qa_perm_apply")
        if pre:
            mark("prefetch", side)
        if stats:
            ss = self.stats_streams[self._stats_rr % len(self.stats_streams)]
            self._stats_rr += 1
            ss.wait_stream(stream)                       # the tensor's input is in place (H2D copy / previous pass)
            check(L.qa_tile_stats(slot["x"].data_ptr(), _lib.QA_DT_BF16, slot["rows"], slot["cols"], slot["cols"], 0,
                                  0xF, STATS_FAST if self.metric == "mae" else STATS_FAST_APPROX_ABS,
                                  slot["table"].data_ptr(), ss.cuda_stream), "qa_tile_stats")
            stream.wait_stream(ss)
            mark("stats")
        if assign:
            slot["rng"].copy_(self._rng0, non_blocking=True)      # every tensor restarts the seeded stream
            args = (slot["table"].data_ptr(), slot["ntiles"], float(slot["numel"]),
                    METRIC_CODE[self.metric], self.threshold, self._order, len(self.tile_formats),
                    slot["rng"].data_ptr(), slot["assignment"].data_ptr(), slot["counts"].data_ptr(),
                    slot["state"].data_ptr(), slot["work"].data_ptr())
            if self.metric == "atol":
                check(L.qa_greedy_assign(*args, sp), "qa_greedy_assign")
            else:
                # initial sums + delta records on the main stream (overlaps the prefetch), then the chain
                iargs = (slot["table"].data_ptr(), slot["ntiles"], METRIC_CODE[self.metric], self._order,
                         len(self.tile_formats), slot["init"].data_ptr())
                if stats:
                    # the delta records are a plain grid kernel: on the tile-stat stream right behind this tensor's pass,
                    # next to the (latency-bound, one-cluster) sums rather than in front of them
                    check(L.qa_greedy_init_deltas(*iargs, ss.cuda_stream), "qa_greedy_init_deltas")
                    check(L.qa_greedy_init_sums(*iargs, sp), "qa_greedy_init_sums")
                    stream.wait_stream(ss)
                else:
                    check(L.qa_greedy_init(*iargs, sp), "qa_greedy_init")
                mark("init")
                if pre:
                    stream.wait_stream(side)
                    if len(self.tile_formats) >= 3:
                        stream.wait_stream(side2)
                check(L.qa_greedy_assign_par_pre(*args, slot["pre_order"].data_ptr() if pre else None,
                                                 slot["rngs"][1].data_ptr() if pre else None, slot["init"].data_ptr(), sp),
                      "qa_greedy_assign_par_pre")
            mark("chain")
            if self.metric == "atol":      # the cluster kernel leaves the final sums (and max) in `state`
                check(L.qa_assignment_sums(slot["table"].data_ptr(), slot["ntiles"], slot["assignment"].data_ptr(), -1,
                                           slot["sums"].data_ptr(), sp), "qa_assignment_sums")

    def run(self, stats: bool = True, assign: bool = True) -> None:
        """Enqueue one pass over every tensor; returns immediately (no host sync)."""
        cur = torch.cuda.current_stream(self.device)
        order = sorted(range(len(self.slots)), key=lambda i: -self.slots[i]["ntiles"])   # longest chain first
        for k, i in enumerate(order):
            st = self.streams[k % len(self.streams)]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                self._enqueue(self.slots[i], st, stats, assign, side=(self.side_streams[k % len(self.side_streams)], self.side2_streams[k % len(self.side2_streams)]))
        for st in self.streams:
            cur.wait_stream(st)

    def capture(self, stats: bool = True, assign: bool = True) -> None:
        """Record one pass (all tensors, all streams) into a CUDA graph; ``run_graph`` replays it with one launch.
        The pass is ~25 small launches per tensor, so eager enqueueing costs about as much host time as the GPU
        needs to run it; the inputs are the resident ``slot['x']`` buffers, so replays see whatever was loaded last."""
        self.run(stats, assign)                       # eager once: module load, kernel attributes, allocator warm-up
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.run(stats, assign)
        self._graphs[(stats, assign)] = g

    def run_graph(self, stats: bool = True, assign: bool = True) -> None:
        if (stats, assign) not in self._graphs:
            self.capture(stats, assign)
        self._graphs[(stats, assign)].replay()

    def timeline(self) -> list[dict]:
        """After an eager run() with ``self.trace = {}``: per tensor, ms from its stream's start mark to each stage's end."""
        torch.cuda.synchronize(self.device)
        out = []
        t0 = min((tr["start"] for tr in self.trace.values()), key=lambda e: 0)     # any start (all follow the same fork)
        for s in self.slots:
            tr = self.trace.get(id(s), {})
            row = {"shape": (s["rows"], s["cols"])}
            for tag, ev in tr.items():
                if tag != "start":
                    row[tag] = t0.elapsed_time(ev)
            out.append(row)
        return out

    def run_from_host(self, host_tensors) -> list[dict]:
        """End-to-end pass: pinned host bf16 -> device, quantize+score+assign, results back to host."""
        cur = torch.cuda.current_stream(self.device)
        order = sorted(range(len(self.slots)), key=lambda i: -self.slots[i]["ntiles"])
        for k, i in enumerate(order):
            st = self.streams[k % len(self.streams)]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                self.slots[i]["x"].copy_(host_tensors[i].reshape(-1), non_blocking=True)
                self._enqueue(self.slots[i], st, side=(self.side_streams[k % len(self.side_streams)], self.side2_streams[k % len(self.side2_streams)]))
        for st in self.streams:
            cur.wait_stream(st)
        return self.collect()

    def collect(self) -> list[dict]:
        """Device -> host: assignment maps, counts and exact pcc/mae/atol per tensor (synchronises)."""
        out = []
        packs = []
        for s in self.slots:
            packs.append((s["assignment"].to("cpu", non_blocking=True), s["counts"].to("cpu", non_blocking=True),
                          (s["sums"] if self.metric == "atol" else s["state"]).to("cpu", non_blocking=True)))
        torch.cuda.current_stream(self.device).synchronize()
        for s, (a, c, sums) in zip(self.slots, packs):
            counts = {f: int(c[i]) for i, f in enumerate(MIXED)}
            v = sums.numpy()
            if self.metric != "atol":      # state = {sx, sx2, sy, sy2, sxy, sabs, flags, value, cycles x3, max|x-y|, ...}
                v = np.array([v[0], v[1], v[2], v[3], v[4], v[5], v[11]])
            out.append({"assignment": a.numpy().reshape(-(-s["rows"] // 32), -(-s["cols"] // 32)), "counts": counts,
                        "metrics": engine.metrics_from_sums(v, s["numel"]), "state": sums.numpy().copy()})
        return out

    def d2h_bytes(self) -> int:
        return sum(s["ntiles"] + 8 * NFMT + 8 * 8 for s in self.slots)
