"""Multi-GPU partitioning of the quantize-and-score path (one process per GPU, torch.distributed).

The path shards by independent units, so there is no data-path collective:
  * tensors   - the matched-tensor list is bin-packed over ranks by element count (cfg5: 96 of the
                768 expert matrices per GPU);
  * row stripes - one large [R, C] tensor is cut into contiguous stripes of 32*k rows; a stripe is a
                contiguous bf16 range of the input and a contiguous tile range of the tile-stat table.
The only exchange is for a *global* assignment over a striped tensor: the per-stripe tables
(<= 176 B per 1024 elements) are all-gathered in tile order and the greedy runs once.
Works with NCCL (CUDA tensors) and gloo (CPU tensors; used by the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

TILE = 32


def partition_tensors(sizes, world: int) -> list[list[int]]:
    """Longest-processing-time bin packing: indices of `sizes` per rank, deterministic."""
    bins = [[] for _ in range(world)]
    load = [0] * world
    for i in sorted(range(len(sizes)), key=lambda k: (-int(sizes[k]), k)):
        r = min(range(world), key=lambda q: (load[q], q))
        bins[r].append(i)
        load[r] += int(sizes[i])
    return [sorted(b) for b in bins]


def row_stripes(rows: int, world: int) -> list[tuple[int, int]]:
    """Contiguous [start, end) row ranges, each a multiple of 32 rows (the last takes the ragged rest)."""
    tiles_h = -(-rows // TILE)
    base, extra = divmod(tiles_h, world)
    out, t = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((min(rows, t * TILE), min(rows, (t + n) * TILE)))
        t += n
    return out


def gather_tables(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather per-stripe tables [NSTAT, ntiles_r] into the full table [NSTAT, sum ntiles_r], tile order.
    One collective for the tile counts and one for the (zero-padded) tables: `all_gather_into_tensor` where the backend has
    it (NCCL: a single ring/NVLS operation over the whole payload), the list form otherwise (gloo on CPU)."""
    world = dist.get_world_size(group)
    n_local = torch.tensor([local.shape[1]], dtype=torch.int64, device=local.device)
    fused = str(dist.get_backend(group)).lower() == "nccl"     # decided by the backend: the same on every rank
    if fused:
        allc = torch.empty(world, dtype=torch.int64, device=local.device)
        dist.all_gather_into_tensor(allc, n_local, group=group)
        counts = [int(c) for c in allc.tolist()]
    else:
        parts_c = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(parts_c, n_local, group=group)
        counts = [int(c.item()) for c in parts_c]
    width = max(counts)
    if local.shape[1] == width:
        padded = local.contiguous()
    else:
        padded = torch.zeros((local.shape[0], width), dtype=local.dtype, device=local.device)
        padded[:, : local.shape[1]] = local
    if fused:
        out = torch.empty((world,) + tuple(padded.shape), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, padded, group=group)
        parts = [out[r] for r in range(world)]
    else:
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded.contiguous(), group=group)
    return torch.cat([p[:, :c] for p, c in zip(parts, counts)], dim=1).contiguous()


def gather_rows(rows, dst: int = 0, group=None):
    """Per-tensor result rows (small Python objects) to `dst`; returns the flat list there, else None."""
    world = dist.get_world_size(group)
    out = [None] * world if dist.get_rank(group) == dst else None
    dist.gather_object(rows, out, dst=dst, group=group)
    if out is None:
        return None
    return [row for part in out for row in part]


def striped_greedy(x_stripe: torch.Tensor, cols: int, total_numel: int, metric: str, threshold: float, seed: int,
                   tile_formats, group=None):
    """Global greedy over a tensor whose row stripes live on different ranks (CUDA, NCCL).
    Every rank computes its stripe's table; the tables are all-gathered in tile order; every rank then runs the
    (deterministic) greedy on the full table and keeps its slice of the map - cheaper than a broadcast."""
    from . import engine
    p = engine.prepare_tiles(x_stripe.reshape(-1, cols))
    local = engine.tile_stats(p, engine.MIXED_FORMATS, exact_abs=(metric == "mae"))
    full = gather_tables(local, group)
    rng = engine.make_rng(seed, full.device)
    assignment, counts, state = engine.greedy_assign(full, total_numel, metric, threshold, list(tile_formats), rng)
    return assignment, counts, state, full
