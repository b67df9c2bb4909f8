"""CPU: the N>1 host logic on a world_size-2 gloo group (partitioning, table all-gather, result gather)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from quantization_analysis_b200 import sharding


def test_partition_is_balanced_and_complete():
    sizes = [11010048, 37748736, 4128768, 16777216, 117440512] + [14680064] * 24
    for world in (1, 2, 4, 8):
        parts = sharding.partition_tensors(sizes, world)
        assert sorted(i for p in parts for i in p) == list(range(len(sizes)))
        loads = [sum(sizes[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(sizes)


def test_row_stripes_cover_rows_in_tile_multiples():
    for rows in (7168, 70, 32, 1000):
        for world in (1, 2, 3, 8):
            st = sharding.row_stripes(rows, world)
            assert st[0][0] == 0 and st[-1][1] == rows
            for (a, b), (c, d) in zip(st, st[1:]):
                assert b == c
            assert all(a % 32 == 0 or a == rows for a, _ in st)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import qa_oracle as orc
        from quantization_analysis_b200 import synthetic
        x = synthetic.heterogeneous_f32_np((160, 96), 21)
        a, b = sharding.row_stripes(x.shape[0], world)[rank]
        t = orc.tile_stat_table(x[a:b])
        cols = [t["sx"], t["sx2"]] + [t[f][k] for f in orc.MIXED_FORMATS for k in ("sy", "sy2", "sxy", "sabs", "amax")]
        local = torch.from_numpy(np.stack(cols, axis=0))
        full = sharding.gather_tables(local)
        tf = orc.tile_stat_table(x)
        want = np.stack([tf["sx"], tf["sx2"]] + [tf[f][k] for f in orc.MIXED_FORMATS
                                                  for k in ("sy", "sy2", "sxy", "sabs", "amax")], axis=0)
        ok_table = bool(np.array_equal(full.numpy(), want))
        # the global greedy over the gathered table equals the greedy over the unsharded tensor
        tbl = {"th": tf["th"], "tw": tf["tw"], "numel": x.size, "sx": full[0].numpy(), "sx2": full[1].numpy()}
        for i, f in enumerate(orc.MIXED_FORMATS):
            tbl[f] = {k: full[2 + 5 * i + j].numpy() for j, k in enumerate(("sy", "sy2", "sxy", "sabs", "amax"))}
        a1, c1 = orc.greedy_assign(tbl, list(orc.MIXED_FORMATS), "pcc", 0.995, 5)
        a2, c2 = orc.greedy_assign(tf, list(orc.MIXED_FORMATS), "pcc", 0.995, 5)
        rows = sharding.gather_rows([{"rank": rank, "counts": c1}], dst=0)
        if rank == 0:
            q.put((ok_table, bool(np.array_equal(a1, a2)), c1 == c2, len(rows) == world))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_table_gather_and_global_greedy():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == (True, True, True, True)
