"""torchrun worker: row-striped tensor across ranks -> NCCL all-gather of tile tables -> global greedy.
Rank 0 checks the map against the single-GPU result on the full tensor."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist

from quantization_analysis_b200 import engine, sharding, synthetic


def main() -> int:
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    shape = (2048, 1536)
    x = synthetic.randn_bf16_cpu(shape, 77)                      # same tensor on every rank; each keeps its stripe
    a, b = sharding.row_stripes(shape[0], world)[rank]
    stripe = x[a:b].cuda()
    assignment, counts, state, full = sharding.striped_greedy(stripe, shape[1], x.numel(), "pcc", 0.999, 123,
                                                              list(engine.MIXED_FORMATS))
    ok = True
    if rank == 0:
        p = engine.prepare_tiles(x.cuda())
        table = engine.tile_stats(p, engine.MIXED_FORMATS, exact_abs=False)
        a1, c1, _s = engine.greedy_assign(table, p.numel, "pcc", 0.999, list(engine.MIXED_FORMATS), engine.make_rng(123))
        ok = bool(torch.equal(full, table) and torch.equal(a1, assignment) and torch.equal(c1, counts))
        print("striped_greedy_ok" if ok else "striped_greedy_MISMATCH", counts.tolist(), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
