"""world_size-2 CPU worker for test_two_rank_gloo_top2_merge: exercises the N>1 host path of bench.py
(shard ranges, all-gather of the per-shard top-2 records, deterministic merge, frame sharding)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402
from viorb_b200 import sharding, synth  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    M, Q = 20000, 64
    dmap = synth.descriptor_map(M, seed=1234)          # every rank can regenerate the whole map (test only)
    dmap[::11] = dmap[3]
    q = synth.queries_from_map(dmap, Q, seed=7)
    b, e = sharding.shard_range(M, rank, world)
    part = O.hamming_top2(q, dmap[b:e], index_base=b)
    mine = torch.from_numpy(part.view(np.int32).reshape(Q, 4).copy())
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    parts = np.stack([g.numpy() for g in gathered]).view(O.TOP2).reshape(world, Q)
    merged = O.top2_merge(parts)
    full = O.hamming_top2(q, dmap)
    assert (merged == full).all()
    # frame sharding: the union of the shards is the batch, in order
    f0, f1 = sharding.shard_range(10, rank, world)
    counts = torch.tensor([f1 - f0])
    dist.all_reduce(counts)
    assert counts.item() == 10
    print("RANK%d OK" % rank, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
