"""Host-side partitioning of the two data-parallel axes of the hot path.

  extraction : frames are independent -> contiguous blocks of the batch per rank, no collective;
  matching   : the map descriptor matrix is split by rows; every rank scans its shard for all queries,
               the per-shard top-2 records (16 B per query) are all-gathered and merged with the total
               order (distance, global index) -- identical to one sequential scan of the whole map.
"""


def shard_range(n, rank, world):
    """contiguous [begin, end) of item indices owned by `rank`"""
    return n * rank // world, n * (rank + 1) // world


def all_shards(n, world):
    return [shard_range(n, r, world) for r in range(world)]
