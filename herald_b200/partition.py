"""Row-range sharding of one embedding table over the ranks of a group — the rule ps-lite's
AveragePartitioner applies to servers (ps-lite/include/ps/partitioner.h:46-62): ``length // G``
rows per shard, the first ``length % G`` shards one more; a sorted key list therefore splits into
contiguous per-owner slices (ps-lite/include/ps/worker/PSAgent.h:541-559).  Host-side mirror of
`owner_of` / `shard_begin` in csrc/hb_cache.cuh; used by the launcher, the bench and the tests.
"""
import numpy as np


def shard_range(rank, world, length):
    """(row_begin, nrows) of `rank`'s shard."""
    per, rem = divmod(int(length), int(world))
    begin = rank * per + min(rank, rem)
    return begin, per + (1 if rank < rem else 0)


def owner_of(keys, world, length):
    """Owner rank and row inside the owner's shard for every key."""
    keys = np.asarray(keys, np.uint64)
    per, rem = divmod(int(length), int(world))
    cut = np.uint64(rem * (per + 1))
    head = keys < cut
    owner = np.where(head, keys // np.uint64(per + 1),
                     np.uint64(rem) + (keys - cut) // np.uint64(max(per, 1)))
    begin = np.where(owner < rem, owner * np.uint64(per + 1),
                     cut + (owner - np.uint64(rem)) * np.uint64(per))
    return owner.astype(np.int64), (keys - begin).astype(np.uint64)


def split_sorted(sorted_keys, world, length):
    """Slice bounds ``lo[0..world]`` of an ascending key array: owner o gets [lo[o], lo[o+1])."""
    starts = np.array([shard_range(o, world, length)[0] for o in range(world)] + [int(length)],
                      np.uint64)
    return np.searchsorted(np.asarray(sorted_keys, np.uint64), starts, side="left")
