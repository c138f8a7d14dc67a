"""World sharding across the GPUs of one box.

Worlds never interact (each reference WorldPhysics owns its own dWorldID, src/WorldPhysics.hpp:99),
so a batch shards by contiguous world-id ranges with no collective on the step path.  Per-world
randomisation is keyed by the GLOBAL world id, which makes results independent of the GPU count.
"""


def shard_range(total_worlds, rank, world_size):
    """Contiguous [begin, end) of global world ids owned by `rank`; sizes differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(total_worlds, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def owner_of(world_id, total_worlds, world_size):
    base, rem = divmod(total_worlds, world_size)
    split = rem * (base + 1)
    if world_id < split:
        return world_id // (base + 1)
    return rem + (world_id - split) // base
