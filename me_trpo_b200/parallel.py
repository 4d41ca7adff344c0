"""Row sharding of the imaginary rollout across GPUs: every row is independent for
the whole horizon, so rank r of G owns a contiguous block of rows and the rollout needs no
collective.  Noise streams are indexed by GLOBAL row, so a sharded run equals the unsharded one."""


def shard_rows(n_rows, rank, world_size):
    """Contiguous [lo, hi) block of rank; the first n_rows % world_size ranks get one extra row."""
    base, extra = divmod(int(n_rows), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)
