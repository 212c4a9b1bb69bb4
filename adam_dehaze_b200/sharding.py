"""Image sharding across the GPUs of one box: every image is independent in eval mode, so rank r of W
takes a contiguous slice of the batch, classifies, buckets and dehazes it locally and writes its slice of the output.
No data-path collective exists at inference."""


def shard_bounds(total, rank, world):
    """[lo, hi) of rank's contiguous shard; the first total % world ranks carry one extra image."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(int(total), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(total, world):
    return [shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world)]
