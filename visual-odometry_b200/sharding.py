"""Query sharding for the multi-GPU nearest-neighbour sweep (BASELINE.json config 4).

The path shards over independent units: each query's answer depends only on the (replicated)
map, so rank r answers the contiguous block [r*Q/W, (r+1)*Q/W) and the only exchange is an
all-gather of the int32 match indices (400 KB for Q = 1e5) — NCCL over NVLink on the GPUs,
gloo in the CPU tests.  No collective touches the data path itself.
"""


def shard_bounds(n_queries, world_size, rank):
    """[lo, hi) of rank's queries; blocks differ by at most one query."""
    return rank * n_queries // world_size, (rank + 1) * n_queries // world_size


def shard_counts(n_queries, world_size):
    return [shard_bounds(n_queries, world_size, r)[1] - shard_bounds(n_queries, world_size, r)[0]
            for r in range(world_size)]


def gather_indices(dist, idx_shard, idx_all, n_queries):
    """all-gather of every rank's int32 indices into idx_all (a torch tensor of n_queries).
    Equal shards use the flat collective directly; ragged ones (Q not a multiple of the world
    size) are padded to the longest shard, gathered, and un-padded."""
    import torch

    world = dist.get_world_size()
    counts = shard_counts(n_queries, world)
    if len(set(counts)) == 1:
        dist.all_gather_into_tensor(idx_all, idx_shard)
        return idx_all
    longest = max(counts)
    padded = torch.full((longest,), -1, dtype=idx_shard.dtype, device=idx_shard.device)
    padded[: idx_shard.numel()] = idx_shard
    buf = torch.empty(world * longest, dtype=idx_shard.dtype, device=idx_shard.device)
    dist.all_gather_into_tensor(buf, padded)
    lo = 0
    for r, c in enumerate(counts):
        idx_all[lo:lo + c] = buf[r * longest:r * longest + c]
        lo += c
    return idx_all
