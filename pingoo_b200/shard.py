"""Batch sharding for multi-GPU runs: requests are independent, so rank g of G takes the
contiguous range [g*N/G, (g+1)*N/G) and no data-path collective is needed; verdicts are concatenated by rank."""
from .batch import RequestBatch


def shard_range(n_total: int, rank: int, world: int):
    lo = n_total * rank // world
    hi = n_total * (rank + 1) // world
    return lo, hi


def shard_batch(batch: RequestBatch, rank: int, world: int) -> RequestBatch:
    lo, hi = shard_range(batch.n, rank, world)
    return batch.slice(lo, hi)
