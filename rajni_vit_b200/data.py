"""Loader-side data parallelism: every rank decodes only ITS shard of each global batch.

The reference builds one ``DataLoader(ImageFolder, batch_size=B, shuffle=False)`` (``rajni/run.py:72-82``) and has no
notion of ranks.  ``evaluate_model`` can shard any foreign loader by slicing each batch after it was loaded — correct,
but every rank then decodes and resizes the WHOLE batch and throws (N-1)/N of it away.  When the loader is ours, the
sharding moves into the sampler: ``RankShardedBatchSampler`` yields, for global batch k = indices [kB, (k+1)B), exactly
the contiguous slice ``shard_bounds`` would have cut for this rank, so the images a rank sees (and their order) are the
same as on the slice path, with 1/N of the decode work.  A loader built this way carries ``rajni_sharded = True`` and
``evaluate_model`` then leaves its batches alone (the final all-reduce of the counters is unchanged).
"""
from __future__ import annotations

from typing import Iterator, List

import torch

from .eval import shard_bounds


class RankShardedBatchSampler(torch.utils.data.Sampler):
    """Batch sampler over ``range(n)`` in order: global batches of ``batch_size`` (last one partial unless ``drop_last``),
    each cut into ``world`` contiguous shards; this rank's shard is yielded.  Shards that come out empty (fewer images
    than ranks in the last batch) are skipped."""

    def __init__(self, n: int, batch_size: int, rank: int, world: int, drop_last: bool = False):
        if not 0 <= rank < world:
            raise ValueError(f"rank {rank} outside world of {world}")
        self.n, self.batch_size, self.rank, self.world, self.drop_last = n, batch_size, rank, world, drop_last

    def _batches(self) -> Iterator[List[int]]:
        for start in range(0, self.n, self.batch_size):
            size = min(self.batch_size, self.n - start)
            if size < self.batch_size and self.drop_last:
                return
            lo, hi = shard_bounds(size, self.rank, self.world)
            if hi > lo:
                yield list(range(start + lo, start + hi))

    def __iter__(self):
        return self._batches()

    def __len__(self):
        return sum(1 for _ in self._batches())


def sharded_loader(dataset, batch_size: int, rank: int, world: int, **kw) -> torch.utils.data.DataLoader:
    """``DataLoader(dataset, batch_size, shuffle=False)`` whose batches are this rank's shards (see module docstring)."""
    kw.pop("shuffle", None)
    drop_last = kw.pop("drop_last", False)
    loader = torch.utils.data.DataLoader(
        dataset, batch_sampler=RankShardedBatchSampler(len(dataset), batch_size, rank, world, drop_last), **kw)
    loader.rajni_sharded = True
    return loader
