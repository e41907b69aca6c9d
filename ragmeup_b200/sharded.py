"""Row-sharded flat index across the GPUs of one box.

The reference has no multi-device path; this is the B200-native equivalent the north star asks for:
corpus rows are split contiguously over the ranks of a ``torch.distributed`` group (one process per
GPU), every rank scans only its shard, and ONE all-gather of the per-shard ``[Q, k]`` candidates
(fp32 score + int64 global id, <= 150 KB per rank) over NCCL/NVLink feeds the on-device merge
(``rmu_topk_merge``).  Global ids are ``row + rank offset``, so results equal the unsharded search.
"""
from __future__ import annotations

from typing import Any, Callable, Optional, Tuple


class ShardedFlatIndex:
    def __init__(self, local_index: Any, group: Any = None, merge_fn: Optional[Callable] = None):
        """``local_index``: a FlatIndex holding this rank's rows.  ``merge_fn(scores[R,Q,k], ids[R,Q,k],
        metric) -> (scores[Q,k], ids[Q,k])`` defaults to the CUDA merge kernel (tests inject another)."""
        import torch.distributed as dist
        self.dist = dist
        self.index = local_index
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if merge_fn is None:
            from .index import topk_merge
            merge_fn = topk_merge
        self.merge_fn = merge_fn
        self.offset = 0
        self.total = len(local_index)
        self._sizes = [len(local_index)]

    def sync_offsets(self) -> None:
        """Exchange shard sizes; rank r's ids start at sum(sizes[:r]).  Call after (re)loading shards."""
        import torch
        n = len(self.index)
        if self.world == 1:
            self._sizes, self.offset, self.total = [n], 0, n
            return
        dev = getattr(self.index, "device", torch.device("cpu"))
        mine = torch.tensor([n], dtype=torch.int64, device=dev)
        sizes = [torch.zeros_like(mine) for _ in range(self.world)]
        self.dist.all_gather(sizes, mine, group=self.group)
        self._sizes = [int(s.item()) for s in sizes]
        self.offset = sum(self._sizes[: self.rank])
        self.total = sum(self._sizes)

    def search(self, queries, k: int) -> Tuple[Any, Any]:
        """Same queries on every rank -> the same merged (scores [Q,k], global ids [Q,k]) on every rank."""
        import torch
        s, i = self.index.search(queries, k, id_offset=self.offset)
        if self.world == 1:
            return s, i
        Q = s.shape[0]
        # ONE all-gather: score bits + 64-bit id packed as 3 x int32 per candidate
        pack = torch.empty((Q, k, 3), dtype=torch.int32, device=s.device)
        pack[..., 0] = s.contiguous().view(torch.int32)
        pack[..., 1:] = i.contiguous().view(torch.int32).view(Q, k, 2)
        parts = [torch.empty_like(pack) for _ in range(self.world)]
        self.dist.all_gather(parts, pack, group=self.group)
        allp = torch.stack(parts)                                    # [R, Q, k, 3]
        gs = allp[..., 0].contiguous().view(torch.float32)
        gi = allp[..., 1:].contiguous().view(torch.int64).view(self.world, Q, k)
        return self.merge_fn(gs, gi, self.index.metric)
