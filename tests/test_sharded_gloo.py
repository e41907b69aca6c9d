"""CPU, world_size 2 over gloo: the N>1 plumbing of ShardedFlatIndex (offset exchange, the single
packed all-gather, merge call) with the oracle standing in for the per-shard CUDA search/merge."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class OracleShard:
    """FlatIndex look-alike over numpy (the product index needs a GPU; only the plumbing is under test)."""

    def __init__(self, x, metric):
        self.x, self.metric, self.device = x, metric, torch.device("cpu")

    def __len__(self):
        return self.x.shape[0]

    def add(self, v):
        self.x = np.concatenate([self.x, np.asarray(v, np.float32)])

    def search(self, q, k, id_offset=0):
        from oracle import flat_ref
        s, i = flat_ref.flat_search(q.numpy(), self.x, k, self.metric, id_offset=id_offset)
        return torch.from_numpy(s), torch.from_numpy(i)


def oracle_merge(gs, gi, metric):
    from oracle import flat_ref
    s, i = flat_ref.shard_merge(gs.numpy(), gi.numpy(), metric)
    return torch.from_numpy(s), torch.from_numpy(i)


def _worker(rank, world, port, metric, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import flat_ref
        from ragmeup_b200.sharded import ShardedFlatIndex
        rng = np.random.default_rng(7)
        x = rng.standard_normal((1001, 48)).astype(np.float32)
        x[900] = x[5]                                 # duplicate across shards -> tie broken by global id
        q = rng.standard_normal((6, 48)).astype(np.float32)
        q[0] = x[5]
        bounds = [0, 333, 1001]                        # uneven shards
        shard = OracleShard(x[bounds[rank]:bounds[rank + 1]], metric)
        sh = ShardedFlatIndex(shard, merge_fn=oracle_merge)
        sh.sync_offsets()
        assert sh.offset == bounds[rank] and sh.total == 1001
        s, i = sh.search(torch.from_numpy(q), 12)
        fs, fi = flat_ref.flat_search(q, x, 12, metric)
        ok = bool((i.numpy() == fi).all()) and bool(np.allclose(s.numpy(), fs, atol=1e-6))
        out[rank] = ok
    finally:
        dist.destroy_process_group()


def _ingest_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import flat_ref
        from ragmeup_b200.sharded import ShardedFlatIndex
        rng = np.random.default_rng(11)
        batches = [rng.standard_normal((n, 24)).astype(np.float32) for n in (7, 10, 1, 5, 1000)]   # the reference adds 1000 at a time
        sh = ShardedFlatIndex(OracleShard(np.zeros((0, 24), np.float32), "ip"), merge_fn=oracle_merge)
        sh.sync_offsets()
        kept = 0
        for b in batches:
            lo, hi = sh.add(b)
            kept += hi - lo
        x = np.concatenate(batches)
        q = rng.standard_normal((5, 24)).astype(np.float32)
        assert len(sh.index) == kept and sh.total == len(x)
        s, i = sh.search(torch.from_numpy(q), 9)
        fs, fi = flat_ref.flat_search(q, x, 9, "ip")
        got = sh.to_insertion_order(i)
        out[rank] = bool((got == fi).all()) and bool(np.allclose(s.numpy(), fs, atol=1e-6))
    finally:
        dist.destroy_process_group()


def test_sharded_ingest_block_partition_gloo():
    """ShardedFlatIndex.add: every rank keeps its block of every batch, offsets stay in sync without a collective,
    and merged ids map back to the order the documents were added in (server/RAGHelper.py:423-434 is the loop)"""
    world = 2
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_ingest_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


@pytest.mark.parametrize("metric", ["l2", "ip"])
def test_sharded_search_equals_unsharded_gloo(metric):
    world = 2
    import socket
    with socket.socket() as sk:                      # a port the OS says is free right now
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, metric, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
