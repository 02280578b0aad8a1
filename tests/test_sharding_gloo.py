"""CPU, world_size 2 over gloo: DB sharding and the final hit gather (the only multi-rank step of the path)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from reseek_b200.shard import gather_hits, partition_by_residues


def test_partition_balances_residues():
    rng = np.random.default_rng(0)
    lens = rng.integers(30, 1400, size=5000)
    for world in (1, 2, 3, 8):
        parts = partition_by_residues(lens, world)
        assert parts[0][0] == 0 and parts[-1][1] == len(lens)
        assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
        tot = [int(lens[lo:hi].sum()) for lo, hi in parts]
        assert max(tot) - min(tot) <= 2 * lens.max()
    assert partition_by_residues([5, 5], 4)[-1][1] == 2  # more ranks than chains: empty shards are allowed


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from reseek_b200.lib import HIT_DTYPE
    lens = np.arange(10, 110)
    lo, hi = partition_by_residues(lens, world)[rank]
    # fake "hits": one per local DB chain x 2 queries, score encodes the global pair
    n = (hi - lo) * 2
    hits = np.zeros(n, HIT_DTYPE)
    hits["a"] = np.repeat(np.arange(hi - lo), 2)
    hits["b"] = np.tile(np.arange(2), hi - lo)
    hits["score"] = (hits["a"] + lo) * 10 + hits["b"]
    allh = gather_hits(hits, lo, dist, dst=0)
    if rank == 0:
        np.save(out, allh)
    else:
        assert allh is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_hits_world2(tmp_path):
    out = str(tmp_path / "hits.npy")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    allh = np.load(out)
    assert len(allh) == 200
    assert np.array_equal(allh["score"], allh["a"] * 10 + allh["b"])  # A indices were shifted to the unsharded DB
    assert sorted(set(allh["a"].tolist())) == list(range(100))
