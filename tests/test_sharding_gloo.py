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


def test_cabi_partition_equals_python_helper(built_lib):
    """rsk_partition_by_residues (what DBSearcher -gpus N and bench.py cut the DB with; exact 128-bit compare) gives the blocks of
    shard.partition_by_residues, including more ranks than chains, a dominating chain, and residue totals beyond 2^32."""
    from reseek_b200 import lib
    rng = np.random.default_rng(3)
    cases = [rng.integers(30, 1400, size=5000), np.array([5, 5]), np.array([7]), np.array([100000, 1, 1, 1, 1]),
             np.full(9000, 600000, np.int64)]  # 5.4e9 residues
    for lens in cases:
        for world in (1, 2, 3, 8):
            want = partition_by_residues(lens, world)
            got = lib.partition_by_residues(lens, world)
            assert got == want, (len(lens), world)


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


# ---- `-fast -db` on a sharded DB: the merged RankedScoresBag (the one real exchange step) ----
def _triples(seed, nq=6, nt=400, density=0.7):
    """Synthetic (target, query, score) stream in target order with many score ties around the bag cut-off."""
    rng = np.random.default_rng(seed)
    t, q, s = [], [], []
    for ti in range(nt):
        for qi in range(nq):
            if rng.random() < density:
                t.append(ti); q.append(qi); s.append(int(rng.integers(40, 48)))
    return np.array(t, np.uint32), np.array(q, np.uint32), np.array(s, np.uint16)


def _oracle_bag(port, nq, t, q, s, B):
    import ctypes as C
    L = port(1).lib
    L.orc_rsb_new.restype = C.c_void_p
    L.orc_rsb_targets.restype = C.POINTER(C.c_uint32)
    rsb = C.c_void_p(L.orc_rsb_new(nq, B))
    for k in range(len(t)):
        L.orc_rsb_add(rsb, int(q[k]), int(t[k]), int(s[k]))
    L.orc_rsb_finish(rsb)
    out = {}
    for qi in range(nq):
        tg = L.orc_rsb_targets(rsb, qi)
        for k in range(L.orc_rsb_count(rsb, qi)):
            out.setdefault(int(tg[k]), []).append(qi)
    L.orc_rsb_free(rsb)
    return out


@pytest.mark.parametrize("B", [5, 40, 1500])
def test_prefilter_bag_matches_oracle(built_lib, port, B):
    """rsk_prefilter_bag (host part of the product) against the oracle's RankedScoresBag restatement, ties included."""
    import reseek_b200 as rb
    t, q, s = _triples(3)
    r = rb.prefilter_bag(6, t, q, s, rsb_size=B)
    assert r.as_dict() == {k: sorted(v) for k, v in _oracle_bag(port, 6, t, q, s, B).items()}
    assert list(r.targets) == sorted(r.targets)
    sel = r.select(100, 250)
    assert sel.as_dict() == {k - 100: v for k, v in r.as_dict().items() if 100 <= k < 250}


def _bag_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from reseek_b200.shard import merge_prefilter_triples
    t, q, s = _triples(11)
    lens = np.full(400, 10)
    lo, hi = partition_by_residues(lens, world)[rank]
    m = (t >= lo) & (t < hi)

    class Raw:  # what Context.prefilter(raw_only=True) returns on this rank: local target indices
        targets, queries, scores = t[m] - lo, q[m], s[m]
    merged = merge_prefilter_triples(6, Raw, lo, dist, rsb_size=25)
    np.savez(out + f".{rank}.npz", t=merged.targets, q=merged.queries, s=merged.scores)
    dist.barrier()
    dist.destroy_process_group()


def test_merged_bag_world2_equals_unsharded(built_lib, tmp_path):
    import reseek_b200 as rb
    out = str(tmp_path / "bag")
    mp.spawn(_bag_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    t, q, s = _triples(11)
    ref = rb.prefilter_bag(6, t, q, s, rsb_size=25)
    for rank in (0, 1):
        d = np.load(out + f".{rank}.npz")
        assert np.array_equal(d["t"], ref.targets) and np.array_equal(d["q"], ref.queries) and np.array_equal(d["s"], ref.scores)
