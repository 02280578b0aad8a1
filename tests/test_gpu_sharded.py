"""GPU: DB-sharded searches through the product's C ABI (rsk_search_cross_sharded / rsk_search_fast_db_sharded: device hit
sink + NCCL gather, device bag) give the single-GPU answer.

One-GPU part (always runs at round end): the sink path with one rank against the host pipeline; the blocks of a split DB
processed one after the other (incl. an empty block) and concatenated.  Two-GPU part (needs >= 2 devices, `gpurun --gpus
2`): two NCCL ranks in two processes, and two host threads of one process over ncclCommInitAll; the root's gathered result
must have the digest of the single-GPU search."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIELDS = ("a", "b", "score", "lo_a", "lo_b", "hi_a", "hi_b", "ids", "gaps", "lddt", "ts", "pvalue", "evalue", "qual",
          "mu_score", "mu_fwd", "mu_rev", "flags", "path_len")


def _sets(nq=8, ndb=60):
    from reseek_b200 import synth
    q = synth.make_chains(nq, 150, seed=61, length_jitter=0.3)
    db = synth.make_chains(ndb, 160, seed=62, length_jitter=0.5)
    synth.plant_homologs(db, q, 0.4, seed=63)
    return q, db


def _up(ctx, s):
    return ctx.upload(s.lens, s.prof, s.mu, s.xyz, s.selfrev)


def _assert_same_hits(res, ref, ordered=True):
    """Records and paths equal; unordered: compared in (a, b) order (explicit-pair schedules differ per block)."""
    h, r = res.hits, ref.hits
    assert len(h) == len(r), (len(h), len(r))
    o = np.arange(len(h)) if ordered else np.lexsort((h["b"], h["a"]))
    o_ref = np.arange(len(r)) if ordered else np.lexsort((r["b"], r["a"]))
    for f in FIELDS:
        assert np.array_equal(h[f][o].view(np.uint32), r[f][o_ref].view(np.uint32)), f
    for k, kr in zip(o.tolist(), o_ref.tolist()):
        assert res.path(k) == ref.path(kr)
    assert res.digest() == ref.digest()


@pytest.mark.parametrize("mode", ["verysensitive", "sensitive", "fast"])
@pytest.mark.parametrize("batch_pairs", [0, 64])
def test_sink_one_rank_equals_host_pipeline(built_lib, mode, batch_pairs, monkeypatch):
    """comm = NULL: the device-compacted path alone.  batch_pairs=64 cuts the search into many small batches, so the sink
    grows while batches are in flight."""
    import reseek_b200 as rb
    if batch_pairs:
        monkeypatch.setenv("RSK_BATCH_PAIRS", str(batch_pairs))
    q, db = _sets()
    ctx = rb.Context(0, {"verysensitive": rb.MODE_VERYSENSITIVE, "sensitive": rb.MODE_SENSITIVE, "fast": rb.MODE_FAST}[mode])
    Q, D = _up(ctx, q), _up(ctx, db)
    ref = ctx.search_cross(D, Q, keep=rb.KEEP_HITS, want_paths=True)
    assert len(ref.hits) > 0
    res = ctx.search_cross_sharded(None, D, Q, 0, keep=rb.KEEP_HITS, want_paths=True)
    _assert_same_hits(res, ref)
    st = ctx.stats()
    assert st["hits"] == len(ref.hits) and st["pairs"] == q.n * db.n
    # KEEP_ALL through the sink: one record per scheduled pair, in schedule order = (a, b) order of a cross search
    ref_all = ctx.search_cross(D, Q, keep=rb.KEEP_ALL, want_paths=True)
    res_all = ctx.search_cross_sharded(None, D, Q, 0, keep=rb.KEEP_ALL, want_paths=True)
    _assert_same_hits(res_all, ref_all)
    # records only
    res_np = ctx.search_cross_sharded(None, D, Q, 0, keep=rb.KEEP_HITS, want_paths=False)
    assert np.array_equal(res_np.hits["score"].view(np.uint32), ref.hits["score"].view(np.uint32)) and len(res_np.paths) == 0
    ctx.close()


@pytest.mark.parametrize("nblocks", [2, 5])
def test_blocks_on_one_gpu_equal_unsharded(built_lib, nblocks):
    """The ranks' code path block after block on one device: a_base offsets, an empty block, one-rank communicators."""
    import reseek_b200 as rb
    q, db = _sets()
    ctx = rb.Context(0, rb.MODE_SENSITIVE)
    comm = rb.Comm(ctx, 1, 0)
    Q, D = _up(ctx, q), _up(ctx, db)
    ref = ctx.search_cross(D, Q, keep=rb.KEEP_HITS, want_paths=True)
    parts = rb.partition_by_residues(db.lens, nblocks)
    parts = parts[:1] + [(parts[0][1], parts[0][1])] + parts[1:]  # plus an empty block
    hits, paths = [], []
    for lo, hi in parts:
        T = _up(ctx, db.subset(range(lo, hi))) if hi > lo else None
        res = ctx.search_cross_sharded(comm, T, Q, lo, keep=rb.KEEP_HITS, want_paths=True)
        hits.append(res.hits.copy())
        paths += [res.path(k) for k in range(len(res.hits))]
    hits = np.concatenate(hits)
    assert len(hits) == len(ref.hits) > 0
    for f in FIELDS:
        assert np.array_equal(hits[f].view(np.uint32), ref.hits[f].view(np.uint32)), f
    assert paths == [ref.path(k) for k in range(len(ref.hits))]
    assert comm.stats()["bytes_sent"] == 0
    comm.close()
    ctx.close()


@pytest.mark.parametrize("rsb_size", [0, 3])
def test_fast_db_sharded_one_rank_and_blocks(built_lib, rsb_size):
    """`-fast -db`: the sharded entry with one rank equals rsk_search_fast_db; the device bag over the concatenated raw
    triples of two blocks (what the all-gather delivers) equals the unsharded candidate list, cut-off ties included."""
    import reseek_b200 as rb
    q, db = _sets()
    ctx = rb.Context(0, rb.MODE_FAST)
    Q, T = _up(ctx, q), _up(ctx, db)
    pf = ctx.prefilter(Q, T, rsb_size=rsb_size)
    ref = ctx.search_fast_db(Q, T, rsb_size=rsb_size, keep=rb.KEEP_HITS, want_paths=True)
    assert len(pf.targets) > 0 and len(ref.hits) > 0
    res, cands = ctx.search_fast_db_sharded(None, Q, T, 0, rsb_size=rsb_size, keep=rb.KEEP_HITS, want_paths=True)
    assert np.array_equal(cands.targets, pf.targets) and np.array_equal(cands.queries, pf.queries) and np.array_equal(cands.scores, pf.scores)
    _assert_same_hits(res, ref, ordered=False)
    # host restatement of the bag on the raw triples (rsk_prefilter_bag) = device bag
    raw = ctx.prefilter(Q, T, rsb_size=rsb_size, raw_only=True)
    hb = rb.prefilter_bag(q.n, raw.targets, raw.queries, raw.scores, rsb_size)
    assert np.array_equal(hb.targets, pf.targets) and np.array_equal(hb.queries, pf.queries) and np.array_equal(hb.scores, pf.scores)
    # blocks: raw triples per block, concatenated in block order, equal the unsharded stream
    parts = rb.partition_by_residues(db.lens, 3)
    ts, qs, ss = [], [], []
    for lo, hi in parts:
        Tb = _up(ctx, db.subset(range(lo, hi)))
        r = ctx.prefilter(Q, Tb, rsb_size=rsb_size, raw_only=True)
        ts.append(r.targets + np.uint32(lo)); qs.append(r.queries); ss.append(r.scores)
    assert np.array_equal(np.concatenate(ts), raw.targets) and np.array_equal(np.concatenate(qs), raw.queries)
    assert np.array_equal(np.concatenate(ss), raw.scores)
    ctx.close()


@pytest.mark.parametrize("mode", ["sensitive", "fast"])
def test_self_search_sink_one_rank(built_lib, mode, monkeypatch):
    """RunSelf through the sharded entry with one rank (rows in several chunks, so the sink is appended to) equals rsk_search_self."""
    import reseek_b200 as rb
    monkeypatch.setenv("RSK_SELF_CHUNK_PAIRS", "500")
    q, db = _sets(8, 70)
    ctx = rb.Context(0, rb.MODE_SENSITIVE if mode == "sensitive" else rb.MODE_FAST)
    S = _up(ctx, db)
    ref = ctx.search_self(S, keep=rb.KEEP_HITS, want_paths=True)
    res = ctx.search_self_sharded(None, S, keep=rb.KEEP_HITS, want_paths=True)
    assert len(ref.hits) >= db.n
    _assert_same_hits(res, ref, ordered=False)
    h = res.hits
    assert np.all((h["a"][1:] > h["a"][:-1]) | ((h["a"][1:] == h["a"][:-1]) & (h["b"][1:] > h["b"][:-1])))  # (a, b) ascending
    ctx.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    import reseek_b200 as rb
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    q, db = _sets()
    lo, hi = rb.partition_by_residues(db.lens, world)[rank]
    ctx = rb.Context(rank, rb.MODE_FAST)
    comm = rb.Comm.from_torch_dist(ctx, dist)
    Q, T = _up(ctx, q), _up(ctx, db.subset(range(lo, hi)))
    res, cands = ctx.search_fast_db_sharded(comm, Q, T, lo, keep=rb.KEEP_HITS, want_paths=True)
    ctx.set_params(rb.params_preset(rb.MODE_SENSITIVE))
    r2 = ctx.search_cross_sharded(comm, T, Q, lo, keep=rb.KEEP_HITS, want_paths=True)
    if rank == 0:
        np.savez(out, t=cands.targets, q=cands.queries, s=cands.scores, hits=res.hits, h2=r2.hits,
                 d1=np.uint64(res.digest()), d2=np.uint64(r2.digest()), moved=np.uint64(comm.stats()["bytes_recv"]))
    else:
        assert res is None and r2 is None
    dist.barrier()
    comm.close()
    dist.destroy_process_group()
    ctx.close()


def test_two_ranks_nccl_equal_single_gpu(built_lib, tmp_path):
    import torch.multiprocessing as mp
    import reseek_b200 as rb
    if rb.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2); bench.py --gpus N asserts the same digests on the driver's scaling run")
    out = str(tmp_path / "sharded.npz")
    mp.spawn(_nccl_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    d = np.load(out)
    q, db = _sets()
    ctx = rb.Context(0, rb.MODE_FAST)
    Q, T = _up(ctx, q), _up(ctx, db)
    pf = ctx.prefilter(Q, T)
    ref = ctx.search_fast_db(Q, T, keep=rb.KEEP_HITS, want_paths=True)
    assert np.array_equal(d["t"], pf.targets) and np.array_equal(d["q"], pf.queries) and np.array_equal(d["s"], pf.scores)
    assert int(d["d1"]) == ref.digest() and len(d["hits"]) == len(ref.hits) > 0
    ctx.set_params(rb.params_preset(rb.MODE_SENSITIVE))
    ref2 = ctx.search_cross(T, Q, keep=rb.KEEP_HITS, want_paths=True)
    assert int(d["d2"]) == ref2.digest() and len(d["h2"]) == len(ref2.hits) > 0
    for f in FIELDS:
        assert np.array_equal(d["h2"][f].view(np.uint32), ref2.hits[f].view(np.uint32)), f
    assert int(d["moved"]) > 0
    ctx.close()


def test_two_gpus_one_process_comm_init_all(built_lib):
    """DBSearcher's layout: one process, one host thread per GPU, communicators from ncclCommInitAll."""
    import threading
    import reseek_b200 as rb
    if rb.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    q, db = _sets()
    ctxs = [rb.Context(k, rb.MODE_SENSITIVE) for k in range(2)]
    comms = rb.Comm.create_all(ctxs)
    parts = rb.partition_by_residues(db.lens, 2)
    out = [None, None]

    def work(r):
        lo, hi = parts[r]
        Q, T = _up(ctxs[r], q), _up(ctxs[r], db.subset(range(lo, hi)))
        out[r] = ctxs[r].search_cross_sharded(comms[r], T, Q, lo, keep=rb.KEEP_HITS, want_paths=True)

    th = [threading.Thread(target=work, args=(r,)) for r in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert out[1] is None and out[0] is not None
    Q, T = _up(ctxs[0], q), _up(ctxs[0], db)
    ref = ctxs[0].search_cross(T, Q, keep=rb.KEEP_HITS, want_paths=True)
    _assert_same_hits(out[0], ref)
    # RunSelf with the rows of the pair triangle interleaved over the two GPUs
    sets = [_up(ctxs[r], db) for r in range(2)]

    def work_self(r):
        out[r] = ctxs[r].search_self_sharded(comms[r], sets[r], keep=rb.KEEP_HITS, want_paths=True)

    th = [threading.Thread(target=work_self, args=(r,)) for r in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert out[1] is None
    _assert_same_hits(out[0], ctxs[0].search_self(sets[0], keep=rb.KEEP_HITS, want_paths=True), ordered=False)
    for c in comms:
        c.close()
    for c in ctxs:
        c.close()
