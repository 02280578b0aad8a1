"""GPU: DB-sharded searches give the single-GPU answer.

One-GPU part (always runs at round end): the two blocks of a split DB are processed one after the other on the same
device and merged with the same code the ranks use.  Two-GPU part (needs >= 2 devices, `gpurun --gpus 2`): two NCCL ranks,
one block each, triples all-gathered over NVLink, hits gathered on rank 0."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sets():
    from reseek_b200 import synth
    q = synth.make_chains(8, 150, seed=61, length_jitter=0.3)
    db = synth.make_chains(60, 160, seed=62, length_jitter=0.5)
    synth.plant_homologs(db, q, 0.4, seed=63)
    return q, db


def _single(rb, q, db, rsb_size):
    ctx = rb.Context(0, rb.MODE_FAST)
    Q = ctx.upload(q.lens, q.prof, q.mu, q.xyz, q.selfrev)
    T = ctx.upload(db.lens, db.prof, db.mu, db.xyz, db.selfrev)
    pf = ctx.prefilter(Q, T, rsb_size=rsb_size)
    res = ctx.postfilter(Q, T, pf, keep=rb.KEEP_HITS, want_paths=True)
    out = (pf.targets.copy(), pf.queries.copy(), pf.scores.copy(), res.hits.copy(), [res.path(k) for k in range(len(res.hits))])
    ctx.close()
    return out


@pytest.mark.parametrize("rsb_size", [0, 3])
def test_two_blocks_on_one_gpu_equal_unsharded(built_lib, rsb_size):
    import reseek_b200 as rb
    from reseek_b200.shard import partition_by_residues
    if rb.device_count() < 1:
        pytest.fail("no CUDA device")
    q, db = _sets()
    t_ref, q_ref, s_ref, hits_ref, paths_ref = _single(rb, q, db, rsb_size)
    assert len(t_ref) > 0 and len(hits_ref) > 0
    ctx = rb.Context(0, rb.MODE_FAST)
    Q = ctx.upload(q.lens, q.prof, q.mu, q.xyz, q.selfrev)
    parts = partition_by_residues(db.lens, 2)
    blocks, raws = [], []
    for lo, hi in parts:
        d = db.subset(range(lo, hi))
        T = ctx.upload(d.lens, d.prof, d.mu, d.xyz, d.selfrev)
        blocks.append(T)
        raws.append(ctx.prefilter(Q, T, rsb_size=rsb_size, raw_only=True))
    merged = rb.prefilter_bag(q.n, np.concatenate([r.targets + np.uint32(lo) for r, (lo, hi) in zip(raws, parts)]),
                              np.concatenate([r.queries for r in raws]), np.concatenate([r.scores for r in raws]), rsb_size)
    assert np.array_equal(merged.targets, t_ref) and np.array_equal(merged.queries, q_ref) and np.array_equal(merged.scores, s_ref)
    hits, paths = [], []
    for T, (lo, hi) in zip(blocks, parts):
        res = ctx.postfilter(Q, T, merged.select(lo, hi), keep=rb.KEEP_HITS, want_paths=True)
        h = res.hits.copy()
        h["b"] += np.uint32(lo)
        hits.append(h)
        paths += [res.path(k) for k in range(len(h))]
    hits = np.concatenate(hits)
    # KEEP_HITS lists come in the library's schedule order (per call: query-major), which differs between one call over
    # the whole DB and one call per block: compare in the canonical (target, query) order
    o, o_ref = _canon(hits), _canon(hits_ref)
    for f in ("a", "b", "score", "lo_a", "lo_b", "hi_a", "hi_b", "ids", "gaps", "lddt", "ts", "evalue", "mu_fwd", "mu_rev", "flags", "path_len"):
        assert np.array_equal(hits[f][o], hits_ref[f][o_ref]), f
    assert [paths[k] for k in o] == [paths_ref[k] for k in o_ref]
    ctx.close()


def _canon(h):
    return np.lexsort((h["a"], h["b"]))


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
    from reseek_b200.shard import partition_by_residues, search_fast_db_sharded, gather_hits
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    q, db = _sets()
    lo, hi = partition_by_residues(db.lens, world)[rank]
    d = db.subset(range(lo, hi))
    ctx = rb.Context(rank, rb.MODE_FAST)
    Q = ctx.upload(q.lens, q.prof, q.mu, q.xyz, q.selfrev)
    T = ctx.upload(d.lens, d.prof, d.mu, d.xyz, d.selfrev)
    merged, hits, res = search_fast_db_sharded(ctx, Q, T, lo, dist)
    # RunQuery-style sharding too: streamed side = this rank's DB block (slot A), queries replicated
    ctx.set_params(rb.params_preset(rb.MODE_SENSITIVE))
    r2 = ctx.search_cross(T, Q, keep=rb.KEEP_HITS, want_paths=False)
    h2 = gather_hits(r2.hits, lo, dist, dst=0, field="a")
    if rank == 0:
        np.savez(out, t=merged.targets, q=merged.queries, s=merged.scores, hits=hits, h2=h2)
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()


def test_two_ranks_nccl_equal_single_gpu(built_lib, tmp_path):
    import torch
    import torch.multiprocessing as mp
    import reseek_b200 as rb
    if rb.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    out = str(tmp_path / "sharded.npz")
    mp.spawn(_nccl_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    d = np.load(out)
    q, db = _sets()
    t_ref, q_ref, s_ref, hits_ref, _ = _single(rb, q, db, 0)
    assert np.array_equal(d["t"], t_ref) and np.array_equal(d["q"], q_ref) and np.array_equal(d["s"], s_ref)
    o, o_ref = _canon(d["hits"]), _canon(hits_ref)
    for f in ("a", "b", "score", "lo_a", "lo_b", "hi_a", "hi_b", "ts", "evalue", "path_len"):
        assert np.array_equal(d["hits"][f][o], hits_ref[f][o_ref]), f
    ctx = rb.Context(0, rb.MODE_SENSITIVE)
    Q = ctx.upload(q.lens, q.prof, q.mu, q.xyz, q.selfrev)
    T = ctx.upload(db.lens, db.prof, db.mu, db.xyz, db.selfrev)
    res2 = ctx.search_cross(T, Q, keep=rb.KEEP_HITS, want_paths=False)  # keep the Results alive: .hits is a view of its memory
    ref2 = res2.hits.copy()
    o, o_ref = np.lexsort((d["h2"]["b"], d["h2"]["a"])), np.lexsort((ref2["b"], ref2["a"]))
    assert len(ref2) > 0
    for f in ("a", "b", "score", "ts", "evalue"):
        assert np.array_equal(d["h2"][f][o], ref2[f][o_ref]), f
    ctx.close()
