"""GPU, BASELINE.json full size (config 5 at L=300: 100 queries x 100 000 chains = 1e7 pairs, 9e11 DP cells per pass):
parity through a random sample against the oracle and through size-independent properties of the whole result set."""
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NQ, NDB, L = 100, 100_000, 300


@pytest.fixture(scope="module")
def world(built_lib):
    import reseek_b200 as rb
    from reseek_b200 import synth
    if rb.device_count() < 1:
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    q = synth.make_chains(NQ, L, seed=20260122)
    db = synth.make_chains(NDB, L, seed=20260122 + 1000)
    synth.plant_homologs(db, q, 0.01, seed=20260122 + 7)
    ctx = rb.Context(0, rb.MODE_VERYSENSITIVE)
    Q = ctx.upload(q.lens, q.prof, q.mu, q.xyz, q.selfrev)
    D = ctx.upload(db.lens, db.prof, db.mu, db.xyz, db.selfrev)
    res = ctx.search_cross(D, Q, keep=rb.KEEP_ALL, want_paths=True)
    yield rb, ctx, q, db, Q, D, res
    ctx.close()


def _digest(h):
    """Order-sensitive checksum of the fields that define an alignment."""
    c = 0
    for f in ("a", "b", "score", "lo_a", "lo_b", "hi_a", "hi_b", "ids", "gaps", "lddt", "ts", "path_len"):
        c = zlib.crc32(np.ascontiguousarray(h[f]).view(np.uint8), c)
    return c


def test_full_size_sample_matches_oracle(world, port):
    from tests.util import assert_hit_matches_oracle, to_oracle_chains
    rb, ctx, q, db, Q, D, res = world
    assert len(res.hits) == NQ * NDB
    rng = np.random.default_rng(11)
    ks = rng.choice(len(res.hits), 250, replace=False)
    # make sure planted homologs (long paths) are in the sample: take the 50 best-scoring pairs as well
    ks = np.concatenate([ks, np.argsort(res.hits["score"])[-50:]])
    oq = to_oracle_chains(q)
    p = port(3)
    for k in ks.tolist():
        h = res.hits[k]
        a = int(h["a"])
        oa = to_oracle_chains(db.subset([a]))[0]
        r, rpath = p.align_pair(oa, oq[int(h["b"])])
        assert_hit_matches_oracle(h, res.path(k), r, rpath, ctx=f"pair {k}")
    assert int(res.hits["path_len"][ks[-1]]) > 100, "the best pair is a planted homolog with a long path"


def test_full_size_record_invariants(world):
    rb, ctx, q, db, Q, D, res = world
    h = res.hits
    assert np.array_equal(h["a"], np.repeat(np.arange(NDB, dtype=np.uint32), NQ)) and np.array_equal(h["b"], np.tile(np.arange(NQ, dtype=np.uint32), NDB))
    assert (h["score"] >= 0).all() and np.isfinite(h["score"]).all()
    has = h["path_len"] > 0
    assert has.mean() > 0.99, "-verysensitive: practically every pair has a positive local alignment"
    assert ((h["score"] > 0) == has).all()
    ev = (h["flags"] & rb.HIT_HAS_EVALUE) != 0
    assert ev.all()  # MinFwdScore = 0 in this mode: CalcEvalue runs even for the few pairs without a positive cell
    hh = h[has]
    assert (hh["ids"] + hh["gaps"] == hh["path_len"]).all()
    assert (hh["hi_a"] >= hh["lo_a"]).all() and (hh["hi_b"] >= hh["lo_b"]).all()
    assert (hh["hi_a"] < L).all() and (hh["hi_b"] < L).all()
    # M columns consume both chains, D only A, I only B (sw.cpp:8-77): spans follow from the path
    sub = np.nonzero(has)[0][:: max(1, int(has.sum()) // 20000)]
    for k in sub.tolist():
        p = res.path(k)
        m, d, i = p.count("M"), p.count("D"), p.count("I")
        r = h[k]
        assert m + d + i == int(r["path_len"]) and m == int(r["ids"]) and d + i == int(r["gaps"])
        assert int(r["hi_a"]) - int(r["lo_a"]) + 1 == m + d and int(r["hi_b"]) - int(r["lo_b"]) + 1 == m + i
        assert p[0] == "M" and p[-1] == "M"
    assert np.allclose(h["evalue"][has], h["pvalue"][has].astype(np.float64) * 8340, rtol=1e-6)


def test_full_size_sharding_and_orientation_invariance(world):
    """Checksum of checksums: two half-DB searches reproduce the whole (what a 2-GPU run computes), and explicit pair
    lists (rows = DB chain, the non-transposed kernels) reproduce the cross search (rows = query, transposed kernels)."""
    rb, ctx, q, db, Q, D, res = world
    whole = [_digest(res.hits[: NDB // 2 * NQ]), _digest(res.hits[NDB // 2 * NQ:])]
    parts = []
    for lo, hi in ((0, NDB // 2), (NDB // 2, NDB)):
        d = db.subset(range(lo, hi))
        Dk = ctx.upload(d.lens, d.prof, d.mu, d.xyz, d.selfrev)
        r = ctx.search_cross(Dk, Q, keep=rb.KEEP_ALL, want_paths=False)
        h = r.hits.copy()
        h["a"] += np.uint32(lo)
        parts.append(_digest(h))
        Dk.free()
    assert parts == whole
    rng = np.random.default_rng(5)
    ks = np.sort(rng.choice(NQ * NDB, 20000, replace=False))
    ia, ib = (ks // NQ).astype(np.uint32), (ks % NQ).astype(np.uint32)
    r = ctx.search_pairs(D, Q, ia, ib, keep=rb.KEEP_ALL, want_paths=True)
    assert _digest(r.hits) == _digest(res.hits[ks])
    for j in range(0, len(ks), 97):
        assert r.path(j) == res.path(int(ks[j]))


def test_full_size_sensitive_is_a_filtered_subset(world):
    """-sensitive on the same pairs: whatever passes the Mu filter carries the very same SW alignment."""
    rb, ctx, q, db, Q, D, res = world
    ctx.set_params(rb.params_preset(rb.MODE_SENSITIVE))
    rs = ctx.search_cross(D, Q, keep=rb.KEEP_ALL, want_paths=False)
    st = ctx.stats()
    ctx.set_params(rb.params_preset(rb.MODE_VERYSENSITIVE))
    hs, hv = rs.hits, res.hits
    rej = (hs["flags"] & rb.HIT_MU_REJECTED) != 0
    assert 0.80 < rej.mean() < 0.99 and int(rej.sum()) == st["mu_filter_rejected"]
    assert (hs["mu_score"][~rej] >= 12).all() and (hs["path_len"][rej] == 0).all()
    for f in ("score", "lo_a", "lo_b", "path_len"):
        assert np.array_equal(hs[f][~rej], hv[f][~rej]), f
    both = ~rej & ((hs["flags"] & rb.HIT_HAS_EVALUE) != 0)
    assert (hs["score"][~rej & ~both] < 7).all(), "CalcEvalue is skipped below MinFwdScore = 7 (dssaligner.cpp:861)"
    for f in ("hi_a", "hi_b", "ids", "gaps", "lddt", "ts", "evalue"):
        assert np.array_equal(hs[f][both], hv[f][both]), f


@pytest.mark.parametrize("length,ndb", [(100, 100_000), (800, 12_500)])
def test_full_size_other_lengths_sample_matches_oracle(built_lib, port, length, ndb):
    """Config 5 at L = 100 (half-warp wavefronts, 1e7 pairs) and L = 800 (three passes of 32 x 9 rows per chain; the DB is cut to
    12 500 chains = 8e11 cells to bound the test time - the kernels see the same shapes): random sample + best-scoring pairs against
    the oracle, and the record invariants over the whole result."""
    import reseek_b200 as rb
    from reseek_b200 import synth
    from tests.util import assert_hit_matches_oracle, to_oracle_chains
    q = synth.make_chains(NQ, length, seed=20260122 + length)
    db = synth.make_chains(ndb, length, seed=20260122 + 1000 + length)
    synth.plant_homologs(db, q, 0.01, seed=20260122 + 7 + length)
    ctx = rb.Context(0, rb.MODE_VERYSENSITIVE)
    Q = ctx.upload(q.lens, q.prof, q.mu, q.xyz, q.selfrev)
    D = ctx.upload(db.lens, db.prof, db.mu, db.xyz, db.selfrev)
    res = ctx.search_cross(D, Q, keep=rb.KEEP_ALL, want_paths=True)
    h = res.hits
    assert len(h) == NQ * ndb
    rng = np.random.default_rng(length)
    ks = np.concatenate([rng.choice(len(h), 120 if length == 100 else 40, replace=False), np.argsort(h["score"])[-20:]])
    oq = to_oracle_chains(q)
    p = port(3)
    for k in ks.tolist():
        oa = to_oracle_chains(db.subset([int(h[k]["a"])]))[0]
        r, rpath = p.align_pair(oa, oq[int(h[k]["b"])])
        assert_hit_matches_oracle(h[k], res.path(k), r, rpath, ctx=f"L={length} pair {k}")
    assert int(h["path_len"][ks[-1]]) > length // 3, "the best pair is a planted homolog with a long path"
    has = h["path_len"] > 0
    hh = h[has]
    assert has.mean() > 0.99 and (hh["ids"] + hh["gaps"] == hh["path_len"]).all()
    assert (hh["hi_a"] < length).all() and (hh["hi_b"] < length).all() and (hh["hi_a"] >= hh["lo_a"]).all()
    # the device-compacted path returns the same records
    r2 = ctx.search_cross_sharded(None, D, Q, 0, keep=rb.KEEP_HITS, want_paths=False)
    rep = (h["flags"] & rb.HIT_REPORTED) != 0
    assert len(r2.hits) == int(rep.sum()) and np.array_equal(r2.hits["score"].view(np.uint32), h["score"][rep].view(np.uint32))
    ctx.close()
