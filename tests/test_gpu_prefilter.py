"""GPU parity of the `-search Q -db DB -fast` pipeline through the C ABI: candidate lists against the
reference binary's own prefilter TSV (tests/golden/golden_prefilter.npz) and against the CPU oracle on synthetic chains."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb(built_lib):
    import reseek_b200
    if reseek_b200.device_count() < 1:
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    return reseek_b200


def _mu_only_set(ctx, mus):
    """Chain set that carries Mu letters only (the prefilter reads nothing else)."""
    lens = np.array([len(m) for m in mus], np.uint32)
    tot = int(lens.sum())
    return ctx.upload(lens, np.zeros((8, tot), np.uint8), np.concatenate(mus).astype(np.uint8), np.zeros((3, tot), np.float32), None)


@pytest.mark.parametrize("name,kw", [("idxq", {}), ("idxt", {"index_mode": 2}), ("rsb5", {"rsb_size": 5}), ("idxq", {"nofuse": 1})])
def test_prefilter_matches_reference_candidate_tsv(rb, name, kw, monkeypatch):
    """q10.bca vs q100.bca: the reference binary's candidate TSV at -threads 1 (-idxq default, -idxt, -rsb_size 5; the default
    once more through the global-memory form of K7/K8)."""
    from tests.test_oracle_golden import _prefilter_fixture
    kw = dict(kw)
    if kw.pop("nofuse", 0):
        monkeypatch.setenv("RSK_PF_NOFUSE", "1")
    g, mq, mt = _prefilter_fixture()
    ctx = rb.Context(0, rb.MODE_FAST)
    Q, T = _mu_only_set(ctx, mq), _mu_only_set(ctx, mt)
    r = ctx.prefilter(Q, T, **kw)
    got = list(zip(r.targets.tolist(), r.queries.tolist()))
    want = list(zip(g[f"{name}_t"].tolist(), g[f"{name}_q"].tolist()))
    assert got == want and len(want) >= 50
    # the TSV text itself (rankedscoresbag.cpp:219-232)
    lines = r.tsv().splitlines()
    assert lines[0] == f"prefilter\t{len(set(g[f'{name}_t'].tolist()))}"
    flat = [(int(f[0]), int(x)) for f in (ln.split("\t") for ln in lines[1:]) for x in f[2:]]
    assert flat == want
    assert all(int(ln.split("\t")[1]) == len(ln.split("\t")) - 2 for ln in lines[1:])
    # without the K/L exchange of the query letters the reference's list is NOT reproduced
    if name == "idxq":
        r2 = ctx.prefilter(Q, T, kl_swap=False)
        assert list(zip(r2.targets.tolist(), r2.queries.tolist())) != want
    ctx.close()


@pytest.mark.parametrize("path", ["fused", "fused_q16", "fused_q1", "fused_nostage", "fused_nostage_q16", "global"])
@pytest.mark.parametrize("nq,index_mode", [(7, 0), (7, 2), (120, 0)])
def test_prefilter_scores_match_oracle_on_synthetic(rb, port, nq, index_mode, path, monkeypatch):
    """Planted homologs + random chains, ragged lengths (incl. chains shorter than one 7-window): every (target, query)
    two-hit diagonal score equals the oracle's brute-force restatement; > 100 queries switches the index side.  Both forms of
    K7/K8: hits kept in shared-memory bitmaps (the default for these sizes; also with a tiny two-hit queue, so that the overflow
    rounds run) and the global-memory path (sorted keys per target)."""
    from reseek_b200 import synth
    if path == "global":
        monkeypatch.setenv("RSK_PF_NOFUSE", "1")
    elif path.startswith("fused_q"):  # a two-hit queue of 16 / 1 entries: the overflow rounds over the bitmap do the work
        monkeypatch.setenv("RSK_PF_QUEUE", path[7:])
    elif path.startswith("fused_nostage"):  # letters read from global memory (the form for query blocks too large to stage)
        monkeypatch.setenv("RSK_PF_NOSTAGE", "1")
        if path.endswith("q16"):
            monkeypatch.setenv("RSK_PF_QUEUE", "16")
    q = synth.make_chains(nq, [5, 7, 60, 150, 300, 420, 33][:min(nq, 7)] + [90] * max(0, nq - 7), seed=501)
    t = synth.make_chains(60, 140, seed=502, length_jitter=0.7)
    synth.plant_homologs(t, q, 0.5, seed=503, sub=0.25, indel=0.03)
    mq = [q.chain(i)[1] for i in range(q.n)]
    mt = [t.chain(i)[1] for i in range(t.n)]
    qhood = index_mode == 1 or (index_mode == 0 and nq <= 100)
    want, raw = port(1).prefilter(mq, mt, query_neighborhood=qhood)
    ctx = rb.Context(0, rb.MODE_FAST)
    Q = ctx.upload(q.lens, q.prof, q.mu, q.xyz, q.selfrev)
    T = ctx.upload(t.lens, t.prof, t.mu, t.xyz, t.selfrev)
    r = ctx.prefilter(Q, T, index_mode=index_mode)
    assert r.raw_count == len(raw) >= 20
    got = {(int(a), int(b)): int(s) for a, b, s in zip(r.targets, r.queries, r.scores)}
    assert got == raw  # B = 1500 keeps everything here
    assert r.as_dict() == want
    ctx.close()


def test_prefilter_bag_truncation_matches_oracle(rb, port):
    """A bag smaller than the candidate count: lazy truncation at 2B and the unstable cut-off ties of QuickSortOrderDesc."""
    from reseek_b200 import synth
    q = synth.make_chains(3, 120, seed=511)
    t = synth.make_chains(150, 100, seed=512, length_jitter=0.3)
    synth.plant_homologs(t, q, 0.9, seed=513, sub=0.3)
    mq = [q.chain(i)[1] for i in range(q.n)]
    mt = [t.chain(i)[1] for i in range(t.n)]
    ctx = rb.Context(0, rb.MODE_FAST)
    Q = ctx.upload(q.lens, q.prof, q.mu, q.xyz, q.selfrev)
    T = ctx.upload(t.lens, t.prof, t.mu, t.xyz, t.selfrev)
    for B in (4, 9):
        want, raw = port(1).prefilter(mq, mt, rsb_size=B)
        r = ctx.prefilter(Q, T, rsb_size=B)
        assert r.raw_count == len(raw) > 3 * 2 * B, "the test needs more candidates than 2B per query"
        assert r.as_dict() == want
    ctx.close()


def test_search_fast_db_matches_oracle(rb, port):
    """Both stages: prefilter candidates -> AlignBags under the sensitive preset (search.cpp:106-108), A = query."""
    from reseek_b200 import synth
    from tests.util import assert_hit_matches_oracle, to_oracle_chains
    q = synth.make_chains(6, 130, seed=521, length_jitter=0.4)
    t = synth.make_chains(80, 150, seed=522, length_jitter=0.5)
    synth.plant_homologs(t, q, 0.4, seed=523, sub=0.2, indel=0.02)
    mq = [q.chain(i)[1] for i in range(q.n)]
    mt = [t.chain(i)[1] for i in range(t.n)]
    want, _ = port(1).prefilter(mq, mt)
    ctx = rb.Context(0, rb.MODE_FAST)
    Q = ctx.upload(q.lens, q.prof, q.mu, q.xyz, q.selfrev)
    T = ctx.upload(t.lens, t.prof, t.mu, t.xyz, t.selfrev)
    res = ctx.search_fast_db(Q, T, keep=rb.KEEP_ALL)
    pairs = [(tt, qq) for tt in sorted(want) for qq in want[tt]]
    assert [(int(h["b"]), int(h["a"])) for h in res.hits] == pairs and len(pairs) >= 10
    oq, ot = to_oracle_chains(q), to_oracle_chains(t)
    p = port(2)  # DM_AlwaysSensitive
    nrep = 0
    for k, h in enumerate(res.hits):
        r, rpath = p.align_pair(oq[int(h["a"])], ot[int(h["b"])])
        assert_hit_matches_oracle(h, res.path(k), r, rpath, ctx=f"cand {k}")
        nrep += bool(int(h["flags"]) & rb.HIT_REPORTED)
    hits = ctx.search_fast_db(Q, T, keep=rb.KEEP_HITS)
    assert len(hits.hits) == nrep > 0
    assert abs(ctx.params.omega - 22.0) < 1e-6, "the context's own (fast) parameters are restored after the post-filter"
    ctx.close()


def test_prefilter_edge_cases(rb):
    from reseek_b200 import synth
    ctx = rb.Context(0, rb.MODE_FAST)
    tiny = synth.make_chains(3, [1, 6, 3], seed=531)          # no chain has a 7-window
    t = synth.make_chains(5, 50, seed=532)
    A, T = ctx.upload(tiny.lens, tiny.prof, tiny.mu, tiny.xyz, tiny.selfrev), ctx.upload(t.lens, t.prof, t.mu, t.xyz, t.selfrev)
    assert len(ctx.prefilter(A, T)) == 0 and len(ctx.prefilter(T, A)) == 0
    assert ctx.prefilter(A, T).tsv() == "prefilter\t0\n"
    assert len(ctx.search_fast_db(A, T).hits) == 0
    # identical sets: every chain finds itself with the maximal diagonal score
    r = ctx.prefilter(T, T, kl_swap=False)
    d = {(int(a), int(b)): int(s) for a, b, s in zip(r.targets, r.queries, r.scores)}
    assert all((i, i) in d and d[(i, i)] == max(v for (tt, _), v in d.items() if tt == i) for i in range(t.n))
    nomu = ctx.upload(t.lens, t.prof, None, t.xyz, t.selfrev)
    with pytest.raises(rb.ReseekB200Error):
        ctx.prefilter(nomu, T)
    ctx.close()


@pytest.mark.parametrize("B,span", [(3, 5), (50, 7), (100, 40), (1500, 12), (1500, 300), (400, 2)])
def test_device_bag_equals_host_bag_on_tied_streams(built_lib, B, span):
    """RankedScoresBag replayed on the device (warp-parallel exact Hoare partitions + per-lane small ranges) against the host
    restatement (rsk_prefilter_bag, the reference's AddScore / TruncateVecs / QuickSortOrderDesc, rankedscoresbag.cpp:5-51,
    sort.h:71-108): streams long enough for many truncations, scores drawn from `span` values so that the cut-off is always
    inside a run of ties, plus rising and falling trends (the admission threshold then moves a lot / not at all)."""
    import reseek_b200 as rb
    rng = np.random.default_rng(1000 * B + span)
    nq = 7
    n = 60000
    t = np.sort(rng.integers(0, 50000, size=n)).astype(np.uint32)
    q = rng.integers(0, nq, size=n).astype(np.uint32)
    s = rng.integers(1, 1 + span, size=n).astype(np.int64)
    trend = np.linspace(0, 3 * span, n).astype(np.int64)
    s = np.where(q % 3 == 0, s + trend, np.where(q % 3 == 1, s + trend[::-1], s)).astype(np.uint16)
    # (target, query) pairs are unique in a real stream
    key = t.astype(np.int64) * nq + q
    _, first = np.unique(key, return_index=True)
    first.sort()
    t, q, s = t[first], q[first], s[first]
    ctx = rb.Context(0, rb.MODE_FAST)
    dev = ctx.prefilter_bag_device(nq, t, q, s, B)
    host = rb.prefilter_bag(nq, t, q, s, B)
    assert len(host.targets) > 0
    assert np.array_equal(dev.targets, host.targets) and np.array_equal(dev.queries, host.queries) and np.array_equal(dev.scores, host.scores)
    ctx.close()
