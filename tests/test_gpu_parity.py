"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb(built_lib):
    import reseek_b200
    if reseek_b200.device_count() < 1:
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    return reseek_b200


def _check_all(rb, port, res, oa, ob):
    from tests.util import assert_hit_matches_oracle
    assert len(res.hits) > 0
    for k, h in enumerate(res.hits):
        r, rpath = port.align_pair(oa[int(h["a"])], ob[int(h["b"])])
        assert_hit_matches_oracle(h, res.path(k), r, rpath, ctx=f"pair {k} (a={h['a']} b={h['b']})")


@pytest.mark.parametrize("la,lb,seed", [(40, 50, 1), (130, 90, 2), (300, 300, 3), (257, 33, 4), (520, 140, 5)])
def test_cross_verysensitive_matches_oracle(rb, port, la, lb, seed):
    from reseek_b200 import synth
    from tests.util import to_oracle_chains
    a = synth.make_chains(5, la, seed=100 + seed, length_jitter=0.25)
    b = synth.make_chains(19, lb, seed=200 + seed, length_jitter=0.25)
    synth.plant_homologs(a, b, 0.6, seed=300 + seed)
    ctx = rb.Context(0, rb.MODE_VERYSENSITIVE)
    A = ctx.upload(a.lens, a.prof, a.mu, a.xyz, a.selfrev)
    B = ctx.upload(b.lens, b.prof, b.mu, b.xyz, b.selfrev)
    res = ctx.search_cross(A, B, keep=rb.KEEP_ALL, want_paths=True)
    assert len(res.hits) == a.n * b.n
    _check_all(rb, port(3), res, to_oracle_chains(a), to_oracle_chains(b))
    ctx.close()


def test_explicit_pairs_and_self(rb, port):
    from reseek_b200 import synth
    from tests.util import to_oracle_chains
    s = synth.make_chains(12, 80, seed=7, length_jitter=0.5)
    ctx = rb.Context(0, rb.MODE_VERYSENSITIVE)
    S = ctx.upload(s.lens, s.prof, s.mu, s.xyz, s.selfrev)
    oc = to_oracle_chains(s)
    rng = np.random.default_rng(5)
    ia = rng.integers(0, s.n, 40).astype(np.uint32)
    ib = rng.integers(0, s.n, 40).astype(np.uint32)
    res = ctx.search_pairs(S, S, ia, ib, keep=rb.KEEP_ALL)
    assert np.array_equal(res.hits["a"], ia) and np.array_equal(res.hits["b"], ib)
    _check_all(rb, port(3), res, oc, oc)
    res = ctx.search_self(S, keep=rb.KEEP_ALL)
    assert len(res.hits) == s.n * (s.n + 1) // 2
    _check_all(rb, port(3), res, oc, oc)
    ctx.close()


@pytest.mark.parametrize("mode", [2, 1])
def test_cross_with_mu_filter_matches_oracle(rb, port, mode):
    """-sensitive / -fast: Mu int8 SW filter (fwd, reversed, 777/255 saturation rules) -> survivors -> SW."""
    from reseek_b200 import synth
    from tests.util import to_oracle_chains
    a = synth.make_chains(40, 150, seed=900 + mode, length_jitter=0.5)
    b = synth.make_chains(23, 170, seed=950 + mode, length_jitter=0.5)
    synth.plant_homologs(a, b, 0.5, seed=77, sub=0.15)   # close homologs: int8 saturation on the forward pass
    synth.plant_homologs(a, b, 0.3, seed=78, sub=0.45)   # remote ones: around the omega thresholds
    ctx = rb.Context(0, mode)
    A = ctx.upload(a.lens, a.prof, a.mu, a.xyz, a.selfrev)
    B = ctx.upload(b.lens, b.prof, b.mu, b.xyz, b.selfrev)
    res = ctx.search_cross(A, B, keep=rb.KEEP_ALL, want_paths=True)
    st = ctx.stats()
    assert len(res.hits) == a.n * b.n
    _check_all(rb, port(mode), res, to_oracle_chains(a), to_oracle_chains(b))
    nrej = int(np.sum((res.hits["flags"] & rb.HIT_MU_REJECTED) != 0))
    assert 0 < nrej < len(res.hits), "the test must exercise both outcomes of the filter"
    assert st["mu_filter_in"] == len(res.hits) and st["mu_filter_rejected"] == nrej
    assert st["mu_saturated"] == int(np.sum(res.hits["mu_fwd"] == 777)) > 0
    assert st["sw_pairs"] == len(res.hits) - nrej
    # the same pairs as an explicit list (PostMuFilter-style) must give identical records
    ia = np.repeat(np.arange(a.n, dtype=np.uint32), b.n)
    ib = np.tile(np.arange(b.n, dtype=np.uint32), a.n)
    res2 = ctx.search_pairs(A, B, ia, ib, keep=rb.KEEP_ALL, want_paths=True)
    for k in ("score", "lo_a", "lo_b", "hi_a", "hi_b", "ids", "gaps", "ts", "evalue", "mu_fwd", "mu_rev", "flags", "path_len"):
        assert np.array_equal(res.hits[k], res2.hits[k]), k
    # KEEP_HITS returns exactly the reported subset
    res3 = ctx.search_cross(A, B, keep=rb.KEEP_HITS, want_paths=True)
    rep = res.hits[(res.hits["flags"] & rb.HIT_REPORTED) != 0]
    assert len(res3.hits) == len(rep) and np.array_equal(np.sort(res3.hits["score"]), np.sort(rep["score"]))
    ctx.close()


@pytest.mark.parametrize("mode", [3, 2, 1])
def test_golden_reference_fixtures_on_gpu(rb, mode):
    """Real chains (test_data/q100.bca, scop40.bca) with the reference's own answers (tests/golden)."""
    from tests.golden_util import load_chains, load_pairs
    from tests.util import bits
    g = load_pairs(mode)
    chains = load_chains(g["selfrev"])
    ctx = rb.Context(0, mode)
    if mode == 3:  # -verysensitive never loads Mu letters (dbsearcher.cpp:251-252)
        for c in chains:
            c.mu = None
    S = ctx.upload_chains(chains)
    res = ctx.search_cross(S, S, keep=rb.KEEP_ALL, want_paths=True)
    n = nmkf = nmkf_path = 0
    for k, h in enumerate(res.hits):
        assert (int(h["a"]), int(h["b"])) == (int(g["a"][k]), int(g["b"][k]))
        mkf = bool(g["mkf"][k])
        assert bool(int(h["flags"]) & rb.HIT_MKF) == mkf, f"pair {k}: DoMKF routing"
        if mkf:
            nmkf += 1
            nmkf_path += bool(g["path_list"][k])
            assert (int(h["mu_fwd"]), int(h["mu_rev"])) == (int(g["best_hsp"][k]), int(g["best_chain"][k])), f"pair {k} MKF HSP/chain scores"
            if not g["path_list"][k]:
                # no alignment: the reference leaves Hi at Lo + 0 - 1 (wrapped) in one branch; unobservable
                assert res.path(k) == "" and float(h["score"]) == 0.0 and float(h["evalue"]) > 1e38
                continue
        n += 1
        assert res.path(k) == g["path_list"][k], f"pair {k} path"
        assert bits(h["score"]) == bits(g["score"][k]), f"pair {k} score"
        assert (int(h["hi_a"]), int(h["hi_b"]), int(h["ids"]), int(h["gaps"])) == (int(g["hi_a"][k]), int(g["hi_b"][k]), int(g["ids"][k]), int(g["gaps"][k]))
        assert bits(h["ts"]) == bits(g["ts"][k]), f"pair {k} ts"
        assert bits(h["evalue"]) == bits(g["evalue"][k]) and bits(h["pvalue"]) == bits(g["pvalue"][k]), f"pair {k} E/P"
        if res.path(k):
            assert (int(h["lo_a"]), int(h["lo_b"])) == (int(g["lo_a"][k]), int(g["lo_b"][k]))
        if g["evalue"][k] < 1e38:
            assert bits(h["lddt"]) == bits(g["lddt"][k]) and bits(h["qual"]) == bits(g["qual"][k])
    assert n > 150
    if mode != 3:
        assert nmkf > 150 and nmkf_path >= 20 and ctx.stats()["mkf_pairs"] == nmkf
    ctx.close()


def test_smoke_entry(rb):
    import __graft_entry__ as g
    g.smoke()


def test_both_orientations_and_all_row_classes(rb, port):
    """The scheduler puts the smaller set on the kernel's rows (transposed DP) and picks a kernel class by row
    length: cover rows = A and rows = B for lengths that hit every R = 1..8 and multi-pass tables, with exact
    score ties (duplicated chains) so that the first-maximum rule is exercised in both orientations."""
    from reseek_b200 import synth
    from tests.util import to_oracle_chains
    lens = [20, 33, 70, 100, 129, 161, 193, 225, 250, 300, 420, 530]
    small = synth.make_chains(len(lens), lens, seed=4242)
    big = synth.make_chains(40, 140, seed=4343, length_jitter=0.6)
    synth.plant_homologs(big, small, 0.7, seed=4444, sub=0.2)
    # internal repeats -> equal best scores at different cells
    for i in range(0, big.n, 5):
        s, e = int(big.off[i]), int(big.off[i + 1])
        h = (e - s) // 2
        big.prof[:, s + h:s + 2 * h] = big.prof[:, s:s + h]
        big.xyz[:, s + h:s + 2 * h] = big.xyz[:, s:s + h]
    ctx = rb.Context(0, rb.MODE_VERYSENSITIVE)
    S = ctx.upload(small.lens, small.prof, small.mu, small.xyz, small.selfrev)
    Bg = ctx.upload(big.lens, big.prof, big.mu, big.xyz, big.selfrev)
    os_, ob = to_oracle_chains(small), to_oracle_chains(big)
    res = ctx.search_cross(S, Bg, keep=rb.KEEP_ALL)     # A = small (12 chains), B = big (40): rows = A
    _check_all(rb, port(3), res, os_, ob)
    res = ctx.search_cross(Bg, S, keep=rb.KEEP_ALL)     # A = big, B = small: rows = B (transposed kernel)
    _check_all(rb, port(3), res, ob, os_)
    res = ctx.search_self(Bg, keep=rb.KEEP_ALL)
    _check_all(rb, port(3), res, ob, ob)
    ctx.close()


@pytest.mark.parametrize("mode", [2, 1])
def test_long_chain_path_matches_oracle(rb, port, mode):
    """Chains >= MKFL (k-mer seeds -> chain -> banded x-drop) mixed with short ones, cross and explicit."""
    from reseek_b200 import synth
    from tests.util import to_oracle_chains
    a = synth.make_chains(7, [650, 120, 700, 820, 610, 300, 640], seed=61)
    b = synth.make_chains(9, [300, 640, 720, 150, 660, 1000, 90, 605, 2], seed=62)
    synth.plant_homologs(b, a, 0.9, seed=63, sub=0.25, indel=0.03)
    ctx = rb.Context(0, mode)
    A = ctx.upload(a.lens, a.prof, a.mu, a.xyz, a.selfrev)
    B = ctx.upload(b.lens, b.prof, b.mu, b.xyz, b.selfrev)
    oa, ob = to_oracle_chains(a), to_oracle_chains(b)
    res = ctx.search_cross(A, B, keep=rb.KEEP_ALL, want_paths=True)
    p = port(mode)
    nmkf = npath = 0
    from tests.util import bits
    for k, h in enumerate(res.hits):
        r, rpath = p.align_pair(oa[int(h["a"])], ob[int(h["b"])])
        mkf = bool(int(h["flags"]) & rb.HIT_MKF)
        if not mkf:
            from tests.util import assert_hit_matches_oracle
            assert_hit_matches_oracle(h, res.path(k), r, rpath, ctx=f"pair {k}")
            continue
        nmkf += 1
        assert res.path(k) == rpath, f"pair {k} ({oa[int(h['a'])].L} x {ob[int(h['b'])].L}) path"
        assert bits(h["score"]) == bits(r.score), f"pair {k} score {h['score']} vs {r.score}"
        if rpath:
            npath += 1
            assert (int(h["lo_a"]), int(h["lo_b"]), int(h["hi_a"]), int(h["hi_b"])) == (r.lo_a, r.lo_b, r.hi_a, r.hi_b)
            assert bits(h["ts"]) == bits(r.ts) and bits(h["lddt"]) == bits(r.lddt), f"pair {k} ts/lddt"
    assert nmkf >= 40 and npath >= 5
    ia = np.repeat(np.arange(a.n, dtype=np.uint32), b.n)
    ib = np.tile(np.arange(b.n, dtype=np.uint32), a.n)
    res2 = ctx.search_pairs(A, B, ia, ib, keep=rb.KEEP_ALL, want_paths=True)
    for f in ("score", "lo_a", "lo_b", "hi_a", "hi_b", "ts", "flags", "path_len", "mu_fwd", "mu_rev"):
        assert np.array_equal(res.hits[f], res2.hits[f]), f
    ctx.close()


@pytest.mark.parametrize("mode", [3, 2, 1])
def test_selfrev_scores_match_reference(rb, mode):
    """GetSelfRevScore (alignpair.cpp:7-25) under ProfileLoader parameters (omega = 0, profileloader.cpp:22-26):
    chain vs its coordinate-reversed self (reference-made reversed profiles in the fixtures, forward Mu letters)."""
    from tests.golden_util import GOLDEN, load_chains, load_pairs
    from tests.util import bits
    g = load_pairs(mode)
    chains = load_chains()
    d = np.load(GOLDEN / "golden_chains.npz")
    params = rb.params_preset(mode)
    params.omega = 0.0
    ctx = rb.Context(0, params=params)
    if mode == 3:
        for c in chains:
            c.mu = None
    S = ctx.upload_chains(chains)
    lens = np.array([c.L for c in chains], np.uint32)
    mu = None if mode == 3 else np.concatenate([c.mu for c in chains])
    xyz_rev = np.concatenate([c.xyz[:, ::-1] for c in chains], axis=1)
    Srev = ctx.upload(lens, d["rev_prof"], mu, xyz_rev, None)
    sr = ctx.selfrev(S, Srev)
    assert np.array_equal(bits(sr), bits(g["selfrev"])), (sr, g["selfrev"])
    ctx.close()


def test_gapless_mu_prescores_match_oracle(rb, port):
    from reseek_b200 import synth
    from tests.util import bits
    a = synth.make_chains(9, 120, seed=71, length_jitter=0.6)
    b = synth.make_chains(14, 150, seed=72, length_jitter=0.6)
    synth.plant_homologs(b, a, 0.5, seed=73, sub=0.2)
    ctx = rb.Context(0, rb.MODE_SENSITIVE)
    A = ctx.upload(a.lens, a.prof, a.mu, a.xyz, a.selfrev)
    B = ctx.upload(b.lens, b.prof, b.mu, b.xyz, b.selfrev)
    ia = np.repeat(np.arange(a.n, dtype=np.uint32), b.n)
    ib = np.tile(np.arange(b.n, dtype=np.uint32), a.n)
    f, i = ctx.mu_gapless_scores(A, B, ia, ib)
    p = port(2)
    for k in range(len(ia)):
        of, oi = p.mu_gapless(a.chain(int(ia[k]))[1], b.chain(int(ib[k]))[1])
        assert bits(f[k]) == bits(of) and int(i[k]) == oi, (k, f[k], of, i[k], oi)
    ctx.close()
