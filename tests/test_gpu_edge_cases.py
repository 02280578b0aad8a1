"""GPU: the corners of the pair path - tiny and very ragged chains, pairs without any positive cell, identical chains
(full-length diagonal paths across many checkpoint strips and passes), very long chains (many passes), empty requests and
bad arguments (the reference would asserta/Die; the C ABI returns an error and a message)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb(built_lib):
    import reseek_b200
    if reseek_b200.device_count() < 1:
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    return reseek_b200


def _check_all(port, res, oa, ob):
    from tests.util import assert_hit_matches_oracle
    for k, h in enumerate(res.hits):
        r, rpath = port.align_pair(oa[int(h["a"])], ob[int(h["b"])])
        assert_hit_matches_oracle(h, res.path(k), r, rpath, ctx=f"pair {k} (a={h['a']} b={h['b']})")


@pytest.mark.parametrize("mode", [3, 2])
def test_tiny_and_ragged_chains(rb, port, mode):
    from reseek_b200 import synth
    from tests.util import to_oracle_chains
    lens = [1, 2, 3, 4, 7, 8, 31, 32, 33, 1400]
    a = synth.make_chains(len(lens), lens, seed=901)
    b = synth.make_chains(len(lens), lens[::-1], seed=902)
    ctx = rb.Context(0, mode)
    A = ctx.upload(a.lens, a.prof, a.mu, a.xyz, a.selfrev)
    B = ctx.upload(b.lens, b.prof, b.mu, b.xyz, b.selfrev)
    p = port(mode)
    oa, ob = to_oracle_chains(a), to_oracle_chains(b)
    res = ctx.search_cross(A, B, keep=rb.KEEP_ALL)
    assert len(res.hits) == len(lens) ** 2
    _check_all(p, res, oa, ob)
    res = ctx.search_self(A, keep=rb.KEEP_ALL)
    _check_all(p, res, oa, oa)
    ctx.close()


def test_pairs_without_positive_cell(rb, port):
    """Chains whose every cell scores negative: SWFast returns 0 and an empty path (sw.cpp:200-201).  With MinFwdScore = 0
    (-verysensitive) the reference still walks through CalcEvalue (dssaligner.cpp:861): Hi = Lo - 1, no columns, E = 8340."""
    from reseek_b200 import synth
    from tests.util import to_oracle_chains
    a = synth.make_chains(3, 40, seed=911)
    b = synth.make_chains(4, 55, seed=912)
    tab = np.array(rb.params_preset(3).tables[:], np.float32)
    # letters of the most negative entry of every feature table: A gets the row letter everywhere, B the column letter
    offs = [0] + [400 + 256 * k for k in range(7)]
    alph = [20] + [16] * 7
    for f in range(8):
        t = tab[offs[f]:offs[f] + alph[f] ** 2].reshape(alph[f], alph[f])
        i, j = np.unravel_index(np.argmin(t), t.shape)
        assert t[i, j] < 0
        a.prof[f, :] = i
        b.prof[f, :] = j
    ctx = rb.Context(0, rb.MODE_VERYSENSITIVE)
    A = ctx.upload(a.lens, a.prof, a.mu, a.xyz, a.selfrev)
    B = ctx.upload(b.lens, b.prof, b.mu, b.xyz, b.selfrev)
    res = ctx.search_cross(A, B, keep=rb.KEEP_ALL)
    assert (res.hits["score"] == 0).all() and (res.hits["path_len"] == 0).all()
    assert ((res.hits["flags"] & rb.HIT_REPORTED) == 0).all() and ((res.hits["flags"] & rb.HIT_HAS_EVALUE) != 0).all()
    assert (res.hits["evalue"] == 8340).all() and (res.hits["pvalue"] == 1).all()
    assert (res.hits["lo_a"] == 0xFFFFFFFF).all() and (res.hits["hi_a"] == 0xFFFFFFFE).all() and (res.hits["ids"] == 0).all()
    _check_all(port(3), res, to_oracle_chains(a), to_oracle_chains(b))
    assert len(ctx.search_cross(A, B, keep=rb.KEEP_HITS).hits) == 0
    # the same pairs under -sensitive: score 0 < MinFwdScore 7 -> CalcEvalue is skipped, members keep their ClearAlign values
    ctx.set_params(rb.params_preset(rb.MODE_SENSITIVE))
    nomu_a = ctx.upload(a.lens, a.prof, None, a.xyz, a.selfrev)
    nomu_b = ctx.upload(b.lens, b.prof, None, b.xyz, b.selfrev)
    res = ctx.search_cross(nomu_a, nomu_b, keep=rb.KEEP_ALL)
    assert ((res.hits["flags"] & rb.HIT_HAS_EVALUE) == 0).all() and (res.hits["hi_a"] == 0xFFFFFFFF).all() and (res.hits["evalue"] > 1e30).all()
    ctx.close()


def test_identical_and_very_long_chains(rb, port):
    """Self alignments (one long diagonal: the traceback re-runs every strip of every pass) and chains of several thousand
    residues (up to 9 passes of 384 rows; LDDT over thousands of columns)."""
    from reseek_b200 import synth
    from tests.util import to_oracle_chains
    lens = [50, 383, 384, 385, 769, 1500, 3300]
    s = synth.make_chains(len(lens), lens, seed=921)
    ctx = rb.Context(0, rb.MODE_VERYSENSITIVE)
    S = ctx.upload(s.lens, s.prof, s.mu, s.xyz, s.selfrev)
    idx = np.arange(len(lens), dtype=np.uint32)
    res = ctx.search_pairs(S, S, idx, idx, keep=rb.KEEP_ALL)
    for k, h in enumerate(res.hits):
        assert res.path(k) == "M" * lens[k] and int(h["lo_a"]) == 0 and int(h["hi_b"]) == lens[k] - 1
        assert float(h["lddt"]) == 1.0
    oc = to_oracle_chains(s)
    _check_all(port(3), res, oc, oc)
    # every long chain against every other one, both orientations of the kernel
    ia, ib = np.meshgrid(idx, idx, indexing="ij")
    res = ctx.search_pairs(S, S, ia.ravel().astype(np.uint32), ib.ravel().astype(np.uint32), keep=rb.KEEP_ALL)
    _check_all(port(3), res, oc, oc)
    res = ctx.search_cross(S, S, keep=rb.KEEP_ALL)
    _check_all(port(3), res, oc, oc)
    ctx.close()


def test_alignment_longer_than_the_lddt_shared_memory(rb, port):
    """A 9 000-residue chain against itself and against a mutated copy: the alignment has more columns than the LDDT kernel's
    shared-memory buffers hold (~8 000), so the kernel keeps them in global scratch - same records as the oracle (round 1
    returned RSK_ERR_LIMIT for the whole call)."""
    from reseek_b200 import synth
    from tests.util import to_oracle_chains
    s = synth.make_chains(1, [9000], seed=931)
    t = synth.make_chains(1, [9000], seed=932)
    synth.plant_homologs(t, s, 1.0, seed=933, sub=0.1, indel=0.002)
    for mode in (rb.MODE_VERYSENSITIVE, rb.MODE_FAST):
        ctx = rb.Context(0, mode)
        S = ctx.upload(s.lens, s.prof, s.mu, s.xyz, s.selfrev)
        T = ctx.upload(t.lens, t.prof, t.mu, t.xyz, t.selfrev)
        z = np.zeros(1, np.uint32)
        res = ctx.search_pairs(S, S, z, z, keep=rb.KEEP_ALL)
        assert res.path(0) == "M" * 9000 and float(res.hits[0]["lddt"]) == 1.0
        _check_all(port(mode), res, to_oracle_chains(s), to_oracle_chains(s))
        res = ctx.search_pairs(S, T, z, z, keep=rb.KEEP_ALL)
        assert int(res.hits[0]["ids"]) > 8100, "the alignment covers (nearly) the whole chain"
        _check_all(port(mode), res, to_oracle_chains(s), to_oracle_chains(t))
        ctx.close()


def test_empty_requests_and_bad_arguments(rb):
    from reseek_b200 import synth
    s = synth.make_chains(4, 30, seed=931)
    ctx = rb.Context(0, rb.MODE_SENSITIVE)
    S = ctx.upload(s.lens, s.prof, s.mu, s.xyz, s.selfrev)
    e = np.zeros(0, np.uint32)
    assert len(ctx.search_pairs(S, S, e, e).hits) == 0
    with pytest.raises(rb.ReseekB200Error, match="out of range"):
        ctx.search_pairs(S, S, np.array([0], np.uint32), np.array([4], np.uint32))
    with pytest.raises(rb.ReseekB200Error, match="length 0"):
        ctx.upload(np.array([5, 0], np.uint32), s.prof[:, :5], s.mu[:5], s.xyz[:, :5], None)
    L = rb.load_library()
    h = rb.lib.ChainsHost(2, 11, s.lens.ctypes.data_as(C.c_void_p), s.prof.ctypes.data_as(C.c_void_p), None,
                          s.xyz.ctypes.data_as(C.c_void_p), None)
    out = C.c_void_p()
    assert L.rsk_chainset_upload(ctx.handle, C.byref(h), C.byref(out)) == -1 and b"sum(len)" in L.rsk_last_error()
    assert L.rsk_search_cross(ctx.handle, None, S.handle, None, C.byref(out)) == -1
    other = rb.Context(0, rb.MODE_SENSITIVE)
    with pytest.raises(rb.ReseekB200Error, match="different context"):
        other.search_cross(S, S)
    p = rb.params_preset(2)
    p.gap_open = 0.5
    with pytest.raises(rb.ReseekB200Error, match="gap penalties"):
        ctx.set_params(p)  # dssparams.cpp:106-108
    other.close()
    ctx.close()


@pytest.mark.parametrize("mode,keep", [(3, 1), (2, 0)])
def test_self_search_in_row_chunks_equals_one_piece(rb, mode, keep, monkeypatch):
    """rsk_search_self cuts large sets into row chunks (the pair list of 1e5 chains would not fit the host): same records,
    same order, same paths, summed statistics."""
    from reseek_b200 import synth
    s = synth.make_chains(60, 90, seed=941, length_jitter=0.5)
    synth.plant_homologs(s, s.subset(range(6)), 0.3, seed=942)
    ctx = rb.Context(0, mode)
    S = ctx.upload(s.lens, s.prof, s.mu, s.xyz, s.selfrev)
    whole = ctx.search_self(S, keep=keep, want_paths=True)
    st0 = ctx.stats()
    monkeypatch.setenv("RSK_SELF_CHUNK_PAIRS", "200")
    parts = ctx.search_self(S, keep=keep, want_paths=True)
    st1 = ctx.stats()
    assert len(whole.hits) == len(parts.hits) > 0
    for f in whole.hits.dtype.names:
        if f != "path_off":
            assert np.array_equal(whole.hits[f], parts.hits[f]), f
    assert [whole.path(k) for k in range(len(whole.hits))] == [parts.path(k) for k in range(len(parts.hits))]
    for k in ("pairs", "sw_pairs", "sw_cells", "evalue_pairs", "hits", "mu_filter_in", "mu_filter_rejected"):
        assert st0[k] == st1[k], k
    assert st1["sw_kernel_launches"] > st0["sw_kernel_launches"]
    ctx.close()


def test_mu_filter_32bit_kernel_equals_packed_kernel(rb, monkeypatch):
    """The packed 16-bit Mu filter (two chains per warp) is exact up to 8 000 residues; beyond that the host routes the batch
    to the 32-bit kernel (RSK_MU32 forces it).  Both must give the same records."""
    from reseek_b200 import synth
    a = synth.make_chains(9, 170, seed=951, length_jitter=0.5)
    b = synth.make_chains(70, 150, seed=952, length_jitter=0.6)
    synth.plant_homologs(b, a, 0.5, seed=953)
    ctx = rb.Context(0, rb.MODE_SENSITIVE)
    A = ctx.upload(a.lens, a.prof, a.mu, a.xyz, a.selfrev)
    B = ctx.upload(b.lens, b.prof, b.mu, b.xyz, b.selfrev)
    r16 = ctx.search_cross(A, B, keep=rb.KEEP_ALL, want_paths=True)
    s16 = ctx.search_self(B, keep=rb.KEEP_ALL, want_paths=False)
    monkeypatch.setenv("RSK_MU32", "1")
    r32 = ctx.search_cross(A, B, keep=rb.KEEP_ALL, want_paths=True)
    s32 = ctx.search_self(B, keep=rb.KEEP_ALL, want_paths=False)
    for x, y in ((r16, r32), (s16, s32)):
        assert ((x.hits["flags"] & rb.HIT_MU_REJECTED) != 0).sum() > 10
        for f in x.hits.dtype.names:
            if f != "path_off":
                assert np.array_equal(x.hits[f], y.hits[f]), f
    ctx.close()
