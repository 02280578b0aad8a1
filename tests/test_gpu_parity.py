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


def test_smoke_entry(rb):
    import __graft_entry__ as g
    g.smoke()
