"""GPU: -global (rsk_align_global = DSSAligner::AlignQueryTarget_Global, global.cpp:7-33) against the reference's own results
(tests/golden/golden_global.npz, tools/make_golden_global.py) and against the CPU oracle on seeded synthetic chains,
including the shapes the wavefront treats specially (one residue, 31/32/33 rows, row lengths around the 4-byte trace words)."""
import numpy as np
import pytest

from tests.golden_util import GOLDEN, load_chains
from tests.util import bits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb(built_lib):
    import reseek_b200
    if reseek_b200.device_count() < 1:
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    return reseek_b200


@pytest.mark.parametrize("mode", [3, 2])
def test_global_matches_reference_fixtures(rb, mode):
    g = np.load(GOLDEN / "golden_global.npz")
    chains = load_chains()
    ctx = rb.Context(0, mode)
    S = ctx.upload_chains(chains)
    res = ctx.align_global(S, S, g["a"], g["b"])
    want_paths = bytes(g[f"paths_mode{mode}"]).decode()
    off = g[f"path_off_mode{mode}"].astype(np.int64)
    assert len(res.hits) == len(g["a"]) >= 300
    nreject = 0
    for k, h in enumerate(res.hits):
        want = want_paths[off[k]:off[k + 1]]
        assert (int(h["a"]), int(h["b"])) == (int(g["a"][k]), int(g["b"][k]))
        assert int(h["flags"]) & rb.HIT_GLOBAL
        assert bits(h["score"]) == bits(g[f"score_mode{mode}"][k]), f"pair {k}: {h['score']} vs {g[f'score_mode{mode}'][k]}"
        assert res.path(k) == want, f"pair {k} path differs"
        if int(h["flags"]) & rb.HIT_MU_REJECTED:
            nreject += 1
            assert want == "" and float(h["score"]) == -9999.0
        else:
            assert (int(h["lo_a"]), int(h["lo_b"])) == (0, 0)
            LA, LB = chains[int(h["a"])].L, chains[int(h["b"])].L
            assert want.count("M") + want.count("D") == LA and want.count("M") + want.count("I") == LB
    assert (nreject > 300) if mode == 2 else (nreject == 0)
    ctx.close()


def test_global_special_shapes_match_oracle(rb, port):
    from reseek_b200 import synth
    from tests.util import to_oracle_chains
    lens = [1, 2, 3, 4, 5, 31, 32, 33, 63, 64, 65, 127, 300, 1100]
    a = synth.make_chains(len(lens), lens, seed=4101)
    b = synth.make_chains(len(lens), lens[::-1], seed=4102)
    ctx = rb.Context(0, 3)
    A = ctx.upload(a.lens, a.prof, a.mu, a.xyz, a.selfrev)
    B = ctx.upload(b.lens, b.prof, b.mu, b.xyz, b.selfrev)
    ia, ib = np.meshgrid(np.arange(len(lens)), np.arange(len(lens)), indexing="ij")
    res = ctx.align_global(A, B, ia.ravel(), ib.ravel())
    p = port(3)
    oa, ob = to_oracle_chains(a), to_oracle_chains(b)
    for k, h in enumerate(res.hits):
        r, path = p.align_pair_global(oa[int(h["a"])], ob[int(h["b"])])
        assert bits(h["score"]) == bits(r.score), f"pair {k} (LA={oa[int(h['a'])].L}, LB={ob[int(h['b'])].L}): {h['score']} vs {r.score}"
        assert res.path(k) == path, f"pair {k} path"
    # empty request and an out-of-range index
    assert len(ctx.align_global(A, B, [], []).hits) == 0
    with pytest.raises(rb.ReseekB200Error):
        ctx.align_global(A, B, [0], [len(lens)])
    ctx.close()


def test_global_related_chains_and_filter_match_oracle(rb, port):
    """Mutated copies (long diagonal runs, few gaps) under -sensitive: the Mu filter decides first, survivors are aligned."""
    from reseek_b200 import synth
    from tests.util import to_oracle_chains
    rng = np.random.default_rng(4201)
    db = synth.make_chains(48, rng.integers(40, 420, size=48), seed=4202)
    q = synth.make_chains(6, rng.integers(60, 400, size=6), seed=4203)
    synth.plant_homologs(db, q, 0.5, seed=4204, sub=0.15)
    ctx = rb.Context(0, 2)
    Q = ctx.upload(q.lens, q.prof, q.mu, q.xyz, q.selfrev)
    D = ctx.upload(db.lens, db.prof, db.mu, db.xyz, db.selfrev)
    ia, ib = np.meshgrid(np.arange(q.n), np.arange(db.n), indexing="ij")
    res = ctx.align_global(Q, D, ia.ravel(), ib.ravel())
    p = port(2)
    oq, od = to_oracle_chains(q), to_oracle_chains(db)
    kept = 0
    for k, h in enumerate(res.hits):
        r, path = p.align_pair_global(oq[int(h["a"])], od[int(h["b"])])
        assert bool(int(h["flags"]) & rb.HIT_MU_REJECTED) == bool(r.filtered), f"pair {k} filter decision"
        assert bits(h["score"]) == bits(r.score) and res.path(k) == path, f"pair {k}"
        kept += not r.filtered
    assert 0 < kept < len(res.hits)
    ctx.close()
