"""Shared helpers of the test-suite (test infrastructure; may import oracle/)."""
import numpy as np

from oracle.pyoracle import Chain


def to_oracle_chains(sc):
    """reseek_b200.synth.SynthChains -> list of oracle Chain objects."""
    out = []
    for i in range(sc.n):
        p, m, x, sr = sc.chain(i)
        out.append(Chain(p.copy(), m.copy(), x.copy(), sr))
    return out


def bits(x):
    return np.asarray(x, np.float32).view(np.uint32)


def assert_hit_matches_oracle(h, path, r, rpath, ctx=""):
    """h: one rsk_hit record (numpy void), r: OrcResult.  Bit-exact on everything but P/E/Qual,
    which are (float) casts of double libm results computed from a bit-exact test statistic."""
    if r.filtered:
        assert int(h["flags"]) & 1, f"{ctx}: oracle rejects in the Mu filter, library does not"
    else:
        assert not (int(h["flags"]) & 1), f"{ctx}: library rejects in the Mu filter, oracle does not"
    if not (int(h["flags"]) & 8):  # RSK_HIT_MKF pairs carry m_BestHSPScore / m_BestChainScore in mu_fwd / mu_rev instead
        assert (int(h["mu_fwd"]), int(h["mu_rev"])) == (r.mu_fwd, r.mu_rev), f"{ctx} mu fwd/rev {h['mu_fwd']},{h['mu_rev']} vs {r.mu_fwd},{r.mu_rev}"
        assert float(h["mu_score"]) == r.mu_score, f"{ctx} mu score"
    assert bits(h["score"]) == bits(r.score), f"{ctx} score {h['score']} vs {r.score}"
    assert path == rpath, f"{ctx} path differs"
    assert int(h["path_len"]) == r.path_len, ctx
    if r.path_len:
        assert int(h["lo_a"]) == r.lo_a and int(h["lo_b"]) == r.lo_b, ctx
    assert int(h["hi_a"]) == r.hi_a and int(h["hi_b"]) == r.hi_b, f"{ctx} hi"
    assert int(h["ids"]) == r.ids and int(h["gaps"]) == r.gaps, f"{ctx} ids/gaps"
    assert bits(h["ts"]) == bits(r.ts), f"{ctx} ts {h['ts']} vs {r.ts}"
    assert bits(h["lddt"]) == bits(r.lddt), f"{ctx} lddt {h['lddt']} vs {r.lddt}"
    for k, v in (("evalue", r.evalue), ("pvalue", r.pvalue)):
        a, b = float(h[k]), float(v)
        assert a == b or abs(a - b) <= 1e-6 * abs(b), f"{ctx} {k} {a} vs {b}"  # tolerance stated by north_star
