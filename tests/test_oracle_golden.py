"""CPU: the C oracle (oracle/reseek_oracle.c) against the golden vectors produced by the unmodified reference."""
import numpy as np
import pytest

from tests.golden_util import GOLDEN, load_chains, load_pairs
from tests.util import bits


@pytest.mark.parametrize("mode", [3, 2, 1])
def test_align_pair_matches_reference_fixtures(port, mode):
    g = load_pairs(mode)
    chains = load_chains(g["selfrev"])
    p = port(mode)
    n = checked = with_path = 0
    for k in range(len(g["a"])):
        A, B = chains[int(g["a"][k])], chains[int(g["b"][k])]
        if mode == 3:  # verysensitive never loads Mu letters (dbsearcher.cpp:251-252)
            A = type(A)(A.prof, None, A.xyz, A.selfrev)
            B = type(B)(B.prof, None, B.xyz, B.selfrev)
        r, path = p.align_pair(A, B)
        n += 1
        assert path == g["path_list"][k], f"pair {k} path"
        assert bits(r.score) == bits(g["score"][k]), f"pair {k} score {r.score} vs {g['score'][k]}"
        if path:
            with_path += 1
            assert (r.lo_a, r.lo_b) == (int(g["lo_a"][k]), int(g["lo_b"][k]))
        assert (r.hi_a, r.hi_b, r.ids, r.gaps) == (int(g["hi_a"][k]), int(g["hi_b"][k]), int(g["ids"][k]), int(g["gaps"][k]))
        assert bits(r.ts) == bits(g["ts"][k]), f"pair {k} ts"
        assert bits(r.evalue) == bits(g["evalue"][k]) and bits(r.pvalue) == bits(g["pvalue"][k]), f"pair {k} E/P"
        if g["evalue"][k] < 1e38:
            assert bits(r.qual) == bits(g["qual"][k]) and bits(r.lddt) == bits(g["lddt"][k]), f"pair {k} qual/lddt"
            checked += 1
        if mode != 3 and not r.filtered:
            assert r.mu_score == g["mu_score"][k], f"pair {k} mu score"
    assert n == len(g['a']) and with_path > 10 and checked > 10
    if mode != 3:
        assert int(np.sum(g['mkf'])) > 100  # the long-chain (MKF / x-drop) path is part of the fixtures


def test_mu_sw_matches_parasail_fixtures(port):
    g = np.load(GOLDEN / "golden_mu_sw.npz")
    p = port(2)
    oa = np.concatenate([[0], np.cumsum(g["la"])]).astype(np.int64)
    ob = np.concatenate([[0], np.cumsum(g["lb"])]).astype(np.int64)
    nsat = 0
    for k in range(len(g["la"])):
        a, b = g["a"][oa[k]:oa[k + 1]], g["b"][ob[k]:ob[k + 1]]
        s, sat = p.mu_sw(a, b)
        assert sat == int(g["sat"][k]), f"case {k} saturation"
        if sat:
            nsat += 1
            assert int(g["score"][k]) == 255  # parasail reports INT8_MAX - bias when saturated
        else:
            assert s == int(g["score"][k]), f"case {k}: {s} vs {g['score'][k]}"
    assert nsat > 10


def test_swfast_matches_reference_fixtures(port):
    g = np.load(GOLDEN / "golden_swfast.npz")
    p = port(3)
    paths = bytes(g["paths"]).decode()
    po = g["path_off"].astype(np.int64)
    off = 0
    for k in range(len(g["la"])):
        la, lb = int(g["la"][k]), int(g["lb"][k])
        S = g["mats"][off:off + la * lb].reshape(la, lb)
        off += la * lb
        open_, ext = (-0.685533, -0.051881) if k % 2 else (-1.5, -0.42)
        s, lo_a, lo_b, path = p.swfast_matrix(S, open_, ext)
        assert bits(s) == bits(g["score"][k]), f"case {k}"
        assert path == paths[po[k]:po[k + 1]], f"case {k} path"
        if path:
            assert (lo_a, lo_b) == (int(g["lo_a"][k]), int(g["lo_b"][k]))


def test_statsig_formulas(port):
    p = port(3)
    # elbow at 0.11, P clamps at 1, E = P * 8340 (statsig.cpp:27-50, statsig.h:3)
    pv, ev, q = p.statsig(-1.0)
    assert pv == 1.0 and ev == 8340.0
    pv, ev, q = p.statsig(0.5)
    assert pv == 10 ** (-52 * 0.5 - 3.7) and ev == pv * 8340 and q == 1 / (1 + 10 ** ((5.0 - 40.0 * 0.5) / 10) / 2)
    assert p.statsig(0.7)[2] == 1.0  # logE < -20
    pv, _, _ = p.statsig(0.05)
    assert pv == 10 ** (-80 * 0.05 - 0.58)


def _prefilter_fixture():
    g = np.load(GOLDEN / "golden_prefilter.npz")
    qo = np.concatenate([[0], np.cumsum(g["q_len"])]).astype(np.int64)
    to = np.concatenate([[0], np.cumsum(g["t_len"])]).astype(np.int64)
    mq = [g["q_mu"][qo[i]:qo[i + 1]] for i in range(len(g["q_len"]))]
    mt = [g["t_mu"][to[i]:to[i + 1]] for i in range(len(g["t_len"]))]
    return g, mq, mt


@pytest.mark.parametrize("name,kw", [("idxq", {}), ("idxt", {"query_neighborhood": False}), ("rsb5", {"rsb_size": 5})])
def test_prefilter_matches_reference_candidate_lists(port, name, kw):
    """MuPreFilter (5-mer index probe, two-hit diagonals, FindHSP, top-B bag) against the candidate TSV the reference
    binary wrote for q10.bca vs q100.bca at -threads 1: query-neighbourhood, target-neighbourhood and a tiny bag."""
    g, mq, mt = _prefilter_fixture()
    got, _ = port(1).prefilter(mq, mt, **kw)
    pairs = [(t, q) for t in sorted(got) for q in got[t]]
    want = list(zip(g[f"{name}_t"].tolist(), g[f"{name}_q"].tolist()))
    assert pairs == want and len(want) >= 50


@pytest.mark.parametrize("mode", [3, 2])
def test_global_alignment_matches_reference_fixtures(port, mode):
    """-global: the restatement of ViterbiFastMem + TraceBackBitMem under AlignQueryTarget_Global (global.cpp:7-33) against
    the reference's results (tools/make_golden_global.py): m_GlobalScore bit for bit and the same path; under -sensitive the
    Mu filter rejects first (score stays -9999, no path)."""
    g = np.load(GOLDEN / "golden_global.npz")
    chains = load_chains()
    p = port(mode)
    paths = bytes(g[f"paths_mode{mode}"]).decode()
    off = g[f"path_off_mode{mode}"].astype(np.int64)
    step = 1 if mode == 2 else 3  # every third pair without the filter keeps the CPU suite short
    n = 0
    for k in range(0, len(g["a"]), step):
        r, path = p.align_pair_global(chains[int(g["a"][k])], chains[int(g["b"][k])])
        assert bits(r.score) == bits(g[f"score_mode{mode}"][k]), f"pair {k}"
        assert path == paths[off[k]:off[k + 1]], f"pair {k} path"
        n += 1
    assert n >= 100
