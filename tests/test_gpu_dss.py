"""GPU: DSS feature extraction on the device (dss_kernel.cu) is letter-exact.

Committed fixtures: the 21 real chains of tests/golden/golden_chains.npz carry the reference's own 8 feature planes, Mu
letters, reversed-chain planes and self-reverse scores (tools/make_golden.py).  When the reference's SCOP40 file is at hand
(build/data/scop40.bca, not committed) all 11 211 chains are compared with the reference itself (oracle/_ref) and with the
shipped known-answer file test_data/scop40.mu.fa (dss.cpp:629-644)."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
FEATURES = ["AA", "NENDist", "Conf", "NENConf", "RENDist", "DstNxtHlx", "StrandDens", "NormDens"]


@pytest.fixture(scope="module")
def rb(built_lib):
    import reseek_b200
    if reseek_b200.device_count() < 1:
        pytest.fail("no CUDA device")
    return reseek_b200


def test_dss_golden_chains(rb):
    g = np.load(GOLDEN / "golden_chains.npz", allow_pickle=True)
    ctx = rb.Context(0, rb.MODE_SENSITIVE)
    S = rb.ChainSet.from_coords(ctx, g["lens"], g["seq"], g["xyz"], with_mu=True)
    prof, mu, sr = S.download_features()
    for f in range(8):
        bad = np.nonzero(prof[f] != g["prof"][f])[0]
        assert len(bad) == 0, f"{FEATURES[f]}: {len(bad)} letters differ, first at residue {bad[:5]}"
    assert np.array_equal(mu, g["mu"]), "Mu letters"
    assert np.all(sr > 1e38)  # self-reverse scores start unset
    R = S.reversed()
    rprof, rmu, _ = R.download_features()
    for f in range(8):
        assert np.array_equal(rprof[f], g["rev_prof"][f]), f"reversed chains: {FEATURES[f]}"
    assert np.array_equal(rmu, g["mu"])  # forward letters on purpose (alignpair.cpp:22)
    ctx.close()


@pytest.mark.parametrize("mode", [3, 2, 1])
def test_dss_then_selfrev_equals_reference(rb, mode):
    """coordinates -> device DSS -> reversed set -> self-reverse scores, with the loader's parameters (omega = 0,
    profileloader.cpp:22-26): the float bits of the reference's GetSelfRevScore for the 21 golden chains."""
    from tests.golden_util import load_pairs
    from tests.util import bits
    g = np.load(GOLDEN / "golden_chains.npz", allow_pickle=True)
    want = load_pairs(mode)["selfrev"]
    p = rb.params_preset(mode)
    p.omega = 0
    ctx = rb.Context(0, params=p)
    S = rb.ChainSet.from_coords(ctx, g["lens"], g["seq"], g["xyz"], with_mu=(mode != 3))
    R = S.reversed()
    got = ctx.selfrev(S, R)
    assert np.array_equal(bits(got), bits(want)), np.nonzero(bits(got) != bits(want))
    _, _, sr = S.download_features(want_mu=False)
    assert np.array_equal(bits(sr), bits(want))
    ctx.close()


def test_dss_scop40_all_chains(rb):
    from reseek_b200 import chainio
    bca, mufa = ROOT / "build" / "data" / "scop40.bca", ROOT / "build" / "data" / "scop40.mu.fa"
    if not bca.exists():
        pytest.skip("build/data/scop40.bca (the reference's test_data, not committed) is not here")
    import time
    labels, seqs, xyzs = chainio.read_bca(bca)
    lens = np.array([len(s) for s in seqs], np.uint32)
    aa = np.frombuffer(b"".join(seqs), np.uint8)
    xyz = np.concatenate(xyzs, axis=1)
    ctx = rb.Context(0, rb.MODE_SENSITIVE)
    S = rb.ChainSet.from_coords(ctx, lens, aa, xyz, with_mu=True)  # warm-up: tables, scratch
    S.free()
    t0 = time.perf_counter()
    S = rb.ChainSet.from_coords(ctx, lens, aa, xyz, with_mu=True)
    dt = time.perf_counter() - t0
    prof, mu, _ = S.download_features()
    print(f"device DSS: {len(lens)} chains, {int(lens.sum())} residues in {dt * 1e3:.1f} ms (upload + kernel)")
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    # (1) the shipped known-answer file (dss.cpp:629-644): records are in chain order, letters 'A' + letter.  It was NOT made
    # with the strict-IEEE build: the reference compiled with -O2 -ffp-contract=off (the canonical oracle) differs
    # from it in 239 of the 11 211 chains (FP flags flip NEN argmins and bin thresholds) - so must a letter-exact DSS.
    if mufa.exists():
        recs = []
        for line in mufa.read_text().splitlines():
            if line.startswith(">"):
                recs.append([line[1:].strip(), []])
            elif recs:
                recs[-1][1].append(line.strip())
        assert [r[0] for r in recs] == labels
        nbad = 0
        for i, (_, parts) in enumerate(recs):
            v = np.frombuffer("".join(parts).encode(), np.uint8).astype(np.int32)
            v = np.where(v >= ord("a"), v - ord("a") + 26, v - ord("A")).astype(np.uint8)
            nbad += not np.array_equal(v, mu[off[i]:off[i + 1]])
        print(f"chains whose Mu letters differ from the shipped scop40.mu.fa: {nbad} (strict reference build: 239)")
        assert nbad == 239
    # (2) the reference itself: all 8 planes + Mu letters, every chain
    from oracle.pyoracle import Ref
    if Ref.available():
        ref = Ref(2)
        nbad = 0
        for i in range(len(labels)):
            rp, rm, _ = ref.dss(seqs[i], xyzs[i])
            s, e = off[i], off[i + 1]
            if not (np.array_equal(rp, prof[:, s:e]) and np.array_equal(rm, mu[s:e])):
                nbad += 1
                if nbad <= 3:
                    f = [FEATURES[k] for k in range(8) if not np.array_equal(rp[k], prof[k, s:e])]
                    print(f"chain {i} {labels[i]} L={e - s}: planes {f} differ")
        assert nbad == 0, f"{nbad} of {len(labels)} chains differ from the reference's DSS"
    ctx.close()
