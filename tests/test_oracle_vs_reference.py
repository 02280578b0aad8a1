"""CPU, build container only: the C oracle against the live reference library on fresh random inputs."""
import numpy as np
import pytest

from oracle.pyoracle import Ref
from tests.util import bits

pytestmark = pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built (needs /root/reference)")


def test_tables_and_presets(port):
    for mode in (1, 2, 3):
        ref = Ref(mode)
        p = port(mode)
        sc, tb = ref.get_params()
        assert np.array_equal(bits(tb), bits(p.tables()))
        P = p.params
        mine = [P.gap_open, P.gap_ext, P.min_fwd_score, P.omega, P.omega_fwd, P.mkfl, P.mkf_x1, P.mkf_x2,
                P.mkf_min_hsp_score, P.mkf_min_mega_hsp_score, P.mu_gap_open, P.mu_gap_ext]
        assert np.array_equal(bits(np.array(mine, np.float32)), bits(sc))
    f, a, b = ref.mu_matrices()
    assert np.array_equal(bits(f), bits(p.mu_f32())) and np.array_equal(a, p.mu_i8()) and np.array_equal(b, p.mu_kmer_i8())


@pytest.mark.parametrize("mode", [3, 2])
def test_random_synthetic_pairs(port, built_lib, mode):
    from reseek_b200 import synth
    from tests.util import to_oracle_chains
    ref = Ref(mode)
    p = port(mode)
    a = synth.make_chains(6, 90, seed=31 + mode, length_jitter=0.4)
    b = synth.make_chains(12, 110, seed=41 + mode, length_jitter=0.4)
    synth.plant_homologs(b, a, 0.5, seed=51)
    ca, cb = to_oracle_chains(a), to_oracle_chains(b)
    nalign = 0
    for A in ca:
        for B in cb:
            rr, rpath = ref.align_pair(A, B, use_mu=(mode != 3), use_kmers=False)
            if mode == 3:
                A2, B2 = type(A)(A.prof, None, A.xyz, A.selfrev), type(B)(B.prof, None, B.xyz, B.selfrev)
            else:
                A2, B2 = A, B
            r, path = p.align_pair(A2, B2)
            assert path == rpath and bits(r.score) == bits(rr.score)
            assert bits(r.ts) == bits(rr.ts) and bits(r.evalue) == bits(rr.evalue)
            nalign += bool(path)
    assert nalign > 5


def test_mu_filter_random(port):
    ref = Ref(2)
    p = port(2)
    rng = np.random.default_rng(3)
    for t in range(300):
        la, lb = int(rng.integers(1, 400)), int(rng.integers(1, 400))
        a = rng.integers(0, 36, la).astype(np.uint8)
        b = rng.integers(0, 36, lb).astype(np.uint8)
        if t % 3 == 0:
            n = min(la, lb)
            b[:n] = a[:n]
        s, fwd, rev = p.mu_filter_score(a, b)
        assert s == ref.mu_score(a, b), f"case {t}"


@pytest.mark.parametrize("mode", [2, 1])
def test_mkf_xdrop_path_random_long_chains(port, built_lib, mode):
    """Chains >= MKFL: 3-mer seeds, ordered HSP gating, chaining, mega-HSP scores, 8-mer seed, banded x-drop
    forward/backward with traceback and merge against the live reference."""
    from reseek_b200 import synth
    from tests.util import to_oracle_chains
    from oracle.pyoracle import Chain
    ref = Ref(mode)
    p = port(mode)
    a = synth.make_chains(6, [650, 700, 820, 610, 900, 640], seed=61)
    b = synth.make_chains(8, [300, 640, 720, 150, 660, 1000, 90, 605], seed=62)
    synth.plant_homologs(b, a, 0.9, seed=63, sub=0.25, indel=0.03)
    ca, cb = to_oracle_chains(a), to_oracle_chains(b)
    for c in ca + cb:  # 3-mers as DSS::GetMuKmers makes them (pattern "111")
        m = c.mu.astype(np.uint32)
        c.kmers = (m[:-2] * 36 + m[1:-1]) * 36 + m[2:]
    npath = nmkf = 0
    for A in ca:
        for B in cb:
            rr, rpath = ref.align_pair(A, B)
            r, path = p.align_pair(A, B)
            nmkf += rr.mkf
            assert path == rpath, (A.L, B.L)
            assert bits(r.score) == bits(rr.score)
            assert bits(r.ts) == bits(rr.ts) and bits(r.evalue) == bits(rr.evalue)
            if path:
                npath += 1
                assert (r.lo_a, r.lo_b, r.hi_a, r.hi_b) == (rr.lo_a, rr.lo_b, rr.hi_a, rr.hi_b)
    assert nmkf == len(ca) * len(cb) and npath >= 5


def test_gapless_variants_random(port):
    """SWFastGaplessProfb / SWFastPinopGapless against the live reference functions."""
    ref = Ref(2)
    p = port(2)
    rng = np.random.default_rng(11)
    for t in range(200):
        la, lb = int(rng.integers(1, 300)), int(rng.integers(1, 300))
        a = rng.integers(0, 36, la).astype(np.uint8)
        b = rng.integers(0, 36, lb).astype(np.uint8)
        if t % 3 == 0:
            n = min(la, lb)
            b[:n] = a[:n]
        x, y = p.mu_gapless(a, b), ref.mu_gapless(a, b)
        assert bits(x[0]) == bits(y[0]) and x[1] == y[1], (la, lb, x, y)
