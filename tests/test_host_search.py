"""Whole searches from .bca files through the host C++ layer: .bca reader -> DSS look-alike (feature extraction) ->
GPU search -> hit writer, against golden outputs of the reference BINARY on the same files (tools/make_golden_search.py).

CPU part: the .bca reader and the DSS stage reproduce the reference's feature letters exactly (no GPU involved).
GPU part: `-search X`, `-search Q -db DB` and `-search Q -db DB -fast` give the reference's TSV, line for line."""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from tests.golden_util import ALN_COLUMNS, GLOBAL_COLUMNS, GOLDEN, SEARCH_COLUMNS, aln_blocks, fasta2_records, golden_bca, golden_bca_short

ROOT = Path(__file__).resolve().parent.parent
DEMO = ROOT / "reseek_b200" / "rsk_host_demo"
HOSTLIB = ROOT / "reseek_b200" / "libreseek_b200_host.so"


def _run(*args, **extra_env):
    env = dict(os.environ, RSK_BLOCK_CHAINS="8", **{k: str(v) for k, v in extra_env.items()})  # streamed side in several blocks
    return subprocess.run([str(DEMO), *map(str, args)], capture_output=True, text=True, timeout=900, env=env)


def _golden(name):
    return (GOLDEN / name).read_text().splitlines()


@pytest.mark.gpu
def test_dss_lookalike_reproduces_reference_letters(built_lib):
    """rskh_dss_features (DSS::GetProfile / GetMuLetters restated on the host) on 21 real chains (49..1231 residues):
    all 8 feature planes, the Mu letters and the profile of the coordinate-reversed chain, letter for letter."""
    H = C.CDLL(str(HOSTLIB))
    g = np.load(GOLDEN / "golden_chains.npz", allow_pickle=True)
    lens = g["lens"].astype(np.int64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    for i in range(len(lens)):
        s, e = int(off[i]), int(off[i + 1])
        n = e - s
        x, y, z = (np.ascontiguousarray(g["xyz"][k, s:e], np.float32) for k in range(3))
        prof, mu, rev = np.zeros((8, n), np.uint8), np.zeros(n, np.uint8), np.zeros((8, n), np.uint8)
        rc = H.rskh_dss_features(bytes(g["seq"][s:e]), x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p),
                                 z.ctypes.data_as(C.c_void_p), n, prof.ctypes.data_as(C.c_void_p),
                                 mu.ctypes.data_as(C.c_void_p), rev.ctypes.data_as(C.c_void_p))
        assert rc == 0
        assert np.array_equal(prof, g["prof"][:, s:e]), f"chain {i}: profile"
        assert np.array_equal(mu, g["mu"][s:e]), f"chain {i}: Mu letters"
        assert np.array_equal(rev, g["rev_prof"][:, s:e]), f"chain {i}: reversed-chain profile"


@pytest.mark.gpu
def test_dss_lookalike_vs_live_reference(built_lib):
    """Same check against the compiled reference on 300 SCOP40 chains (build container only)."""
    from oracle.pyoracle import Ref
    bca = Path("/root/reference/test_data/scop40.bca")
    if not Ref.available() or not bca.exists():
        pytest.skip("needs oracle/_ref and the reference's test data")
    H = C.CDLL(str(HOSTLIB))
    ref = Ref(2)
    n = ref.bca_open(bca)
    rng = np.random.default_rng(7)
    for i in sorted(rng.choice(n, 300, replace=False).tolist()):
        label, seq, xyz = ref.bca_chain(i)
        L = xyz.shape[1]
        p_ref, mu_ref, _ = ref.dss(seq, xyz)
        rp_ref = ref.rev_profile(seq, xyz)
        x, y, z = (np.ascontiguousarray(xyz[k]) for k in range(3))
        prof, mu, rev = np.zeros((8, L), np.uint8), np.zeros(L, np.uint8), np.zeros((8, L), np.uint8)
        assert H.rskh_dss_features(seq, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p), L,
                                   prof.ctypes.data_as(C.c_void_p), mu.ctypes.data_as(C.c_void_p), rev.ctypes.data_as(C.c_void_p)) == 0
        assert np.array_equal(prof, p_ref) and np.array_equal(mu, mu_ref) and np.array_equal(rev, rp_ref), label


@pytest.mark.gpu
def test_bca_reader_and_features_command(built_lib, tmp_path):
    """.bca written by reseek_b200.chainio -> BCAData::Open/ReadChain -> DSS -> .rskc dump; no GPU is touched."""
    from reseek_b200 import chainio
    g6, g21 = golden_bca(tmp_path)
    r = _run("features", g21, tmp_path / "g21.rskc")
    assert r.returncode == 0, r.stderr
    d, labels, seq = chainio.read_rskc(tmp_path / "g21.rskc")
    g = np.load(GOLDEN / "golden_chains.npz", allow_pickle=True)
    assert np.array_equal(d["lens"], g["lens"]) and labels == [str(x) for x in g["labels"]] and np.array_equal(seq, g["seq"])
    assert np.array_equal(d["xyz"], g["xyz"]), "integer coordinates decode like PDBChain::ICToCoord"
    assert np.array_equal(d["prof"], g["prof"]) and np.array_equal(d["mu"], g["mu"])
    r = _run("features", tmp_path / "missing.bca", tmp_path / "x.rskc")
    assert r.returncode == 1 and "---Fatal error---" in r.stderr
    (tmp_path / "bad.bca").write_bytes(b"\0" * 64)
    r = _run("features", tmp_path / "bad.bca", tmp_path / "x.rskc")
    assert r.returncode == 1 and "Bad magic" in r.stderr  # bcadata.cpp:73-75


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fast", "sensitive", "verysensitive"])
def test_selfsearch_matches_reference_binary(built_lib, tmp_path, mode):
    g6, g21 = golden_bca(tmp_path)
    r = _run("selfsearch", mode, g21, tmp_path / "out.tsv", SEARCH_COLUMNS)
    assert r.returncode == 0, r.stderr
    got = sorted((tmp_path / "out.tsv").read_text().splitlines())
    want = _golden(f"golden_search_self_{mode}.tsv")
    assert len(want) > 30 and got == want


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["sensitive", "verysensitive"])
def test_search_db_matches_reference_binary(built_lib, tmp_path, mode):
    g6, g21 = golden_bca(tmp_path)
    r = _run("search", mode, g6, g21, tmp_path / "out.tsv", SEARCH_COLUMNS)
    assert r.returncode == 0, r.stderr
    got = sorted((tmp_path / "out.tsv").read_text().splitlines())
    want = _golden(f"golden_search_db_{mode}.tsv")
    assert len(want) >= 8 and got == want


@pytest.mark.gpu
def test_search_fast_db_matches_reference_binary(built_lib, tmp_path):
    g6, g21 = golden_bca(tmp_path)
    r = _run("searchfast", g6, g21, tmp_path / "cands.tsv", tmp_path / "out.tsv", SEARCH_COLUMNS)
    assert r.returncode == 0, r.stderr
    got = sorted((tmp_path / "out.tsv").read_text().splitlines())
    want = _golden("golden_search_fastdb.tsv")
    assert len(want) >= 6 and got == want


@pytest.mark.gpu
def test_reseek_command_lines_on_two_gpus_match_reference_binary(built_lib, tmp_path):
    """`-gpus 2` (DBSearcher::m_GpuCount: rsk_comm_create_all, rows / DB blocks sharded, hits gathered over NCCL) through the
    reference's own command lines: the sorted output equals the reference binary's, as on one GPU."""
    import reseek_b200 as rb
    if rb.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2); the one-GPU forms of the same commands run above")
    g6, g21 = golden_bca(tmp_path)
    out = tmp_path / "out.tsv"
    for args, golden in (
            (["-search", g21, "-fast"], "golden_search_self_fast.tsv"),
            (["-search", g21, "-sensitive"], "golden_search_self_sensitive.tsv"),
            (["-search", g6, "-db", g21, "-sensitive"], "golden_search_db_sensitive.tsv"),
            (["-search", g6, "-db", g21, "-fast"], "golden_search_fastdb.tsv")):
        r = _run(*args, "-gpus", 2, "-output", out, "-columns", SEARCH_COLUMNS)
        assert r.returncode == 0, r.stderr
        assert sorted(out.read_text().splitlines()) == _golden(golden), " ".join(map(str, args))


@pytest.mark.gpu
def test_selfsearch_aln_fasta2_and_row_columns_match_reference_binary(built_lib, tmp_path):
    """-aln, -fasta2 and the row columns (qrow, trow, qrowg, trowg, muscore, ...) of a whole `-search X -sensitive` run:
    identical to the reference binary's files (tools/make_golden_aln.py), hits in canonical order."""
    g4, gs = golden_bca_short(tmp_path)
    r = _run("selfsearch", "sensitive", gs, tmp_path / "out.tsv", ALN_COLUMNS, RSK_ALN=tmp_path / "out.aln", RSK_FASTA2=tmp_path / "out.fa2")
    assert r.returncode == 0, r.stderr
    assert sorted((tmp_path / "out.tsv").read_text().splitlines()) == _golden("golden_aln_self_sensitive.tsv")
    assert aln_blocks((tmp_path / "out.aln").read_text()) == aln_blocks((GOLDEN / "golden_aln_self_sensitive.aln").read_text())
    assert fasta2_records((tmp_path / "out.fa2").read_text()) == fasta2_records((GOLDEN / "golden_aln_self_sensitive.fa2").read_text())


@pytest.mark.gpu
def test_search_db_aln_rowlen_unaligned_match_reference_binary(built_lib, tmp_path):
    """`-search g4 -db gshort -verysensitive -aln -fasta2 -unaligned -rowlen 60` through the streamed-DB path."""
    g4, gs = golden_bca_short(tmp_path)
    r = _run("search", "verysensitive", g4, gs, tmp_path / "out.tsv", SEARCH_COLUMNS, RSK_ALN=tmp_path / "out.aln",
             RSK_FASTA2=tmp_path / "out.fa2", RSK_UNALIGNED=1, RSK_ROWLEN=60)
    assert r.returncode == 0, r.stderr
    got = aln_blocks((tmp_path / "out.aln").read_text())
    assert len(got) >= 50 and got == aln_blocks((GOLDEN / "golden_aln_db_verysensitive.aln").read_text())
    assert fasta2_records((tmp_path / "out.fa2").read_text()) == fasta2_records((GOLDEN / "golden_aln_db_verysensitive.fa2").read_text())


@pytest.mark.gpu
def test_selfsearch_global_matches_reference_binary(built_lib, tmp_path):
    """`-search gshort.bca -global -verysensitive` (runself.cpp:48-57): global score, coordinates as the reference prints
    them on this path (hi unset) and CIGAR, line for line."""
    g4, gs = golden_bca_short(tmp_path)
    r = _run("selfsearch", "verysensitive", gs, tmp_path / "out.tsv", GLOBAL_COLUMNS, RSK_GLOBAL=1)
    assert r.returncode == 0, r.stderr
    got = sorted((tmp_path / "out.tsv").read_text().splitlines())
    want = _golden("golden_global_self.tsv")
    assert len(want) == 196 and got == want


@pytest.mark.gpu
def test_file_name_signatures_of_the_fast_db_stages(built_lib, tmp_path):
    """MuPreFilter / PostMuFilter called with file names (PostMuFilter's reference argument list, search.cpp:14-18) give
    the candidate TSV and the hit lines of the in-memory versions."""
    g6, g21 = golden_bca(tmp_path)
    a = _run("searchfast", g6, g21, tmp_path / "c1.tsv", tmp_path / "o1.tsv")
    b = _run("searchfastfiles", g6, g21, tmp_path / "c2.tsv", tmp_path / "o2.tsv")
    assert a.returncode == 0 and b.returncode == 0, a.stderr + b.stderr
    assert (tmp_path / "c1.tsv").read_text() == (tmp_path / "c2.tsv").read_text()
    lines = (tmp_path / "o1.tsv").read_text()
    assert len(lines.splitlines()) >= 6 and lines == (tmp_path / "o2.tsv").read_text()


def _cli(*args):
    return subprocess.run([str(DEMO), *map(str, args)], capture_output=True, text=True, timeout=900,
                          env=dict(os.environ, RSK_BLOCK_CHAINS="8"))


@pytest.mark.gpu
def test_reseek_command_line_is_a_drop_in(built_lib, tmp_path):
    """The reference's own command lines (`reseek -search ...`, search.cpp:20-111) given to rsk_host_demo: self search,
    streamed -db search, -fast -db, -noself/-evalue, -global, -aln/-fasta2 — outputs equal the reference binary's."""
    from tests.golden_util import NOSELF_COLUMNS
    g6, g21 = golden_bca(tmp_path)
    g4, gs = golden_bca_short(tmp_path)
    out = tmp_path / "o.tsv"

    def lines():
        return sorted(out.read_text().splitlines())
    r = _cli("-search", g21, "-sensitive", "-output", out, "-columns", SEARCH_COLUMNS, "-threads", 1)
    assert r.returncode == 0, r.stderr
    assert lines() == _golden("golden_search_self_sensitive.tsv")
    r = _cli("-search", g6, "-db", g21, "-verysensitive", "-output", out, "-columns", SEARCH_COLUMNS)
    assert r.returncode == 0, r.stderr
    assert lines() == _golden("golden_search_db_verysensitive.tsv")
    r = _cli("-search", g6, "-db", g21, "-fast", "-output", out, "-columns", SEARCH_COLUMNS)
    assert r.returncode == 0, r.stderr
    assert lines() == _golden("golden_search_fastdb.tsv")
    r = _cli("-search", g21, "-sensitive", "-noself", "-evalue", 1, "-output", out, "-columns", NOSELF_COLUMNS)
    assert r.returncode == 0, r.stderr
    want = _golden("golden_search_self_noself_evalue.tsv")
    assert len(want) >= 2 and lines() == want
    r = _cli("-search", gs, "-global", "-verysensitive", "-output", out, "-columns", GLOBAL_COLUMNS)
    assert r.returncode == 0, r.stderr
    assert lines() == _golden("golden_global_self.tsv")
    r = _cli("-search", g4, "-db", gs, "-verysensitive", "-output", out, "-aln", tmp_path / "o.aln", "-fasta2", tmp_path / "o.fa2",
             "-unaligned", "-rowlen", 60)
    assert r.returncode == 0, r.stderr
    assert aln_blocks((tmp_path / "o.aln").read_text()) == aln_blocks((GOLDEN / "golden_aln_db_verysensitive.aln").read_text())
    assert fasta2_records((tmp_path / "o.fa2").read_text()) == fasta2_records((GOLDEN / "golden_aln_db_verysensitive.fa2").read_text())
    r = _cli("-search", g21, "-output", out)
    assert r.returncode == 1 and "Must set -fast, -sensitive or -verysensitive" in r.stderr


@pytest.mark.gpu
def test_alignpair_command_line_matches_reference_binary(built_lib, tmp_path):
    """`reseek -alignpair Q.bca -input2 T.bca -aln F [-global]` (cmd_alignpair, alignpair.cpp:164-228): the best pair of two
    chain files and its alignment block, byte for byte (tools/make_golden_aln.py)."""
    from tests.golden_util import golden_bca_disjoint
    q4, gsx = golden_bca_disjoint(tmp_path)
    for name, extra in (("golden_alignpair.aln", []), ("golden_alignpair_global.aln", ["-global"])):
        r = _cli("-alignpair", q4, "-input2", gsx, "-aln", tmp_path / "ap.aln", *extra)
        assert r.returncode == 0, r.stderr
        assert (tmp_path / "ap.aln").read_text() == (GOLDEN / name).read_text(), name
    r = _cli("-alignpair", q4)
    assert r.returncode == 1 and "Must specify -input2" in r.stderr


@pytest.mark.gpu
def test_history_dependent_pairs_of_the_reference(built_lib, tmp_path):
    """A handful of the SCOP40 chain pairs on which the reference's all-vs-all output depends on what its aligner did before
    (uncleared x-drop trace matrix, xdpmem.h:96-107; tools/check_history_pairs.py, profiles/r2_history_pairs.md): the lines the
    reference prints for each pair ALONE (fixture) are the lines this engine prints."""
    from reseek_b200 import chainio
    g = np.load(GOLDEN / "history_pairs.npz", allow_pickle=True)
    lens = g["lens"].astype(np.int64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    cols = str(g["columns"])
    assert len(g["lines"]) >= 3
    for k, members in enumerate(g["members"]):
        ids = [int(m) for m in members if m >= 0]
        chainio.write_bca(tmp_path / "pair.bca", [str(g["labels"][i]) for i in ids],
                          [bytes(g["seq"][off[i]:off[i + 1]]) for i in ids], [g["xyz"][:, off[i]:off[i + 1]] for i in ids])
        out = tmp_path / "pair.tsv"
        r = _cli("-search", tmp_path / "pair.bca", "-fast", "-output", out, "-columns", cols)
        assert r.returncode == 0, r.stderr
        assert sorted(out.read_text().splitlines()) == sorted(str(g["lines"][k]).splitlines()), f"pair {k}"


def test_bca_per_rank_ranges(built_lib, tmp_path):
    """ChainReader2::OpenRange: every rank of a multi-process run reads its own contiguous, residue-balanced block of a shared
    .bca straight from the file's length table (no GPU involved); the blocks tile the file and equal rsk_partition_by_residues."""
    import reseek_b200 as rb
    g6, g21 = golden_bca(tmp_path)
    g = np.load(GOLDEN / "golden_chains.npz", allow_pickle=True)
    lens = g["lens"]
    for nranks in (1, 2, 5, 30):
        want = rb.partition_by_residues(lens, nranks)
        for r in range(nranks):
            out = _run("bcarange", g21, r, nranks)
            assert out.returncode == 0, out.stderr
            lo, hi, first, last, res = out.stdout.split() if want[r][1] > want[r][0] else (out.stdout.split() + ["", ""])[:2] + ["", "", "0"]
            assert (int(lo), int(hi)) == want[r]
            if want[r][1] > want[r][0]:
                assert first == str(g["labels"][want[r][0]]) and last == str(g["labels"][want[r][1] - 1])
                assert int(res) == int(lens[want[r][0]:want[r][1]].sum())
