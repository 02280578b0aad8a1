"""The host C++ look-alikes (reseek_b200/csrc/host: DSSAligner, DBSearcher, MuPreFilter/PostMuFilter) over the C ABI.

CPU part: the shim library builds, exports the reference's class surface and dies loudly (Die(): message + exit 1) without
a GPU.  GPU part: the harness drives RunSelf / RunQuery / AlignQueryTarget / the -fast -db pair of functions and its TSV
output must equal, byte for byte, the lines formatted from the C-ABI results that the other parity tests pin to the oracle."""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
DEMO = ROOT / "reseek_b200" / "rsk_host_demo"
HOSTLIB = ROOT / "reseek_b200" / "libreseek_b200_host.so"


def _sets(tmp_path, with_long=False):
    from reseek_b200 import synth, chainio
    q = synth.make_chains(9, 110, seed=41, length_jitter=0.4)
    db = synth.make_chains(23, 140, seed=42, length_jitter=0.5)
    synth.plant_homologs(db, q, 0.5, seed=43)
    lq, sq = chainio.write_rskc(tmp_path / "q.rskc", q, chainio.default_labels(q.n, "qry"))
    ld, sd = chainio.write_rskc(tmp_path / "db.rskc", db, chainio.default_labels(db.n, "dbc"))
    return q, db, lq, ld, sq, sd


def _seqs(chains, seq):
    return [bytes(seq[int(chains.off[i]):int(chains.off[i + 1])]) for i in range(chains.n)]


def _run(*args):
    env = dict(os.environ, RSK_BLOCK_CHAINS="7")  # streamed side in several blocks
    return subprocess.run([str(DEMO), *map(str, args)], capture_output=True, text=True, timeout=600, env=env)


def test_host_library_exports_reference_surface(built_lib):
    assert HOSTLIB.exists() and DEMO.exists()
    syms = subprocess.run(["nm", "-DC", "--defined-only", str(HOSTLIB)], capture_output=True, text=True).stdout
    for name in ("reseek_b200::DSSAligner::SetParams(", "reseek_b200::DSSAligner::SetQuery(", "reseek_b200::DSSAligner::SetTarget(",
                 "reseek_b200::DSSAligner::UnsetQuery(", "reseek_b200::DSSAligner::AlignQueryTarget(", "reseek_b200::DSSAligner::Align_NoAccel(",
                 "reseek_b200::DSSAligner::ToTsv(", "reseek_b200::DSSAligner::DoMKF(", "reseek_b200::DSSAligner::ClearAlign(",
                 "reseek_b200::DSSAligner::ToAln(", "reseek_b200::DSSAligner::ToFasta2(", "reseek_b200::DSSAligner::GetKabsch(",
                 "reseek_b200::DSSAligner::AlignBags(", "reseek_b200::DSSAligner::AlignQueryTarget_Global(",
                 "reseek_b200::DSSAligner::Stats(", "reseek_b200::DBSearcher::Setup(", "reseek_b200::DBSearcher::RunSelf(",
                 "reseek_b200::DBSearcher::RunQuery(", "reseek_b200::DBSearcher::BaseOnAln(", "reseek_b200::DBSearcher::Reject(",
                 "reseek_b200::DBSearcher::AddChain(", "reseek_b200::MuPreFilter(", "reseek_b200::PostMuFilter(",
                 "reseek_b200::GetSelfRevScore(", "reseek_b200::DSSParams::SetDSSParams(", "reseek_b200::Die("):
        assert name in syms, f"libreseek_b200_host.so does not define {name}"


def test_host_shim_dies_without_gpu(built_lib, tmp_path):
    import reseek_b200 as rb
    if rb.device_count() > 0:
        pytest.skip("a GPU is present")
    _sets(tmp_path)
    r = _run("pair", "verysensitive", tmp_path / "q.rskc", 0, 1, tmp_path / "o.tsv")
    assert r.returncode == 1 and "---Fatal error---" in r.stderr and "no CUDA device" in r.stderr  # Die(), myutils.cpp:785
    r = _run("self", "bogusmode", tmp_path / "q.rskc", tmp_path / "o.tsv")
    assert r.returncode == 1 and "Must set -fast, -sensitive or -verysensitive" in r.stderr  # dssparams.cpp:90


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["verysensitive", "sensitive", "fast"])
def test_runself_matches_cabi(built_lib, tmp_path, mode):
    import reseek_b200 as rb
    q, db, lq, ld, sq, sd = _sets(tmp_path)
    r = _run("self", mode, tmp_path / "db.rskc", tmp_path / "self.tsv")
    assert r.returncode == 0, r.stderr
    got = (tmp_path / "self.tsv").read_text().splitlines()
    ctx = rb.Context(0, {"fast": 1, "sensitive": 2, "verysensitive": 3}[mode])
    S = ctx.upload(db.lens, db.prof, db.mu, db.xyz, db.selfrev)
    # DBSearcher::RunSelf goes through the sharded entry (hits compacted on the device, returned in (a, b) order), also on one GPU
    res = ctx.search_self_sharded(None, S, keep=rb.KEEP_HITS, want_paths=True)
    seqs = _seqs(db, sd)
    want = []
    for k, h in enumerate(res.hits):
        a, b = int(h["a"]), int(h["b"])
        for up in ([True, False] if a != b else [True]):  # runself.cpp:60-66
            want.append(rb.format_tsv(h, res.path(k), ld[a], ld[b], db.lens[a], db.lens[b], up=up, seq_a=seqs[a], seq_b=seqs[b]))
    assert len(want) > 0 and got == want
    assert f"hits {len(want)}" in r.stderr and f"OnAln calls {len(want)}" in r.stderr
    ctx.close()


@pytest.mark.gpu
def test_runquery_matches_cabi(built_lib, tmp_path):
    import reseek_b200 as rb
    q, db, lq, ld, sq, sd = _sets(tmp_path)
    # stream = db.rskc (DSSAligner query slot A), in-memory set = q.rskc (slot B); hits are written with Up = false
    r = _run("query", "sensitive", tmp_path / "db.rskc", tmp_path / "q.rskc", tmp_path / "query.tsv")
    assert r.returncode == 0, r.stderr
    got = (tmp_path / "query.tsv").read_text().splitlines()
    ctx = rb.Context(0, rb.MODE_SENSITIVE)
    A = ctx.upload(db.lens, db.prof, db.mu, db.xyz, db.selfrev)
    B = ctx.upload(q.lens, q.prof, q.mu, q.xyz, q.selfrev)
    res = ctx.search_cross(A, B, keep=rb.KEEP_HITS, want_paths=True)
    sa, sb = _seqs(db, sd), _seqs(q, sq)
    want = [rb.format_tsv(h, res.path(k), ld[int(h["a"])], lq[int(h["b"])], db.lens[int(h["a"])], q.lens[int(h["b"])], up=False,
                          seq_a=sa[int(h["a"])], seq_b=sb[int(h["b"])]) for k, h in enumerate(res.hits)]
    assert len(want) > 0 and got == want
    ctx.close()


@pytest.mark.gpu
def test_alignquerytarget_batch_of_one(built_lib, tmp_path, port):
    import reseek_b200 as rb
    from tests.util import to_oracle_chains
    q, db, lq, ld, sq, sd = _sets(tmp_path)
    oc = to_oracle_chains(db)
    seqs = _seqs(db, sd)
    ctx = rb.Context(0, rb.MODE_VERYSENSITIVE)
    S = ctx.upload(db.lens, db.prof, db.mu, db.xyz, db.selfrev)
    for i, j in ((0, 1), (3, 3), (5, 2)):
        r = _run("pair", "verysensitive", tmp_path / "db.rskc", i, j, tmp_path / "pair.tsv")
        assert r.returncode == 0, r.stderr
        got = (tmp_path / "pair.tsv").read_text().splitlines()
        res = ctx.search_pairs(S, S, np.array([i], np.uint32), np.array([j], np.uint32), keep=rb.KEEP_ALL)
        h = res.hits[0]
        orc, opath = port(3).align_pair(oc[i], oc[j])
        assert res.path(0) == opath and np.float32(h["score"]).view(np.uint32) == np.float32(orc.score).view(np.uint32)
        want = [rb.format_tsv(h, res.path(0), ld[i], ld[j], db.lens[i], db.lens[j], up=True, seq_a=seqs[i], seq_b=seqs[j])] if opath else []
        assert got == want
    ctx.close()


@pytest.mark.gpu
def test_fastdb_prefilter_postfilter_matches_cabi(built_lib, tmp_path):
    import reseek_b200 as rb
    q, db, lq, ld, sq, sd = _sets(tmp_path)
    r = _run("fastdb", tmp_path / "q.rskc", tmp_path / "db.rskc", tmp_path / "cands.tsv", tmp_path / "hits.tsv")
    assert r.returncode == 0, r.stderr
    ctx = rb.Context(0, rb.MODE_FAST)
    Q = ctx.upload(q.lens, q.prof, q.mu, q.xyz, q.selfrev)
    T = ctx.upload(db.lens, db.prof, db.mu, db.xyz, db.selfrev)
    pf = ctx.prefilter(Q, T)
    assert (tmp_path / "cands.tsv").read_text() == pf.tsv()
    res = ctx.postfilter(Q, T, pf, keep=rb.KEEP_ALL, want_paths=True)
    sa, sb = _seqs(q, sq), _seqs(db, sd)
    want = [rb.format_tsv(h, res.path(k), lq[int(h["a"])], ld[int(h["b"])], q.lens[int(h["a"])], db.lens[int(h["b"])], up=True,
                          seq_a=sa[int(h["a"])], seq_b=sb[int(h["b"])])
            for k, h in enumerate(res.hits) if float(h["evalue"]) <= 10]
    assert (tmp_path / "hits.tsv").read_text().splitlines() == want
    ctx.close()


@pytest.mark.gpu
def test_alignquerytarget_global_batch_of_one(built_lib, tmp_path, port):
    """DSSAligner::AlignQueryTarget_Global of the look-alike (alignpair.cpp:110-114 with -global): gscore, dpscore (stays 0
    on this path) and CIGAR as the reference's writer prints them, from the oracle's global alignment."""
    import reseek_b200 as rb
    from tests.util import to_oracle_chains
    q, db, lq, ld, sq, sd = _sets(tmp_path)
    oc = to_oracle_chains(db)
    for i, j in ((0, 1), (3, 3), (5, 2)):
        r = _run("pairglobal", "verysensitive", tmp_path / "db.rskc", i, j, tmp_path / "g.tsv")
        assert r.returncode == 0, r.stderr
        orc, opath = port(3).align_pair_global(oc[i], oc[j])
        want = "\t".join([ld[i], ld[j], "%.1f" % orc.score, "0", rb.path_to_cigar(opath, up=True)])
        assert (tmp_path / "g.tsv").read_text().splitlines() == [want]


@pytest.mark.gpu
def test_alignbags_equals_alignquerytarget(built_lib, tmp_path):
    """DSSAligner::AlignBags (chainbag.cpp:44-84, what PostMuFilter calls per candidate) gives the record AlignQueryTarget
    gives for the same two chains."""
    _sets(tmp_path)
    for mode in ("sensitive", "verysensitive"):
        for i, j in ((0, 1), (3, 3), (5, 2), (2, 7)):
            a = _run("pair", mode, tmp_path / "db.rskc", i, j, tmp_path / "p.tsv")
            b = _run("pairbags", mode, tmp_path / "db.rskc", i, j, tmp_path / "b.tsv")
            assert a.returncode == 0 and b.returncode == 0, a.stderr + b.stderr
            assert (tmp_path / "p.tsv").read_text() == (tmp_path / "b.tsv").read_text()
