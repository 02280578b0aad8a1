"""CPU: the C-ABI library loads without a GPU, exports every declared symbol and fails loudly on compute."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_exports_every_declared_symbol(built_lib):
    hdr = (ROOT / "include" / "reseek_b200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(rsk_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 25
    for n in names:
        assert hasattr(built_lib, n), f"libreseek_b200.so does not export {n}"


def test_presets_match_oracle(built_lib, port):
    import reseek_b200 as rb
    for mode in (1, 2, 3):
        p = rb.params_preset(mode)
        o = port(mode).params
        assert np.array_equal(np.array(p.tables[:], np.float32).view(np.uint32), np.array(o.tables[:], np.float32).view(np.uint32))
        for k in ("gap_open", "gap_ext", "min_fwd_score", "omega", "omega_fwd", "mu_gap_open", "mu_gap_ext", "mkfl",
                  "mkf_x1", "mkf_x2", "mkf_min_hsp_score", "mkf_min_mega_hsp_score"):
            assert getattr(p, k) == getattr(o, k), k
    assert rb.params_preset(3).max_evalue > 1e300 and rb.params_preset(2).max_evalue == 10
    with pytest.raises(rb.ReseekB200Error):
        rb.params_preset(7)


def test_statsig_host_functions(built_lib, port):
    o = port(3)
    for ts in (-0.3, 0.0, 0.05, 0.11, 0.2, 0.7, 1.5):
        pv, ev, q = o.statsig(ts)
        assert built_lib.rsk_pvalue(ts) == pv and built_lib.rsk_evalue(ts) == ev and built_lib.rsk_qual(ts) == q


def test_no_silent_cpu_fallback(built_lib):
    import reseek_b200 as rb
    if rb.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(rb.ReseekB200Error, match="no CUDA device"):
        rb.Context(0, rb.MODE_VERYSENSITIVE)


def test_product_never_imports_oracle():
    for f in (ROOT / "reseek_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".cuh", ".cpp", ".h", ".hpp") and f.is_file():
            txt = f.read_text(errors="ignore")
            assert "pyoracle" not in txt and "reseek_oracle" not in txt and "oracle/" not in txt.replace("oracle/_ref", ""), f


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_tsv_writer_matches_reference_lines(built_lib, mode):
    """rsk_format_tsv against lines printed by the reference's own DSSAligner::ToTsv (golden fixtures), both
    directions (Up / !Up: query/target swap and the D<->I swap of the CIGAR)."""
    import reseek_b200 as rb
    from tests.golden_util import load_chains, load_pairs
    g = load_pairs(mode)
    chains = load_chains()
    cols = str(g["tsv_cols"])
    names = cols.split("+")
    assert len(g["tsv_lines"]) >= 20
    for k, up, line in zip(g["tsv_k"], g["tsv_up"], g["tsv_lines"]):
        k = int(k)
        h = np.zeros(1, rb.HIT_DTYPE)[0]
        for f in ("score", "lo_a", "lo_b", "hi_a", "hi_b", "ids", "gaps", "lddt", "ts", "pvalue", "evalue", "qual"):
            h[f] = g[f][k]
        h["path_len"] = len(g["path_list"][k])
        mkf = bool(g["mkf"][k])
        if mkf:
            h["flags"] = rb.HIT_MKF
            h["mu_fwd"], h["mu_rev"] = g["best_hsp"][k], g["best_chain"][k]
        A, B = chains[int(g["a"][k])], chains[int(g["b"][k])]
        mine = rb.format_tsv(h, g["path_list"][k], A.label, B.label, A.L, B.L, up=bool(up), columns=cols, seq_a=A.seq, seq_b=B.seq)
        want, got = str(line).split("\t"), mine.split("\t")
        assert len(want) == len(got) == len(names)
        for nm, w, m in zip(names, want, got):
            if nm in ("muhsp", "muchain") and not mkf:
                continue  # the reference prints whatever the previous long-chain alignment left in m_MKF
            assert w == m, f"pair {k} up={up} column {nm}: reference '{w}' vs '{m}'"
    assert rb.path_to_cigar("MMDDIM", up=True) == "2M2D1I1M" and rb.path_to_cigar("MMDDIM", up=False) == "2M2I1D1M"


def _oracle_self_hits(mode, max_len=500):
    """(hit record, path, A, B) for every pair (i <= j) of the short golden chains the reference's RunSelf would report
    (E <= 10, dbsearcher.cpp:258-265), computed by the CPU oracle."""
    import reseek_b200 as rb
    from oracle.pyoracle import Port
    from tests.golden_util import load_chains
    chains = [c for c in load_chains() if c.L < max_len]
    port = Port(mode=mode)
    out = []
    for i in range(len(chains)):
        for j in range(i, len(chains)):
            r, path = port.align_pair(chains[i], chains[j])
            if r.filtered or r.path_len == 0 or not r.evalue <= 10:
                continue
            h = np.zeros(1, rb.HIT_DTYPE)[0]
            for f in ("score", "lo_a", "lo_b", "hi_a", "hi_b", "ids", "gaps", "lddt", "ts", "pvalue", "evalue", "qual",
                      "mu_score", "mu_fwd", "mu_rev", "path_len"):
                h[f] = getattr(r, f)
            out.append((h, path, chains[i], chains[j]))
    return out


def test_aln_fasta2_and_row_columns_match_reference_files(built_lib):
    """rsk_format_aln / rsk_format_fasta2 / the qrow, trow, qrowg, trowg, ts, muscore, aq, qcovpct, tcovpct columns against the
    files the reference binary wrote for `-search gshort.bca -sensitive -aln -fasta2 -columns ...` (tools/make_golden_aln.py);
    the alignments themselves come from the CPU oracle here (the GPU path writes the same files in tests/test_host_search.py).
    Both directions of every off-diagonal hit, as runself.cpp:60-66 emits them."""
    import reseek_b200 as rb
    from tests.golden_util import ALN_COLUMNS, GOLDEN, aln_blocks, fasta2_records
    alns, recs, lines = [], [], []
    for h, path, A, B in _oracle_self_hits(2):
        for up in ([True] if A is B else [True, False]):
            alns.append(rb.format_aln(h, path, A.label, B.label, A.seq, B.seq, up=up))
            recs.append(rb.format_fasta2(h, path, A.label, B.label, A.seq, B.seq, up=up))
            lines.append(rb.format_tsv(h, path, A.label, B.label, A.L, B.L, up=up, columns=ALN_COLUMNS, seq_a=A.seq, seq_b=B.seq))
    assert len(lines) >= 18
    assert sorted(lines) == (GOLDEN / "golden_aln_self_sensitive.tsv").read_text().splitlines()
    assert aln_blocks("".join(alns)) == aln_blocks((GOLDEN / "golden_aln_self_sensitive.aln").read_text())
    assert fasta2_records("".join(recs)) == fasta2_records((GOLDEN / "golden_aln_self_sensitive.fa2").read_text())


def test_aln_writer_rowlen_and_unaligned_flanks(built_lib):
    """-rowlen 60 blocks and -unaligned rows (lower-case flanks, '.' padding) against the reference binary's files for
    `-search g4.bca -db gshort.bca -verysensitive`."""
    import reseek_b200 as rb
    from oracle.pyoracle import Port
    from tests.golden_util import GOLDEN, SHORT_QUERIES, aln_blocks, fasta2_records, load_chains
    chains = load_chains()
    query = [chains[i] for i in SHORT_QUERIES]
    db = [c for c in chains if c.L < 500]
    port = Port(mode=3)
    alns, recs = [], []
    for q in query:
        for t in db:
            r, path = port.align_pair(q, t)
            if r.path_len == 0:
                continue
            h = np.zeros(1, rb.HIT_DTYPE)[0]
            for f in ("score", "lo_a", "lo_b", "hi_a", "hi_b", "ids", "gaps", "lddt", "ts", "pvalue", "evalue", "qual", "path_len"):
                h[f] = getattr(r, f)
            alns.append(rb.format_aln(h, path, q.label, t.label, q.seq, t.seq, up=True, rowlen=60))
            recs.append(rb.format_fasta2(h, path, q.label, t.label, q.seq, t.seq, up=True, unaligned=True))
    assert len(alns) >= 50
    assert aln_blocks("".join(alns)) == aln_blocks((GOLDEN / "golden_aln_db_verysensitive.aln").read_text())
    assert fasta2_records("".join(recs)) == fasta2_records((GOLDEN / "golden_aln_db_verysensitive.fa2").read_text())


def test_kabsch_matches_reference_superpositions(built_lib):
    """rsk_kabsch (Horn's quaternion method) against the reference's Kabsch() (tools/make_golden_kabsch.py): the same optimum,
    so rotation, translation and residual agree to rounding.  Tolerances: |du| <= 1e-8, |dt| <= 1e-6 A, residual within
    1e-6 * max(1, msd); the applied transform must also reproduce the residual it reports."""
    import reseek_b200 as rb
    from tests.golden_util import GOLDEN, load_chains
    g = np.load(GOLDEN / "golden_kabsch.npz")
    ch = load_chains()
    assert len(g["a"]) >= 100
    for k in range(len(g["a"])):
        A, B = ch[int(g["a"][k])], ch[int(g["b"][k])]
        path, up = str(g["paths"][k]), bool(g["up"][k])
        msd, t, u = rb.kabsch(A.xyz, B.xyz, g["lo_a"][k], g["lo_b"][k], path, up=up)
        assert np.abs(u - g["u"][k]).max() <= 1e-8, f"case {k} rotation"
        assert np.abs(t - g["t"][k]).max() <= 1e-6, f"case {k} translation"
        assert abs(msd - g["msd"][k]) <= 1e-6 * max(1.0, g["msd"][k]), f"case {k} residual"
        assert abs(np.linalg.det(u) - 1) < 1e-12
        i, j, X, Y = int(g["lo_a"][k]), int(g["lo_b"][k]), [], []
        for c in path:
            if c == "M":
                X.append(A.xyz[:, i]); Y.append(B.xyz[:, j]); i += 1; j += 1
            elif c == "D":
                i += 1
            else:
                j += 1
        X, Y = np.array(X, np.float64), np.array(Y, np.float64)
        if not up:
            X, Y = Y, X
        assert abs(((X @ u.T + t - Y) ** 2).sum() / len(X) - msd) <= 1e-9 * max(1.0, msd)
    with pytest.raises(rb.ReseekB200Error):
        rb.kabsch(ch[0].xyz, ch[1].xyz, 0, 0, "DDII")  # no aligned pair


def test_global_record_columns_match_reference_binary(built_lib):
    """A record of the -global path through rsk_format_tsv (gscore; dpscore stays 0; hi unset, so qhi/thi print 0; E-value
    unset prints 99.0) against the reference binary's `-search gshort.bca -global -verysensitive` output
    (tools/make_golden_global.py); the alignments come from the CPU oracle here, from the GPU in tests/test_host_search.py."""
    import reseek_b200 as rb
    from oracle.pyoracle import Port
    from tests.golden_util import GLOBAL_COLUMNS, GOLDEN, load_chains
    chains = [c for c in load_chains() if c.L < 500]
    port = Port(mode=3)
    FLT_MAX = np.finfo(np.float32).max
    lines = []
    for i in range(len(chains)):
        for j in range(i, len(chains)):
            r, path = port.align_pair_global(chains[i], chains[j])
            h = np.zeros(1, rb.HIT_DTYPE)[0]
            h["score"], h["lo_a"], h["lo_b"], h["path_len"], h["flags"] = r.score, 0, 0, len(path), rb.HIT_GLOBAL
            h["hi_a"] = h["hi_b"] = h["ids"] = h["gaps"] = 0xFFFFFFFF
            h["evalue"] = h["pvalue"] = h["qual"] = FLT_MAX
            h["ts"] = -FLT_MAX
            A, B = chains[i], chains[j]
            for up in ([True] if i == j else [True, False]):
                lines.append(rb.format_tsv(h, path, A.label, B.label, A.L, B.L, up=up, columns=GLOBAL_COLUMNS, seq_a=A.seq, seq_b=B.seq))
    assert sorted(lines) == (GOLDEN / "golden_global_self.tsv").read_text().splitlines()
