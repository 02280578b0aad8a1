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
