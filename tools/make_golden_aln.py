#!/usr/bin/env python3
"""Golden -aln / -fasta2 files and row columns of the reference BINARY (build container only; output committed).

    -search gshort.bca -sensitive -columns <ALN_COLUMNS> -aln -fasta2        -> golden_aln_self_sensitive.{tsv,aln,fa2}
    -search g4.bca -db gshort.bca -verysensitive -aln -fasta2 -unaligned -rowlen 60
                                                                             -> golden_aln_db_verysensitive.{aln,fa2}
    -alignpair g4.bca -input2 gsx.bca [-global] -aln                          -> golden_alignpair[_global].aln
    -search g21.bca -sensitive -noself -evalue 1                             -> golden_search_self_noself_evalue.tsv

gshort.bca holds the golden chains shorter than 500 residues, g4.bca four of them (tests/golden_util.golden_bca_short).  Hits are put into a
canonical order (sorted lines / blocks / records) because the reference's order depends on thread timing.
tests/test_host_search.py runs reseek_b200/rsk_host_demo on the same files and compares byte for byte."""
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tests.golden_util import ALN_COLUMNS, NOSELF_COLUMNS, ALN_RULE, GOLDEN, aln_blocks, fasta2_records, golden_bca, golden_bca_disjoint, golden_bca_short  # noqa: E402

REF = ROOT / "oracle" / "_ref" / "reseek_ref"


def run(args):
    cmd = [str(REF)] + [str(a) for a in args] + ["-threads", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise SystemExit(f"{' '.join(cmd)} failed:\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")


def canon_aln(path):
    return "".join("\n" + ALN_RULE + "\n" + b for b in aln_blocks(Path(path).read_text()))


def canon_fa2(path):
    return "".join(r + "\n\n" for r in fasta2_records(Path(path).read_text()))


def main():
    with tempfile.TemporaryDirectory() as t:
        tmp = Path(t)
        g4, gs = golden_bca_short(tmp)
        run(["-search", gs, "-sensitive", "-output", tmp / "o.tsv", "-columns", ALN_COLUMNS, "-aln", tmp / "o.aln",
             "-fasta2", tmp / "o.fa2"])
        lines = sorted((tmp / "o.tsv").read_text().splitlines())
        (GOLDEN / "golden_aln_self_sensitive.tsv").write_text("\n".join(lines) + "\n")
        (GOLDEN / "golden_aln_self_sensitive.aln").write_text(canon_aln(tmp / "o.aln"))
        (GOLDEN / "golden_aln_self_sensitive.fa2").write_text(canon_fa2(tmp / "o.fa2"))
        print("self sensitive", len(lines))
        run(["-search", g4, "-db", gs, "-verysensitive", "-output", tmp / "o2.tsv", "-aln", tmp / "o2.aln", "-fasta2",
             tmp / "o2.fa2", "-unaligned", "-rowlen", "60"])
        (GOLDEN / "golden_aln_db_verysensitive.aln").write_text(canon_aln(tmp / "o2.aln"))
        (GOLDEN / "golden_aln_db_verysensitive.fa2").write_text(canon_fa2(tmp / "o2.fa2"))
        print("db verysensitive", len((tmp / "o2.tsv").read_text().splitlines()))
        # -alignpair (cmd_alignpair, alignpair.cpp:164-228) on two files without a common chain, local and -global
        q4, gsx = golden_bca_disjoint(tmp)
        for name, extra in (("golden_alignpair.aln", []), ("golden_alignpair_global.aln", ["-global"])):
            r = subprocess.run([str(REF), "-alignpair", str(q4), "-input2", str(gsx), "-aln", str(tmp / "ap.aln")] + extra,
                               capture_output=True, text=True)
            if r.returncode != 0:
                raise SystemExit(r.stdout[-2000:] + r.stderr[-2000:])
            (GOLDEN / name).write_text((tmp / "ap.aln").read_text())
            print(name, len((tmp / "ap.aln").read_text().splitlines()), "lines")
        # -noself and -evalue (dssaligner.cpp:1020-1021, runself.cpp:39-40, dbsearcher.cpp:75-76)
        run(["-search", golden_bca(tmp)[1], "-sensitive", "-noself", "-evalue", "1", "-output", tmp / "o3.tsv", "-columns", NOSELF_COLUMNS])
        lines = sorted((tmp / "o3.tsv").read_text().splitlines())
        (GOLDEN / "golden_search_self_noself_evalue.tsv").write_text("\n".join(lines) + "\n")
        print("self -noself -evalue 1", len(lines))


if __name__ == "__main__":
    main()
