#!/usr/bin/env python3
"""Golden outputs of the reference BINARY for whole searches from .bca files (build container only; output committed).

The 21 real chains of tests/golden/golden_chains.npz are written as .bca files with our own writer
(reseek_b200.chainio.write_bca, checked to round-trip through the reference's reader) and the unmodified reference
(oracle/_ref/reseek_ref, -threads 1) is run on them:

    -search g21.bca                      -fast / -sensitive / -verysensitive     -> golden_search_self_<mode>.tsv
    -search g6.bca -db g21.bca           -sensitive / -verysensitive             -> golden_search_db_<mode>.tsv
    -search g6.bca -db g21.bca -fast     (prefilter + post-filter)               -> golden_search_fastdb.tsv

Lines are sorted (the reference's order depends on thread timing).  tests/test_host_search.py re-creates the .bca files and
compares the output of reseek_b200/rsk_host_demo (DSS look-alike -> GPU search -> hit writer) line by line."""
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from reseek_b200 import chainio  # noqa: E402

REF = ROOT / "oracle" / "_ref" / "reseek_ref"
GOLD = ROOT / "tests" / "golden"
COLUMNS = "query+target+dpscore+newts+lddt+evalue+pvalue+qlo+qhi+ql+tlo+thi+tl+ids+gaps+pctid+cigar"


def golden_bca(tmp):
    g = np.load(GOLD / "golden_chains.npz", allow_pickle=True)
    lens = g["lens"].astype(np.int64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    labels = [str(x) for x in g["labels"]]
    seqs = [bytes(g["seq"][int(off[i]):int(off[i + 1])]) for i in range(len(lens))]
    xyzs = [g["xyz"][:, int(off[i]):int(off[i + 1])] for i in range(len(lens))]
    pick = [0, 3, 8, 11, 14, 19]  # short, medium and two long (>= MKFL) chains
    chainio.write_bca(tmp / "g21.bca", labels, seqs, xyzs)
    chainio.write_bca(tmp / "g6.bca", [labels[i] for i in pick], [seqs[i] for i in pick], [xyzs[i] for i in pick])
    return tmp / "g6.bca", tmp / "g21.bca"


def run(args, out):
    cmd = [str(REF)] + [str(a) for a in args] + ["-output", str(out), "-columns", COLUMNS, "-threads", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise SystemExit(f"{' '.join(cmd)} failed:\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    return sorted(Path(out).read_text().splitlines())


def main():
    with tempfile.TemporaryDirectory() as t:
        tmp = Path(t)
        g6, g21 = golden_bca(tmp)
        for mode in ("fast", "sensitive", "verysensitive"):
            lines = run(["-search", g21, f"-{mode}"], tmp / "o.tsv")
            (GOLD / f"golden_search_self_{mode}.tsv").write_text("\n".join(lines) + "\n")
            print("self", mode, len(lines))
        for mode in ("sensitive", "verysensitive"):
            lines = run(["-search", g6, "-db", g21, f"-{mode}"], tmp / "o.tsv")
            (GOLD / f"golden_search_db_{mode}.tsv").write_text("\n".join(lines) + "\n")
            print("db", mode, len(lines))
        lines = run(["-search", g6, "-db", g21, "-fast"], tmp / "o.tsv")
        (GOLD / "golden_search_fastdb.tsv").write_text("\n".join(lines) + "\n")
        print("fastdb", len(lines))


if __name__ == "__main__":
    main()
