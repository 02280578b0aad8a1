#!/usr/bin/env python3
"""The reference's own SCOP40 regression (test_scripts/scop40.bash + check_scop40.py) applied to THIS engine's output.

scop40.bash runs `reseek -search scop40.bca -db scop40.bca -output X.tsv -columns query+target+evalue` three times (-fast,
-sensitive, -fast -evalue 1); check_scop40.py evaluates the TSVs with test_scripts/scop40.py (level sf2, E-values) and requires
  fast      SEPQ0.1 >= 0.2100  SEPQ1 >= 0.3140  SEPQ10 >= 0.4200   (each -0.01)   check_scop40.py:49-51
  sensitive SEPQ0.1 >= 0.2170  SEPQ1 >= 0.3410  SEPQ10 >= 0.4740
  evalue1   SEPQ0.1 >= 0.2100  SEPQ1 >= 0.3100  SEPQ10 >= 0.3500
- the only known answers for the search path the reference ships.  Here the same three command lines are given to rsk_host_demo
on the GPU box (tools/scop40_regression.sh), the TSVs come back gzipped and are evaluated with the reference's scop40.py
imported from /root/reference/test_scripts (build container only; nothing of it is copied)."""
import gzip
import shutil
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REFSCRIPTS = Path("/root/reference/test_scripts")
WANT = {"fast": (0.2100, 0.3140, 0.4200), "sensitive": (0.2170, 0.3410, 0.4740), "evalue1": (0.2100, 0.3100, 0.3500)}


def main():
    src = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "gpurun_out" / "scop40_regression"
    sys.path.insert(0, str(REFSCRIPTS))
    import scop40  # the reference's own evaluation code
    sc = scop40.Scop40("e", "sf2", str(ROOT / "build" / "data" / "dom_scopid.tsv"))
    rows, errors = [], 0
    with tempfile.TemporaryDirectory() as tmp:
        for name, want in WANT.items():
            gz = src / f"scop40-{name}.tsv.gz"
            if not gz.exists():
                print("missing", gz)
                continue
            tsv = Path(tmp) / f"scop40-{name}.tsv"
            with gzip.open(gz, "rb") as f, open(tsv, "wb") as g:
                shutil.copyfileobj(f, g)
            nlines = sum(1 for _ in open(tsv))
            sc.eval_file(str(tsv), 0, 1, 2, False)
            got = (sc.tpr_at_fpepq0_1, sc.tpr_at_fpepq1, sc.tpr_at_fpepq10)
            ok = all(g - w >= -0.01 for g, w in zip(got, want))
            errors += not ok
            rows.append((name, nlines, got, want, ok))
            print(f"{name}: {nlines} lines  SEPQ0.1={got[0]:.4f}({got[0] - want[0]:+.4f}) SEPQ1={got[1]:.4f}({got[1] - want[1]:+.4f}) "
                  f"SEPQ10={got[2]:.4f}({got[2] - want[2]:+.4f})  {'PASSED' if ok else 'FAILED'}")
    out = ["# SCOP40 accuracy regression of the reference (test_scripts/check_scop40.py) on this engine\n",
           "`rsk_host_demo -search scop40.bca -db scop40.bca -columns query+target+evalue` with -fast / -sensitive / -fast -evalue 1 on one B200 "
           "(tools/scop40_regression.sh), evaluated with the reference's own `test_scripts/scop40.py` (level sf2, E-values; tools/check_scop40_sepq.py).\n",
           "| run | hit lines | SEPQ0.1 | SEPQ1 | SEPQ10 | reference thresholds (each -0.01) | result |", "|---|---|---|---|---|---|---|"]
    for name, n, got, want, ok in rows:
        out.append(f"| {name} | {n} | {got[0]:.4f} | {got[1]:.4f} | {got[2]:.4f} | {want[0]:.4f} / {want[1]:.4f} / {want[2]:.4f} | {'PASSED' if ok else 'FAILED'} |")
    (ROOT / "profiles" / "r2_scop40_sepq.md").write_text("\n".join(out) + "\n")
    return 1 if errors or not rows else 0


if __name__ == "__main__":
    sys.exit(main())
