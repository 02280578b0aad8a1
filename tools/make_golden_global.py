#!/usr/bin/env python3
"""Golden vectors for -global: the reference's DSSAligner::AlignQueryTarget_Global (global.cpp:7-33 -> ViterbiFastMem +
TraceBackBitMem) run through oracle/_ref/libreseek_ref.so on pairs of the 21 real golden chains (build container only;
output committed as tests/golden/golden_global.npz).

Ordered pairs (a, b) with at least one chain of at most 450 residues, under the -verysensitive preset (no filter) and the
-sensitive preset (Mu filter first: a rejected pair keeps m_GlobalScore = -9999 and has no path).
Also the reference binary's own output of `-search gshort.bca -global -verysensitive` (golden_global_self.tsv)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.pyoracle import Ref  # noqa: E402
from tests.golden_util import GOLDEN, load_chains  # noqa: E402


def main():
    chains = load_chains()
    pairs = [(i, j) for i in range(len(chains)) for j in range(len(chains)) if not (chains[i].L > 450 and chains[j].L > 450)]
    out = {"a": np.array([p[0] for p in pairs], np.uint32), "b": np.array([p[1] for p in pairs], np.uint32)}
    for mode in (2, 3):
        ref = Ref(mode=mode)
        scores, paths, off = [], [], [0]
        for i, j in pairs:
            r, path = ref.align_pair(chains[i], chains[j], noaccel=2)
            assert r.path_len == len(path)
            scores.append(r.score)
            paths.append(path)
            off.append(off[-1] + len(path))
        out[f"score_mode{mode}"] = np.array(scores, np.float32)
        out[f"paths_mode{mode}"] = np.frombuffer("".join(paths).encode(), np.uint8)
        out[f"path_off_mode{mode}"] = np.array(off, np.uint64)
        print("mode", mode, "pairs", len(pairs), "with path", sum(1 for p in paths if p), "path bytes", off[-1])
    np.savez_compressed(GOLDEN / "golden_global.npz", **out)
    # the reference BINARY: `-search gshort.bca -global -verysensitive` (runself.cpp:48-57), sorted lines
    import subprocess
    import tempfile
    from tests.golden_util import GLOBAL_COLUMNS, golden_bca_short
    with tempfile.TemporaryDirectory() as t:
        g4, gs = golden_bca_short(t)
        cmd = [str(ROOT / "oracle" / "_ref" / "reseek_ref"), "-search", str(gs), "-global", "-verysensitive", "-output",
               f"{t}/o.tsv", "-columns", GLOBAL_COLUMNS, "-threads", "1"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit(r.stdout[-2000:] + r.stderr[-2000:])
        lines = sorted(Path(f"{t}/o.tsv").read_text().splitlines())
        (GOLDEN / "golden_global_self.tsv").write_text("\n".join(lines) + "\n")
        print("self -global lines", len(lines))


if __name__ == "__main__":
    main()
