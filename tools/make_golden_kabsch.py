#!/usr/bin/env python3
"""Golden vectors for the superposition: the reference's Kabsch(ChainA, ChainB, LoA, LoB, Path, t, u) (kabsch.cpp:330-387)
through oracle/_ref/libreseek_ref.so on -verysensitive alignments of the 21 real golden chains, both directions
(build container only; output committed as tests/golden/golden_kabsch.npz).  Alignments with fewer than three M columns are
left out: the rotation is not unique there."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.pyoracle import Ref  # noqa: E402
from tests.golden_util import GOLDEN, load_chains  # noqa: E402


def main():
    ch = load_chains()
    ref = Ref(mode=3)
    rec = []
    for i in range(len(ch)):
        for j in range(len(ch)):
            if (ch[i].L > 450 and ch[j].L > 450) or (i * 21 + j) % 7:
                continue
            r, path = ref.align_pair(ch[i], ch[j])
            if path.count("M") < 3:
                continue
            for up in (True, False):
                if up:
                    m, t, u = ref.kabsch(ch[i], ch[j], r.lo_a, r.lo_b, path)
                else:  # dssaligner.cpp:1380-1384
                    m, t, u = ref.kabsch(ch[j], ch[i], r.lo_b, r.lo_a, path.translate(str.maketrans("DI", "ID")))
                rec.append((i, j, int(up), r.lo_a, r.lo_b, path, m, t, u))
    np.savez_compressed(GOLDEN / "golden_kabsch.npz", a=np.array([x[0] for x in rec], np.uint32),
                        b=np.array([x[1] for x in rec], np.uint32), up=np.array([x[2] for x in rec], np.uint8),
                        lo_a=np.array([x[3] for x in rec], np.uint32), lo_b=np.array([x[4] for x in rec], np.uint32),
                        paths=np.array([x[5] for x in rec]), msd=np.array([x[6] for x in rec]),
                        t=np.array([x[7] for x in rec]), u=np.array([x[8] for x in rec]))
    print("cases", len(rec))


if __name__ == "__main__":
    main()
