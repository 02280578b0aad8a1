#!/usr/bin/env python3
"""The pairs on which the reference's SCOP40 all-vs-all (`reseek -search scop40.bca -fast`) differs from this engine are the
reference's own history-dependent pairs: aligned in ISOLATION (a .bca holding just the two chains, -threads 1) the reference
prints this engine's lines.

What carries over inside the reference is not pinned down completely.  The x-drop trace matrix is re-malloc'ed per call and
never cleared (xdpmem.h:96-107, mx.h:38-54; INIT_TRACE empty outside TRACE builds, xdropfwd.cpp:101-104), so trace cells
outside the computed band hold recycled heap content - but a run with a zero-filled heap (GLIBC_TUNABLES=glibc.malloc.
tcache_count=0 MALLOC_PERTURB_=255) still differs on about as many pairs, some with a different DP score, so aligner state
that survives between pairs (the long-chain k-mer path) is involved as well.  What is established: for every such pair the
reference, asked for that pair as the first work of a fresh aligner, prints this engine's lines.

Inputs: the reference binary's all-vs-all TSV (oracle/_ref/reseek_ref, run here on the CPU), this engine's TSV of the same
command (rsk_host_demo on the GPU box, brought back gzipped), both with
  -columns query+target+dpscore+newts+evalue+qlo+qhi+tlo+thi+cigar
Writes a log (profiles/r2_history_pairs.md) and, with --fixture, a small committed fixture of a few such pairs
(tests/golden/history_pairs.npz) for tests/test_host_search.py."""
import argparse
import gzip
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
COLS = "query+target+dpscore+newts+evalue+qlo+qhi+tlo+thi+cigar"


def read_tsv(path):
    op = gzip.open if str(path).endswith(".gz") else open
    d = {}
    with op(path, "rt") as f:
        for line in f:
            line = line.rstrip("\n")
            p = line.split("\t")
            if len(p) < 2:
                continue
            d[(p[0], p[1])] = line
    return d


def fresh_pair(bca, a, b):
    """The pair as the FIRST alignment of a fresh DSSAligner in a fresh process (oracle/_ref through oracle/ref_driver.cpp:
    features from the reference's DSS, self-reverse scores from its loader aligner, which has its own scratch memory)."""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from reseek_b200 import chainio\n"
        "from oracle.pyoracle import Ref, Chain\n"
        "labels, seqs, xyzs = chainio.read_bca(%r)\n"
        "ia, ib = sorted([labels.index(%r), labels.index(%r)])\n"
        "ref = Ref(1)\n"
        "def load(i):\n"
        "    p, m, k = ref.dss(seqs[i], xyzs[i])\n"
        "    return Chain(p, m, xyzs[i], ref.selfrev(seqs[i], xyzs[i], loader=True), label=labels[i], seq=seqs[i], kmers=k)\n"
        "A, B = load(ia), load(ib)\n"
        "print(ref.align_pair_tsv(A, B, %r, True)); print(ref.align_pair_tsv(A, B, %r, False))\n") % (str(ROOT), str(bca), a, b, COLS, COLS)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True)
    d = {}
    for line in r.stdout.splitlines():
        p = line.split("\t")
        if len(p) >= 2:
            d[(p[0], p[1])] = line
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/tmp/ref_self.tsv")
    ap.add_argument("--ours", default=str(ROOT / "gpurun_out" / "ours_self.tsv.gz"))
    ap.add_argument("--bca", default=str(ROOT / "build" / "data" / "scop40.bca"))
    ap.add_argument("--refbin", default=str(ROOT / "oracle" / "_ref" / "reseek_ref"))
    ap.add_argument("--log", default=str(ROOT / "profiles" / "r2_history_pairs.md"))
    ap.add_argument("--fixture", default=str(ROOT / "tests" / "golden" / "history_pairs.npz"))
    ap.add_argument("--nfixture", type=int, default=4)
    args = ap.parse_args()
    from reseek_b200 import chainio
    ref, ours = read_tsv(args.ref), read_tsv(args.ours)
    labels, seqs, xyzs = chainio.read_bca(args.bca)
    index = {}
    for i, lab in enumerate(labels):
        index.setdefault(lab, i)
    keys = set(ref) | set(ours)
    diff = sorted(k for k in keys if ref.get(k) != ours.get(k))
    pairs = sorted({tuple(sorted(k)) for k in diff})
    out = [f"# History-dependent pairs of the reference's SCOP40 all-vs-all (`-search scop40.bca -fast`)\n",
           f"Reference binary (oracle/_ref/reseek_ref, strict build, -threads 8): {len(ref)} lines.  This engine (rsk_host_demo on one B200): "
           f"{len(ours)} lines.  Lines that differ or exist on one side only: {len(diff)}, in {len(pairs)} unordered chain pairs.\n",
           "Each of those pairs was then written to a `.bca` of its own and given to the reference alone "
           "(`reseek_ref -search pair.bca -fast -threads 1 -columns " + COLS + "`):\n",
           "| pair | lengths | lines differing in the all-vs-all | reference in isolation == this engine |", "|---|---|---|---|"]
    nsame = 0
    fixture = []
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        for (a, b) in pairs:
            ia, ib = index[a], index[b]
            ids = [ia] if ia == ib else [ia, ib]
            chainio.write_bca(tmp / "pair.bca", [labels[i] for i in ids], [seqs[i] for i in ids], [xyzs[i] for i in ids])
            subprocess.run([args.refbin, "-search", str(tmp / "pair.bca"), "-fast", "-threads", "1", "-columns", COLS, "-output",
                            str(tmp / "pair.tsv")], capture_output=True, check=True)
            iso = read_tsv(tmp / "pair.tsv")
            want = {k: v for k, v in ours.items() if set(k) <= {a, b}}
            same = iso == want
            how = "yes"
            if not same:
                # in the two-chain file the aligner has already aligned (a, a) and the self-reverse pairs when it reaches (a, b):
                # ask a fresh process for the pair as its very first alignment
                fr = fresh_pair(args.bca, a, b)
                same = fr == {k: v for k, v in ours.items() if set(k) == {a, b}}
                how = "yes (first alignment of a fresh process; the two-chain run still carried history)" if same else "NO"
            nsame += same
            nd = sum(1 for k in diff if set(k) == {a, b})
            out.append(f"| {a} {b} | {len(seqs[ia])} {len(seqs[ib])} | {nd} | {how} |")
            if how != "yes":
                continue
            if same and len(fixture) < args.nfixture and len(seqs[ia]) + len(seqs[ib]) < 1400:
                fixture.append((ids, sorted(iso.values())))
    out.insert(3, f"**{nsame} of {len(pairs)} pairs: in isolation the reference prints exactly this engine's lines.**\n")
    out.append("\nNotes.  (1) The reference's all-vs-all output is not reproducible from run to run on these pairs (two runs here gave 113 "
               "and 142 differing pairs against the same engine output); every other line is identical.  (2) A reference run with a "
               "zero-filled heap (`GLIBC_TUNABLES=glibc.malloc.tcache_count=0 MALLOC_PERTURB_=255`) still differs from this engine on "
               "289 lines, some with a different DP score: uninitialised trace cells (xdpmem.h:96-107) are not the whole story, state "
               "of the long-chain path that survives from pair to pair inside one `DSSAligner` is involved.  (3) This engine has no "
               "such state: every pair is computed from its two chains alone, which is what the isolated runs of the reference give.")
    Path(args.log).write_text("\n".join(out) + "\n")
    print(f"{len(diff)} differing lines, {len(pairs)} pairs, {nsame} explained; log in {args.log}")
    if fixture:
        lab, sq, xy, lines, members = [], [], [], [], []
        for ids, ls in fixture:
            members.append([len(lab) + k for k in range(len(ids))])
            for i in ids:
                lab.append(labels[i]); sq.append(np.frombuffer(seqs[i], np.uint8)); xy.append(xyzs[i])
            lines.append("\n".join(ls))
        np.savez_compressed(args.fixture, labels=np.array(lab), lens=np.array([len(s) for s in sq], np.uint32),
                            seq=np.concatenate(sq), xyz=np.concatenate(xy, axis=1), lines=np.array(lines),
                            members=np.array([m + [-1] * (2 - len(m)) for m in members], np.int32), columns=np.array(COLS))
        print(f"fixture with {len(fixture)} pairs in {args.fixture}")
    return 0 if nsame == len(pairs) else 1


if __name__ == "__main__":
    sys.exit(main())
