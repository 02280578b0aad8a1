#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libreseek_ref.so).

Run in the build container only (needs /root/reference/test_data and `make -C oracle ref`):
    python tools/make_golden.py
The fixtures are small and committed; the GPU box never needs /root/reference.

golden_chains.npz    20 real chains of test_data/q100.bca + scop40.bca as the aligner sees them:
                     DSS feature planes, Mu letters, Mu 3-mers, decoded coordinates, ProfileLoader self-reverse scores
golden_pairs_mode{1,2,3}.npz   DSSAligner::AlignQueryTarget on all ordered pairs in -fast / -sensitive / -verysensitive:
                     fp32 bits of score/TS/LDDT/P/E/Qual, Lo/Hi/ids/gaps, Mu filter score, MKF flag, path strings
golden_mu_sw.npz     raw parasail int8 striped SW scores + saturation flags (parasail.cpp:515) on Mu-letter pairs
golden_swfast.npz    SWFast (sw.cpp:79) on explicit random float matrices incl. ties and all-negative cases
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.pyoracle import Ref  # noqa: E402

OUT = ROOT / "tests" / "golden"
TD = Path("/root/reference/test_data")


def pack_chains(chains):
    lens = np.array([c.L for c in chains], np.uint32)
    d = {
        "lens": lens,
        "prof": np.concatenate([c.prof for c in chains], axis=1),
        "mu": np.concatenate([c.mu for c in chains]),
        "xyz": np.concatenate([c.xyz for c in chains], axis=1),
        "selfrev": np.array([c.selfrev for c in chains], np.float32),
        "nkmers": np.array([len(c.kmers) for c in chains], np.uint32),
        "kmers": np.concatenate([c.kmers for c in chains]).astype(np.uint32),
        "labels": np.array([c.label for c in chains]),
        "seq": np.frombuffer(b"".join(c.seq for c in chains), np.uint8),
    }
    return d


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    ref = Ref(2)
    chains = []
    n = ref.bca_open(TD / "q100.bca")
    lens = [ref.lib.ref_bca_len(i) for i in range(n)]
    order = np.argsort(lens)
    pick = sorted(set([int(order[0]), int(order[1]), int(order[n // 4]), int(order[n // 2]), int(order[-1]), int(order[-2])] + list(range(10))))
    for i in pick:
        chains.append(ref.load_chain(i, loader_selfrev=True))
    # a few long chains (>= 600) from scop40 so that the MKF / x-drop path has fixtures too
    n2 = ref.bca_open(TD / "scop40.bca")
    lens2 = np.array([ref.lib.ref_bca_len(i) for i in range(n2)])
    longs = [int(i) for i in np.where(lens2 >= 600)[0][:3]] + [int(i) for i in np.where((lens2 >= 500) & (lens2 < 600))[0][:2]]
    for i in longs:
        chains.append(ref.load_chain(i, loader_selfrev=True))
    print("chains:", [(c.label, c.L) for c in chains])
    packed = pack_chains(chains)
    # profiles of the coordinate-reversed chains (PDBChain::GetReverse + DSS): inputs of the self-reverse score (alignpair.cpp:7-25)
    packed["rev_prof"] = np.concatenate([ref.rev_profile(c.seq, c.xyz) for c in chains], axis=1)
    np.savez_compressed(OUT / "golden_chains.npz", **packed)

    nc = len(chains)
    for mode in (1, 2, 3):
        r = Ref(mode)
        # self-reverse scores as the ProfileLoader of this mode computes them (MKFL differs per mode)
        srs = []
        for c in chains:
            srs.append(r.selfrev(c.seq, c.xyz, loader=True, with_mu=(mode != 3)))
        for c, s in zip(chains, srs):
            c.selfrev = s
        rec = {k: [] for k in ("a", "b", "score", "lo_a", "lo_b", "hi_a", "hi_b", "ids", "gaps", "lddt", "ts", "pvalue",
                               "evalue", "qual", "mu_score", "mkf", "best_hsp", "best_chain", "xdrop", "path_len")}
        paths = []
        for a in range(nc):
            for b in range(nc):
                # verysensitive: DBSearcher::LoadDB does not even load Mu letters (dbsearcher.cpp:251-252)
                rr, path = r.align_pair(chains[a], chains[b], use_mu=(mode != 3))
                rec["a"].append(a); rec["b"].append(b)
                for k in ("score", "lddt", "ts", "pvalue", "evalue", "qual", "mu_score"):
                    rec[k].append(getattr(rr, k))
                for k in ("lo_a", "lo_b", "hi_a", "hi_b", "ids", "gaps", "mkf", "path_len"):
                    rec[k].append(getattr(rr, k))
                rec["best_hsp"].append(rr.best_hsp_score); rec["best_chain"].append(rr.best_chain_score)
                rec["xdrop"].append(rr.xdrop_score)
                paths.append(path)
        out = {}
        for k, v in rec.items():
            if k in ("score", "lddt", "ts", "pvalue", "evalue", "qual", "mu_score", "xdrop"):
                out[k] = np.array(v, np.float32)
            elif k in ("mkf", "best_hsp", "best_chain"):
                out[k] = np.array(v, np.int32)
            else:
                out[k] = np.array(v, np.uint32)
        # the reference's own TSV writer on a sample of pairs, both directions (DSSAligner::ToTsv)
        cols = "query+target+evalue+pvalue+ql+tl+qlo+qhi+tlo+thi+qcovpct+tcovpct+pctid+newts+raw+dpscore+lddt+ids+gaps+aq+muhsp+muchain+cigar"
        tsv_k, tsv_up, tsv_line = [], [], []
        for k in range(0, nc * nc, 3):
            if not paths[k] or rec["evalue"][k] > 1e38:
                continue
            a, b2 = rec["a"][k], rec["b"][k]
            for up in (1, 0):
                line = r.align_pair_tsv(chains[a], chains[b2], cols, up, use_mu=(mode != 3))
                tsv_k.append(k); tsv_up.append(up); tsv_line.append(line)
        out["tsv_cols"] = np.array(cols)
        out["tsv_k"] = np.array(tsv_k, np.uint32)
        out["tsv_up"] = np.array(tsv_up, np.uint8)
        out["tsv_lines"] = np.array(tsv_line)
        out["selfrev"] = np.array(srs, np.float32)
        out["path_off"] = np.concatenate([[0], np.cumsum([len(p) for p in paths])]).astype(np.uint64)
        out["paths"] = np.frombuffer("".join(paths).encode(), np.uint8)
        np.savez_compressed(OUT / f"golden_pairs_mode{mode}.npz", **out)
        print(f"mode {mode}: {nc * nc} pairs, with path {sum(1 for p in paths if p)}, mkf {int(np.sum(out['mkf']))}")

    # -fast -db prefilter: the reference binary's own candidate TSV (search.cpp:87-89, -keeptmp) at -threads 1
    import os, re, subprocess
    from oracle.pyoracle import REF_BIN

    def mus(fn):
        n = ref.bca_open(fn)
        out = []
        for i in range(n):
            label, seq, xyz = ref.bca_chain(i)
            out.append(ref.dss(seq, xyz)[1])
        return out
    mq, mt = mus(TD / "q10.bca"), mus(TD / "q100.bca")
    pf = {"q_len": np.array([len(m) for m in mq], np.uint32), "q_mu": np.concatenate(mq),
          "t_len": np.array([len(m) for m in mt], np.uint32), "t_mu": np.concatenate(mt)}
    for name, extra in (("idxq", []), ("idxt", ["-idxt"]), ("rsb5", ["-rsb_size", "5"])):
        subprocess.run([str(REF_BIN), "-search", str(TD / "q10.bca"), "-db", str(TD / "q100.bca"), "-fast", "-keeptmp",
                        "-threads", "1", "-output", "/tmp/_pf_hits.tsv", "-log", "/tmp/_pf.log"] + extra,
                       capture_output=True, check=True)
        fn = re.search(r"MuFilterTsvFN=(\S+)", open("/tmp/_pf.log").read()).group(1)
        tg, qs = [], []
        for ln in open(fn).read().splitlines()[1:]:
            f = ln.split("\t")
            for x in f[2:]:
                tg.append(int(f[0])); qs.append(int(x))
        os.remove(fn)
        pf[f"{name}_t"] = np.array(tg, np.uint32)
        pf[f"{name}_q"] = np.array(qs, np.uint32)
        print("prefilter", name, len(tg), "candidate pairs")
    np.savez_compressed(OUT / "golden_prefilter.npz", **pf)

    # raw parasail
    rng = np.random.default_rng(42)
    r = Ref(2)
    A, B, S, SAT = [], [], [], []
    mus = [c.mu for c in chains]
    for _ in range(300):
        a = mus[rng.integers(nc)]
        b = mus[rng.integers(nc)]
        if rng.random() < 0.3:  # near-identical pairs drive the int8 lanes into saturation
            b = a.copy()
            k = rng.integers(0, len(b), size=max(1, len(b) // 10))
            b[k] = rng.integers(0, 36, size=len(k))
        if rng.random() < 0.2:
            a = a[:rng.integers(1, min(40, len(a)))]
        s, sat = r.parasail_sw(a, b)
        A.append(a); B.append(b); S.append(s); SAT.append(sat)
    np.savez_compressed(OUT / "golden_mu_sw.npz",
                        a=np.concatenate(A), b=np.concatenate(B),
                        la=np.array([len(x) for x in A], np.uint32), lb=np.array([len(x) for x in B], np.uint32),
                        score=np.array(S, np.int32), sat=np.array(SAT, np.int32))
    print("parasail: saturated", int(np.sum(SAT)), "of", len(SAT))

    # SWFast on explicit matrices
    mats, LA, LB, SC, LOA, LOB, P = [], [], [], [], [], [], []
    for t in range(60):
        la, lb = int(rng.integers(1, 70)), int(rng.integers(1, 70))
        if t % 6 == 0:
            Smx = -np.abs(rng.normal(size=(la, lb))).astype(np.float32)  # all negative: no alignment
        elif t % 6 == 1:
            Smx = rng.integers(-2, 3, size=(la, lb)).astype(np.float32)  # small integers: many exact ties
        else:
            Smx = (rng.normal(size=(la, lb)) - 0.3).astype(np.float32)
            k = min(la, lb)
            Smx[np.arange(k), np.arange(k)] += 1.0
        s, lo_a, lo_b, path = r.swfast(Smx, -0.685533, -0.051881) if t % 2 else r.swfast(Smx, -1.5, -0.42)
        mats.append(Smx.ravel()); LA.append(la); LB.append(lb); SC.append(s); LOA.append(lo_a); LOB.append(lo_b); P.append(path)
    np.savez_compressed(OUT / "golden_swfast.npz", mats=np.concatenate(mats), la=np.array(LA, np.uint32), lb=np.array(LB, np.uint32),
                        score=np.array(SC, np.float32), lo_a=np.array(LOA, np.uint32), lo_b=np.array(LOB, np.uint32),
                        path_off=np.concatenate([[0], np.cumsum([len(p) for p in P])]).astype(np.uint64),
                        paths=np.frombuffer("".join(P).encode(), np.uint8))
    print("swfast: no-alignment cases", sum(1 for p in P if not p))


if __name__ == "__main__":
    main()
