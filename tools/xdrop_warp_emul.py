#!/usr/bin/env python3
"""Lane-level emulation (numpy, float32) of mkf_xdrop_warp_kernel's row loop and run traceback, checked against a literal
sequential restatement of the same DP (xdropfwd.cpp:71-386 as in mkf_kernel.cu::xdrop_item).

Developer tool: the warp kernel reorganises a strictly sequential DP (prefix scans, rule-based previous-row reads, speculative
row extension); this harness finds disagreements on the CPU, cell by cell, where a GPU run only shows a different path.
Usage: python tools/xdrop_warp_emul.py [ntrials]   (needs build/data/scop40.bca and oracle/_ref for real feature letters;
falls back to random score matrices otherwise)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

F = np.float32
NEG = F(-9e9)
XB_DM, XB_IM, XB_MD, XB_MI = 1, 2, 4, 8
NONE = 0xFFFFFFFF


def seq_xdrop(S, open_, ext, X):
    """xdrop_item: returns (best, besti, bestj, tb matrix, path string reversed-as-staged)."""
    LA, LB = S.shape
    absopen, absext = F(-open_), F(-ext)
    M = np.zeros(LB + 4, F)  # M[-1] -> index 0
    Dr = np.zeros(LB + 4, F)
    W = LB + 3
    tb = np.zeros((LA + 3, W), np.uint8)

    def Mg(j):
        return M[j + 1]

    def Ms(j, v):
        M[j + 1] = v
    Ms(-1, NEG)
    Dr[0] = NEG
    Dr[1] = NEG
    best = F(0)
    besti = bestj = 0
    prev_jlo = prev_jhi = 0
    jlo = jhi = 1
    M0 = best
    for i in range(1, LA + 1):
        if jlo == prev_jlo:
            Ms(jlo - 1, NEG)
            Dr[jlo] = NEG
        endj = min(prev_jhi + 1, LB)
        for j in range(endj + 1, min(jhi + 1, LB) + 1):
            Ms(j - 1, NEG)
            Dr[j] = NEG
        next_jlo = next_jhi = NONE
        I0 = NEG
        j = jlo
        while j <= jhi:
            bits = 0
            saved = M0
            x = M0
            if Dr[j] > x:
                x = Dr[j]; bits = XB_DM
            if I0 > x:
                x = I0; bits = XB_IM
            M0 = Mg(j)
            s = F(S[i - 1, j - 1] + x)
            Ms(j, s)
            h = F(F(s - best) + X)
            if h > 0:
                next_jlo = min(next_jlo, j + 1); next_jhi = j + 1
            if h > absopen:
                next_jlo = min(next_jlo, j)
            if h > absext and j == jhi and jhi + 1 < LB:
                jhi += 1
                ne = max(min(jhi + 1, LB), endj)
                for j2 in range(endj + 1, ne + 1):
                    if j2 - 1 > j:
                        Ms(j2 - 1, NEG)
                    Dr[j2] = NEG
                endj = ne
            if s >= best:
                best = s; besti = i; bestj = j
            if j != jlo:
                md = F(saved + open_)
                dn = F(Dr[j] + ext)
                if md >= dn:
                    dn = md; bits |= XB_MD
                Dr[j] = dn
                h = F(F(dn - best) + X)
                if h > 0:
                    next_jlo = min(next_jlo, j - 1); next_jhi = max(next_jhi, j - 1)
            mi = F(saved + open_)
            I0 = F(I0 + ext)
            if mi >= I0:
                I0 = mi; bits |= XB_MI
            h = F(F(I0 - best) + X)
            if h > 0:
                next_jlo = min(next_jlo, j + 1); next_jhi = max(next_jhi, j + 1)
            if h > absext and j == jhi and jhi + 1 < LB:
                jhi += 1
                ne = max(min(jhi + 1, LB), endj)
                for j2 in range(endj + 1, ne + 1):
                    Ms(j2 - 1, NEG)
                    Dr[j2] = NEG
                endj = ne
            tb[i, j] = bits
            j += 1
        if jhi < LB:
            j1 = jhi + 1
            b1 = 0
            md = F(M0 + open_)
            dn = F(Dr[j1] + ext)
            if md >= dn:
                dn = md; b1 = XB_MD
            Dr[j1] = dn
            tb[i, j1] = b1
        if next_jlo == NONE:
            break
        prev_jlo, prev_jhi = jlo, jhi
        jlo, jhi = min(next_jlo, LB), min(next_jhi, LB)
        if jlo == prev_jlo:
            M0 = NEG; Dr[jlo] = NEG
        else:
            M0 = Mg(jlo - 1)
    return best, besti, bestj, tb


def traceback_seq(tb, besti, bestj):
    i, j, st, out = besti, bestj, 0, []
    while True:
        out.append("MDI"[st])
        if i == 1 or j == 1:
            break
        if st == 0:
            c = tb[i, j]
            nx = 1 if c & XB_DM else 2 if c & XB_IM else 0
            i -= 1; j -= 1
        elif st == 1:
            nx = 0 if tb[i, j + 1] & XB_MD else 1
            i -= 1
        else:
            nx = 0 if tb[i + 1, j] & XB_MI else 2
            j -= 1
        st = nx
    return "".join(out)


def shfl_up(v, d, fill=None):
    out = v.copy()
    out[d:] = v[:-d]
    return out  # lanes < d keep their own value, as __shfl_up_sync does


def warp_xdrop(S, open_, ext, X):
    """mkf_xdrop_warp_kernel, lanes as numpy vectors of 32."""
    LA, LB = S.shape
    absopen, absext = F(-open_), F(-ext)
    M = np.zeros(LB + 40, F)
    Dr = np.zeros(LB + 40, F)
    W = LB + 3
    tb = np.zeros((LA + 3, W), np.uint8)
    lane = np.arange(32)
    M[0] = NEG  # M[-1]
    Dr[0] = NEG
    Dr[1] = NEG
    best = F(0)
    besti = bestj = 0
    prev_jlo = prev_jhi = 0
    jlo = jhi = 1
    diag0 = F(0)
    for i in range(1, LA + 1):
        next_jlo = next_jhi = NONE
        I_carry, diag_carry, best_run = NEG, diag0, best
        jhi_cur, j0, jhi_final = jhi, jlo, jhi
        M0_end = NEG
        jstar = jhi
        quirk_row = jstar >= prev_jhi + 1
        wipe = wipe_site1 = False
        while True:
            j = j0 + lane
            inb = j <= LB
            jj = np.minimum(j, LB + 30)
            oldM = np.where(inb & (j >= prev_jlo) & (j <= prev_jhi), M[jj + 1], NEG).astype(F)
            oldD = np.where(inb & (j > prev_jlo) & (j <= prev_jhi + 1), Dr[jj], NEG).astype(F)
            diag = shfl_up(oldM, 1)
            diag[0] = diag_carry
            mi = (diag + F(open_)).astype(F)
            A = mi.copy()
            t0 = F(I_carry + ext)
            A[0] = mi[0] if mi[0] >= t0 else t0
            for d in range(5):
                v = shfl_up(A, 1 << d)
                for _ in range(1 << d):
                    v = (v + F(ext)).astype(F)
                A = np.where(lane >= (1 << d), np.maximum(A, v), A).astype(F)
            I0 = shfl_up(A, 1)
            I0[0] = I_carry
            tI = (I0 + F(ext)).astype(F)
            bMI = mi >= tI
            Inew = np.where(bMI, mi, tI).astype(F)
            x = diag.copy()
            bits = np.zeros(32, np.uint8)
            m = oldD > x
            x = np.where(m, oldD, x); bits = np.where(m, XB_DM, bits)
            m = I0 > x
            x = np.where(m, I0, x).astype(F); bits = np.where(m, XB_IM, bits)
            srow = np.where(inb, S[i - 1, np.minimum(j, LB) - 1], F(0)).astype(F)
            s = np.where(inb, (srow + x).astype(F), NEG).astype(F)
            pm = s.copy()
            for d in range(5):
                v = shfl_up(pm, 1 << d)
                pm = np.where(lane >= (1 << d), np.maximum(pm, v), pm).astype(F)
            bb = shfl_up(pm, 1)
            best_before = np.maximum(F(best_run), bb).astype(F)
            best_before[0] = best_run
            best_after = np.maximum(best_before, s).astype(F)
            h1 = ((s - best_before).astype(F) + F(X)).astype(F)
            hasD = j != jlo
            md = (diag + F(open_)).astype(F)
            dn = (oldD + F(ext)).astype(F)
            bMD = md >= dn
            dn = np.where(bMD, md, dn).astype(F)
            h2 = ((dn - best_after).astype(F) + F(X)).astype(F)
            h3 = ((Inew - best_after).astype(F) + F(X)).astype(F)
            E = ((h1 > absext) | (h3 > absext)) & (j + 1 < LB)
            if quirk_row and j0 <= jstar < j0 + 32:
                ls = jstar - j0
                if E[ls]:
                    wipe = True
                    wipe_site1 = bool(h1[ls] > absext)
                    if wipe_site1 and jstar >= prev_jhi + 2:
                        dnl = F(NEG + F(ext))
                        bMD[ls] = md[ls] >= dnl
                        dn[ls] = md[ls] if bMD[ls] else dnl
                        h2[ls] = F(F(dn[ls] - best_after[ls]) + F(X))
            stop = ~inb | ((j >= jhi_cur) & ~E)
            last_chunk = bool(stop.any())
            f = int(np.argmax(stop)) if last_chunk else 31
            valid = lane <= f
            for l in range(32):
                if valid[l]:
                    M[j[l] + 1] = s[l]
                    bt = int(bits[l]) | (XB_MI if bMI[l] else 0)
                    if hasD[l]:
                        Dr[j[l]] = dn[l]
                        if bMD[l]:
                            bt |= XB_MD
                    tb[i, j[l]] = bt
            upd = valid & (s >= best_before)
            if upd.any():
                bestj = j0 + int(np.nonzero(upd)[0].max()); besti = i
            best_run = max(F(best_run), pm[f])
            lo_c = np.full(32, NONE, np.int64)
            mx = np.zeros(32, np.int64)
            for l in range(32):
                if valid[l]:
                    if h1[l] > 0: lo_c[l] = j[l] + 1
                    if h1[l] > absopen: lo_c[l] = min(lo_c[l], j[l])
                    if hasD[l] and h2[l] > 0:
                        lo_c[l] = min(lo_c[l], j[l] - 1); mx[l] = j[l] - 1
                    if h3[l] > 0:
                        lo_c[l] = min(lo_c[l], j[l] + 1); mx[l] = j[l] + 1
            next_jlo = min(next_jlo, int(lo_c.min()))
            mA = valid & (h1 > 0)
            if mA.any():
                lA = int(np.nonzero(mA)[0].max())
                next_jhi = max(j0 + lA + 1, int(mx[lA:].max()))
            else:
                mm = int(mx.max())
                if mm:
                    next_jhi = max(next_jhi, mm)
            if last_chunk:
                jhi_final = j0 + f
                M0_end = oldM[f]
                break
            I_carry = Inew[31]
            diag_carry = oldM[31]
            if j0 + 32 > jhi_cur:
                jhi_cur = j0 + 32
            j0 += 32
        if jhi_final < LB:
            j1 = jhi_final + 1
            pd = Dr[j1] if (j1 > prev_jlo and j1 <= prev_jhi + 1) else NEG
            md = F(M0_end + F(open_))
            dn = F(pd + F(ext))
            b1 = 0
            if md >= dn:
                dn = md; b1 = XB_MD
            Dr[j1] = dn
            tb[i, j1] = b1
        if wipe:
            for q in range(prev_jhi + 2, (jstar - 1 if wipe_site1 else jstar) + 1):
                Dr[q] = NEG
            if not wipe_site1:
                for q in range(prev_jhi + 1, jstar + 1):
                    M[q + 1] = NEG
        best = F(best_run)
        if next_jlo == NONE:
            break
        prev_jlo, prev_jhi = jlo, jhi_final
        jlo, jhi = min(next_jlo, LB), min(next_jhi, LB)
        diag0 = NEG if jlo == prev_jlo else M[jlo - 1 + 1]
    return best, besti, bestj, tb


def traceback_runs(tb, besti, bestj):
    W = tb.shape[1]
    flat = tb.reshape(-1)
    lane = np.arange(32)
    ti, tj, st, out = besti, bestj, 0, []
    while True:
        di = lane if st != 2 else 0 * lane
        dj = lane if st != 1 else 0 * lane
        inr = (ti > di) & (tj > dj)
        pi, pj = ti - di, tj - dj
        boundary = inr & ((pi == 1) | (pj == 1))
        nx = np.full(32, st)
        for l in range(32):
            if inr[l] and not boundary[l]:
                if st == 0:
                    c = flat[pi[l] * W + pj[l]]
                    nx[l] = 1 if c & XB_DM else 2 if c & XB_IM else 0
                elif st == 1:
                    nx[l] = 0 if flat[pi[l] * W + pj[l] + 1] & XB_MD else 1
                else:
                    nx[l] = 0 if flat[(pi[l] + 1) * W + pj[l]] & XB_MI else 2
        endm = ~inr | boundary | (nx != st)
        e = int(np.argmax(endm)) if endm.any() else 31
        out.append("MDI"[st] * (e + 1))
        if not endm.any():
            ti -= 32 if st != 2 else 0
            tj -= 32 if st != 1 else 0
            continue
        if boundary[e]:
            break
        nst = int(nx[e])
        ei = ti - (e if st != 2 else 0)
        ej = tj - (e if st != 1 else 0)
        ti = ei - 1 if st != 2 else ei
        tj = ej - 1 if st != 1 else ej
        st = nst
    return "".join(out)


def compare(S, open_, ext, X, tag):
    b0, i0, j0, tb0 = seq_xdrop(S, open_, ext, X)
    b1, i1, j1, tb1 = warp_xdrop(S, open_, ext, X)
    ok = (np.float32(b0).tobytes() == np.float32(b1).tobytes()) and (i0, j0) == (i1, j1) and np.array_equal(tb0, tb1)
    if ok and b0 > 0:
        p0, p1 = traceback_seq(tb0, i0, j0), traceback_runs(tb1, i1, j1)
        ok = p0 == p1
    if not ok:
        d = np.argwhere(tb0 != tb1)
        print(f"MISMATCH {tag}: best {b0} {b1} end ({i0},{j0}) ({i1},{j1}) first differing trace cells {d[:5].tolist()} of {len(d)}")
    return ok


def main():
    ntr = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(7)
    open_, ext, X = F(-0.685533), F(-0.051881), F(8)
    bad = 0
    mats = []
    bca = ROOT / "build" / "data" / "scop40.bca"
    try:
        from oracle.pyoracle import Port, Ref
        from reseek_b200 import chainio
        if bca.exists() and Ref.available():
            labels, seqs, xyzs = chainio.read_bca(bca)
            ref = Ref(2)
            port = Port(2)
            sc, tbl = ref.get_params()
            tabs = []
            k = 0
            for f in range(8):
                n = 20 if f == 0 else 16
                tabs.append(tbl[k:k + n * n].reshape(n, n)); k += n * n
            want = ["d1nqka_", "d1ofda2"]
            idx = [labels.index(w) for w in want]
            longs = [i for i in range(len(labels)) if len(seqs[i]) >= 400][:12]
            profs = {i: ref.dss(seqs[i], xyzs[i])[0] for i in set(idx + longs)}

            def smat(a, b):
                pa, pb = profs[a], profs[b]
                t = np.zeros((pa.shape[1], pb.shape[1]), F)
                for f in range(8):  # per-cell feature order, starting from 0 (xdrophsp.cpp:8-33)
                    t = (t + tabs[f][pa[f][:, None], pb[f][None, :]]).astype(F)
                return t
            full = smat(idx[0], idx[1])
            # sub-matrices from many seeds of the pair that differed on the GPU, both directions
            for _ in range(ntr):
                la0, lb0 = int(rng.integers(0, full.shape[0] - 40)), int(rng.integers(0, full.shape[1] - 40))
                mats.append((full[la0:, lb0:], f"{want[0]}x{want[1]} fwd seed ({la0},{lb0})"))
                mats.append((full[:la0 + 30, :lb0 + 30][::-1, ::-1], f"{want[0]}x{want[1]} bwd seed ({la0 + 30},{lb0 + 30})"))
            for a in longs[:4]:
                for b in longs[4:8]:
                    fm = smat(a, b)
                    la0, lb0 = int(rng.integers(0, fm.shape[0] // 2)), int(rng.integers(0, fm.shape[1] // 2))
                    mats.append((fm[la0:, lb0:], f"{labels[a]}x{labels[b]} seed ({la0},{lb0})"))
    except Exception as e:  # noqa: BLE001
        print("real chains unavailable:", e)
    for t in range(ntr):
        LA, LB = int(rng.integers(2, 200)), int(rng.integers(2, 300))
        S = rng.normal(-0.3, 1.2, size=(LA, LB)).astype(F)
        # a noisy diagonal of positives with gaps, so that bands wander, widen and shrink
        d = int(rng.integers(-20, 20))
        for i in range(LA):
            jj = i + d + int(rng.integers(-1, 2)) * int(rng.integers(0, 15))
            if 0 <= jj < LB:
                S[i, jj] += F(rng.uniform(1.0, 4.0))
        mats.append((S, f"random {t} {LA}x{LB}"))
    for S, tag in mats:
        if S.shape[0] < 2 or S.shape[1] < 2:
            continue
        bad += not compare(np.ascontiguousarray(S), open_, ext, X, tag)
    print(f"{len(mats)} DPs, {bad} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
