"""Seeded synthetic chains for the benchmark and the parity tests (SURVEY.md §8d "kernel-level" inputs).

Feature letters are drawn i.i.d. from the trained background vectors (trained_features.cpp X_f_i, exposed by
the library as rsk_feature_bgfreq), Mu letters from a first-order Markov chain with a uniform-ish stationary
law, coordinates are a 3.8 A random walk rounded to 0.1 A and decoded like PDBChain::ICToCoord
(pdbchain.h:89-90: float(ic/10.0f) - 1000).  A fraction of the "DB" chains are mutated copies of queries
(30 % letter substitutions, 5 % indels) so that the filter -> SW mix and non-trivial alignments are exercised.
Pure numpy; no compute of the hot path happens here.
"""
import ctypes as C

import numpy as np

from . import lib as _lib

NFEAT = 8
ALPHA = [20, 16, 16, 16, 16, 16, 16, 16]


def background():
    L = _lib.load_library()
    out = []
    for f in range(NFEAT):
        n = L.rsk_feature_alpha(f)
        p = np.array(L.rsk_feature_bgfreq(f)[:n], np.float64)
        out.append(p / p.sum())
    return out


class SynthChains:
    """Structure-of-arrays chain set on the host: lens [n], prof [8][total], mu [total], xyz [3][total], selfrev [n]."""

    def __init__(self, lens, prof, mu, xyz, selfrev):
        self.lens = np.ascontiguousarray(lens, np.uint32)
        self.prof = np.ascontiguousarray(prof, np.uint8)
        self.mu = np.ascontiguousarray(mu, np.uint8)
        self.xyz = np.ascontiguousarray(xyz, np.float32)
        self.selfrev = np.ascontiguousarray(selfrev, np.float32)
        self.off = np.concatenate([[0], np.cumsum(self.lens, dtype=np.int64)])

    @property
    def n(self):
        return len(self.lens)

    @property
    def total(self):
        return int(self.off[-1])

    def chain(self, i):
        """(prof [8][L], mu [L], xyz [3][L], selfrev) views of chain i."""
        s, e = int(self.off[i]), int(self.off[i + 1])
        return self.prof[:, s:e], self.mu[s:e], self.xyz[:, s:e], float(self.selfrev[i])

    def subset(self, idx):
        idx = list(idx)
        lens = self.lens[idx]
        cols = np.concatenate([np.arange(self.off[i], self.off[i + 1]) for i in idx]) if idx else np.zeros(0, np.int64)
        return SynthChains(lens, self.prof[:, cols], self.mu[cols], self.xyz[:, cols], self.selfrev[idx])

    def nbytes(self):
        return self.lens.nbytes + self.prof.nbytes + self.mu.nbytes + self.xyz.nbytes + self.selfrev.nbytes


def _walk(rng, L):
    """3.8 A C-alpha random walk with direction persistence, quantised like the .bca integer coordinates."""
    d = rng.normal(size=(L, 3))
    for i in range(1, L):
        d[i] = 0.75 * d[i - 1] + 0.66 * d[i]
    d /= np.linalg.norm(d, axis=1, keepdims=True) + 1e-12
    pos = np.cumsum(3.8 * d, axis=0)
    pos -= pos.mean(axis=0)
    ic = np.floor((pos + 1000.0) * 10.0 + 0.5).astype(np.uint16)
    return (ic.astype(np.float32) / np.float32(10.0) - np.float32(1000.0)).astype(np.float32).T  # [3][L]


def _walk_fast(rng, L):
    """Vectorised variant of _walk for large sets (AR(1) filter via cumulative products is not needed:
    we only want plausible, bounded inter-residue distances, so a smoothed random direction suffices)."""
    d = rng.normal(size=(L + 8, 3))
    k = np.array([0.05, 0.1, 0.2, 0.3, 0.2, 0.1, 0.05])
    d = np.stack([np.convolve(d[:, c], k, mode="same") for c in range(3)], axis=1)[4:4 + L]
    d /= np.linalg.norm(d, axis=1, keepdims=True) + 1e-12
    pos = np.cumsum(3.8 * d, axis=0)
    pos -= pos.mean(axis=0)
    ic = np.floor((pos + 1000.0) * 10.0 + 0.5).astype(np.uint16)
    return (ic.astype(np.float32) / np.float32(10.0) - np.float32(1000.0)).astype(np.float32).T


def _walk_all(rng, lens):
    """All chains at once: smoothed random unit steps of 3.8 A, positions restarted at every chain start,
    quantised to the .bca grid (0.1 A) and decoded as ICToCoord does."""
    lens = np.asarray(lens, np.int64)
    total = int(lens.sum())
    d = rng.standard_normal(size=(total + 8, 3), dtype=np.float32)
    k = np.array([0.05, 0.1, 0.2, 0.3, 0.2, 0.1, 0.05], np.float32)
    d = np.stack([np.convolve(d[:, c], k, mode="same") for c in range(3)], axis=1)[4:4 + total]
    d /= np.linalg.norm(d, axis=1, keepdims=True) + 1e-12
    pos = np.cumsum(3.8 * d, axis=0, dtype=np.float64)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    base = np.repeat(pos[starts] , lens, axis=0)
    pos = pos - base
    ic = np.floor((pos + 1000.0) * 10.0 + 0.5)
    ic = np.clip(ic, 0, 65535).astype(np.uint16)
    return np.ascontiguousarray((ic.astype(np.float32) / np.float32(10.0) - np.float32(1000.0)).astype(np.float32).T)


def make_chains(n, length, seed, length_jitter=0.0, bg=None, selfrev_scale=0.08):
    """n chains of (about) `length` residues.  length may be an int or an array of n lengths."""
    rng = np.random.default_rng(seed)
    bg = bg or background()
    if np.isscalar(length):
        if length_jitter > 0:
            lens = np.maximum(8, rng.normal(length, length * length_jitter, size=n).astype(np.int64))
        else:
            lens = np.full(n, int(length), np.int64)
    else:
        lens = np.asarray(length, np.int64)
    total = int(lens.sum())
    prof = np.empty((NFEAT, total), np.uint8)
    for f in range(NFEAT):
        # inverse CDF through a 16-bit lookup table (quantisation 2^-16, far below the frequencies' precision)
        cdf = np.cumsum(bg[f])
        lut = np.minimum(np.searchsorted(cdf, (np.arange(65536) + 0.5) / 65536.0, side="right"), ALPHA[f] - 1).astype(np.uint8)
        prof[f] = lut[rng.integers(0, 65536, size=total, dtype=np.uint16)]
    # Mu letters: sticky first-order chain (secondary structure persists), 36 letters
    mu = rng.integers(0, 36, size=total).astype(np.uint8)
    stick = rng.random(total) < 0.55
    stick[0] = False
    idx = np.arange(total)
    last = np.maximum.accumulate(np.where(~stick, idx, 0))
    mu = mu[last]
    xyz = _walk_all(rng, lens)
    # self-reverse scores are per-chain inputs of the hot path (alignpair.cpp:7-25); plausible magnitude ~ 0.08*L
    selfrev = (selfrev_scale * lens * (0.5 + rng.random(n))).astype(np.float32)
    return SynthChains(lens.astype(np.uint32), prof, mu, xyz, selfrev)


def plant_homologs(db, queries, frac, seed, sub=0.30, indel=0.05, bg=None):
    """Overwrite a fraction of db chains (same length class) with mutated copies of query chains.
    Works chain by chain, keeping each db chain's length, so the SoA layout is unchanged."""
    rng = np.random.default_rng(seed)
    bg = bg or background()
    nplant = int(round(frac * db.n))
    targets = rng.choice(db.n, size=nplant, replace=False) if nplant else []
    for t in targets:
        q = int(rng.integers(0, queries.n))
        qp, qm, qx, _ = queries.chain(q)
        Lq = qp.shape[1]
        s, e = int(db.off[t]), int(db.off[t + 1])
        Lt = e - s
        # walk through the query with indels until Lt residues are produced
        src = []
        i = int(rng.integers(0, max(1, Lq // 8)))
        while len(src) < Lt:
            r = rng.random()
            if r < indel / 2:
                i += int(rng.integers(1, 4))  # deletion in the copy
            elif r < indel:
                src.append(-1)  # insertion: a random residue
                continue
            src.append(i if i < Lq else -1)
            i += 1
        src = np.array(src[:Lt])
        rnd = (src < 0) | (rng.random(Lt) < sub)
        keep = ~rnd
        for f in range(NFEAT):
            col = db.prof[f, s:e]
            col[keep] = qp[f, src[keep]]
        m = db.mu[s:e]
        m[keep] = qm[src[keep]]
        # coordinates: copy the query geometry where aligned so LDDT is meaningful
        x = db.xyz[:, s:e]
        x[:, keep] = qx[:, src[keep]]
    return db
