"""Developer tool: the five BASELINE.json configurations on one GPU (per-GPU share for the 8-GPU ones), synthetic data.
Prints one line per configuration: wall-clock of the public search call (host buffers in, hits out) and device-resident time."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import reseek_b200 as rb
from reseek_b200 import synth


def run_cross(name, mode, nq, ndb, L, seed, reps=2):
    q = synth.make_chains(nq, L, seed=seed)
    db = synth.make_chains(ndb, L, seed=seed + 1)
    synth.plant_homologs(db, q, 0.01, seed=seed + 2)
    ctx = rb.Context(0, mode)
    Q = ctx.upload(q.lens, q.prof, q.mu, q.xyz, q.selfrev)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        D = ctx.upload(db.lens, db.prof, db.mu, db.xyz, db.selfrev)
        res = ctx.search_cross(D, Q, keep=rb.KEEP_HITS, want_paths=True)
        wall = time.perf_counter() - t0
        st = ctx.stats()
        nh = len(res.hits)
        del res
        ctx.search_cross_device(D, Q)
        sd = ctx.stats()
        D.free()
        best = (wall, st, sd, nh) if best is None or wall < best[0] else best
    wall, st, sd, nh = best
    pairs = nq * ndb
    cells = float(np.sum(q.lens, dtype=np.float64)) * float(np.sum(db.lens, dtype=np.float64))
    print(f"{name}: {pairs:.3g} pairs, {cells:.3g} cells | e2e {wall*1e3:.1f} ms = {pairs/wall:.3e} pairs/s, {cells/wall:.3e} cells/s | "
          f"device {sd['total_ms']:.1f} ms = {pairs/(sd['total_ms']*1e-3):.3e} pairs/s, {cells/(sd['total_ms']*1e-3):.3e} cells/s "
          f"(mu {sd['mu_kernel_ms']:.1f}, sw {sd['sw_kernel_ms']:.1f}, lddt {sd['lddt_kernel_ms']:.1f}) | sw pairs {st['sw_pairs']}, hits {nh}", flush=True)
    ctx.close()


def run_self(name, mode, n, seed):
    rng = np.random.default_rng(seed)
    lens = np.clip(np.exp(rng.normal(np.log(143), 0.55, size=n)), 30, 1419).astype(np.int64)
    s = synth.make_chains(n, lens, seed=seed)
    synth.plant_homologs(s, s.subset(range(50)), 0.01, seed=seed + 1)
    ctx = rb.Context(0, mode)
    S = ctx.upload(s.lens, s.prof, s.mu, s.xyz, s.selfrev)
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        res = ctx.search_self(S, keep=rb.KEEP_HITS, want_paths=True)
        wall = time.perf_counter() - t0
        st = ctx.stats()
        nh = len(res.hits)
        del res
        best = (wall, st, nh) if best is None or wall < best[0] else best
    wall, st, nh = best
    print(f"{name}: {st['pairs']:.3g} pairs (mean L {lens.mean():.0f}) | e2e {wall*1e3:.1f} ms = {st['pairs']/wall:.3e} pairs/s | kernels: mu {st['mu_kernel_ms']:.0f} ms, "
          f"sw {st['sw_kernel_ms']:.0f}, long-chain {st['mkf_kernel_ms']:.0f}, lddt {st['lddt_kernel_ms']:.0f} | sw pairs {st['sw_pairs']}, long-chain pairs {st['mkf_pairs']}, hits {nh}", flush=True)
    ctx.close()


if __name__ == "__main__":
    run_cross("c2  -search 1 x 1e3, L=300, -sensitive", rb.MODE_SENSITIVE, 1, 1000, 300, 20260119)
    run_self("c3  all-vs-all 11211 chains (SCOP40-like lengths), -fast", rb.MODE_FAST, 11211, 20260120)
    run_cross("c4  -search 100 x 1.25e5 (1/8 of the 1e6-chain DB), L=300, -sensitive", rb.MODE_SENSITIVE, 100, 125000, 300, 20260121)
    run_cross("c5a -verysensitive 100 x 1e5, L=100", rb.MODE_VERYSENSITIVE, 100, 100000, 100, 20260122)
    run_cross("c5b -verysensitive 100 x 1e5, L=300", rb.MODE_VERYSENSITIVE, 100, 100000, 300, 20260122)
    run_cross("c5c -verysensitive 100 x 1e5, L=800", rb.MODE_VERYSENSITIVE, 100, 100000, 800, 20260122, reps=1)
