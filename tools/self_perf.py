"""Developer tool: all-vs-all (RunSelf) timing on a SCOP40-like synthetic set (BASELINE config 3)."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import reseek_b200 as rb
from reseek_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 11211
mode = int(sys.argv[2]) if len(sys.argv) > 2 else rb.MODE_FAST
rng = np.random.default_rng(20260120)
# SCOP40-like length law (mean 174, median 143, max 1419): log-normal clipped
lens = np.clip(np.exp(rng.normal(np.log(143), 0.55, size=n)), 30, 1419).astype(np.int64)
s = synth.make_chains(n, lens, seed=20260120)
synth.plant_homologs(s, s.subset(range(50)), 0.01, seed=5)
ctx = rb.Context(0, mode)
S = ctx.upload(s.lens, s.prof, s.mu, s.xyz, s.selfrev)
api = sys.argv[3] if len(sys.argv) > 3 else "self"   # "self": rsk_search_self, "sink": rsk_search_self_sharded with one rank
for rep in range(3):
    t0 = time.perf_counter()
    if api == "sink":
        res = ctx.search_self_sharded(None, S, keep=rb.KEEP_HITS, want_paths=True)
    else:
        res = ctx.search_self(S, keep=rb.KEEP_HITS, want_paths=True)
    dt = time.perf_counter() - t0
    st = ctx.stats()
    print(f"{api} rep {rep}: n {n} mean L {lens.mean():.0f} pairs {st['pairs']} wall {dt:.2f} s -> {st['pairs']/dt:.3e} pairs/s; device total {st['total_ms']:.0f} ms "
          f"(mu {st['mu_kernel_ms']:.0f}, sw {st['sw_kernel_ms']:.0f}, mkf {st['mkf_kernel_ms']:.0f}, lddt {st['lddt_kernel_ms']:.0f}); "
          f"sw_pairs {st['sw_pairs']} mkf_pairs {st['mkf_pairs']} hits {len(res.hits)}", flush=True)
    del res
