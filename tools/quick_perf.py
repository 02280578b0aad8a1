"""Quick device-side timing of the search kernels (developer tool; the contract benchmark is bench.py)."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import reseek_b200 as rb
from reseek_b200 import synth

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ndb = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
L = int(sys.argv[3]) if len(sys.argv) > 3 else 300
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
mode = int(sys.argv[5]) if len(sys.argv) > 5 else rb.MODE_VERYSENSITIVE
q = synth.make_chains(nq, L, seed=1)
db = synth.make_chains(ndb, L, seed=2)
synth.plant_homologs(db, q, 0.01, seed=3)
ctx = rb.Context(0, mode)
t0 = time.time()
A = ctx.upload(db.lens, db.prof, db.mu, db.xyz, db.selfrev)
B = ctx.upload(q.lens, q.prof, q.mu, q.xyz, q.selfrev)
print("upload s", time.time() - t0)
for rep in range(reps):
    t0 = time.time()
    ctx.search_cross_device(A, B)
    w = time.time() - t0
    st = ctx.stats()
    cells = st["sw_cells"]
    print(f"rep {rep}: wall {w:.3f}s total_ms {st['total_ms']:.1f} mu_ms {st['mu_kernel_ms']:.1f} sw_ms {st['sw_kernel_ms']:.1f} lddt_ms {st['lddt_kernel_ms']:.1f} "
          f"pairs {st['pairs']} cells {cells:.3e} sw cells/s {cells / (st['sw_kernel_ms'] * 1e-3):.3e} "
          f"e2e-dev cells/s {cells / (st['total_ms'] * 1e-3):.3e} pairs/s {st['pairs'] / (st['total_ms'] * 1e-3):.3e} "
          f"mu cells/s {2 * float(np.sum(db.lens, dtype=np.float64)) * float(np.sum(q.lens, dtype=np.float64)) / max(st['mu_kernel_ms'], 1e-9) / 1e-3:.3e}")
t0 = time.time()
res = ctx.search_cross(A, B, keep=rb.KEEP_HITS, want_paths=True)
print("full search with D2H wall", time.time() - t0, "hits", len(res.hits), ctx.stats())
