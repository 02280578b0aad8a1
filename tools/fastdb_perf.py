"""Developer tool: device timing of the `-fast -db` pipeline (prefilter K6-K8 + post-filter) on synthetic chains."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import reseek_b200 as rb
from reseek_b200 import synth

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ndb = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
L = int(sys.argv[3]) if len(sys.argv) > 3 else 300
q = synth.make_chains(nq, L, seed=1)
db = synth.make_chains(ndb, L, seed=2)
synth.plant_homologs(db, q, 0.01, seed=3)
ctx = rb.Context(0, rb.MODE_FAST)
Q = ctx.upload(q.lens, q.prof, q.mu, q.xyz, q.selfrev)
T = ctx.upload(db.lens, db.prof, db.mu, db.xyz, db.selfrev)
for rep in range(2):
    ctx.sync()
    t0 = time.perf_counter()
    pf = ctx.prefilter(Q, T)
    ctx.sync()
    t1 = time.perf_counter()
    res = ctx.postfilter(Q, T, pf, keep=rb.KEEP_HITS, want_paths=True)
    t2 = time.perf_counter()
    st = ctx.stats()
    print(f"rep {rep}: prefilter {1e3*(t1-t0):.1f} ms ({nq*ndb/(t1-t0):.3e} pairs/s, raw {pf.raw_count}, cands {len(pf)}), "
          f"postfilter {1e3*(t2-t1):.1f} ms, hits {len(res.hits)}, total pairs/s {nq*ndb/(t2-t0):.3e}")
