#!/usr/bin/env python3
"""Throughput of rsk_align_global (K9) on synthetic chains: python tools/global_perf.py [npairs] [L]."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import reseek_b200 as rb  # noqa: E402
from reseek_b200 import synth  # noqa: E402

npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
n = 400
a = synth.make_chains(n, L, seed=11, length_jitter=0.3)
ctx = rb.Context(0, 3)
A = ctx.upload(a.lens, a.prof, a.mu, a.xyz, a.selfrev)
rng = np.random.default_rng(5)
ia = rng.integers(0, n, size=npairs).astype(np.uint32)
ib = rng.integers(0, n, size=npairs).astype(np.uint32)
cells = float(np.sum(a.lens[ia].astype(np.float64) * a.lens[ib]))
for rep in range(3):
    t0 = time.time()
    res = ctx.align_global(A, A, ia, ib)
    dt = time.time() - t0
    kms = ctx.stats()["sw_kernel_ms"]
    print(f"rep {rep}: {npairs} pairs, {cells:.3e} cells, wall {dt:.3f}s, {cells / dt:.3e} cells/s (end to end), kernel {kms:.1f} ms = {cells / (kms * 1e-3):.3e} cells/s, hits {len(res.hits)}")

# CPU side of the same thing: the oracle port (one thread) on a bounded sample of the same pairs
from oracle.pyoracle import Port  # noqa: E402
from tests.util import to_oracle_chains  # noqa: E402
oc = to_oracle_chains(a)
port = Port(mode=3)
m = min(300, npairs)
t0 = time.time()
for k in range(m):
    port.align_pair_global(oc[int(ia[k])], oc[int(ib[k])])
dt = time.time() - t0
c = float(np.sum(a.lens[ia[:m]].astype(np.float64) * a.lens[ib[:m]]))
print(f"cpu oracle (1 thread): {m} pairs, {c:.3e} cells, {dt:.3f}s, {c / dt:.3e} cells/s")
