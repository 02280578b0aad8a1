"""Developer tool: wall-clock breakdown of the e2e path of bench.py (upload, search, result handling)."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import reseek_b200 as rb
import bench

q, db = bench.workload(0)
stream = torch.cuda.current_stream().cuda_stream
ctx = rb.Context(0, rb.MODE_VERYSENSITIVE, stream=stream)
pin = {k: torch.from_numpy(getattr(db, k)).pin_memory() for k in ("lens", "prof", "mu", "xyz", "selfrev")}
dbp = {k: v.numpy() for k, v in pin.items()}
Q = ctx.upload(q.lens, q.prof, q.mu, q.xyz, q.selfrev)
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    Dk = ctx.upload(dbp["lens"], dbp["prof"], dbp["mu"], dbp["xyz"], dbp["selfrev"])
    t1 = time.perf_counter()
    res = ctx.search_cross(Dk, Q, keep=rb.KEEP_HITS, want_paths=True)
    t2 = time.perf_counter()
    st = ctx.stats()
    n = len(res.hits)
    t3 = time.perf_counter()
    Dk.free()
    del res
    t4 = time.perf_counter()
    print(f"rep {rep}: upload {1e3*(t1-t0):.1f} ms, search {1e3*(t2-t1):.1f} ms (device total {st['total_ms']:.1f}, sw {st['sw_kernel_ms']:.1f}), "
          f"results view {1e3*(t3-t2):.1f} ms, free {1e3*(t4-t3):.1f} ms, all {1e3*(t4-t0):.1f} ms, hits {n}")
