#!/usr/bin/env python3
"""bench.py - benchmark of the B200 Reseek hot path (contract: see the task statement / DESIGN.md §6).

  python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...   the reference's own CPU implementation of the path

Headline workload (BASELINE.json north_star / configs[4], SURVEY §8 "c5"): `-verysensitive` full float SW + traceback +
LDDT/E-value of Q=100 synthetic query chains (L=300) against a synthetic DB of 1e5 chains (L=300) PER GPU (weak scaling:
every rank owns one 1e5-chain DB shard, queries replicated, no data-path collective).  One step = one pass of the hot path
over that shard: 1e7 chain pairs, 9e11 DP cells.

Printed JSON (one line, rank 0):
  value    = SW residue-cells/s, whole job, inputs resident in HBM, device-timed (CUDA events, max over ranks)
  e2e      = same metric through the public C-ABI call with HOST buffers: per step the DB shard is uploaded from pinned
             host memory and the hit records + paths are read back (wall clock around the call)
  roofline = dominant kernel (sw_affine_f32_tb) algorithmic bytes / its CUDA-event time vs measured HBM peak
  cpu_baseline = the reference's CPU path (oracle/_ref, all host threads) on a bounded sample of the workload
  sensitive    = chain-pairs/s of the same shard under -sensitive (Mu filter decides), device + e2e + CPU sample
  legs     = the other BASELINE configs, each device-timed + e2e (+ roofline where the SW kernel dominates):
             c5_L100 / c5_L800 (weak), c4_strong (100 x ONE fixed 1e6-chain DB, -sensitive, block-partitioned over the
             ranks, hits gathered on rank 0 over NVLink INSIDE the timed region, digest compared with the single-GPU
             search), fastdb_sharded (-fast -db: prefilter triples all-gathered, merged bag, hits gathered; digest check),
             c3 (SCOP40-length all-vs-all -fast, rows of the pair triangle interleaved over the ranks; digest check)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SEED = 20260117 + 5  # seed = 20260117 + config_id
SEED_C4 = 20260117 + 4
SEED_C3 = 20260117 + 3
NQ, NDB, L = 100, 100_000, 300
METRIC = "sw_residue_cells_per_s"
UNIT = "cells/s"
# CPU sample (cpu_baseline and --impl reference): first CPU_NQ queries x first CPU_NDB DB chains of the same workload
CPU_NQ, CPU_NDB = 8, 5000
C4_NDB, C4_SEG = 1_000_000, 25_000
FAST_NDB = 200_000


def workload(rank, nq=NQ, ndb=NDB, length=L):
    from reseek_b200 import synth
    q = synth.make_chains(nq, length, seed=SEED)
    db = synth.make_chains(ndb, length, seed=SEED + 1000 * (rank + 1) + length)
    synth.plant_homologs(db, q, 0.01, seed=SEED + 7 + rank)
    return q, db


def concat_chains(parts):
    from reseek_b200.synth import SynthChains
    return SynthChains(np.concatenate([p.lens for p in parts]), np.concatenate([p.prof for p in parts], axis=1),
                       np.concatenate([p.mu for p in parts]), np.concatenate([p.xyz for p in parts], axis=1),
                       np.concatenate([p.selfrev for p in parts]))


def slice_chains(s, lo, hi):
    from reseek_b200.synth import SynthChains
    a, b = int(s.off[lo]), int(s.off[hi])
    return SynthChains(s.lens[lo:hi], s.prof[:, a:b], s.mu[a:b], s.xyz[:, a:b], s.selfrev[lo:hi])


def c4_queries(nq=NQ, length=L):
    from reseek_b200 import synth
    return synth.make_chains(nq, length, seed=SEED_C4)


def c4_block(q, lo, hi, seg=C4_SEG, length=L):
    """Chains [lo, hi) of the ONE fixed c4 database: segment s (seg chains) is generated from seed SEED_C4 + 1 + s whatever
    the rank count, so every rank count partitions the same database."""
    from reseek_b200 import synth
    s0, s1 = lo // seg, (hi + seg - 1) // seg

    def make(s):
        d = synth.make_chains(seg, length, seed=SEED_C4 + 1 + s)
        synth.plant_homologs(d, q, 0.01, seed=SEED_C4 + 100003 + s)
        return d
    with ThreadPoolExecutor(max(1, min(host_threads(), s1 - s0, 16))) as ex:
        parts = list(ex.map(make, range(s0, s1)))
    full = concat_chains(parts) if len(parts) > 1 else parts[0]
    return slice_chains(full, lo - s0 * seg, hi - s0 * seg)


def algorithmic_bytes(lens_a, lens_b):
    """per pair 8*(LA+LB) profile bytes + LA*LB trace (we pack 4 bits/cell = 0.5 B) + <=(LA+LB) path + 64 B record."""
    sa, sb = float(np.sum(lens_a, dtype=np.float64)), float(np.sum(lens_b, dtype=np.float64))
    na, nb = len(lens_a), len(lens_b)
    cells = sa * sb
    per_pair_lin = 9.0 * (sa * nb + sb * na)
    return 0.5 * cells + per_pair_lin + 64.0 * na * nb


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu = gpu
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_sample_run(q, db, steps=1, warmup=0, mode=3, fast_build=False):
    """The reference's CPU path (oracle/_ref when present, else the scalar oracle port) on the bounded sample, all host
    threads.  mode 3 = -verysensitive (every pair: SetSMx + SWFast + traceback + LDDT + E-value), 2 = -sensitive (Mu filter
    first).  Returns a dict with cells/s, pairs/s, kind, cores, sample."""
    from oracle.pyoracle import Port, Ref
    qs = q.subset(range(min(CPU_NQ, q.n)))
    ds = db.subset(range(min(CPU_NDB, db.n)))
    ia = np.repeat(np.arange(ds.n, dtype=np.uint32), qs.n)
    ib = np.tile(np.arange(qs.n, dtype=np.uint32), ds.n)
    cells = float(np.sum(ds.lens[ia].astype(np.float64) * qs.lens[ib].astype(np.float64)))
    what = "-verysensitive: SetSMx+SWFast+traceback+LDDT+E-value per pair" if mode == 3 else \
        "-sensitive: parasail Mu filter, then SetSMx+SWFast+traceback+LDDT+E-value for the survivors"
    sample = f"first {ds.n} DB chains x first {qs.n} queries of the same workload = {len(ia)} pairs, {cells:.3g} cells per step ({what})"
    if Ref.available(fast_build):
        ref = Ref(mode, fast=fast_build)
        cores = host_threads()
        kind = "reference"
        build = "-O3 -march=x86-64-v3 (oracle/Makefile ref_fast)" if fast_build else "-O2 -ffp-contract=off -mavx2 (strict, canonical)"

        def run():
            ref.align_batch(ds, qs, ia, ib, cores)
    else:
        if fast_build:
            return None
        port = Port(mode)
        cores = 1
        kind = "port"
        build = "gcc -O2 -ffp-contract=off (oracle/reseek_oracle.c)"
        from tests.util import to_oracle_chains
        ca, cb = to_oracle_chains(ds), to_oracle_chains(qs)

        def run():
            port.align_pairs(ca, cb, ia, ib)
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = (time.perf_counter() - t0) / steps
    return {"cells_per_s": cells / dt, "pairs_per_s": len(ia) / dt, "kind": kind, "cores": cores, "sample": sample,
            "ms_per_step": dt * 1e3, "build": build}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    q, db = workload(0, nq=CPU_NQ, ndb=CPU_NDB)
    r = cpu_sample_run(q, db, steps=args.steps, warmup=args.warmup, mode=3)
    out = {"impl": "reference", "metric": METRIC, "value": r["cells_per_s"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "chain_pairs_per_s": r["pairs_per_s"],
           "config": {"workload": f"c5 -verysensitive full SW: Q={NQ} x DB={NDB}/GPU, L={L} (CPU arm runs a bounded sample per step)",
                      "mode": "verysensitive", "sample": r["sample"]},
           "cpu_baseline": {"value": r["cells_per_s"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                            "build": r["build"]},
           "e2e": {"value": r["cells_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    # the other half of the metric on the same host cores: chain pairs/s under -sensitive (the Mu filter decides)
    steps_s = max(1, min(args.steps, 3))
    rs = cpu_sample_run(q, db, steps=steps_s, warmup=min(args.warmup, 1), mode=2)
    out["sensitive"] = {"chain_pairs_per_s": rs["pairs_per_s"], "ms_per_step": rs["ms_per_step"], "cores": rs["cores"],
                        "kind": rs["kind"], "sample": rs["sample"], "build": rs["build"]}
    rf = cpu_sample_run(q, db, steps=steps_s, warmup=min(args.warmup, 1), mode=3, fast_build=True)
    if rf:
        out["fast_build"] = {"value": rf["cells_per_s"], "unit": UNIT, "chain_pairs_per_s": rf["pairs_per_s"], "cores": rf["cores"],
                             "build": rf["build"], "note": "same unmodified sources, speed-fairness build; not used for parity"}
        rfs = cpu_sample_run(q, db, steps=steps_s, warmup=0, mode=2, fast_build=True)
        if rfs:
            out["fast_build"]["sensitive_chain_pairs_per_s"] = rfs["pairs_per_s"]
    emit(out)


def emit(obj):
    """The one JSON line of the contract, on the process's ORIGINAL stdout (see main)."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


_REAL_STDOUT = 1


class Bench:
    """Per-rank state shared by the legs."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import reseek_b200 as rb
        self.torch, self.dist, self.rb, self.args = torch, dist, rb, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if rb.device_count() < 1:
            raise SystemExit("bench.py: no CUDA device; libreseek_b200 has no CPU fallback")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.ctx = rb.Context(self.local, rb.MODE_VERYSENSITIVE)
        self.comm = rb.Comm.from_torch_dist(self.ctx, dist) if self.world > 1 else rb.Comm(self.ctx, 1, 0)
        self.launches = 0

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, vals, op="max"):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return t.tolist()

    def pin(self, s):
        t = {k: self.torch.from_numpy(getattr(s, k)).pin_memory() for k in ("lens", "prof", "mu", "xyz", "selfrev")}
        return t, {k: v.numpy() for k, v in t.items()}

    def upload(self, d):
        return self.ctx.upload(d["lens"], d["prof"], d["mu"], d["xyz"], d["selfrev"])

    def upload_chains(self, s):
        return self.ctx.upload(s.lens, s.prof, s.mu, s.xyz, s.selfrev)

    def timed(self, fn, steps, warmup):
        """W untimed calls, then K calls between barrier+synchronize; device time by CUDA events, max over ranks (ms per step).
        Every library call ends synchronised with its stream, so the events bracket exactly the K calls."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        wall = (time.perf_counter() - t0) / steps
        dev = e0.elapsed_time(e1) / steps
        dev_max, wall_max = self.reduce([dev, wall * 1e3])
        return dev_max, wall_max

    def mode(self, m):
        self.ctx.set_params(self.rb.params_preset(m))


def collective_block(B, comm, steps, step_ms):
    """Per-step figures of the communicator's data-path collectives since reset_stats(): payload put on NVLink by all ranks, and
    the device time of the gather rounds (CUDA events on the context stream).  A round cannot finish before the slowest rank
    has arrived, so the time on a rank that arrived early is mostly waiting: `ms_per_step_min` (the rank that arrived last) is
    the transfer itself, `ms_per_step_max` includes the load imbalance between the ranks."""
    cs = comm.stats()
    ms = cs["collective_ms"] / max(1, steps)
    mx = B.reduce([ms])[0]
    mn = -B.reduce([-ms])[0]
    sent = B.reduce([float(cs["bytes_sent"]) / max(1, steps)], "sum")[0]
    return {"ms_per_step_min": mn, "ms_per_step_max": mx, "nvlink_bytes_per_step": sent, "rounds_per_step": cs["collectives"] / max(1, steps),
            "share_of_step": mn / step_ms if step_ms else None}


def roofline_block(B, db_lens, q_lens, sw_ms_step, sw_launches_step, dev_ms_step, clocks, cells_rank):
    peak_src, peak = "fallback (B200_PROFILING.md)", 6650.0
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peak = float(json.loads(pk.read_text()).get("hbm_gbs", peak))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    sw_launches_step = max(1, sw_launches_step)
    alg_bytes = algorithmic_bytes(db_lens, q_lens) / sw_launches_step
    sw_ms_launch = sw_ms_step / sw_launches_step
    achieved = alg_bytes / (sw_ms_launch * 1e-3) / 1e9
    smem_peak = 148 * 128 * clocks["sm_mhz"] * 1e6 / 1e9 if clocks and clocks.get("sm_mhz") else None  # GB/s: 128 B/clk/SM
    # measured DRAM bytes of the kernel (ncu dram__bytes_read + write, profiles/sw_kernel_traffic.json): per cell for the chain
    # length that was captured, scaled to this leg's launch size; null for lengths without a capture
    traffic = None
    tf = ROOT / "profiles" / "sw_kernel_traffic.json"
    if tf.exists():
        try:
            per_cell = json.loads(tf.read_text()).get("dram_bytes_per_cell_by_chain_length", {})
            key = str(int(round(float(np.mean(q_lens)))))
            if key in per_cell:
                traffic = float(per_cell[key]) * cells_rank / sw_launches_step
        except Exception:
            traffic = None
    smem_ach = 32.0 * cells_rank / (sw_ms_step * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "kernel": "sw_affine_f32_tb_kernel", "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms_per_launch": sw_ms_launch,
            "launches_per_step": sw_launches_step, "kernel_share_of_step": sw_ms_step / dev_ms_step,
            "kernel_cells_per_s": cells_rank / (sw_ms_step * 1e-3),
            "binding_resource": {"name": "shared-memory bandwidth (score-table reads, 32 B/cell)", "unit": "GB/s",
                                 "achieved": smem_ach, "peak": smem_peak, "frac": (smem_ach / smem_peak) if smem_peak else None},
            "note": "the DP is bound by shared-memory table reads, not HBM (profiles/); HBM fraction reported as SURVEY §8d requires"}


def leg_c5(B, length, ndb, steps, warmup, e2e_steps, sample_clocks=False):
    """-verysensitive full SW over this rank's shard (weak scaling).  Returns (leg dict, q, db)."""
    rb, ctx, args = B.rb, B.ctx, B.args
    q, db = workload(B.rank, args.nq, ndb, length)
    B.mode(rb.MODE_VERYSENSITIVE)
    pin, dbp = B.pin(db)
    Q = B.upload_chains(q)
    D = B.upload(dbp)
    cells_rank = float(np.sum(db.lens, dtype=np.float64)) * float(np.sum(q.lens, dtype=np.float64))
    pairs_rank = db.n * q.n
    acc = {"sw_ms": 0.0, "lddt_ms": 0.0, "launches": 0, "sw_launches": 0, "n": 0}

    def step():
        ctx.search_cross_device(D, Q)
        st = ctx.stats()
        acc["sw_ms"] += st["sw_kernel_ms"]; acc["lddt_ms"] += st["lddt_kernel_ms"]
        acc["launches"] += st["kernel_launches"]; acc["sw_launches"] += st["sw_kernel_launches"]; acc["n"] += 1

    for _ in range(warmup):
        ctx.search_cross_device(D, Q)
    sampler = ClockSampler(B.local) if sample_clocks else None
    if sampler:
        B.barrier()
        sampler.start()
    dev_ms, _ = B.timed(step, steps, 0)
    clocks = sampler.stop() if sampler else None
    sw_ms_step = B.reduce([acc["sw_ms"] / steps])[0]
    cells_all, pairs_all, launches_all = B.reduce([cells_rank, float(pairs_rank), float(acc["launches"])], "sum")
    B.launches += acc["launches"]
    value = cells_all / (dev_ms * 1e-3)

    # e2e: host buffers in, hits + paths out, every step
    D.free()
    io = {"h2d": 0, "d2h": 0, "hits": 0}

    def e2e_step():
        Dk = B.upload(dbp)
        res = ctx.search_cross(Dk, Q, keep=rb.KEEP_HITS, want_paths=True)
        st = ctx.stats()
        io["h2d"] += Dk.h2d_bytes + st["h2d_bytes"]; io["d2h"] += st["d2h_bytes"]; io["hits"] += len(res.hits)
        B.launches += st["kernel_launches"]
        Dk.free()
        del res

    for _ in range(max(1, min(warmup, 2))):  # untimed: pinned staging buffers and host result blocks get allocated
        e2e_step()
    io.update(h2d=0, d2h=0, hits=0)
    _, e2e_ms = B.timed(e2e_step, e2e_steps, 0)
    h2d, d2h, nhits = B.reduce([float(io["h2d"]), float(io["d2h"]), float(io["hits"])], "sum")
    leg = {"workload": f"-verysensitive full SW+traceback+LDDT/E-value: Q={q.n} x DB={ndb} chains per GPU, L={length}",
           "scaling": "weak", "value": value, "unit": UNIT, "ms_per_step": dev_ms, "chain_pairs_per_s": pairs_all / (dev_ms * 1e-3),
           "steps": steps, "pairs_per_step": pairs_all, "cells_per_step": cells_all,
           "e2e": {"value": cells_all / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": h2d / e2e_steps, "d2h_bytes_per_step": d2h / e2e_steps, "hits_per_step": nhits / e2e_steps,
                   "steps": e2e_steps, "api": "rsk_chainset_upload + rsk_search_cross(keep=HITS, paths) from pinned host buffers"},
           "roofline": roofline_block(B, db.lens, q.lens, sw_ms_step, acc["sw_launches"] // max(1, acc["n"]), dev_ms, clocks, cells_rank)}
    return leg, q, db, dbp, Q, clocks, launches_all, pin


def leg_sensitive(B, q, db, dbp, Q, steps, warmup, e2e_steps):
    """Second half of the metric: chain pairs/s where the Mu filter decides (`-search Q -db DB -sensitive`), same shard."""
    rb, ctx = B.rb, B.ctx
    B.mode(rb.MODE_SENSITIVE)
    D = B.upload(dbp)
    cells_rank = float(np.sum(db.lens, dtype=np.float64)) * float(np.sum(q.lens, dtype=np.float64))
    acc = {"mu_ms": 0.0, "sw_pairs": 0.0}

    def step():
        ctx.search_cross_device(D, Q)
        st = ctx.stats()
        acc["mu_ms"] += st["mu_kernel_ms"]; acc["sw_pairs"] += st["sw_pairs"]
        B.launches += st["kernel_launches"]

    for _ in range(warmup):
        step()
    acc["mu_ms"] = 0.0; acc["sw_pairs"] = 0.0
    dev_ms, _ = B.timed(step, steps, 0)
    mu_ms = B.reduce([acc["mu_ms"] / steps])[0]
    pairs_all, cells_all = B.reduce([float(db.n * q.n), cells_rank], "sum")
    D.free()
    io = {"h2d": 0, "d2h": 0, "hits": 0}

    def e2e_step():
        Dk = B.upload(dbp)
        res = ctx.search_cross_sharded(None, Dk, Q, 0, keep=rb.KEEP_HITS, want_paths=True)
        st = ctx.stats()
        io["h2d"] += Dk.h2d_bytes + st["h2d_bytes"]; io["d2h"] += st["d2h_bytes"]; io["hits"] += len(res.hits)
        B.launches += st["kernel_launches"]
        Dk.free()
        del res

    e2e_step()
    io.update(h2d=0, d2h=0, hits=0)
    _, e2e_ms = B.timed(e2e_step, e2e_steps, 0)
    h2d, d2h, nhits = B.reduce([float(io["h2d"]), float(io["d2h"]), float(io["hits"])], "sum")
    B.mode(rb.MODE_VERYSENSITIVE)
    return {"chain_pairs_per_s": pairs_all / (dev_ms * 1e-3), "ms_per_step": dev_ms, "steps": steps,
            "mu_filter_cells_per_s": 2.0 * cells_all / (mu_ms * 1e-3) if mu_ms > 0 else None,
            "mu_filter_share_of_step": mu_ms / dev_ms, "sw_pairs_share": acc["sw_pairs"] / (db.n * q.n * steps),
            "e2e": {"chain_pairs_per_s": pairs_all / (e2e_ms * 1e-3), "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d / e2e_steps,
                    "d2h_bytes_per_step": d2h / e2e_steps, "hits_per_step": nhits / e2e_steps, "steps": e2e_steps,
                    "api": "rsk_chainset_upload + rsk_search_cross_sharded(comm=NULL, keep=HITS, paths): hits compacted on the device, "
                           "only the hit set crosses PCIe"},
            "note": "same shard and queries under the -sensitive preset (Mu int8 SW filter fwd+rev, then float SW for the survivors), "
                    "device-resident, CUDA events; filter cells counted as 2*LA*LB per pair"}


def leg_c4_strong(B, steps, warmup):
    """c4: 100 queries x ONE fixed 1e6-chain DB, -sensitive, block-partitioned by residues over the ranks (strong scaling).
    The hit gather to rank 0 (NCCL, exact sizes, device to device) is inside both timed regions."""
    rb, ctx, comm, args = B.rb, B.ctx, B.comm, B.args
    ndb = args.c4_ndb
    q = c4_queries(args.nq)
    # fixed-length chains: the residue-balanced partition is the chain-count partition
    lo, hi = ndb * B.rank // B.world, ndb * (B.rank + 1) // B.world
    t0 = time.perf_counter()
    blk = c4_block(q, lo, hi, seg=min(C4_SEG, ndb))
    gen_s = time.perf_counter() - t0
    pin, bp = B.pin(blk)
    B.mode(rb.MODE_SENSITIVE)
    Q = B.upload_chains(q)
    D = B.upload(bp)
    pairs_all = float(ndb) * q.n
    out = {"digest": None, "hits": 0}

    def dev_step():
        res = ctx.search_cross_sharded(comm, D, Q, lo, keep=rb.KEEP_HITS, want_paths=True)
        B.launches += ctx.stats()["kernel_launches"]
        if res is not None:
            out["digest"], out["hits"] = res.digest(), len(res.hits)
        del res

    for _ in range(warmup):  # also establishes the NCCL point-to-point connections
        dev_step()
    comm.reset_stats()
    dev_ms, _ = B.timed(dev_step, steps, 0)
    coll = collective_block(B, comm, steps, dev_ms)
    D.free()
    io = {"h2d": 0, "d2h": 0}

    def e2e_step():
        Dk = B.upload(bp)
        res = ctx.search_cross_sharded(comm, Dk, Q, lo, keep=rb.KEEP_HITS, want_paths=True)
        st = ctx.stats()
        io["h2d"] += Dk.h2d_bytes + st["h2d_bytes"]; io["d2h"] += st["d2h_bytes"]
        B.launches += st["kernel_launches"]
        Dk.free()
        del res

    e2e_step()
    io.update(h2d=0, d2h=0)
    _, e2e_ms = B.timed(e2e_step, steps, 0)
    h2d, d2h = B.reduce([float(io["h2d"]), float(io["d2h"])], "sum")
    # rank 0: the same search on ONE GPU through the host pipeline (rsk_search_cross), digests must agree
    check = None
    if B.rank == 0 and not args.no_digest_check:
        full = blk if B.world == 1 else c4_block(q, 0, ndb, seg=min(C4_SEG, ndb))
        Df = B.upload_chains(full)
        ref = ctx.search_cross(Df, Q, keep=rb.KEEP_HITS, want_paths=True)
        check = {"single_gpu_digest": "%016x" % ref.digest(), "sharded_digest": "%016x" % (out["digest"] or 0),
                 "single_gpu_hits": len(ref.hits), "sharded_hits": out["hits"], "equal": ref.digest() == out["digest"] and len(ref.hits) == out["hits"]}
        del ref
        Df.free()
        if not check["equal"]:
            raise SystemExit(f"bench.py: c4 sharded digest differs from the single-GPU search: {check}")
    B.barrier()
    B.mode(rb.MODE_VERYSENSITIVE)
    return {"workload": f"c4 -sensitive: Q={q.n} x ONE DB of {ndb} chains (L={L}), block-partitioned over {B.world} GPU(s)",
            "scaling": "strong", "metric": "chain_pairs_per_s", "value": pairs_all / (dev_ms * 1e-3), "unit": "pairs/s",
            "ms_per_step": dev_ms, "steps": steps, "pairs_per_step": pairs_all, "per_gpu_pairs": float(hi - lo) * q.n,
            "api": "rsk_search_cross_sharded (device hit sink + NCCL gather on rank 0), DB block resident",
            "collective": dict(coll, what="hit gather: 32-byte count all-gather + exact-size ncclSend/ncclRecv of records and path bytes to rank 0"),
            "e2e": {"value": pairs_all / (e2e_ms * 1e-3), "unit": "pairs/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d / steps,
                    "d2h_bytes_per_step": d2h / steps, "steps": steps,
                    "api": "rsk_chainset_upload(block, pinned host) + rsk_search_cross_sharded incl. gather + read-out on rank 0"},
            "digest_check": check, "db_generation_s": gen_s}, q, blk, bp, Q, lo, pin


def leg_fastdb(B, q, blk, lo, steps, warmup):
    """`-search Q -db DB -fast` on the first FAST_NDB chains of the c4 database, block-partitioned: prefilter triples
    all-gathered over NVLink, merged top-1500 bag on every rank, post-filter of the own candidates, hits gathered."""
    rb, ctx, comm, args = B.rb, B.ctx, B.comm, B.args
    ndb = min(args.fast_ndb, args.c4_ndb)
    flo, fhi = ndb * B.rank // B.world, ndb * (B.rank + 1) // B.world
    # this rank's share of [0, ndb) out of the c4 chains it already holds, else regenerate
    if flo >= lo and fhi <= lo + blk.n:
        mine = slice_chains(blk, flo - lo, fhi - lo)
    else:
        mine = c4_block(q, flo, fhi, seg=min(C4_SEG, args.c4_ndb))
    B.mode(rb.MODE_FAST)
    Q = B.upload_chains(q)
    T = B.upload_chains(mine)
    out = {"digest": None, "hits": 0, "cands": 0}

    def step():
        res, cands = ctx.search_fast_db_sharded(comm, Q, T, flo, keep=rb.KEEP_HITS, want_paths=True, want_cands=B.rank == 0)
        B.launches += ctx.stats()["kernel_launches"]
        if res is not None:
            out["digest"], out["hits"] = res.digest(), len(res.hits)
        if cands is not None:
            out["cands"] = len(cands)
        del res, cands

    for _ in range(warmup):
        step()
    comm.reset_stats()
    dev_ms, wall_ms = B.timed(step, steps, 0)
    coll = collective_block(B, comm, steps, dev_ms)
    check = None
    if B.rank == 0 and not args.no_digest_check:
        full = mine if B.world == 1 else c4_block(q, 0, ndb, seg=min(C4_SEG, args.c4_ndb))
        Tf = B.upload_chains(full)
        ref = ctx.search_fast_db(Q, Tf, keep=rb.KEEP_HITS, want_paths=True)
        check = {"single_gpu_digest": "%016x" % ref.digest(), "sharded_digest": "%016x" % (out["digest"] or 0),
                 "single_gpu_hits": len(ref.hits), "sharded_hits": out["hits"], "equal": ref.digest() == out["digest"] and len(ref.hits) == out["hits"]}
        del ref
        Tf.free()
        if not check["equal"]:
            raise SystemExit(f"bench.py: -fast -db sharded digest differs from the single-GPU search: {check}")
    B.barrier()
    T.free()
    B.mode(rb.MODE_VERYSENSITIVE)
    pairs_all = float(ndb) * q.n
    return {"workload": f"-fast -db: Q={q.n} x DB of {ndb} chains (L={L}), block-partitioned over {B.world} GPU(s): 5-mer prefilter, "
                        "merged top-1500 bag, post-filter under the sensitive preset",
            "scaling": "strong", "metric": "chain_pairs_per_s", "value": pairs_all / (wall_ms * 1e-3), "unit": "pairs/s",
            "ms_per_step": wall_ms, "device_ms_per_step": dev_ms, "steps": steps, "candidates": out["cands"], "hits": out["hits"],
            "api": "rsk_search_fast_db_sharded (inputs resident; candidate list + hits read out on rank 0)",
            "collective": dict(coll, what="all-gather of (query, target<<16|score) triples in rank order + hit gather on rank 0"),
            "digest_check": check}


def leg_c3(B, steps, warmup):
    """c3: all-vs-all `-search X -fast` on a SCOP40-sized synthetic set (11 211 chains, SCOP40-like lengths).  Every rank holds
    the set; the rows of the pair triangle are interleaved over the ranks, hits gathered on rank 0 (strong scaling)."""
    rb, ctx, comm, args = B.rb, B.ctx, B.comm, B.args
    from reseek_b200 import synth
    rng = np.random.default_rng(SEED_C3)
    # SCOP40-like lengths: log-normal fitted to mean 174 / median 143, clipped to [30, 1419] (SURVEY §8 sizes)
    lens = np.clip(np.exp(rng.normal(np.log(143.0), 0.62, size=11211)), 30, 1419).astype(np.int64)
    s = synth.make_chains(len(lens), lens, seed=SEED_C3 + 1)
    B.mode(rb.MODE_FAST)
    S = B.upload_chains(s)
    npairs = s.n * (s.n + 1) // 2
    info = {"digest": None, "hits": 0}

    def step():
        res = ctx.search_self_sharded(comm, S, keep=rb.KEEP_HITS, want_paths=True)
        st = ctx.stats()
        info.update(sw_pairs=st["sw_pairs"], sw_cells=st["sw_cells"], mkf_pairs=st["mkf_pairs"],
                    mu_ms=st["mu_kernel_ms"], sw_ms=st["sw_kernel_ms"], mkf_ms=st["mkf_kernel_ms"])
        B.launches += st["kernel_launches"]
        if res is not None:
            info["digest"], info["hits"] = res.digest(), len(res.hits)
        del res

    for _ in range(warmup):
        step()
    comm.reset_stats()
    dev_ms, wall_ms = B.timed(step, steps, 0)
    coll = collective_block(B, comm, steps, dev_ms)
    tot = B.reduce([float(info.get("sw_pairs", 0)), float(info.get("sw_cells", 0)), float(info.get("mkf_pairs", 0))], "sum")
    kms = B.reduce([info.get("mu_ms", 0.0), info.get("sw_ms", 0.0), info.get("mkf_ms", 0.0)])
    check = None
    if B.rank == 0 and not args.no_digest_check:
        ref = ctx.search_self(S, keep=rb.KEEP_HITS, want_paths=True)
        check = {"single_gpu_digest": "%016x" % ref.digest(), "sharded_digest": "%016x" % (info["digest"] or 0),
                 "single_gpu_hits": len(ref.hits), "sharded_hits": info["hits"], "equal": ref.digest() == info["digest"] and len(ref.hits) == info["hits"]}
        del ref
        if not check["equal"]:
            raise SystemExit(f"bench.py: c3 sharded digest differs from the single-GPU search: {check}")
    B.barrier()
    S.free()
    B.mode(rb.MODE_VERYSENSITIVE)
    return {"workload": f"c3 all-vs-all -search X -fast: {s.n} chains, SCOP40-like lengths (mean {float(np.mean(lens)):.0f}, max {int(lens.max())}), "
                        f"{npairs} pairs i<=j, rows interleaved over {B.world} GPU(s)", "scaling": "strong", "metric": "chain_pairs_per_s",
            "value": npairs / (wall_ms * 1e-3), "unit": "pairs/s", "ms_per_step": wall_ms, "device_ms_per_step": dev_ms, "steps": steps,
            "timing": "wall clock around rsk_search_self_sharded (plan + kernels + hit gather + read-out on rank 0), max over ranks",
            "sw_pairs": tot[0], "sw_cells": tot[1], "long_chain_pairs": tot[2], "hits": info["hits"],
            "kernel_ms_max_over_ranks": {"mu_filter": kms[0], "sw": kms[1], "long_chain": kms[2]},
            "collective": dict(coll, what="hit gather on rank 0"),
            "digest_check": check}


def main():
    # Libraries chat on stdout (NCCL prints its version banner there): keep stdout for the JSON line alone by pointing
    # file descriptor 1 at stderr for everything else.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--nq", type=int, default=NQ)
    ap.add_argument("--ndb", type=int, default=NDB)
    ap.add_argument("--len", type=int, default=L, dest="length")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--legs", default="sensitive,c5_L100,c5_L800,c4_strong,fastdb_sharded,c3",
                    help="comma list of the extra legs to run ('' = headline only)")
    ap.add_argument("--leg-steps", type=int, default=2)
    ap.add_argument("--c4-ndb", type=int, default=C4_NDB)
    ap.add_argument("--fast-ndb", type=int, default=FAST_NDB)
    ap.add_argument("--l800-ndb", type=int, default=25_000)
    ap.add_argument("--no-digest-check", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    legs_wanted = [x for x in args.legs.split(",") if x]
    t_start = time.perf_counter()
    B = Bench(args)
    ls = max(1, min(args.steps, args.leg_steps))

    head, q, db, dbp, Q, clocks, launches_head, pin = leg_c5(B, args.length, args.ndb, args.steps, args.warmup, max(1, args.e2e_steps),
                                                             sample_clocks=True)
    out = None
    legs = {}
    sens = None
    if "sensitive" in legs_wanted:
        sens = leg_sensitive(B, q, db, dbp, Q, max(1, min(args.steps, 5)), 1, ls)
    cpu = cpu_s = None
    if B.rank == 0 and not args.no_cpu_baseline:
        cpu = cpu_sample_run(q, db, steps=1, warmup=0, mode=3)
        if sens is not None:
            cpu_s = cpu_sample_run(q, db, steps=1, warmup=0, mode=2)
    del pin, dbp, db
    B.barrier()
    if "c5_L100" in legs_wanted:
        legs["c5_L100"] = leg_c5(B, 100, args.ndb, ls, 1, ls)[0]
    if "c5_L800" in legs_wanted:
        legs["c5_L800"] = leg_c5(B, 800, args.l800_ndb, ls, 1, ls)[0]
        legs["c5_L800"]["note"] = (f"DB of {args.l800_ndb} chains per GPU instead of 1e5 to bound the bench time "
                                   "(6.4e5 cells per pair: the per-cell rate does not depend on the DB size)")
    if "c4_strong" in legs_wanted or "fastdb_sharded" in legs_wanted:
        c4, q4, blk, bp, Q4, lo, pin4 = leg_c4_strong(B, ls, 1)
        if "c4_strong" in legs_wanted:
            legs["c4_strong"] = c4
        del bp, pin4
        if "fastdb_sharded" in legs_wanted:
            legs["fastdb_sharded"] = leg_fastdb(B, q4, blk, lo, ls, 1)
        del blk
    if "c3" in legs_wanted:
        legs["c3"] = leg_c3(B, 1, 1)

    launches_all = B.reduce([float(B.launches)], "sum")[0]
    if B.rank == 0:
        out = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": B.world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "chain_pairs_per_s": head["chain_pairs_per_s"],
            "config": {"workload": f"c5 -verysensitive full SW+traceback+LDDT/E-value: Q={args.nq} queries x DB={args.ndb} chains per GPU, L={args.length}",
                       "mode": "verysensitive", "pairs_per_step": head["pairs_per_step"], "cells_per_step": head["cells_per_step"],
                       "db_chains_total": args.ndb * B.world, "sharding": "DB shard per rank, queries replicated, no data-path collective "
                       "in the headline (legs.c4_strong / legs.fastdb_sharded time the gather collectives)",
                       "l2": "inputs larger than L2 (DB shard %.0f MB + %.0f MB checkpoint scratch per step)" % (631.0 * args.ndb / 1e5 * args.length / 300, 220.0),
                       "seed": SEED},
            "clocks": clocks, "e2e": head["e2e"], "gpu_launches": int(launches_all), "roofline": head["roofline"],
            "legs": legs, "bench_wall_s": None,
        }
        if sens is not None:
            out["sensitive"] = sens
        if cpu is not None:
            out["cpu_baseline"] = {"value": cpu["cells_per_s"], "unit": UNIT, "cores": cpu["cores"], "kind": cpu["kind"], "sample": cpu["sample"],
                                   "chain_pairs_per_s": cpu["pairs_per_s"], "build": cpu["build"]}
            if cpu_s is not None:
                out["cpu_baseline"]["sensitive_chain_pairs_per_s"] = cpu_s["pairs_per_s"]
                out["sensitive"]["cpu_baseline"] = {"chain_pairs_per_s": cpu_s["pairs_per_s"], "cores": cpu_s["cores"], "kind": cpu_s["kind"],
                                                    "sample": cpu_s["sample"], "build": cpu_s["build"]}
        out["bench_wall_s"] = time.perf_counter() - t_start
        emit(out)
    if B.world > 1:
        B.dist.barrier()
    B.comm.close()
    if B.world > 1:
        B.dist.destroy_process_group()
    B.ctx.close()


if __name__ == "__main__":
    main()
