#!/usr/bin/env python3
"""bench.py - headline benchmark of the B200 Reseek hot path (contract: see the task statement / DESIGN.md §6).

  python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...   the reference's own CPU implementation of the path

Workload (BASELINE.json north_star / configs[4], SURVEY §8 "c5"): `-verysensitive` full float SW + traceback +
LDDT/E-value of Q=100 synthetic query chains (L=300) against a synthetic DB of 1e5 chains (L=300) PER GPU
(weak scaling: every rank owns one 1e5-chain DB shard, queries replicated, no data-path collective).
One step = one pass of the hot path over that shard: 1e7 chain pairs, 9e11 DP cells.

Printed JSON (one line, rank 0):
  value   = SW residue-cells/s, whole job, inputs resident in HBM, device-timed (CUDA events, max over ranks)
  e2e     = same metric through the public C-ABI call with HOST buffers: per step the DB shard is uploaded from
            pinned host memory and the hit records + paths are read back (wall clock around the call)
  roofline= dominant kernel (sw_affine_f32_tb) algorithmic bytes / its CUDA-event time vs measured HBM peak
  cpu_baseline = the reference's CPU path (oracle/_ref, all host threads) on a bounded sample of the workload
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SEED = 20260117 + 5  # seed = 20260117 + config_id
NQ, NDB, L = 100, 100_000, 300
METRIC = "sw_residue_cells_per_s"
UNIT = "cells/s"
# CPU sample (cpu_baseline and --impl reference): first CPU_NQ queries x first CPU_NDB DB chains of the same workload
CPU_NQ, CPU_NDB = 8, 5000


def workload(rank, nq=NQ, ndb=NDB, length=L):
    from reseek_b200 import synth
    q = synth.make_chains(nq, length, seed=SEED)
    db = synth.make_chains(ndb, length, seed=SEED + 1000 * (rank + 1))
    synth.plant_homologs(db, q, 0.01, seed=SEED + 7 + rank)
    return q, db


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


def cpu_sample_run(q, db, steps=1, warmup=0):
    """The reference's CPU path (oracle/_ref when present, else the scalar oracle port) on the bounded sample.
    Returns (cells/s, pairs/s, kind, cores, sample description, ms per step)."""
    from oracle.pyoracle import Port, Ref
    qs = q.subset(range(min(CPU_NQ, q.n)))
    ds = db.subset(range(min(CPU_NDB, db.n)))
    ia = np.repeat(np.arange(ds.n, dtype=np.uint32), qs.n)
    ib = np.tile(np.arange(qs.n, dtype=np.uint32), ds.n)
    cells = float(np.sum(ds.lens[ia].astype(np.float64) * qs.lens[ib].astype(np.float64)))
    sample = (f"first {ds.n} DB chains x first {qs.n} queries of the same workload = {len(ia)} pairs, "
              f"{cells:.3g} cells per step (-verysensitive: SetSMx+SWFast+traceback+LDDT+E-value per pair)")
    if Ref.available():
        ref = Ref(3)
        cores = host_threads()
        kind = "reference"

        def run():
            ref.align_batch(ds, qs, ia, ib, cores)
    else:
        port = Port(3)
        cores = 1
        kind = "port"
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
    return cells / dt, len(ia) / dt, kind, cores, sample, dt * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    q, db = workload(0, nq=CPU_NQ, ndb=CPU_NDB)
    cps, pps, kind, cores, sample, ms = cpu_sample_run(q, db, steps=args.steps, warmup=args.warmup)
    out = {"impl": "reference", "metric": METRIC, "value": cps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "chain_pairs_per_s": pps,
           "config": {"workload": f"c5 -verysensitive full SW: Q={NQ} x DB={NDB}/GPU, L={L} (CPU arm runs a bounded sample per step)",
                      "mode": "verysensitive", "sample": sample},
           "cpu_baseline": {"value": cps, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": cps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


def emit(obj):
    """The one JSON line of the contract, on the process's ORIGINAL stdout (see main)."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


_REAL_STDOUT = 1


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
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import reseek_b200 as rb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rb.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; libreseek_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    q, db = workload(rank, args.nq, args.ndb, args.length)
    stream = torch.cuda.current_stream().cuda_stream
    ctx = rb.Context(local, rb.MODE_VERYSENSITIVE, stream=stream)
    # host buffers of the streamed side live in pinned memory (e2e uploads them every step)
    pin = {k: torch.from_numpy(getattr(db, k)).pin_memory() for k in ("lens", "prof", "mu", "xyz", "selfrev")}
    dbp = {k: v.numpy() for k, v in pin.items()}
    Q = ctx.upload(q.lens, q.prof, q.mu, q.xyz, q.selfrev)        # in-memory side (DSSAligner target slot B)
    D = ctx.upload(dbp["lens"], dbp["prof"], dbp["mu"], dbp["xyz"], dbp["selfrev"])  # streamed side (slot A)
    cells_rank = float(np.sum(db.lens, dtype=np.float64)) * float(np.sum(q.lens, dtype=np.float64))
    pairs_rank = db.n * q.n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident inputs, device-timed ----
    for _ in range(args.warmup):
        ctx.search_cross_device(D, Q)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sw_ms = lddt_ms = 0.0
    launches = sw_launches = 0
    ev0.record()
    for _ in range(args.steps):
        ctx.search_cross_device(D, Q)
        st = ctx.stats()
        sw_ms += st["sw_kernel_ms"]
        lddt_ms += st["lddt_kernel_ms"]
        launches += st["kernel_launches"]
        sw_launches += st["sw_kernel_launches"]
    ev1.record()
    barrier()
    clocks = sampler.stop()
    dev_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([dev_ms, sw_ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([cells_rank, float(pairs_rank), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms_max, sw_ms_max = t.tolist()
    cells_all, pairs_all, launches_all = tot.tolist()
    value = cells_all * args.steps / (dev_ms_max * 1e-3)
    pairs_per_s = pairs_all * args.steps / (dev_ms_max * 1e-3)

    # ---- second leg of the metric: chain pairs/s where the Mu filter decides (-sensitive: `-search Q -db DB -sensitive`) ----
    ctx.set_params(rb.params_preset(rb.MODE_SENSITIVE))
    ctx.search_cross_device(D, Q)
    barrier()
    evs0, evs1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mu_ms = sens_sw_pairs = 0.0
    evs0.record()
    for _ in range(args.steps):
        ctx.search_cross_device(D, Q)
        st = ctx.stats()
        mu_ms += st["mu_kernel_ms"]
        sens_sw_pairs += st["sw_pairs"]
    evs1.record()
    barrier()
    ts = torch.tensor([evs0.elapsed_time(evs1), mu_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
    sens_ms, mu_ms_max = ts.tolist()
    sens_pairs_per_s = pairs_all * args.steps / (sens_ms * 1e-3)
    ctx.set_params(rb.params_preset(rb.MODE_VERYSENSITIVE))

    # ---- e2e: host buffers in, hits out, every step ----
    e2e_steps = max(1, args.e2e_steps)
    D.free()
    h2d = d2h = 0
    nhits = 0

    def e2e_step():
        Dk = ctx.upload(dbp["lens"], dbp["prof"], dbp["mu"], dbp["xyz"], dbp["selfrev"])
        res = ctx.search_cross(Dk, Q, keep=rb.KEEP_HITS, want_paths=True)
        st = ctx.stats()
        out = (Dk.h2d_bytes + st["h2d_bytes"], st["d2h_bytes"], len(res.hits))
        Dk.free()
        del res
        return out

    for _ in range(max(1, min(args.warmup, 2))):  # untimed: pinned staging buffers and host result blocks get allocated
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        a, b_, c = e2e_step()
        h2d += a
        d2h += b_
        nhits += c
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    io = torch.tensor([float(h2d), float(d2h), float(nhits)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(io, op=dist.ReduceOp.SUM)  # bytes moved and hits of the whole job
    h2d, d2h, nhits = io.tolist()
    e2e_value = cells_all / te.item()

    if rank == 0:
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        peak_src = "fallback (B200_PROFILING.md)"
        peak = 6650.0
        if pk.exists():
            peaks = json.loads(pk.read_text())
            peak = float(peaks.get("hbm_gbs", peak))
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        # one step = sw_launches_step launches of the SW kernel (one per batch of <= 2M pairs) that together cover the
        # rank's pairs once: per-launch figures are the step's divided by that count
        sw_launches_step = max(1, sw_launches // args.steps)
        alg_bytes = algorithmic_bytes(db.lens, q.lens) / sw_launches_step
        sw_ms_launch = sw_ms_max / args.steps / sw_launches_step
        achieved = alg_bytes / (sw_ms_launch * 1e-3) / 1e9
        smem_peak = 148 * 128 * clocks["sm_mhz"] * 1e6 / 1e9 if clocks.get("sm_mhz") else None  # GB/s: 128 B/clk/SM
        traffic = None
        tf = ROOT / "profiles" / "sw_kernel_traffic.json"
        if tf.exists():
            try:
                traffic = json.loads(tf.read_text()).get("dram_bytes_per_launch_at_bench_shape")
            except Exception:
                traffic = None
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "chain_pairs_per_s": pairs_per_s,
            "sensitive": {"chain_pairs_per_s": sens_pairs_per_s, "ms_per_step": sens_ms / args.steps,
                          "mu_filter_cells_per_s": 2.0 * cells_all * args.steps / (mu_ms_max * 1e-3) if mu_ms_max > 0 else None,
                          "mu_filter_share_of_step": mu_ms_max / sens_ms, "sw_pairs_share": sens_sw_pairs / (pairs_rank * args.steps),
                          "note": "same shard and queries under the -sensitive preset (Mu int8 SW filter fwd+rev, then float SW for the "
                                  "survivors), device-resident, CUDA events; filter cells counted as 2*LA*LB per pair"},
            "config": {"workload": f"c5 -verysensitive full SW+traceback+LDDT/E-value: Q={args.nq} queries x DB={args.ndb} chains per GPU, L={args.length}",
                       "mode": "verysensitive", "pairs_per_step": pairs_all, "cells_per_step": cells_all,
                       "db_chains_total": args.ndb * world, "sharding": "DB shard per rank, queries replicated, no data-path collective",
                       "l2": "inputs larger than L2 (DB shard %.0f MB + %.0f MB checkpoint scratch per step)" % (db.nbytes() / 1e6, 220.0),
                       "seed": SEED},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d / e2e_steps, "d2h_bytes_per_step": d2h / e2e_steps,
                    "ms_per_step": te.item() * 1e3, "hits_per_step": nhits / e2e_steps, "steps": e2e_steps,
                    "api": "rsk_chainset_upload + rsk_search_cross(keep=HITS, paths) from pinned host buffers"},
            "gpu_launches": int(launches_all),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "sw_affine_f32_tb_kernel", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms_per_launch": sw_ms_launch,
                         "launches_per_step": sw_launches_step,
                         "kernel_share_of_step": sw_ms_max / dev_ms_max,
                         "kernel_cells_per_s": cells_rank / (sw_ms_max / args.steps * 1e-3),
                         "binding_resource": {"name": "shared-memory bandwidth (score-table reads, 32 B/cell)", "unit": "GB/s",
                                              "achieved": 32.0 * cells_rank / (sw_ms_max / args.steps * 1e-3) / 1e9,
                                              "peak": smem_peak,
                                              "frac": (32.0 * cells_rank / (sw_ms_max / args.steps * 1e-3) / 1e9 / smem_peak) if smem_peak else None},
                         "note": "the DP is bound by shared-memory table reads, not HBM (profiles/r1_sw_kernel_v4_ncu.md); "
                                 "HBM fraction reported as SURVEY §8d requires"},
        }
        if not args.no_cpu_baseline:
            cps, pps, kind, cores, sample, ms = cpu_sample_run(q, db, steps=1, warmup=0)
            out["cpu_baseline"] = {"value": cps, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                                   "chain_pairs_per_s": pps}
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
