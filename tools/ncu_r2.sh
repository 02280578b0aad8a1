#!/bin/bash
# Round-2 ncu captures (run on the GPU box from the repo root; outputs under gpurun_out/ncu_r2/).
# One kernel per capture, --set full for the unit tables, plus a dram-bytes-only pass of K1 at the bench shape.
set -u
OUT=gpurun_out/ncu_r2
mkdir -p $OUT
NCU="ncu --clock-control none --csv --page raw"
METRICS_SHORT=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed.sum

# 1. K4 warp-cooperative x-drop on the 165 real SCOP40 chains >= 500 residues (13 695 long-chain pairs)
( cd build/data && $NCU --set full -k regex:mkf_xdrop_warp -c 1 --log-file ../../$OUT/k4_xdrop_warp_full.csv ../../reseek_b200/rsk_host_demo -search long165.bca -fast -gpus 1 -output /tmp/l.tsv > /dev/null 2>&1 )
# 2. K10 DSS on all 11 211 SCOP40 chains
$NCU --set full -k regex:dss_kernel -c 1 --log-file $OUT/k10_dss_full.csv python -c "
import numpy as np, reseek_b200 as rb
from reseek_b200 import chainio
l,s,x=chainio.read_bca('build/data/scop40.bca')
ctx=rb.Context(0, rb.MODE_SENSITIVE)
S=rb.ChainSet.from_coords(ctx, np.array([len(a) for a in s],np.uint32), np.frombuffer(b''.join(s),np.uint8), np.concatenate(x,axis=1))
" > /dev/null 2>&1
# 3. K1: full set on the first launch of the bench shape, then the dram bytes of the same launch at clock-control none
$NCU --set full -k regex:sw_affine -c 1 --log-file $OUT/k1_sw_full.csv python bench.py --steps 1 --warmup 0 --legs '' --no-cpu-baseline > /dev/null 2>&1
# 4. short chains (half-warp classes): L = 100
$NCU --set full -k regex:sw_affine -c 1 --log-file $OUT/k1_sw_L100_full.csv python bench.py --steps 1 --warmup 0 --legs '' --no-cpu-baseline --len 100 > /dev/null 2>&1
# 5. prefilter kernels (K6 neighbourhood, K7 probe, K8 extend, bag)
$NCU --set full -k regex:pf_ -c 14 --log-file $OUT/k6_k8_prefilter_full.csv python tools/fastdb_perf.py 100 20000 > /dev/null 2>&1
# 6. launch list of one bench step (serialised, cold caches: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_bench.csv python bench.py --steps 1 --warmup 1 --legs '' --no-cpu-baseline > /dev/null 2>&1
ls -la $OUT
