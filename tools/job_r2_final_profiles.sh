#!/bin/bash
# Round-2 closing captures on one B200 (outputs under gpurun_out/ncu_r2b/ and gpurun_out/r2_final.log)
set -u
OUT=gpurun_out/ncu_r2b
mkdir -p $OUT
NCU="ncu --clock-control none --csv --page raw"
{
echo "== validate_real (whole-process wall clock, md5 of sorted output) =="
bash tools/validate_real.sh
echo "== scop40 all-vs-all with RSK_TIMING =="
( cd build/data && RSK_TIMING=1 ../../reseek_b200/rsk_host_demo -search scop40.bca -fast -output /tmp/self.tsv 2>&1 | grep -v "^$" | tail -40 )
echo "== fastdb timing =="
RSK_TIMING=1 python tools/fastdb_perf.py 100 20000 2>&1 | tail -12
} > gpurun_out/r2_final.log 2>&1
# K4 on the 165 long chains, current grid
( cd build/data && $NCU --set full -k regex:mkf_xdrop_warp -c 1 --log-file ../../$OUT/k4_xdrop_warp_full.csv ../../reseek_b200/rsk_host_demo -search long165.bca -fast -gpus 1 -output /tmp/l.tsv > /dev/null 2>&1 )
# prefilter kernels, final forms
$NCU --set full -k regex:pf_ -c 12 --log-file $OUT/prefilter_full.csv python tools/fastdb_perf.py 100 20000 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_fastdb.csv python tools/fastdb_perf.py 100 20000 > /dev/null 2>&1
# launch list of the all-vs-all
( cd build/data && ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file ../../$OUT/launches_scop40_self.csv ../../reseek_b200/rsk_host_demo -search scop40.bca -fast -gpus 1 -output /tmp/self2.tsv > /dev/null 2>&1 )
ls -la $OUT
