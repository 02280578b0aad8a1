#!/bin/bash
# The reference's SCOP40 regression commands (test_scripts/scop40.bash) through rsk_host_demo on the GPU box; outputs gzipped
# under gpurun_out/scop40_regression/ for tools/check_scop40_sepq.py.  Needs build/data/scop40.bca (+ dom_scopid.tsv for the
# SCOP40Bench subclass run).
set -u
OUT=$PWD/gpurun_out/scop40_regression
mkdir -p $OUT
cd build/data
run() {  # name, extra options...
	name=$1; shift
	s=$(date +%s%N)
	../../reseek_b200/rsk_host_demo -search scop40.bca -db scop40.bca -output /tmp/scop40-$name.tsv -columns query+target+evalue "$@" > /tmp/scop40-$name.log 2>&1
	t=$(date +%s%N)
	echo "scop40-$name: wall_ms $(( (t - s) / 1000000 )) lines $(wc -l < /tmp/scop40-$name.tsv) rc $?"
	gzip -9 -c /tmp/scop40-$name.tsv > $OUT/scop40-$name.tsv.gz
}
run fast -fast
run sensitive -sensitive
run evalue1 -fast -evalue 1
for m in fast sensitive; do
	../../reseek_b200/rsk_host_demo -scop40bench scop40.bca -lookup dom_scopid.tsv -$m 2>/dev/null | tail -1 | sed "s/^/SCOP40Bench -$m: /" | tee -a $OUT/scop40bench.txt
done
ls -la $OUT
