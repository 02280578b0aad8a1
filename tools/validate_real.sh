# Real-data runs of the reference's own command lines through rsk_host_demo (data under build/data, not committed).
cd build/data
COLS=query+target+dpscore+newts+lddt+evalue+pvalue+qlo+qhi+ql+tlo+thi+tl+ids+gaps+pctid+cigar
for rep in 1 2; do
  s=$(date +%s%N)
  ../../reseek_b200/rsk_host_demo -search q100.bca -db scop40.bca -sensitive -columns $COLS -output /tmp/q.tsv > /tmp/q.log 2>&1
  t=$(date +%s%N)
  echo "q100 x scop40 -sensitive rep $rep: wall_ms $(( (t - s) / 1000000 )) lines $(wc -l < /tmp/q.tsv) md5 $(sort /tmp/q.tsv | md5sum | cut -c1-12)"
done
for rep in 1 2; do
  s=$(date +%s%N)
  ../../reseek_b200/rsk_host_demo -search scop40.bca -fast -output /tmp/self.tsv > /tmp/self.log 2>&1
  t=$(date +%s%N)
  echo "scop40 all-vs-all -fast rep $rep: wall_ms $(( (t - s) / 1000000 )) lines $(wc -l < /tmp/self.tsv) md5 $(sort /tmp/self.tsv | md5sum | cut -c1-12)"
done
