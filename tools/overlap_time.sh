cd build/data
for e in 0 1; do
  if [ $e = 1 ]; then export RSK_NO_OVERLAP=1; else unset RSK_NO_OVERLAP; fi
  for rep in 1 2; do
    s=$(date +%s%N)
    timeout 150 ../../reseek_b200/rsk_host_demo search sensitive scop40.bca scop40.bca /tmp/o$e.tsv > /tmp/log$e.txt 2>&1
    t=$(date +%s%N)
    echo "no_overlap=$e rep $rep wall_ms $(( (t - s) / 1000000 )) lines $(wc -l < /tmp/o$e.tsv)"
  done
done
