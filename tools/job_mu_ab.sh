#!/bin/bash
# A/B of the Mu filter (K3) between two builds: the tree's library and build/libreseek_b200_head.so, which has to be built
# first from the commit to compare against (git archive <commit> reseek_b200/csrc include | tar -x -C /tmp/x; make OBJDIR=... OUT=...).
for L in 150 300; do
  for v in head new head new; do
    lib=$PWD/reseek_b200/libreseek_b200.so; [ $v = head ] && lib=$PWD/build/libreseek_b200_head.so
    RSK_LIB=$lib python tools/quick_perf.py 100 $(( 12000000 / L )) $L 3 2 2>&1 | grep "^rep 2" | sed "s/^.*mu_ms/L=$L $v mu_ms/; s/sw_ms.*mu cells/ mu cells/"
  done
done
