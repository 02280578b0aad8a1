#!/bin/bash
# A/B of the number of column chains a warp sweeps as one wavefront (RSK_SW_CHAIN = 4 / 6 / 8); log in gpurun_out/chain_ab.log
{
echo "== parity of the chain-8 variant =="
RSK_LIB=$PWD/build/libreseek_b200_c8.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -x -q 2>&1 | tail -2
for L in 100 300 800; do
  for v in c4 c6 c8 c4 c6 c8; do
    lib=$PWD/reseek_b200/libreseek_b200.so; [ $v = c6 ] && lib=$PWD/build/libreseek_b200_c6.so; [ $v = c8 ] && lib=$PWD/build/libreseek_b200_c8.so
    echo "== L=$L $v =="
    RSK_LIB=$lib python tools/quick_perf.py 100 $(( 6000000 / L )) $L 3 2>&1 | grep "^rep [12]" | sed 's/mu_ms.*sw cells/sw cells/; s/e2e-dev.*//'
  done
done
} > gpurun_out/chain_ab.log 2>&1
