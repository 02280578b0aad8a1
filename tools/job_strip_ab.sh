#!/bin/bash
# A/B of the checkpoint strip length of the SW kernel (RSK_SW_STRIP = 8 / 16 / 32); log in gpurun_out/strip_ab.log
{
echo "== parity of the strip-8 variant =="
RSK_LIB=$PWD/build/libreseek_b200_s8.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -x -q 2>&1 | tail -2
for L in 100 150 300; do
  for v in s16 s8 s32 s16 s8 s32; do
    lib=$PWD/reseek_b200/libreseek_b200.so; [ $v = s8 ] && lib=$PWD/build/libreseek_b200_s8.so; [ $v = s32 ] && lib=$PWD/build/libreseek_b200_s32.so
    echo "== L=$L $v =="
    RSK_LIB=$lib python tools/quick_perf.py 100 $(( 4000000 / L )) $L 3 2>&1 | grep "^rep [12]" | sed 's/mu_ms.*sw cells/sw cells/; s/e2e-dev.*//'
  done
done
} > gpurun_out/strip_ab.log 2>&1
