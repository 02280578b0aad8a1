#!/bin/bash
# A/B of the TMA checkpoint-store experiment (RSK_SW_TMA_CKPT): the same timing run with the default library and the variant.
# Run on the GPU box from the repo root; log in gpurun_out/tma_ab.log.
{
echo "== parity of the variant (SW alignment tests) =="
RSK_LIB=$PWD/build/libreseek_b200_tma.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -x -q 2>&1 | tail -3
for L in 100 120 250; do
  for v in default tma default tma; do
    lib=$PWD/reseek_b200/libreseek_b200.so; [ $v = tma ] && lib=$PWD/build/libreseek_b200_tma.so
    echo "== L=$L $v =="
    RSK_LIB=$lib python tools/quick_perf.py 100 $(( 4000000 / L )) $L 4 2>&1 | grep "^rep [123]" | sed 's/mu_ms.*sw cells/sw cells/; s/e2e-dev.*//'
  done
done
} > gpurun_out/tma_ab.log 2>&1
