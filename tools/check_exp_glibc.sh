#!/bin/sh
# builds and runs the bit-for-bit check of exp_glibc.cuh against this host's libm (no contraction by the compiler: every fused
# operation is an explicit fma() call)
set -e
cd "$(dirname "$0")"
gcc -O2 -ffp-contract=off -mfma -o /tmp/check_exp_glibc check_exp_glibc.c -lm
/tmp/check_exp_glibc "$@"
