#!/bin/bash
# last short call: wh_split2 with the tail columns shared by the two splits -- parity subset + timing
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
S=gpurun_out/r2u_summary.txt
GCCNMF_WH_SPLIT2=1 timeout 100 python -m pytest tests/test_gpu_tma.py -q -x -k "klnmf" > gpurun_out/r2u_tests.log 2>&1
echo "tests rc=$?" > $S; tail -2 gpurun_out/r2u_tests.log >> $S
timeout 120 python tools/tma_gemm_check.py variants > gpurun_out/r2u_variants.log 2>&1
echo "variants rc=$?" >> $S; tail -4 gpurun_out/r2u_variants.log >> $S
cat $S
