#!/bin/bash
# Round-2 GPU call 1: validate the cta_group::2 pair kernel, stamp it inside the live loop, new full-size parity tests, whole suite, bench.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
echo "== pair tests" > gpurun_out/r2a_summary.txt
GCCNMF_TEST_EXPERIMENTAL=1 timeout 420 python -m pytest tests/test_gpu_tma.py -k pairs -x -q > gpurun_out/r2a_pairs.log 2>&1
echo "pairs rc=$?" >> gpurun_out/r2a_summary.txt
tail -3 gpurun_out/r2a_pairs.log >> gpurun_out/r2a_summary.txt
if grep -q "passed" gpurun_out/r2a_pairs.log && ! grep -q "failed" gpurun_out/r2a_pairs.log; then
  GEMM_PAIR=1 timeout 300 python tools/tma_gemm_check.py stamps > gpurun_out/r2a_stamps_pair.log 2>&1
  echo "stamps pair rc=$?" >> gpurun_out/r2a_summary.txt
  GEMM_PAIR=1 TIME_VARIANTS=short timeout 300 python tools/tma_gemm_check.py time > gpurun_out/r2a_time_pair.log 2>&1
  GEMM_PAIR=1 timeout 400 python tools/tma_gemm_check.py nmf > gpurun_out/r2a_nmf_pair.log 2>&1
  echo "nmf pair rc=$?" >> gpurun_out/r2a_summary.txt
fi
timeout 300 python tools/tma_gemm_check.py stamps > gpurun_out/r2a_stamps_base.log 2>&1
TIME_VARIANTS=short timeout 300 python tools/tma_gemm_check.py time > gpurun_out/r2a_time_base.log 2>&1
echo "== full-size parity" >> gpurun_out/r2a_summary.txt
timeout 1200 python -m pytest tests/test_gpu_parity_full.py -q -s > gpurun_out/r2a_parity_full.log 2>&1
echo "parity_full rc=$?" >> gpurun_out/r2a_summary.txt
tail -5 gpurun_out/r2a_parity_full.log >> gpurun_out/r2a_summary.txt
echo "== whole gpu suite" >> gpurun_out/r2a_summary.txt
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity_full.py > gpurun_out/r2a_pytest.log 2>&1
echo "suite rc=$?" >> gpurun_out/r2a_summary.txt
tail -5 gpurun_out/r2a_pytest.log >> gpurun_out/r2a_summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?" >> gpurun_out/r2a_summary.txt
cat gpurun_out/r2a_summary.txt
