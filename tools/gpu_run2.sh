#!/bin/bash
# Round-2 GPU call 2: pair kernel along x, the fused real-time path (tests, sanitizer, latency), config-1 parity thresholds.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
S=gpurun_out/r2b_summary.txt
echo "== pair tests" > $S
GCCNMF_TEST_EXPERIMENTAL=1 timeout 420 python -m pytest tests/test_gpu_tma.py -k pairs -x -q > gpurun_out/r2b_pairs.log 2>&1
echo "pairs rc=$?" >> $S; tail -3 gpurun_out/r2b_pairs.log >> $S
if grep -q "passed" gpurun_out/r2b_pairs.log && ! grep -q "failed" gpurun_out/r2b_pairs.log; then
  GEMM_PAIR=1 timeout 300 python tools/tma_gemm_check.py stamps > gpurun_out/r2b_stamps_pair.log 2>&1
  echo "stamps pair rc=$?" >> $S
  GEMM_PAIR=1 TIME_VARIANTS=short timeout 300 python tools/tma_gemm_check.py time > gpurun_out/r2b_time_pair.log 2>&1
  GEMM_PAIR=1 timeout 400 python tools/tma_gemm_check.py nmf > gpurun_out/r2b_nmf_pair.log 2>&1
  echo "nmf pair rc=$?" >> $S
fi
echo "== online / realtime tests" >> $S
timeout 900 python -m pytest tests/test_gpu_online.py -q -s > gpurun_out/r2b_online.log 2>&1
echo "online rc=$?" >> $S; tail -8 gpurun_out/r2b_online.log >> $S
echo "== sanitizer on the realtime tests" >> $S
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_online.py -q -k "realtime_processor_against or overlap_add_ring or coefficient_inference" > gpurun_out/r2b_sanitizer.log 2>&1
echo "sanitizer rc=$?" >> $S; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2b_sanitizer.log | tail -3 >> $S
echo "== rt latency" >> $S
timeout 900 python tools/rt_latency.py --chunks 2000 --json gpurun_out/r2b_rt_latency.json > gpurun_out/r2b_rt_latency.log 2>&1
echo "latency rc=$?" >> $S; cat gpurun_out/r2b_rt_latency.log >> $S
echo "== full-size parity" >> $S
timeout 1500 python -m pytest tests/test_gpu_parity_full.py -q -s > gpurun_out/r2b_parity_full.log 2>&1
echo "parity_full rc=$?" >> $S; tail -4 gpurun_out/r2b_parity_full.log >> $S
echo "== whole gpu suite" >> $S
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity_full.py --deselect tests/test_gpu_online.py > gpurun_out/r2b_pytest.log 2>&1
echo "suite rc=$?" >> $S; tail -4 gpurun_out/r2b_pytest.log >> $S
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
echo "bench rc=$?" >> $S
python - <<'PY' >> $S 2>&1
import json
d=json.load(open('gpurun_out/r2b_bench.json'))
print('bench value', d['value'], 'e2e', d['e2e']['value'], d['stage_ms'])
PY
cat $S
