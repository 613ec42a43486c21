#!/bin/bash
# Round-2 GPU call 5: CTA pairs + dual-N for the W.H contractions, candidate refinement of the argmax: correctness, stamps, time, bench.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
S=gpurun_out/r2e_summary.txt
echo "== tests" > $S
timeout 900 python -m pytest tests/test_gpu_tma.py tests/test_gpu_parity.py -q -x > gpurun_out/r2e_tests.log 2>&1
echo "tma+parity rc=$?" >> $S; tail -4 gpurun_out/r2e_tests.log >> $S
timeout 1500 python -m pytest tests/test_gpu_parity_full.py -q -s > gpurun_out/r2e_parity_full.log 2>&1
echo "parity_full rc=$?" >> $S; tail -3 gpurun_out/r2e_parity_full.log >> $S
echo "== stamps / time" >> $S
timeout 300 python tools/tma_gemm_check.py stamps > gpurun_out/r2e_stamps.log 2>&1
grep -A4 "pdl=0" gpurun_out/r2e_stamps.log >> $S
TIME_VARIANTS=short timeout 300 python tools/tma_gemm_check.py time > gpurun_out/r2e_time.log 2>&1
head -4 gpurun_out/r2e_time.log >> $S
echo "== bench" >> $S
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
echo "bench rc=$?" >> $S
python - <<'PY' >> $S 2>&1
import json
d=json.load(open('gpurun_out/r2e_bench.json'))
print('bench value', d['value'], 'e2e', d['e2e']['value'], d['stage_ms'], 'launches', d['gpu_launches'])
PY
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity_full.py --deselect tests/test_gpu_parity.py --deselect tests/test_gpu_tma.py > gpurun_out/r2e_pytest.log 2>&1
echo "rest of suite rc=$?" >> $S; tail -3 gpurun_out/r2e_pytest.log >> $S
cat $S
