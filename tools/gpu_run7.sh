#!/bin/bash
# Round-2 GPU call 7: runtime variants of the W.H contractions (tile 104 / 112 / 128, preload, pairs) in one process; tests of the narrow tiles.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
S=gpurun_out/r2g_summary.txt
echo "== tests" > $S
timeout 900 python -m pytest tests/test_gpu_tma.py -q -x > gpurun_out/r2g_tests.log 2>&1
echo "tma rc=$?" >> $S; tail -3 gpurun_out/r2g_tests.log >> $S
echo "== variants" >> $S
timeout 900 python tools/tma_gemm_check.py variants > gpurun_out/r2g_variants.log 2>&1
echo "variants rc=$?" >> $S; cat gpurun_out/r2g_variants.log >> $S
WH_TILE=104 timeout 300 python tools/tma_gemm_check.py stamps > gpurun_out/r2g_stamps104.log 2>&1
grep -A4 "pdl=0" gpurun_out/r2g_stamps104.log | cut -c1-330 >> $S
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
python - <<'PY' >> $S 2>&1
import json
d=json.load(open('gpurun_out/r2g_bench.json'))
print('bench value', d['value'], 'e2e', d['e2e']['value'], d['stage_ms'], 'launches', d['gpu_launches'])
PY
cat $S
