#!/bin/bash
# Round-2 GPU call 13: wh_split2 with batched DSMEM loads -- parity, timing, stamps, bench (short)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
S=gpurun_out/r2n_summary.txt
echo "== tests with wh_split2" > $S
GCCNMF_WH_SPLIT2=1 timeout 200 python -m pytest tests/test_gpu_tma.py tests/test_gpu_parity.py -q -x -k "klnmf or debug_timing or full_size or pipeline" > gpurun_out/r2n_tests.log 2>&1
echo "tests rc=$?" >> $S; tail -3 gpurun_out/r2n_tests.log >> $S
echo "== variants" >> $S
timeout 200 python tools/tma_gemm_check.py variants > gpurun_out/r2n_variants.log 2>&1
echo "variants rc=$?" >> $S; tail -4 gpurun_out/r2n_variants.log >> $S
WH_SPLIT2=1 timeout 100 python tools/tma_gemm_check.py stamps > gpurun_out/r2n_stamps.log 2>&1
grep -A5 "pdl=0" gpurun_out/r2n_stamps.log | cut -c1-330 >> $S
GCCNMF_WH_SPLIT2=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench_split2.json 2> gpurun_out/r2n_bench_split2.err
python - <<'PY' >> $S 2>&1
import json
d=json.load(open('gpurun_out/r2n_bench_split2.json'))
print('bench split2 value', round(d['value']), 'e2e', round(d['e2e']['value']), d['stage_ms'])
PY
cat $S
