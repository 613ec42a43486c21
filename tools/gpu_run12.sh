#!/bin/bash
# Round-2 GPU call 12: W.H contractions as k-split pairs of plain 208-column tiles (wh_split2) -- parity, timing, stamps, bench
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
S=gpurun_out/r2l_summary.txt
echo "== tests with wh_split2" > $S
GCCNMF_WH_SPLIT2=1 timeout 240 python -m pytest tests/test_gpu_tma.py -q -x -k "klnmf or debug_timing" > gpurun_out/r2l_tma.log 2>&1
echo "tma rc=$?" >> $S; tail -3 gpurun_out/r2l_tma.log >> $S
GCCNMF_WH_SPLIT2=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "full_size or klnmf or pipeline" > gpurun_out/r2l_parity.log 2>&1
echo "parity rc=$?" >> $S; tail -3 gpurun_out/r2l_parity.log >> $S
echo "== variants" >> $S
timeout 300 python tools/tma_gemm_check.py variants > gpurun_out/r2l_variants.log 2>&1
echo "variants rc=$?" >> $S; cat gpurun_out/r2l_variants.log | tail -6 >> $S
WH_SPLIT2=1 timeout 120 python tools/tma_gemm_check.py stamps > gpurun_out/r2l_stamps.log 2>&1
grep -A5 "pdl=0" gpurun_out/r2l_stamps.log | cut -c1-330 >> $S
GCCNMF_WH_SPLIT2=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench_split2.json 2> gpurun_out/r2l_bench_split2.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench_dual.json 2> gpurun_out/r2l_bench_dual.err
python - <<'PY' >> $S 2>&1
import json
for n in ('split2', 'dual'):
    d=json.load(open('gpurun_out/r2l_bench_%s.json' % n))
    print('bench', n, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['stage_ms'])
PY
GCCNMF_WH_SPLIT2=1 timeout 400 python -m pytest tests/test_gpu_parity_full.py -q -s -k "config2 or config1_pipeline" > gpurun_out/r2l_parity_full.log 2>&1
echo "parity_full (split2) rc=$?" >> $S; tail -4 gpurun_out/r2l_parity_full.log | cut -c1-300 >> $S
cat $S
