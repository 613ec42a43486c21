#!/bin/bash
# Round-2 GPU call 10: cluster-reduced k-splits of the W-update numerator (DSMEM), register-tiled phat kernel, L2 window; A/B, bench.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
S=gpurun_out/r2j_summary.txt
echo "== tests" > $S
timeout 300 python -m pytest tests/test_gpu_tma.py -q -x > gpurun_out/r2j_tma.log 2>&1
rc=$?; echo "tma rc=$rc" >> $S; tail -3 gpurun_out/r2j_tma.log >> $S
if [ $rc -ne 0 ]; then
  echo "retrying with w_cluster_reduce off is not automatic: see r2j_tma.log" >> $S
fi
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x > gpurun_out/r2j_parity.log 2>&1
echo "parity rc=$?" >> $S; tail -3 gpurun_out/r2j_parity.log >> $S
echo "== variants" >> $S
timeout 600 python tools/tma_gemm_check.py variants > gpurun_out/r2j_variants.log 2>&1
echo "variants rc=$?" >> $S; cat gpurun_out/r2j_variants.log >> $S
echo "== iteration timeline, 1 GPU" >> $S
timeout 300 python tools/multi_stamps.py > gpurun_out/r2j_stamps1.log 2>&1
echo "multi_stamps rc=$?" >> $S; grep "^rank" gpurun_out/r2j_stamps1.log >> $S
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
python - <<'PY' >> $S 2>&1
import json
d=json.load(open('gpurun_out/r2j_bench.json'))
print('bench value', d['value'], 'e2e', d['e2e']['value'], d['stage_ms'], 'launches', d['gpu_launches'], 'first call ms', d.get('first_call_ms'))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 1400 -c 900 --csv --log-file gpurun_out/r2j_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_ncu_bench.log 2>&1
echo "ncu launches rc=$?" >> $S
timeout 900 python -m pytest tests/test_gpu_parity_full.py -q -s -k "config1_pipeline or config2" > gpurun_out/r2j_parity_full.log 2>&1
echo "parity_full rc=$?" >> $S; tail -3 gpurun_out/r2j_parity_full.log >> $S
cat $S
