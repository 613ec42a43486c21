#!/bin/bash
# Round-2 GPU call 9: persistent argmax GEMM validated first under a short timeout (falls back to the per-tile kernel for the rest of
# the call if it fails), then tests, NMF variants, bench, launch list.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
S=gpurun_out/r2i_summary.txt
echo "== persistent argmax" > $S
timeout 240 python -m pytest tests/test_gpu_parity.py -q -x -k "argmax" > gpurun_out/r2i_argmax.log 2>&1
rc=$?; echo "argmax tests rc=$rc" >> $S; tail -3 gpurun_out/r2i_argmax.log >> $S
if [ $rc -ne 0 ]; then
  export GCCNMF_ARGMAX_PERSISTENT=0
  echo "persistent argmax kernel DISABLED for the rest of this call" >> $S
  nvidia-smi > gpurun_out/r2i_smi_after_fail.txt 2>&1
  timeout 240 python -m pytest tests/test_gpu_parity.py -q -x -k "argmax" > gpurun_out/r2i_argmax_fallback.log 2>&1
  echo "argmax tests (per-tile kernel) rc=$?" >> $S; tail -3 gpurun_out/r2i_argmax_fallback.log >> $S
fi
echo "== tests" >> $S
timeout 600 python -m pytest tests/test_gpu_tma.py tests/test_gpu_parity.py -q -x > gpurun_out/r2i_tests.log 2>&1
echo "tma+parity rc=$?" >> $S; tail -3 gpurun_out/r2i_tests.log >> $S
echo "== variants" >> $S
timeout 600 python tools/tma_gemm_check.py variants > gpurun_out/r2i_variants.log 2>&1
echo "variants rc=$?" >> $S; cat gpurun_out/r2i_variants.log >> $S
timeout 200 python tools/tma_gemm_check.py stamps > gpurun_out/r2i_stamps.log 2>&1
grep -A4 "pdl=0" gpurun_out/r2i_stamps.log | cut -c1-330 >> $S
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
python - <<'PY' >> $S 2>&1
import json
d=json.load(open('gpurun_out/r2i_bench.json'))
print('bench value', d['value'], 'e2e', d['e2e']['value'], d['stage_ms'], 'launches', d['gpu_launches'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 1400 -c 900 --csv --log-file gpurun_out/r2i_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_ncu_bench.log 2>&1
echo "ncu launches rc=$?" >> $S
timeout 900 python -m pytest tests/test_gpu_parity_full.py -q -s -k "argmax or config1_pipeline or config2" > gpurun_out/r2i_parity_full.log 2>&1
echo "parity_full rc=$?" >> $S; tail -3 gpurun_out/r2i_parity_full.log >> $S
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity_full.py --deselect tests/test_gpu_parity.py --deselect tests/test_gpu_tma.py > gpurun_out/r2i_pytest.log 2>&1
echo "rest of suite rc=$?" >> $S; tail -3 gpurun_out/r2i_pytest.log >> $S
cat $S
