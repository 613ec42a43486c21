#!/bin/bash
# Round-2 GPU call 3: where the cta_group::2 pair kernel spends its time (leader / follower stamps), ncu launch list of one bench step,
# ncu --set full of the pair kernel and the argmax GEMM, new tests (masked reconstruction on tensor cores, fused call, peak picking), bench.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
S=gpurun_out/r2c_summary.txt
echo "== pair stamps" > $S
GEMM_PAIR=1 timeout 300 python tools/tma_gemm_check.py stamps > gpurun_out/r2c_stamps_pair.log 2>&1
echo "stamps pair rc=$?" >> $S
timeout 300 python tools/tma_gemm_check.py stamps > gpurun_out/r2c_stamps_base.log 2>&1
GEMM_PAIR=1 TIME_VARIANTS=short timeout 300 python tools/tma_gemm_check.py time > gpurun_out/r2c_time_pair.log 2>&1
TIME_VARIANTS=short timeout 300 python tools/tma_gemm_check.py time > gpurun_out/r2c_time_base.log 2>&1
head -1 gpurun_out/r2c_time_pair.log >> $S; head -1 gpurun_out/r2c_time_base.log >> $S
echo "== tests" >> $S
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tma.py -q -s > gpurun_out/r2c_parity.log 2>&1
echo "parity+tma rc=$?" >> $S; tail -4 gpurun_out/r2c_parity.log >> $S
timeout 1500 python -m pytest tests/test_gpu_parity_full.py -q -s > gpurun_out/r2c_parity_full.log 2>&1
echo "parity_full rc=$?" >> $S; tail -4 gpurun_out/r2c_parity_full.log >> $S
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity_full.py --deselect tests/test_gpu_parity.py --deselect tests/test_gpu_tma.py > gpurun_out/r2c_pytest.log 2>&1
echo "rest of suite rc=$?" >> $S; tail -4 gpurun_out/r2c_pytest.log >> $S
echo "== bench" >> $S
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
echo "bench rc=$?" >> $S
python - <<'PY' >> $S 2>&1
import json
d=json.load(open('gpurun_out/r2c_bench.json'))
print('bench value', d['value'], 'e2e', d['e2e']['value'], d['stage_ms'], 'launches', d['gpu_launches'])
PY
echo "== ncu launch list of one step" >> $S
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 1400 -c 900 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_ncu_bench.log 2>&1
echo "ncu launches rc=$?" >> $S
echo "== ncu full: pair G1 + argmax gemm" >> $S
GEMM_PAIR=1 PROF_ITERS=4 PROF_PDL=1 timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"plane_gemm" -s 12 -c 4 -o gpurun_out/r2c_prof_pair python tools/tma_gemm_check.py prof > gpurun_out/r2c_ncu_pair.log 2>&1
echo "ncu pair rc=$?" >> $S
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"EpiArgmaxTile|build_gcc_planes|refine_argmax" -s 6 -c 3 -o gpurun_out/r2c_prof_argmax python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_ncu_argmax.log 2>&1
echo "ncu argmax rc=$?" >> $S
cat $S
