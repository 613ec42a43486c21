#!/bin/bash
# Round-2 final single-GPU call: whole GPU suite, smoke, stamps at the planned tile width, bench (default arguments), reference arm (short
# budget), ncu launch list, ncu --set full of one KL-NMF iteration and of the argmax stage.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
S=gpurun_out/r2k_summary.txt
echo "== gpu suite" > $S
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2k_pytest.log 2>&1
echo "pytest -m gpu rc=$?" >> $S; tail -4 gpurun_out/r2k_pytest.log >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2k_smoke.log 2>&1
echo "smoke rc=$?" >> $S; tail -2 gpurun_out/r2k_smoke.log >> $S
echo "== stamps (104-column W.H tiles)" >> $S
WH_TILE=104 timeout 200 python tools/tma_gemm_check.py stamps > gpurun_out/r2k_stamps.log 2>&1
grep -A5 "pdl=0" gpurun_out/r2k_stamps.log | cut -c1-330 >> $S
echo "== bench" >> $S
timeout 900 python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
echo "bench rc=$?" >> $S
python - <<'PY' >> $S 2>&1
import json
d=json.load(open('gpurun_out/r2k_bench.json'))
print('bench value', d['value'], 'e2e', d['e2e']['value'], d['stage_ms'], 'launches', d['gpu_launches'], 'first call ms', d.get('first_call_ms'))
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], 'cpu', d.get('cpu_baseline', {}).get('value'), 'clocks', d['clocks'])
PY
GCCNMF_REFERENCE_BUDGET_S=45 timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2k_bench_reference.json 2> gpurun_out/r2k_bench_reference.err
echo "reference arm rc=$?" >> $S; cut -c1-400 gpurun_out/r2k_bench_reference.json >> $S
echo "== ncu" >> $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 1400 -c 900 --csv --log-file gpurun_out/r2k_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_ncu_bench.log 2>&1
echo "ncu launches rc=$?" >> $S
PROF_ITERS=5 PROF_PDL=1 timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"plane_gemm|tma_apply" -s 18 -c 5 -o gpurun_out/r2k_nmf python tools/tma_gemm_check.py prof > gpurun_out/r2k_ncu_nmf.log 2>&1
echo "ncu full nmf rc=$?" >> $S
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"argmax_gemm_persistent|build_gcc_planes|phat_angspec|refine_candidates" -c 4 -o gpurun_out/r2k_argmax python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_ncu_argmax.log 2>&1
echo "ncu full argmax rc=$?" >> $S
ls -la gpurun_out/*.ncu-rep >> $S 2>&1
cat $S
