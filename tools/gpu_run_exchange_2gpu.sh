#!/bin/bash
# 2-GPU call: exchange inside the W update (form 2) -- parity, timeline, bench
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
S=gpurun_out/r2t_summary.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== 2 GPUs: parity" > $S
timeout 900 python -m pytest tests/test_gpu_multi.py -q -s > gpurun_out/r2t_pytest.log 2>&1
echo "multi pytest rc=$?" >> $S; grep -E "collective:|W identical|rel W|MULTI_GPU_CHECK|passed|failed|Error" gpurun_out/r2t_pytest.log | cut -c1-200 >> $S
echo "== timeline" >> $S
GCCNMF_DEBUG_EXCHANGE=1 STAMP_MODES="fused,pull-in-w-update,pull-one-shot" timeout 400 $TR --master-port 29571 tools/multi_stamps.py > gpurun_out/r2t_stamps.log 2>&1
echo "multi_stamps rc=$?" >> $S; grep "^rank 0\|not available\|failed\|Error" gpurun_out/r2t_stamps.log | sed 's/| spans.*| after G4 end:/| after G4 end:/' >> $S
for mode in auto pull1; do
  GCCNMF_COLLECTIVE=$mode timeout 600 $TR --master-port $((29600 + RANDOM % 300)) bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2t_bench_$mode.json 2> gpurun_out/r2t_bench_$mode.err
  echo "bench $mode rc=$?" >> $S
  python - $mode <<'PY' >> $S 2>&1
import json,sys
m=sys.argv[1]
d=json.loads(open('gpurun_out/r2t_bench_%s.json'%m).read().strip().splitlines()[-1])
print(m, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['stage_ms'], d['collective'][:70])
PY
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_bench_1gpu.json 2> gpurun_out/r2t_bench_1gpu.err
python - <<'PY' >> $S 2>&1
import json
d=json.load(open('gpurun_out/r2t_bench_1gpu.json'))
print('1 GPU on this box: value', round(d['value']), 'e2e', round(d['e2e']['value']), d['stage_ms'])
PY
cat $S
