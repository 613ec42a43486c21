#!/bin/bash
# 2-GPU call: iteration timeline with exchange-kernel stamps, release.sys vs light arrival signal; bench of the best forms
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
S=gpurun_out/r2p_summary.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== 2 GPUs: timeline (release.sys arrival signal)" > $S
timeout 400 $TR --master-port 29571 tools/multi_stamps.py > gpurun_out/r2p_stamps.log 2>&1
echo "multi_stamps rc=$?" >> $S; grep "^rank 0" gpurun_out/r2p_stamps.log >> $S
echo "== timeline (light arrival signal of the pack)" >> $S
GCCNMF_MC_LIGHT_SIGNAL=1 timeout 400 $TR --master-port 29572 tools/multi_stamps.py > gpurun_out/r2p_stamps_light.log 2>&1
echo "multi_stamps light rc=$?" >> $S; grep "^rank 0" gpurun_out/r2p_stamps_light.log >> $S
GCCNMF_MC_LIGHT_SIGNAL=1 timeout 600 python -m pytest tests/test_gpu_multi.py -q -s -k "multimem" > gpurun_out/r2p_pytest.log 2>&1
echo "multi pytest (light) rc=$?" >> $S; grep -E "W identical|rel W|MULTI_GPU_CHECK|passed|failed" gpurun_out/r2p_pytest.log >> $S
for mode in multimem multimem1; do
  GCCNMF_MC_LIGHT_SIGNAL=1 GCCNMF_COLLECTIVE=$mode timeout 600 $TR --master-port $((29600 + RANDOM % 300)) bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2p_bench_$mode.json 2> gpurun_out/r2p_bench_$mode.err
  echo "bench light $mode rc=$?" >> $S
  python - $mode <<'PY' >> $S 2>&1
import json,sys
m=sys.argv[1]
d=json.loads(open('gpurun_out/r2p_bench_%s.json'%m).read().strip().splitlines()[-1])
print(m, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['stage_ms'])
PY
done
cat $S
