#!/bin/bash
# 2-GPU call: parity of the sharded pipeline (three forms of the exchange) + the bench line for each
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
S=gpurun_out/r2n_summary.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== 2 GPUs" > $S
timeout 600 python -m pytest tests/test_gpu_multi.py -q -s > gpurun_out/r2n_pytest.log 2>&1
echo "multi pytest rc=$?" >> $S; grep -E "W identical|rel W|MULTI_GPU_CHECK|passed|failed" gpurun_out/r2n_pytest.log >> $S
for mode in multimem multimem1 nccl; do
  GCCNMF_COLLECTIVE=$mode timeout 600 $TR --master-port $((29600 + RANDOM % 300)) bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2n_bench_$mode.json 2> gpurun_out/r2n_bench_$mode.err
  echo "bench $mode rc=$?" >> $S
  python - $mode <<'PY' >> $S 2>&1
import json,sys
m=sys.argv[1]
d=json.loads(open('gpurun_out/r2n_bench_%s.json'%m).read().strip().splitlines()[-1])
print(m, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['stage_ms'])
PY
done
cat $S
