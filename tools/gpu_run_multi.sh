#!/bin/bash
# Multi-GPU call: usage  gpurun --gpus N -- 'bash tools/gpu_run_multi.sh N [extra]'
#   N = 2: sharded-vs-single parity test (multimem + NCCL fallback) and the 2-GPU bench line
#   N = 4: bench + configs[4] (four independent low-latency streams)
#   N = 8: bench + configs[3] (10 min, K = 4096, 200 iterations, frame-sharded)
cd "$(dirname "$0")/.." || exit 1
N=${1:-2}
mkdir -p gpurun_out
S=gpurun_out/r2m${N}_summary.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2m${N}_smi.txt 2>&1
echo "== $N GPUs" > $S
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -q -s > gpurun_out/r2m${N}_pytest.log 2>&1
  echo "multi pytest rc=$?" >> $S; grep -E "collective:|W identical|rel W|rel signal|MULTI_GPU_CHECK|passed|failed" gpurun_out/r2m${N}_pytest.log >> $S
fi
timeout 900 $TR --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2m${N}_bench.json 2> gpurun_out/r2m${N}_bench.err
echo "bench rc=$?" >> $S
python - $N <<'PY' >> $S 2>&1
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r2m%s_bench.json'%n).read().strip().splitlines()[-1])
    print('bench', n, 'GPUs value', d['value'], 'e2e', d['e2e']['value'], d['stage_ms'], d.get('collective'))
except Exception as e:
    print('bench parse failed', e)
PY
if [ "$N" = "2" ]; then
  GCCNMF_COLLECTIVE=nccl timeout 900 $TR --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2m${N}_bench_nccl.json 2> gpurun_out/r2m${N}_bench_nccl.err
  echo "bench nccl rc=$?" >> $S
  python -c "
import json
d=json.loads(open('gpurun_out/r2m2_bench_nccl.json').read().strip().splitlines()[-1]); print('bench nccl value', d['value'], d['stage_ms']['nmf'], d.get('collective'))" >> $S 2>&1
fi
if [ "$N" = "4" ]; then
  timeout 900 $TR --master-port 29551 tools/run_config5.py --json gpurun_out/r2m4_config5.json > gpurun_out/r2m4_config5.log 2>&1
  echo "config5 rc=$?" >> $S; tail -1 gpurun_out/r2m4_config5.log >> $S
fi
if [ "$N" = "8" ]; then
  timeout 1500 $TR --master-port 29541 tools/run_config4.py --json gpurun_out/r2m8_config4.json > gpurun_out/r2m8_config4.log 2>&1
  echo "config4 rc=$?" >> $S; tail -1 gpurun_out/r2m8_config4.log >> $S
fi
cat $S
