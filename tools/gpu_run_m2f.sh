#!/bin/bash
# 2-GPU call: timeline of the pull exchange
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
S=gpurun_out/r2r_summary.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== timeline" > $S
GCCNMF_DEBUG_EXCHANGE=1 STAMP_MODES="fused,pull-one-shot,pull-two-shot,one-shot" timeout 400 $TR --master-port 29571 tools/multi_stamps.py > gpurun_out/r2r_stamps.log 2>&1
echo "multi_stamps rc=$?" >> $S; grep "^rank 0\|not available\|failed\|Error" gpurun_out/r2r_stamps.log | sed 's/| spans.*| after G4 end:/| after G4 end:/' >> $S
cat $S
