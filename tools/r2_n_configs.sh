#!/bin/bash
# BASELINE configs 4 / 5 and the checks on N GPUs (candidate rows over NVLink peer loads)
set -u
N=${1:-2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 tools/multi_gpu_check.py > $O/r2_multi_gpu_check_w$N.log 2>&1; echo "multi_gpu_check rc=$?" >> $O/r2_multi_gpu_check_w$N.log
grep -c ": OK (world" $O/r2_multi_gpu_check_w$N.log; grep "MISMATCH\|rc=\|Error\|error" $O/r2_multi_gpu_check_w$N.log | head
timeout 600 $TR --master-port 29522 tools/bench_configs.py c4 --batch 1 --steps 50 2>/dev/null | tee $O/r2_c4_n${N}_b1.json
timeout 600 $TR --master-port 29523 tools/bench_configs.py c4 --batch 16 --steps 30 2>/dev/null | tee $O/r2_c4_n${N}_b16.json
timeout 900 $TR --master-port 29524 tools/bench_configs.py c5 --batch 1 --steps 20 2>/dev/null | tee $O/r2_c5_n${N}_b1.json
