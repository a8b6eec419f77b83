#!/bin/bash
set -u
N=8
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
bash tools/r2_n2.sh 8
timeout 600 $TR --master-port 29522 tools/bench_configs.py c4 --batch 1 --steps 50 2>/dev/null | tee $O/r2_c4_n${N}_b1.json
timeout 600 $TR --master-port 29523 tools/bench_configs.py c4 --batch 16 --steps 30 2>/dev/null | tee $O/r2_c4_n${N}_b16.json
timeout 900 $TR --master-port 29524 tools/bench_configs.py c5 --batch 1 --steps 20 2>/dev/null | tee $O/r2_c5_n${N}_b1.json
