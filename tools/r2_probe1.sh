#!/bin/bash
# round-2 probe 1 (ONE GPU): where does a 1.25M-row shard step spend its time?
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2p1_smi.txt
S="python tools/step_breakdown.py f16 1250000 384 1024 10 20"
$S > $O/r2p1_step_f16.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv -s 60 -c 120 --log-file $O/r2p1_launches_f16.csv $S > $O/r2p1_ncu_f16.log 2>&1
python tools/step_breakdown.py i8 12500000 384 1 10 20 > $O/r2p1_step_i8.log 2>&1
python tools/step_breakdown.py f16 10000000 384 1024 10 10 > $O/r2p1_step_f16_10m.log 2>&1
python tools/step_breakdown.py i8 10000000 384 1024 10 10 > $O/r2p1_step_i8_10m.log 2>&1
python tools/gemm_stalls.py shard > $O/r2p1_stalls.log 2>&1
python tools/ncu_summary.py launches $O/r2p1_launches_f16.csv $O/r2p1_launches_f16_summary.csv
cat $O/r2p1_step_*.log $O/r2p1_stalls.log; grep crs $O/r2p1_launches_f16_summary.csv
