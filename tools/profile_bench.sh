#!/bin/bash
# Launch list of bench.py + one `ncu --set full` capture of the contraction (ONE GPU, under gpurun).
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B > $O/p_bench_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv -c 900 --log-file $O/launches_bench.csv $B > $O/p_bench_ncu.log 2>&1
$B > $O/p_bench_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 9 -c 1 -o $O/prof_gemm_f16 $B > $O/p_gemm_ncu.log 2>&1
python tools/ncu_summary.py rep $O/prof_gemm_f16.ncu-rep $O/prof_gemm_f16.md
python tools/ncu_summary.py launches $O/launches_bench.csv $O/launches_bench_summary.csv
grep crs $O/launches_bench_summary.csv
head -12 $O/prof_gemm_f16.md
