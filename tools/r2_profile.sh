#!/bin/bash
# Round-2 profiling pass (ONE GPU, under gpurun): launch list of bench.py, ncu --set full of the f16 and int8 contractions.
# Every ncu command is preceded by the same command run plain (exit 0 required).
set -u
O=gpurun_out
mkdir -p $O
NCU_LIST="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
NCU_FULL="ncu --set full --clock-control none --import-source on"
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B > $O/r2_prof_bench_plain.log 2>&1 && $NCU_LIST -c 900 --log-file $O/r2_launches_bench.csv $B > $O/r2_prof_bench_ncu.log 2>&1
$B > $O/r2_prof_bench_plain2.log 2>&1 && $NCU_FULL -k regex:gemm_topk -s 4 -c 1 -o $O/r2_prof_gemm_f16 $B > $O/r2_prof_gemm_f16.log 2>&1
S="python tools/step_breakdown.py i8 10000000 384 1024 10 3"
$S > $O/r2_prof_i8_plain.log 2>&1 && $NCU_FULL -k regex:gemm_topk -s 4 -c 1 -o $O/r2_prof_gemm_i8 $S > $O/r2_prof_gemm_i8.log 2>&1
for r in gemm_f16 gemm_i8; do
  [ -f $O/r2_prof_$r.ncu-rep ] && python tools/ncu_summary.py rep $O/r2_prof_$r.ncu-rep $O/r2_prof_$r.md
done
python tools/ncu_summary.py launches $O/r2_launches_bench.csv $O/r2_launches_bench_summary.csv
grep -i "crs\|gemm_topk\|finalize\|encode\|xmerge\|exact" $O/r2_launches_bench_summary.csv
head -30 $O/r2_prof_gemm_f16.md; head -30 $O/r2_prof_gemm_i8.md
ls -la $O/*.ncu-rep
