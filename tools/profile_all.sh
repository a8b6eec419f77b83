#!/bin/bash
# Round profiling pass (run under gpurun, ONE GPU): launch lists + one `ncu --set full` capture per hot kernel.
# Every ncu command is preceded by the same command run plain (exit 0 required).
set -u
O=gpurun_out
mkdir -p $O
NCU_LIST="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
NCU_FULL="ncu --set full --clock-control none --import-source on"

B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B > $O/p_bench_plain.log 2>&1 && $NCU_LIST -c 900 --log-file $O/launches_bench.csv $B > $O/p_bench_ncu.log 2>&1
$B > $O/p_bench_plain2.log 2>&1 && $NCU_FULL -k regex:gemm_topk -s 4 -c 1 -o $O/prof_gemm_f16 $B > $O/p_gemm_ncu.log 2>&1

S="python tools/profile_scan.py f16 1000000 384 10"
$S > $O/p_scan_f16_plain.log 2>&1 && $NCU_LIST -c 200 --log-file $O/launches_scan_f16.csv $S > /dev/null 2>&1
$S > $O/p_scan_f16_plain2.log 2>&1 && $NCU_FULL -k regex:^scan_kernel -s 3 -c 1 -o $O/prof_scan_f16 $S > $O/p_scan_f16_ncu.log 2>&1

S="python tools/profile_scan.py i8 12500000 384 100"
$S > $O/p_scan_i8_plain.log 2>&1 && $NCU_FULL -k regex:^scan_kernel -s 3 -c 1 -o $O/prof_scan_i8 $S > $O/p_scan_i8_ncu.log 2>&1

S="python tools/profile_scan.py b1 32000000 1024 100"
$S > $O/p_scan_b1_plain.log 2>&1 && $NCU_FULL -k regex:scan_rows_kernel -s 3 -c 1 -o $O/prof_scan_b1 $S > $O/p_scan_b1_ncu.log 2>&1

S="python tools/profile_scan.py b1 32000000 1024 100 4"
$S > $O/p_scan_b1x4_plain.log 2>&1 && $NCU_FULL -k regex:scan_rows_multi -s 3 -c 1 -o $O/prof_scan_b1x4 $S > $O/p_scan_b1x4_ncu.log 2>&1
for r in gemm_f16 scan_f16 scan_i8 scan_b1 scan_b1x4; do
  [ -f $O/prof_$r.ncu-rep ] && python tools/ncu_summary.py rep $O/prof_$r.ncu-rep $O/prof_$r.md
done
python tools/ncu_summary.py launches $O/launches_bench.csv $O/launches_bench_summary.csv
python tools/ncu_summary.py launches $O/launches_scan_f16.csv $O/launches_scan_f16_summary.csv
# keep the reports small enough to travel back (64 MiB cap): only the contraction's full report is kept
rm -f $O/prof_scan_b1.ncu-rep $O/prof_scan_b1x4.ncu-rep
ls -la $O/
