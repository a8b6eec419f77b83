set -u
O=gpurun_out
NCU_FULL="ncu --set full --clock-control none --import-source on"
S="python tools/profile_scan.py f16 1000000 384 10"
$S > $O/p_scan_f16_plain2.log 2>&1 && $NCU_FULL -k regex:^scan_kernel -s 3 -c 1 -o $O/prof_scan_f16 $S > $O/p_scan_f16_ncu.log 2>&1
S="python tools/profile_scan.py i8 12500000 384 100"
$S > $O/p_scan_i8_plain.log 2>&1 && $NCU_FULL -k regex:^scan_kernel -s 3 -c 1 -o $O/prof_scan_i8 $S > $O/p_scan_i8_ncu.log 2>&1
S="python tools/profile_scan.py b1 32000000 1024 100 4"
$S > $O/p_scan_b1x4_plain.log 2>&1 && $NCU_FULL -k regex:scan_rows_multi -s 3 -c 1 -o $O/prof_scan_b1x4 $S > $O/p_scan_b1x4_ncu.log 2>&1
for r in scan_f16 scan_i8 scan_b1x4; do python tools/ncu_summary.py rep $O/prof_$r.ncu-rep $O/prof_$r.md; done
ncu -i $O/prof_scan_b1x4.ncu-rep --page source --csv > $O/prof_scan_b1x4_source.csv 2>/dev/null
rm -f $O/prof_scan_b1x4.ncu-rep
ls -la $O | head -40
