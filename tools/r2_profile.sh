#!/bin/bash
# Round-2 profiling pass (ONE GPU, under gpurun): launch list of bench.py, ncu --set full of the f16 and int8 contractions.
# Every ncu command is preceded by the same command run plain (exit 0 required).
set -u
O=gpurun_out
mkdir -p $O
NCU_LIST="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
NCU_FULL="ncu --set full --clock-control none"
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
# per-instruction stall sampling of the int8 contraction's hot loop, then drop the big reports (64 MiB cap on gpurun_out)
ncu -i $O/r2_prof_gemm_i8.ncu-rep --page source --csv > $O/r2_prof_gemm_i8_source.csv 2>/dev/null
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2_prof_gemm_i8_source.csv", errors="replace")))
hdr = next((r for r in rows if "Source" in r or "# Samples" in " ".join(r)), None)
if hdr:
    i_src = hdr.index("Source") if "Source" in hdr else 1
    i_smp = next((i for i, h in enumerate(hdr) if h.strip().startswith("# Samples")), None)
    if i_smp is not None:
        agg = []
        for r in rows[rows.index(hdr) + 1:]:
            try:
                agg.append((int(r[i_smp].replace(",", "")), r[i_src].strip()[:110]))
            except Exception:
                pass
        tot = sum(a for a, _ in agg) or 1
        with open("gpurun_out/r2_prof_gemm_i8_stalls.md", "w") as f:
            f.write("| share of warp samples | SASS |\n|---|---|\n")
            for n, src in sorted(agg, reverse=True)[:40]:
                f.write(f"| {n / tot:.3f} | `{src}` |\n")
PY
rm -f $O/*.ncu-rep $O/r2_prof_gemm_i8_source.csv
ls -la $O/ | head -40
