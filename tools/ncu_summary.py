"""Turn gpurun_out/*.ncu-rep / launches.csv into the small summaries committed under profiles/."""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors_srcunit_tex_op_read.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second", "sm__inst_executed.sum",
        "smsp__inst_executed.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum"]


def rep_summary(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        for vals in rows[2:]:
            d = dict(zip(hdr, zip(vals, units)))
            f.write(f"## {d.get('Kernel Name', ('?',))[0]}  grid={d.get('Grid Size', ('?',))[0]} block={d.get('Block Size', ('?',))[0]}\n\n")
            f.write("| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k][0]} | {d[k][1]} |\n")
            # pipe mix: which execution pipe bounds a compute-bound kernel
            for k in hdr:
                if k not in KEYS and k.endswith(".avg.pct_of_peak_sustained_active") and (
                        k.startswith("sm__inst_executed_pipe_") or k.startswith("smsp__issue_active")):
                    if d[k][0] not in ("0", "", "n/a"):
                        f.write(f"| {k} | {d[k][0]} | {d[k][1]} |\n")
            f.write("\n")


def launches_summary(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        us = v / 1e3 if u.startswith("n") else (v * 1e3 if u.startswith("m") else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write("kernel,launches,total_us,avg_us,share\n")
        for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{name}\",{n},{us:.1f},{us / n:.1f},{us / tot:.4f}\n")


if __name__ == "__main__":
    kind, src, dst = sys.argv[1:4]
    (rep_summary if kind == "rep" else launches_summary)(src, dst)
