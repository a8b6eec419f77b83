"""Dev tool: shared-pass integer scans (1024-bit rows) — time per batch for group caps 0/2/4/8."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from compressed_rag_suite_b200.index import ShardIndex

def run(n, dim, nq, k, store="b1"):
    ix = ShardIndex(dim, dtype=store, reserve_rows=n)
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    for off in range(0, n, 1 << 20):
        ix.add(torch.randn(min(1 << 20, n - off), dim, device="cuda", generator=g))
    q = torch.randn(nq, dim, device="cuda", generator=g)
    ix.set_option("profiling", 1)
    ix.set_option("force_path", 0)
    for cap, short in ((0, 0), (0, 1), (4, 0), (2, 1), (4, 1), (8, 1)):
        ix.set_option("multi_scan", cap)
        ix.set_option("short_lists", short)
        for _ in range(3):
            ix.search(q, k)
        ks = []
        for _ in range(8):
            ix.search(q, k); ks.append(ix.last_kernel_ms())
        ks.sort()
        ms = ks[len(ks) // 2]
        print(json.dumps({"store": store, "n": n, "dim": dim, "nq": nq, "k": k, "group_cap": cap, "short_lists": short, "kernel_ms": round(ms, 3),
                          "ms_per_query": round(ms / nq, 3), "GBps_per_pass_equiv": round(n * ix.row_bytes * nq / ms / 1e6, 1),
                          "launches": ix.last_stats()["kernel_launches"]}), flush=True)
    ix.close()

if __name__ == "__main__":
    run(32_000_000, 1024, 8, 100)
    run(32_000_000, 1024, 8, 10)
    run(32_000_000, 1024, 1, 100)
    run(12_500_000, 384, 1, 100, store="i8")
    run(16_000_000, 128, 7, 10, store="i8")
