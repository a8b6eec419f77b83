"""Quick device-side timing of the single-query scan path (config 2 shape) — dev tool."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from compressed_rag_suite_b200.index import ShardIndex

def run(store, n, dim, k, iters=20):
    torch.manual_seed(0)
    ix = ShardIndex(dim, dtype=store, reserve_rows=n)
    blk = 1 << 20
    for off in range(0, n, blk):
        m = min(blk, n - off)
        x = torch.randn(m, dim, device="cuda")
        ix.add(x)
    q = torch.randn(1, dim, device="cuda")
    for _ in range(5):
        ix.search(q, k)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record(); ix.search(q, k); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    med = ts[len(ts) // 2]
    gb = n * ix.row_bytes / 1e9
    st = ix.last_stats()
    print(json.dumps({"store": store, "n": n, "dim": dim, "k": k, "ms_med": round(med, 4), "ms_min": round(ts[0], 4),
                      "GBps_med": round(gb / med * 1e3, 1), "GBps_best": round(gb / ts[0] * 1e3, 1),
                      "launches": st["kernel_launches"], "uncert": st["uncertified_total"]}), flush=True)
    ix.close()

if __name__ == "__main__":
    run("f16", 1_000_000, 384, 10)
    run("f16", 4_000_000, 384, 10)
    run("bf16", 4_000_000, 384, 10)
    run("i8", 8_000_000, 384, 10)
    run("i8", 8_000_000, 384, 100)
    run("b1", 32_000_000, 1024, 100)
    run("f16", 2_000_000, 1024, 10)
