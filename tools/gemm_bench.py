"""Dev tool: time the batched (tcgen05) path at the config-3 shape and report fast-vs-exact error."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from compressed_rag_suite_b200.index import ShardIndex

def run(n, dim, nq, k, iters=10, store="f16", cluster=0, warm=3, prefetch=(0,), sample=(65536,)):
    ix = ShardIndex(dim, dtype=store, reserve_rows=n)
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    cen = torch.randn(4096, dim, device="cuda", generator=g); cen /= cen.norm(dim=1, keepdim=True)
    blk = 1 << 20
    for off in range(0, n, blk):
        m = min(blk, n - off)
        z = torch.randn(m, dim, device="cuda", generator=g); z /= z.norm(dim=1, keepdim=True)
        x = 0.6 * cen[torch.arange(off, off + m, device="cuda") % 4096] + 0.8 * z
        ix.add(x)
    z = torch.randn(nq, dim, device="cuda", generator=g); z /= z.norm(dim=1, keepdim=True)
    q = 0.6 * cen[torch.randint(0, 4096, (nq,), device="cuda", generator=g)] + 0.8 * z
    ix.set_option("profiling", 1)
    ix.set_option("gemm_cluster", cluster)
    for pf in prefetch:
      for smp in sample:
        ix.set_option("gemm_prefetch", pf)
        ix.set_option("sample_rows", smp)
        for _ in range(warm):
            ix.search(q, k)
        torch.cuda.synchronize()
        ts, ks = [], []
        for _ in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ix.search(q, k); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b)); ks.append(ix.last_kernel_ms())
        ts.sort(); ks.sort()
        st = ix.last_stats()
        flops = 2.0 * ((nq + 127) // 128 * 128) * n * ix.dim_padded
        print(json.dumps({"cluster": cluster, "prefetch": pf, "sample": smp, "n": n, "dim": dim, "nq": nq, "k": k, "store": store, "ms_med": round(ts[len(ts)//2], 3),
                          "kernel_ms_med": round(ks[len(ks)//2], 3), "qps": round(nq / ts[len(ts)//2] * 1e3),
                          "TFLOPs_kernel": round(flops / ks[len(ks)//2] / 1e9, 1),
                          "launches": st["kernel_launches"], "uncert": st["uncertified_total"]}), flush=True)
    ix.close()

if __name__ == "__main__":
    # A/B on one box, sustained (power-capped) clocks: every variant after ~1 s of load
    run(10_000_000, 384, 1024, 10, cluster=2, iters=30, warm=120, prefetch=(0, 2, 4, 0, 2), sample=(65536,))
    run(10_000_000, 384, 1024, 10, cluster=2, iters=30, warm=120, prefetch=(0,), sample=(0, 65536))
    run(1_250_000, 384, 1024, 10, cluster=2, iters=30, warm=50, prefetch=(0, 2), sample=(0, 65536))
