"""Dev tool: step time vs dominant-kernel time of one search shape (ONE GPU), as the per-shard work of an
N-GPU run looks from one rank.  Run it plain for the event timings, and under
`ncu --metrics gpu__time_duration.sum` for the per-kernel launch list of the same steps.

    python tools/step_breakdown.py [store] [rows] [dim] [nq] [k] [iters]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import compressed_rag_suite_b200._native as _N
if os.environ.get("CRS_LIB"):                      # A/B of two builds of the library on one box
    _N.LIB_PATH = os.path.abspath(os.environ["CRS_LIB"])
from compressed_rag_suite_b200.index import ShardIndex


def build(store, n, dim):
    ix = ShardIndex(dim, dtype=store, reserve_rows=n)
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    cen = torch.randn(4096, dim, device="cuda", generator=g)
    cen /= cen.norm(dim=1, keepdim=True)
    for off in range(0, n, 1 << 20):
        m = min(1 << 20, n - off)
        z = torch.randn(m, dim, device="cuda", generator=g)
        z /= z.norm(dim=1, keepdim=True)
        ix.add(0.6 * cen[torch.arange(off, off + m, device="cuda") % 4096] + 0.8 * z)
    return ix, cen, g


def main():
    a = sys.argv[1:]
    store = a[0] if len(a) > 0 else "f16"
    n = int(a[1]) if len(a) > 1 else 1_250_000
    dim = int(a[2]) if len(a) > 2 else 384
    nq = int(a[3]) if len(a) > 3 else 1024
    k = int(a[4]) if len(a) > 4 else 10
    iters = int(a[5]) if len(a) > 5 else 20
    ix, cen, g = build(store, n, dim)
    z = torch.randn(nq, dim, device="cuda", generator=g)
    z /= z.norm(dim=1, keepdim=True)
    q = (0.6 * cen[torch.randint(0, 4096, (nq,), device="cuda", generator=g)] + 0.8 * z).contiguous()
    ix.set_option("profiling", 1)
    opts = os.environ.get("CRS_OPTS", "")
    for kv in filter(None, opts.split(",")):                  # e.g. CRS_OPTS=share_floor=0,sample_rows=65536
        name, val = kv.split("=")
        ix.set_option(name, int(val))
    for _ in range(5):
        ix.search(q, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ix.search(q, k)
    e1.record()
    torch.cuda.synchronize()
    step = e0.elapsed_time(e1) / iters
    kms = ix.kernel_ms_history()[-iters:]
    st = ix.last_stats()
    out = {"store": store, "rows": n, "dim": dim, "nq": nq, "k": k, "step_ms": round(step, 4),
           "kernel_ms": round(sum(kms) / len(kms), 4), "outside_kernel_ms": round(step - sum(kms) / len(kms), 4),
           "launches": st["kernel_launches"], "path": st["path"], "opts": opts, "uncertified": st["uncertified_total"]}
    # the same step replayed from a CUDA graph (no host launch gaps)
    try:
        gs = ix.capture_search(nq, k)
        gs.queries.copy_(q)
        for _ in range(3):
            gs.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            gs.replay()
        e1.record()
        torch.cuda.synchronize()
        out["graph_step_ms"] = round(e0.elapsed_time(e1) / iters, 4)
    except Exception as ex:  # noqa: BLE001
        out["graph_error"] = str(ex)[:200]
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
