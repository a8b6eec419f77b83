"""Dev tool: where does the MMA issuer wait?  Needs the instrumented build (libcrs_prof.so, -DCRS_GEMM_PROFILE)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compressed_rag_suite_b200._native as N
N.LIB_PATH = os.path.join(os.path.dirname(N.LIB_PATH), "libcrs_prof.so")
import torch
from compressed_rag_suite_b200.index import ShardIndex

def run(n, dim, nq, k, store, cluster):
    ix = ShardIndex(dim, dtype=store, reserve_rows=n)
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    cen = torch.randn(4096, dim, device="cuda", generator=g); cen /= cen.norm(dim=1, keepdim=True)
    for off in range(0, n, 1 << 20):
        m = min(1 << 20, n - off)
        z = torch.randn(m, dim, device="cuda", generator=g); z /= z.norm(dim=1, keepdim=True)
        ix.add(0.6 * cen[torch.arange(off, off + m, device="cuda") % 4096] + 0.8 * z)
    z = torch.randn(nq, dim, device="cuda", generator=g); z /= z.norm(dim=1, keepdim=True)
    q = 0.6 * cen[torch.randint(0, 4096, (nq,), device="cuda", generator=g)] + 0.8 * z
    ix.set_option("gemm_cluster", cluster)
    ix.set_option("profiling", 1)
    lib = N.lib()
    buf = (C.c_ulonglong * 32)()
    for _ in range(5):
        ix.search(q, k)
    torch.cuda.synchronize()
    lib.crs_debug_gemm_profile(buf, 1)
    for _ in range(10):
        ix.search(q, k)
    torch.cuda.synchronize()
    lib.crs_debug_gemm_profile(buf, 1)
    wf, we, tot, cnt = buf[0], buf[1], buf[2], buf[3]
    ewait, etot, ecnt, tiles = buf[4], buf[5], buf[6], buf[7]
    print(f"{store} n={n} nq={nq} k={k} cluster={cluster} grid={ix.last_stats()['grid']}: kernel_ms={ix.last_kernel_ms():.3f} issuers={cnt} "
          f"wait_operands={wf / tot:.3f} wait_accumulator_drained={we / tot:.3f} of issuer time; "
          f"issuer cycles/tile={tot / max(tiles, 1):.0f} (issuing {(tot - wf - we) / max(tiles, 1):.0f}); "
          f"epilogue warp: cycles/tile={etot / max(ecnt, 1) / (tiles / max(cnt, 1)):.0f}, waiting for an accumulator {ewait / max(etot, 1):.3f}",
          flush=True)
    pts = [0] + [1 << j for j in range(14)]
    curve = [(pts[j], round(buf[8 + j] / max(cnt, 1))) for j in range(15) if buf[8 + j]]
    hp = buf[24:32]
    if hp[2]:
        print(f"   floor helper: {hp[0] / hp[2]:.1f} rounds per CTA, {hp[1] / max(hp[0], 1):.0f} cycles per round, first floor stored at cycle "
              f"{hp[3] / max(hp[4], 1):.0f} ({hp[4]} of {hp[2]} helpers stored one)", flush=True)
    thr = cnt * 128.0                     # epilogue threads (issuers x 128 query rows)
    if hp[6]:
        print(f"   insert path per epilogue thread: {hp[5] / thr:.1f} inserts, {hp[6] / thr:.1f} slabs took the insert path "
              f"(first 8 tiles after the warm-up: {(hp[7] >> 32) / thr:.1f} inserts, {(hp[7] & 0xFFFFFFFF) / thr:.1f} slabs)", flush=True)
    print("   progress (tile index: cycles since CTA start when the issuer begins it):", curve, "end:", round(buf[23] / max(cnt, 1)), flush=True)
    ix.close()

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "shard":          # the per-GPU work of the 8-GPU headline run
        run(1_250_000, 384, 1024, 10, "f16", 2)
        run(1_250_000, 384, 1024, 10, "i8", 2)
        run(10_000_000, 384, 1024, 10, "i8", 2)
        sys.exit(0)
    run(10_000_000, 384, 1024, 10, "f16", 2)
    run(10_000_000, 384, 1024, 10, "f16", 22)
    run(10_000_000, 384, 1024, 10, "i8", 2)
    run(10_000_000, 384, 1024, 100, "f16", 2)
    run(10_000_000, 384, 128, 10, "f16", 0)
