"""Dev tool for ncu: a few single-query searches on one store dtype (python tools/profile_scan.py f16 1000000 384 10)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from compressed_rag_suite_b200.index import ShardIndex

store, n, dim, k = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
nq = int(sys.argv[5]) if len(sys.argv) > 5 else 1
ix = ShardIndex(dim, dtype=store, reserve_rows=n)
g = torch.Generator(device="cuda"); g.manual_seed(0)
for off in range(0, n, 1 << 20):
    ix.add(torch.randn(min(1 << 20, n - off), dim, device="cuda", generator=g))
q = torch.randn(nq, dim, device="cuda", generator=g)
ix.set_option("force_path", 0)
for _ in range(6):
    out = ix.search(q, k)
torch.cuda.synchronize()
print("ok", store, n, dim, k, ix.last_stats())
