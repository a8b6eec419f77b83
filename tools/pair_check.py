"""Dev tool: CTA-pair (cta_group::2) contraction vs the cluster-multicast one — results must be identical."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from helpers import clustered, queries_for
from compressed_rag_suite_b200.index import ShardIndex

ok = True
for store, n, dim, nq, k in [("f16", 50000, 384, 256, 10), ("f16", 30011, 384, 130, 10), ("i8", 40000, 384, 300, 10),
                             ("bf16", 20000, 256, 200, 20), ("f16", 9000, 128, 1000, 100), ("f16", 700, 64, 129, 10)]:
    x, centres = clustered(n, dim, seed=n)
    q = queries_for(centres, x, nq, seed=nq)
    ix = ShardIndex(dim, dtype=store)
    ix.add(x)
    ix.set_option("gemm_cluster", 2)
    a = ix.search(q, k)
    ix.set_option("gemm_cluster", 22)
    b = ix.search(q, k)
    st = ix.last_stats()
    same = all(np.array_equal(u, v) for u, v in zip(a, b))
    ok &= same
    print(store, n, dim, nq, k, "path", st["path"], "grid", st["grid"], "SAME" if same else "DIFFERENT", flush=True)
    if not same:
        bad = [i for i in range(nq) if not np.array_equal(a[0][i], b[0][i])]
        print("  first differing queries", bad[:5], a[0][bad[0]].tolist(), b[0][bad[0]].tolist())
sys.exit(0 if ok else 1)
