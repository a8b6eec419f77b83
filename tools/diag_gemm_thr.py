"""Dev tool: locate a GEMM-path vs scan-path mismatch under a threshold and say which one the oracle agrees with."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from helpers import clustered, queries_for
from oracle import encode, search
from compressed_rag_suite_b200.index import ShardIndex

x, centres = clustered(50000, 384, seed=130)
q = queries_for(centres, x, 256, seed=131)
ix = ShardIndex(384)
ix.add(x)
codes = encode.encode_rows(x, "f16", "cosine")
qc = search.encode_queries(q, "f16", "cosine")
for thr in (-np.inf, 0.293, 0.6):
    want = search.search(codes, qc, "f16", 384, 10, thr)
    for rep in range(3):
        ix.set_option("force_path", 1)
        a = ix.search(q, 10, thr)
        sa = ix.last_stats()
        ix.set_option("force_path", 0)
        b = ix.search(q, 10, thr)
        sb = ix.last_stats()
        for name, g in (("gemm", a), ("scan", b)):
            bad = [i for i in range(256) if not (np.array_equal(g[0][i], want[0][i]) and g[2][i] == want[2][i]
                                                  and np.array_equal(g[1][i].view(np.uint32), want[1][i].view(np.uint32)))]
            print(f"thr={thr} rep={rep} {name}: {len(bad)} queries differ from oracle; stats={sa if name=='gemm' else sb}", flush=True)
            for i in bad[:3]:
                print("  q", i, "count got/want", g[2][i], want[2][i])
                print("   ids got ", g[0][i].tolist())
                print("   ids want", want[0][i].tolist())
                print("   sc got ", g[1][i].tolist())
                print("   sc want", want[1][i].tolist())
