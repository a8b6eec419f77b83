"""Dev tool: every kernel of libcrs once at small sizes — meant to run under
`compute-sanitizer --tool memcheck` (one tool per gpurun call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
from helpers import clustered, queries_for
from compressed_rag_suite_b200.index import ShardIndex, merge_topk, select_topk
from compressed_rag_suite_b200.sharded import ShardedMMRSearcher, TwoStageSearcher

n, dim = 9000, 384
x, centres = clustered(n, dim, seed=1)
x[500:560] = x[3]
q = queries_for(centres, x, 140, seed=2)
q[2] = x[3]
for store in ("f16", "bf16", "i8", "b1"):
    ix = ShardIndex(dim, dtype=store)
    ix.add(x[:5000]); ix.add(x[5000:])                      # bulk ingest (+ unaligned second add)
    ix.add(x[:100])                                          # query-sized ingest
    ix.set_option("sample_rows", 1024)
    for k in (10, 100):
        ix.search(q[:1], k)                                  # single-query scan
        ix.search(q[:5], k, 0.2)                             # small batch (tensor cores or shared pass)
        ix.search(q, min(k, 24))                             # > 128 queries: sampled thresholds
        ix.search(q[:9], k, allow=np.random.default_rng(k).random(len(ix)) < 0.3)
    ix.set_option("force_path", 0)
    ix.search(q[:3], 100)
    qd = torch.from_numpy(q[:6]).cuda()
    ids, raw, cnt = ix.search(qd, 20)
    ix.score_rows(qd, ids)
    ShardedMMRSearcher(ix).search_mmr(qd, 5, 20, 0.1)
    vec = ix.fetch_rows(ids[0].cpu().numpy().view(np.uint32))
    ix.mmr(vec, np.linspace(0.9, 0.3, 20), 0.9)
    select_topk(ids, raw, 7)
    merge_topk(torch.stack([ids, ids]), torch.stack([raw, raw]), 10)
    print("ok", store, ix.last_stats()["uncertified_total"], flush=True)
    ix.close()
c, f = ShardIndex(dim, dtype="b1"), ShardIndex(dim, dtype="f16")
c.add(x); f.add(x)
qd = torch.from_numpy(q[:4]).cuda()
TwoStageSearcher(c, f).search(qd, 10, 50)
f.score_vectors(qd, torch.from_numpy(x[:80].reshape(4, 20, dim)).cuda())
torch.cuda.synchronize()
print("all ok")
