"""Multi-GPU check, run under torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/multi_gpu_check.py

Every rank builds its row shard AND (sizes are small) a full single-GPU index; the sharded
result of each pipeline — top-k, top-100 -> MMR -> 10, Hamming top-100 -> fp16 rescoring —
must equal the single-index result bit for bit on every rank.  Prints one line per check.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import torch.distributed as dist

from helpers import clustered, queries_for
from compressed_rag_suite_b200.index import ShardIndex
from compressed_rag_suite_b200.sharded import ShardedMMRSearcher, ShardedSearcher, TwoStageSearcher, shard_bounds


def same(a, b):
    return all(torch.equal(u.view(torch.int32) if u.dtype == torch.float32 else u,
                           v.view(torch.int32) if v.dtype == torch.float32 else v) for u, v in zip(a, b))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok_all = True

    def report(name, ok):
        nonlocal ok_all
        t = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok_all &= bool(t.item())
        if rank == 0:
            print(f"{name}: {'OK' if t.item() else 'MISMATCH'} (world {world})", flush=True)

    for store, n, dim, nq, k in [("f16", 60000, 384, 300, 10), ("i8", 50000, 384, 40, 10), ("b1", 40000, 1024, 5, 100),
                                 ("f16", 7001, 384, 1, 10), ("bf16", 30000, 256, 16, 20)]:
        x, centres = clustered(n, dim, seed=n)
        q = torch.from_numpy(queries_for(centres, x, nq, seed=nq)).cuda()
        lo, hi = shard_bounds(n, world, rank)
        full = ShardIndex(dim, dtype=store, device=local)
        full.add(x)
        shard = ShardIndex(dim, dtype=store, device=local, row_base=lo)
        shard.add(x[lo:hi])
        peer = ShardedSearcher(shard, exchange="peer", max_nq=512)
        nccl = ShardedSearcher(shard, exchange="nccl")
        for thr in (-float("inf"), 0.3):
            want = full.search(q, k, thr)
            report(f"top-{k} {store} n={n} nq={nq} thr={thr} [peer-memory exchange kernel]", same(peer.search(q, k, thr), want))
            report(f"top-{k} {store} n={n} nq={nq} thr={thr} [NCCL allgather + merge]", same(nccl.search(q, k, thr), want))
        # the whole sharded step replayed from one CUDA graph, with fresh queries copied in between replays
        want = full.search(q, k)
        gs = peer.capture(nq, k)
        ok = True
        for rep in range(3):
            qq = torch.roll(q, rep, 0)
            gs.queries.copy_(qq)
            got = gs.replay()
            torch.cuda.synchronize()
            ok &= same(got, full.search(qq, k))
        report(f"top-{k} {store} n={n} nq={nq} [sharded step as one CUDA graph, 3 replays]", ok)
        timed_out, step = peer._peer.status()
        report(f"exchange of {store}: no wait timed out (step {step})", not timed_out)
        if store in ("f16", "i8"):
            w = ShardedMMRSearcher(full, local_only=True)
            want = w.search_mmr(q[:8], 10, 100, 0.1)
            got = ShardedMMRSearcher(shard, exchange="peer", max_nq=512).search_mmr(q[:8], 10, 100, 0.1)
            report(f"top-100 -> MMR -> 10 {store} [candidate rows by NVLink peer loads]", same(got, want))
            got = ShardedMMRSearcher(shard, exchange="nccl").search_mmr(q[:8], 10, 100, 0.1)
            report(f"top-100 -> MMR -> 10 {store} [candidate rows by MAX all-reduce]", same(got, want))
        full.close(); shard.close()

    dim, n = 1024, 50000
    x, centres = clustered(n, dim, seed=77)
    q = torch.from_numpy(queries_for(centres, x, 8, seed=78)).cuda()
    lo, hi = shard_bounds(n, world, rank)
    fc, ff = ShardIndex(dim, dtype="b1", device=local), ShardIndex(dim, dtype="f16", device=local)
    fc.add(x); ff.add(x)
    sc, sf = ShardIndex(dim, dtype="b1", device=local, row_base=lo), ShardIndex(dim, dtype="f16", device=local, row_base=lo)
    sc.add(x[lo:hi]); sf.add(x[lo:hi])
    want = TwoStageSearcher(fc, ff, local_only=True).search(q, 10, 100)      # no collectives involved
    got = TwoStageSearcher(sc, sf, exchange="peer").search(q, 10, 100)
    report("Hamming top-100 -> fp16 rescoring -> 10 [fine rows by NVLink peer loads]", same(got, want))
    got = TwoStageSearcher(sc, sf, exchange="nccl").search(q, 10, 100)
    report("Hamming top-100 -> fp16 rescoring -> 10 [fine scores by MAX all-reduce]", same(got, want))

    dist.barrier()
    dist.destroy_process_group()
    if not ok_all:
        sys.exit(1)


if __name__ == "__main__":
    main()
