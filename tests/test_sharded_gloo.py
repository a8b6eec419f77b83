"""CPU tier: the N>1 host logic (row partition, candidate packing, the ONE allgather) with
world_size 2 on the gloo backend.  The local searches and the merge are stood in for by the
oracle here (there is no GPU); on the GPU box the same exchange feeds libcrs' merge kernel
(tests/test_gpu_parity.py::test_shard_merge_equals_single_index)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from compressed_rag_suite_b200.sharded import (exchange_candidates, pack_candidates, shard_bounds,
                                               unpack_candidates)
from helpers import clustered, queries_for
from oracle import encode, search


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, store, n, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x, centres = clustered(n, 384, seed=5)
        q = queries_for(centres, x, 6, seed=6)
        lo, hi = shard_bounds(n, world, rank)
        codes = encode.encode_rows(x[lo:hi], store)
        qc = search.encode_queries(q, store)
        ids, raw, cnt = search.search(codes, qc, store, 384, k, row_base=lo)        # this rank's local top-k
        packed = pack_candidates(torch.from_numpy(ids.view(np.int32)), torch.from_numpy(raw))
        gathered = exchange_candidates(packed)                                       # the one collective
        assert gathered.shape == (world, 6, k, 2)
        g_ids, g_sc = unpack_candidates(gathered, store in ("i8", "b1"))
        # the product's send layout: one [2, nq, k] buffer (ids block, then raw-score bits), merged in place
        send = torch.stack((torch.from_numpy(ids.view(np.int32)), torch.from_numpy(raw).view(torch.int32)))
        g2 = exchange_candidates(send)
        assert g2.shape == (world, 2, 6, k)
        assert torch.equal(g2[:, 0], g_ids) and torch.equal(g2[:, 1], g_sc.view(torch.int32))
        m = search.merge_topk(g_ids.numpy().view(np.uint32), g_sc.numpy(), k)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ids=m[0], raw=m[1], cnt=m[2])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("store,n,k", [("f16", 3001, 10), ("i8", 2000, 7)])
def test_two_rank_exchange_and_merge(tmp_path, store, n, k):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), store, n, k, str(tmp_path)), nprocs=world, join=True)
    x, centres = clustered(n, 384, seed=5)
    q = queries_for(centres, x, 6, seed=6)
    want = search.search(encode.encode_rows(x, store), search.encode_queries(q, store), store, 384, k)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        assert np.array_equal(got["ids"], want[0])          # every rank holds the global result
        assert np.array_equal(got["raw"], want[1])
        assert np.array_equal(got["cnt"], want[2])


def test_shard_bounds_cover_rows_exactly():
    for n in (0, 1, 7, 1000, 10_000_000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(lo <= hi for lo, hi in spans)


def test_pack_keeps_score_bits():
    ids = torch.tensor([[1, -1, 7]], dtype=torch.int32)
    sc = torch.tensor([[0.25, float("-inf"), -0.0]], dtype=torch.float32)
    p = pack_candidates(ids, sc)
    i2, s2 = unpack_candidates(p[None], False)
    assert torch.equal(i2[0], ids) and torch.equal(s2[0].view(torch.int32), sc.view(torch.int32))


# ---------------------------------------------------------------- candidate-set steps over 2 ranks (configs 4 / 5)
def _worker_stages(rank, world, port, out_dir):
    """Host logic of the sharded MMR / two-stage pipelines on gloo: owners contribute their rows /
    scores, everyone else the neutral element, one MAX all-reduce assembles them (the oracle stands
    in for the local GPU calls)."""
    from compressed_rag_suite_b200.sharded import assemble_over_shards, reference_relevance
    from oracle import pipelines, postprocess
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, dim, nq, fetch, k = 4000, 384, 5, 40, 10
        x, centres = clustered(n, dim, seed=15)
        q = queries_for(centres, x, nq, seed=16)
        lo, hi = shard_bounds(n, world, rank)
        # --- local top-fetch on the int8 shard, the one allgather, merge (as in the first test)
        codes = encode.encode_rows(x, "i8")
        qc = search.encode_queries(q, "i8")
        ids, raw, cnt = search.search(codes[lo:hi], qc, "i8", dim, fetch, row_base=lo)
        g = exchange_candidates(pack_candidates(torch.from_numpy(ids.view(np.int32)), torch.from_numpy(raw)))
        g_ids, g_sc = unpack_candidates(g, True)
        m_ids, m_raw, m_cnt = search.merge_topk(g_ids.numpy().view(np.uint32), g_sc.numpy(), fetch)
        # --- candidate vectors: this rank fills in the rows it owns, zeros elsewhere; MAX assembles
        vec = np.zeros((nq, fetch, dim), dtype=np.uint8)
        own = (m_ids >= lo) & (m_ids < hi)
        vec[own] = codes[m_ids[own].astype(np.int64)].view(np.uint8)
        vec_t = assemble_over_shards(torch.from_numpy(vec))
        full = codes[m_ids.astype(np.int64)].view(np.uint8)
        assert np.array_equal(vec_t.numpy(), full), "MAX over byte codes must reproduce the owners' rows"
        # --- fine scores: owners score, others contribute -inf; MAX assembles
        f_codes = encode.encode_rows(x, "f16")
        fq = search.encode_queries(q, "f16")
        fine = np.full((nq, fetch), -np.inf, dtype=np.float32)
        for i in range(nq):
            sel = np.nonzero(own[i])[0]
            fine[i, sel] = search.raw_scores(f_codes[m_ids[i, sel].astype(np.int64)], fq[i], "f16", dim)
        fine_t = assemble_over_shards(torch.from_numpy(fine))
        # --- relevance transform equals the reference's Python arithmetic
        sims = search.similarity_from_raw(m_raw, "i8", dim)
        rel = reference_relevance(torch.from_numpy(sims)).numpy()
        want_rel = [[postprocess.distance_to_similarity(float(np.float32(1.0) - np.float32(s))) for s in row] for row in sims]
        assert rel.tolist() == want_rel
        np.savez(os.path.join(out_dir, f"stages{rank}.npz"), ids=m_ids, fine=fine_t.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_candidate_assembly(tmp_path):
    world = 2
    mp.spawn(_worker_stages, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a = np.load(os.path.join(str(tmp_path), "stages0.npz"))
    b = np.load(os.path.join(str(tmp_path), "stages1.npz"))
    assert np.array_equal(a["ids"], b["ids"]) and np.array_equal(a["fine"], b["fine"])
    # every candidate got its fine score from exactly one owner
    n, dim = 4000, 384
    x, centres = clustered(n, dim, seed=15)
    q = queries_for(centres, x, 5, seed=16)
    f_codes = encode.encode_rows(x, "f16")
    fq = search.encode_queries(q, "f16")
    for i in range(5):
        want = search.raw_scores(f_codes[a["ids"][i].astype(np.int64)], fq[i], "f16", dim)
        assert np.array_equal(want, a["fine"][i])
