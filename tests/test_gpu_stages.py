"""GPU tier (-m gpu): the candidate-set steps after the top-k — K8 rescoring, unsorted select,
top-100 -> MMR -> 10 (BASELINE config 4) and Hamming top-100 -> fp16 rescoring (config 5) —
bit-exact against oracle/pipelines.py, single shard and row-sharded."""
import numpy as np
import pytest
import torch

from helpers import clustered, queries_for
from oracle import encode, pipelines, search

pytestmark = pytest.mark.gpu

from compressed_rag_suite_b200.index import ShardIndex, select_topk  # noqa: E402
from compressed_rag_suite_b200.sharded import (ShardedMMRSearcher, TwoStageSearcher, reference_relevance,  # noqa: E402
                                               shard_bounds, similarity_of)


@pytest.mark.parametrize("store", ["f16", "bf16", "i8", "b1"])
def test_score_rows_equals_search_scores(store):
    x, centres = clustered(6000, 384, seed=300)
    q = queries_for(centres, x, 12, seed=301)
    ix = ShardIndex(384, dtype=store, row_base=1000)
    ix.add(x)
    ids, raw, cnt = ix.search(q, 50)
    got = ix.score_rows(q, ids)
    assert np.array_equal(got.view(np.uint32), raw.view(np.uint32))
    # rows this shard does not hold and pad ids get the absent score
    other = ids.copy()
    other[:, 0] = 5                     # below row_base
    other[:, 1] = 1000 + 6000           # one past the end
    other[:, 2] = 0xFFFFFFFF
    got = ix.score_rows(q, other)
    absent = -np.inf if store in ("f16", "bf16") else np.iinfo(np.int32).min
    assert (got[:, :3] == absent).all()
    assert np.array_equal(got[:, 3:].view(np.uint32), raw[:, 3:].view(np.uint32))
    # device form
    gd = ix.score_rows(torch.from_numpy(q).cuda(), torch.from_numpy(ids.view(np.int32)).cuda())
    assert np.array_equal(gd.cpu().numpy().view(np.uint32), raw.view(np.uint32))


@pytest.mark.parametrize("is_int", [False, True])
@pytest.mark.parametrize("m,k", [(100, 10), (32, 32), (7, 3), (128, 100)])
def test_select_topk_orders_unsorted_candidates(is_int, m, k):
    rng = np.random.default_rng(m + k)
    nq = 9
    ids = np.stack([rng.choice(100000, m, replace=False) for _ in range(nq)]).astype(np.uint32)
    if is_int:
        sc = rng.integers(-50, 50, (nq, m)).astype(np.int32)        # many ties -> id order decides
    else:
        sc = rng.integers(-50, 50, (nq, m)).astype(np.float32) / 8
    ids[0, ::3] = 0xFFFFFFFF                                         # pads are skipped
    want = search.merge_topk(ids[None], sc[None], k)
    got = select_topk(torch.from_numpy(ids.view(np.int32)).cuda(), torch.from_numpy(sc).cuda(), k)
    assert np.array_equal(got[0].cpu().numpy().view(np.uint32), want[0])
    assert np.array_equal(got[1].cpu().numpy().view(np.uint32), want[1].view(np.uint32))
    assert np.array_equal(got[2].cpu().numpy(), want[2])


def _shards(x, dim, store, g):
    out = []
    for r in range(g):
        lo, hi = shard_bounds(len(x), g, r)
        ix = ShardIndex(dim, dtype=store, row_base=lo)
        ix.add(x[lo:hi])
        out.append(ix)
    return out


@pytest.mark.parametrize("g", [1, 3])
def test_two_stage_hamming_then_fp16_rescoring(g):
    """config 5 in small: 1-bit coarse top-100, fp16 rescoring, top-10."""
    dim, n, nq, k, fetch = 1024, 20000, 16, 10, 100
    x, centres = clustered(n, dim, seed=310)
    q = queries_for(centres, x, nq, seed=311)
    want = pipelines.two_stage(x, q, k, fetch)
    qd = torch.from_numpy(q).cuda()
    coarse, fine = _shards(x, dim, "b1", g), _shards(x, dim, "f16", g)
    if g == 1:
        got = TwoStageSearcher(coarse[0], fine[0]).search(qd, k, fetch)
    else:
        # the row-sharded data path without a process group: the same calls, the two
        # collectives replaced by their definition (gather lists / elementwise max)
        from compressed_rag_suite_b200.index import merge_topk
        loc = [c.search(qd, fetch) for c in coarse]
        ids, _, _ = merge_topk(torch.stack([l[0] for l in loc]), torch.stack([l[1] for l in loc]), fetch)
        fs = torch.stack([f.score_rows(qd, ids) for f in fine]).max(dim=0).values
        got = select_topk(ids, fs, k)
    assert np.array_equal(got[2].cpu().numpy(), want[2])
    assert np.array_equal(got[0].cpu().numpy().view(np.uint32), want[0])
    assert np.array_equal(got[1].cpu().numpy().view(np.uint32), want[1].view(np.uint32))


def test_two_stage_threshold_on_fine_score():
    dim, n = 384, 8000
    x, centres = clustered(n, dim, seed=312)
    q = queries_for(centres, x, 8, seed=313)
    want = pipelines.two_stage(x, q, 10, 64, min_similarity=0.4)
    c, f = ShardIndex(dim, dtype="b1"), ShardIndex(dim, dtype="f16")
    c.add(x); f.add(x)
    got = TwoStageSearcher(c, f).search(torch.from_numpy(q).cuda(), 10, 64, min_similarity=0.4)
    assert np.array_equal(got[2].cpu().numpy(), want[2])
    assert np.array_equal(got[0].cpu().numpy().view(np.uint32), want[0])
    assert (want[2] < 10).any(), "the threshold must bite for this test to mean something"


@pytest.mark.parametrize("store,g", [("i8", 1), ("i8", 4), ("f16", 2)])
def test_top100_then_mmr_to_10(store, g):
    """config 4 in small: int8 top-100 -> reference score transform -> greedy MMR -> first 10."""
    dim, n, nq, k, fetch, pen = 384, 30000, 10, 10, 100, 0.1
    x, centres = clustered(n, dim, seed=320, n_clusters=16)
    q = queries_for(centres, x, nq, seed=321)
    want = pipelines.search_then_mmr(x, q, store, k, fetch, pen)
    qd = torch.from_numpy(q).cuda()
    shards = _shards(x, dim, store, g)
    if g == 1:
        ids, sims, rel, cnt = ShardedMMRSearcher(shards[0]).search_mmr(qd, k, fetch, pen)
    else:
        from compressed_rag_suite_b200.index import merge_topk
        loc = [s.search(qd, fetch) for s in shards]
        cid, raw, ccnt = merge_topk(torch.stack([l[0] for l in loc]), torch.stack([l[1] for l in loc]), fetch)
        vecs = None
        for s in shards:                                      # every owner fills in its rows
            vecs = s.fetch_rows_device(cid, out=vecs)
        s0 = shards[0]
        sm = similarity_of(s0, raw)
        r = reference_relevance(sm)
        order = s0.mmr_device(vecs, r, 1.0 - pen, k).to(torch.int64)
        ids, sims, rel = torch.gather(cid, 1, order), torch.gather(sm, 1, order), torch.gather(r, 1, order)
        cnt = torch.full((nq,), k)
    ids, sims, rel = ids.cpu().numpy(), sims.cpu().numpy(), rel.cpu().numpy()
    for i in range(nq):
        w_ids, w_sims, w_rel = want[i]
        assert int(cnt[i]) == len(w_ids)
        assert ids[i, :len(w_ids)].tolist() == w_ids, i
        assert sims[i, :len(w_ids)].tolist() == w_sims
        assert rel[i, :len(w_ids)].tolist() == w_rel           # float64, exactly the Python arithmetic


def test_mmr_pipeline_with_threshold_and_short_lists():
    dim, n = 384, 5000
    x, centres = clustered(n, dim, seed=330, n_clusters=200)
    q = queries_for(centres, x, 12, seed=331)
    want = pipelines.search_then_mmr(x, q, "f16", 10, 100, 0.3, min_similarity=0.33)
    ix = ShardIndex(dim, dtype="f16")
    ix.add(x)
    ids, sims, rel, cnt = ShardedMMRSearcher(ix).search_mmr(torch.from_numpy(q).cuda(), 10, 100, 0.3, min_similarity=0.33)
    ids = ids.cpu().numpy()
    lens = []
    for i in range(12):
        w_ids = want[i][0]
        lens.append(len(w_ids))
        assert int(cnt[i]) == len(w_ids)
        assert ids[i, :len(w_ids)].tolist() == w_ids
        assert (ids[i, len(w_ids):] == -1).all()
    assert min(lens) < 10, "some query must have fewer than k survivors"


@pytest.mark.parametrize("store", ["f16", "bf16", "i8", "b1"])
def test_score_vectors_equals_score_of_stored_rows(store):
    """Caller-supplied candidate vectors go through the same encoder -> same canonical scores."""
    x, centres = clustered(3000, 384, seed=340)
    q = queries_for(centres, x, 6, seed=341)
    ix = ShardIndex(384, dtype=store)
    ix.add(x)
    ids, raw, cnt = ix.search(q, 40)
    rows = x[ids.astype(np.int64)]                                   # [nq, 40, 384] original fp32 vectors
    empty = ShardIndex(384, dtype=store)                             # carries dim / dtype / metric only
    got = empty.score_vectors(q, rows)
    assert np.array_equal(got.view(np.uint32), raw.view(np.uint32))
    gd = empty.score_vectors(torch.from_numpy(q).cuda(), torch.from_numpy(rows).cuda())
    assert np.array_equal(gd.cpu().numpy().view(np.uint32), raw.view(np.uint32))


def test_two_stage_with_rematerialised_rows_equals_resident_fine_index():
    dim, n = 1024, 12000
    x, centres = clustered(n, dim, seed=350)
    q = torch.from_numpy(queries_for(centres, x, 8, seed=351)).cuda()
    c, f = ShardIndex(dim, dtype="b1"), ShardIndex(dim, dtype="f16")
    c.add(x); f.add(x)
    want = TwoStageSearcher(c, f).search(q, 10, 100)
    xd = torch.from_numpy(x).cuda()
    src = lambda ids: xd[ids.clamp(min=0).to(torch.int64)]          # noqa: E731  (pad ids -> any row, masked later)
    got = TwoStageSearcher(c, ShardIndex(dim, dtype="f16"), row_source=src).search(q, 10, 100)
    for u, v in zip(got, want):
        assert torch.equal(u.view(torch.int32) if u.dtype == torch.float32 else u,
                           v.view(torch.int32) if v.dtype == torch.float32 else v)


def test_captured_search_graph_replays_identically():
    x, centres = clustered(50000, 384, seed=360)
    q = queries_for(centres, x, 6, seed=361)
    for store, nq in (("f16", 1), ("i8", 1), ("f16", 6)):
        ix = ShardIndex(384, dtype=store)
        ix.add(x)
        gs = ix.capture_search(nq, 10, 0.2)
        for rep in range(3):
            qq = torch.from_numpy(q[rep:rep + nq] if nq == 1 else np.roll(q, rep, axis=0)).cuda()
            gs.queries.copy_(qq)
            got = [t.clone() for t in gs.replay()]
            want = ix.search(qq, 10, 0.2)
            torch.cuda.synchronize()
            assert all(torch.equal(u, v) for u, v in zip(got, want))
        ix.close()
