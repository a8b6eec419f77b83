"""CPU tier: the oracle against itself, against independent brute force, and against the
committed golden vectors (which were produced by the reference's own code)."""
import json
import os

import numpy as np
import pytest

from oracle import encode, postprocess, search
from helpers import clustered, queries_for


def test_normalise_is_sequential_fp64():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, 37)).astype(np.float32)
    y = encode.normalise_rows(x)
    for r in range(5):
        n2 = 0.0
        for j in range(37):
            n2 += float(x[r, j]) * float(x[r, j])
        assert np.array_equal(y[r], x[r].astype(np.float64) / np.sqrt(n2))
    assert np.all(encode.normalise_rows(np.zeros((2, 8), np.float32)) == 0.0)


def test_bf16_rounding_is_single_rne():
    from fractions import Fraction
    rng = np.random.default_rng(1)
    y = rng.standard_normal(4000) * np.exp(rng.uniform(-8, 2, 4000))
    bits = encode.f64_to_bf16_bits(y)
    for v, b in zip(y[:1500], bits[:1500]):
        cands = [int(b) - 1, int(b), int(b) + 1]
        errs = [abs(Fraction(float(np.array(np.uint32(c << 16)).view(np.float32))) - Fraction(float(v))) for c in cands]
        assert errs[1] == min(errs)
        if errs[1] == errs[0] or errs[1] == errs[2]:
            assert b % 2 == 0            # ties to even


@pytest.mark.parametrize("store", ["f16", "bf16", "i8", "b1"])
def test_padding_and_layout(store):
    x = np.random.default_rng(2).standard_normal((3, 100)).astype(np.float32)
    c = encode.encode_rows(x, store)
    dp = encode.padded_dim(100, store)
    assert dp >= 100
    if store == "b1":
        assert c.shape == (3, dp // 32) and c.dtype == np.uint32
        y = encode.normalise_rows(x)
        assert ((c[0, 0] >> 5) & 1) == int(y[0, 5] > 0)
        assert np.all(c[:, 4:] == 0)     # pad bits are zero
    else:
        assert c.shape == (3, dp)
        assert np.all(encode.decode_rows(c, store)[:, 100:] == 0)


def test_canonical_vs_shortlist_vs_f32_bruteforce():
    x, centres = clustered(6000, 384, seed=3)
    q = queries_for(centres, x, 6, seed=4)
    codes = encode.encode_rows(x, "f16")
    qc = search.encode_queries(q, "f16")
    a = search.search(codes, qc, "f16", 384, 10)
    b = search.search_f16_shortlist(codes, qc, 10)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
    # plain fp32 brute force on the original embeddings: same ids unless a near-tie, scores within 1e-3
    bi, bs = search.bruteforce_f32(x, q, 10)
    assert np.abs(bs - a[1]).max() < 1e-3
    agree = np.mean(bi == a[0].astype(np.int64))
    assert agree > 0.9


def test_sklearn_cross_check():
    sk = pytest.importorskip("sklearn.neighbors")
    x, centres = clustered(2000, 384, seed=5, dup_frac=0.0)
    q = queries_for(centres, x, 3, seed=6)
    nn = sk.NearestNeighbors(n_neighbors=5, algorithm="brute", metric="cosine").fit(x)
    dist, idx = nn.kneighbors(q)
    ids, raw, _ = search.search(encode.encode_rows(x, "f16"), search.encode_queries(q, "f16"), "f16", 384, 5)
    assert np.abs((1.0 - dist) - raw).max() < 1e-3
    assert np.mean(idx == ids.astype(np.int64)) > 0.8


def test_ties_go_to_lowest_id_and_threshold():
    x, _ = clustered(500, 128, seed=7, dup_frac=0.0)
    x[400] = x[20]
    x[30] = x[20]
    codes = encode.encode_rows(x, "f16")
    qc = search.encode_queries(x[20:21], "f16")
    ids, raw, cnt = search.search(codes, qc, "f16", 128, 3)
    assert list(ids[0]) == [20, 30, 400] and raw[0, 0] == raw[0, 1] == raw[0, 2]
    ids, raw, cnt = search.search(codes, qc, "f16", 128, 10, min_similarity=0.99)
    assert cnt[0] == 3 and ids[0, 3] == 0xFFFFFFFF and raw[0, 3] == -np.inf


def test_int_stores_are_exact():
    x, centres = clustered(1500, 384, seed=8)
    q = queries_for(centres, x, 4, seed=9)
    for store in ("i8", "b1"):
        codes = encode.encode_rows(x, store)
        qc = search.encode_queries(q, store)
        ids, raw, cnt = search.search(codes, qc, store, 384, 20)
        assert raw.dtype == np.int32
        # independent recomputation through the decoded values
        dec = encode.decode_rows(codes, store)[:, :384]
        dq = encode.decode_rows(qc, store)[:, :384]
        full = (dq @ dec.T).astype(np.int64)
        for i in range(4):
            order = np.lexsort((np.arange(1500), -full[i]))[:20]
            assert np.array_equal(order, ids[i].astype(np.int64))
            assert np.array_equal(full[i][order], raw[i].astype(np.int64))


def test_merge_topk_equals_unsharded():
    x, centres = clustered(3000, 384, seed=10)
    q = queries_for(centres, x, 5, seed=11)
    codes = encode.encode_rows(x, "f16")
    qc = search.encode_queries(q, "f16")
    whole = search.search(codes, qc, "f16", 384, 10)
    parts_i, parts_s = [], []
    for g in range(4):
        lo, hi = g * 750, (g + 1) * 750
        i, s, _ = search.search(codes[lo:hi], qc, "f16", 384, 10, row_base=lo)
        parts_i.append(i)
        parts_s.append(s)
    m = search.merge_topk(np.stack(parts_i), np.stack(parts_s), 10)
    for u, v in zip(whole, m):
        assert np.array_equal(u, v)


# ---------------------------------------------------------------- golden vectors (made by the reference)
def test_transform_golden(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "transform_golden.json")))
    for c in g["distance_to_similarity"]:
        assert postprocess.distance_to_similarity(c["distance"], c["metric"]) == c["score"]
    for c in g["rerank"]:
        got = postprocess.rerank(c["query"], [dict(x) for x in c["chunks"]], c["top_k"])
        assert [x["chunk_id"] for x in got] == c["order"]
        assert [x["rerank_score"] for x in got] == c["rerank_scores"]


def test_mmr_dyadic_golden(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "mmr_dyadic_golden.json")))
    assert len(cases) >= 30
    for c in cases:
        v = np.asarray(c["vectors_x64"], dtype=np.float64) / 64.0
        order = postprocess.mmr_order(c["relevance"], postprocess.pairwise_sims_f32(v), 1.0 - c["penalty"])
        assert order == c["order"]
        # prefix property used by the fetch_k extension
        assert postprocess.mmr_order(c["relevance"], postprocess.pairwise_sims_f32(v), 1.0 - c["penalty"],
                                     k_out=max(1, c["m"] // 2)) == c["order"][:max(1, c["m"] // 2)]


def test_threshold_pushdown_is_conservative():
    for t in (0.01, 0.3, 0.5, 0.75, 0.9, 0.999, 1.0):
        lo = postprocess.min_cosine_for_threshold(t)
        for cos in np.linspace(-1, 1, 4001):
            if postprocess.distance_to_similarity(1.0 - cos) >= t:
                assert cos >= lo
    assert postprocess.min_cosine_for_threshold(0.0) == -np.inf
    assert postprocess.min_cosine_for_threshold(1.5) == np.inf


def test_pipelines_reduce_to_their_parts():
    """oracle/pipelines.py: two-stage with fetch = n equals a plain fine search; MMR with zero
    penalty keeps the relevance order; the first MMR pick is always the best hit."""
    from oracle import pipelines
    rng = np.random.default_rng(3)
    x = rng.standard_normal((300, 64)).astype(np.float32)
    q = rng.standard_normal((4, 64)).astype(np.float32)
    ids, raw, cnt = pipelines.two_stage(x, q, 7, 300)
    want = search.search(encode.encode_rows(x, "f16"), search.encode_queries(q, "f16"), "f16", 64, 7)
    assert np.array_equal(ids, want[0]) and np.array_equal(raw, want[1]) and np.array_equal(cnt, want[2])
    plain = search.search(encode.encode_rows(x, "i8"), search.encode_queries(q, "i8"), "i8", 64, 20)
    for pen in (0.0, 0.4):
        out = pipelines.search_then_mmr(x, q, "i8", 5, 20, pen)
        for i, (oi, sims, rel) in enumerate(out):
            assert oi[0] == int(plain[0][i, 0]) and len(oi) == 5 and len(set(oi)) == 5
            if pen == 0.0:
                assert oi == [int(v) for v in plain[0][i, :5]]
