"""CPU tier: property tests (hypothesis) of the oracle's order / threshold / merge rules — the
invariants the GPU tier then demands of the CUDA path at sizes the oracle cannot brute-force."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import encode, search

_store = st.sampled_from(["f16", "bf16", "i8", "b1"])


def _data(seed, n, dim, dup):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, dim)).astype(np.float32)
    if dup and n > 3:
        x[n // 2] = x[0]                      # exact duplicate -> score tie
        x[n - 1] = x[1]
    q = rng.standard_normal((3, dim)).astype(np.float32)
    q[0] = x[0]
    return x, q


@settings(max_examples=40, deadline=None)
@given(_store, st.integers(0, 10_000), st.integers(1, 60), st.sampled_from([8, 64, 100]), st.integers(1, 12), st.booleans())
def test_topk_is_sorted_prefix_and_ties_go_to_lowest_id(store, seed, n, dim, k, dup):
    x, q = _data(seed, n, dim, dup)
    codes, qc = encode.encode_rows(x, store), search.encode_queries(q, store)
    ids, raw, cnt = search.search(codes, qc, store, dim, k)
    big = search.search(codes, qc, store, dim, n)                       # the full ranking
    for i in range(len(q)):
        c = cnt[i]
        assert c == min(k, n)
        assert np.array_equal(ids[i, :c], big[0][i, :c]) and np.array_equal(raw[i, :c], big[1][i, :c])
        assert (ids[i, c:] == 0xFFFFFFFF).all()
        s = raw[i, :c].astype(np.float64)
        assert (np.diff(s) <= 0).all()
        same = np.diff(s) == 0
        assert (np.diff(ids[i, :c].astype(np.int64))[same] > 0).all()
        full = search.raw_scores(codes, qc[i], store, dim)              # every listed score is that row's score
        assert np.array_equal(full[ids[i, :c].astype(np.int64)], raw[i, :c])
        if c < n:                                                       # nothing outside beats the k-th
            rest = np.setdiff1d(np.arange(n), ids[i, :c].astype(np.int64))
            kth = (float(s[-1]), -int(ids[i, c - 1]))
            assert all((float(full[r]), -int(r)) < kth for r in rest)


@settings(max_examples=30, deadline=None)
@given(_store, st.integers(0, 10_000), st.integers(2, 50), st.integers(1, 10), st.floats(-0.5, 0.9))
def test_threshold_keeps_exactly_the_prefix_that_passes(store, seed, n, k, thr):
    x, q = _data(seed, n, 32, True)
    codes, qc = encode.encode_rows(x, store), search.encode_queries(q, store)
    plain = search.search(codes, qc, store, 32, n)
    got = search.search(codes, qc, store, 32, k, min_similarity=thr)
    for i in range(len(q)):
        sims = search.similarity_from_raw(plain[1][i], store, 32)
        keep = plain[0][i][sims >= np.float32(thr)][:k]
        assert got[2][i] == len(keep) and np.array_equal(got[0][i, :len(keep)], keep)


@settings(max_examples=30, deadline=None)
@given(st.sampled_from(["f16", "i8", "b1"]), st.integers(0, 10_000), st.integers(4, 80), st.integers(1, 6), st.integers(1, 8))
def test_shard_and_merge_equals_unsharded_for_any_partition(store, seed, n, g, k):
    x, q = _data(seed, n, 48, True)
    codes, qc = encode.encode_rows(x, store), search.encode_queries(q, store)
    want = search.search(codes, qc, store, 48, k)
    cuts = sorted(np.random.default_rng(seed + 1).integers(0, n + 1, g - 1).tolist())
    bounds = [0] + cuts + [n]                                           # ragged, possibly empty shards
    parts = [search.search(codes[lo:hi], qc, store, 48, k, row_base=lo) for lo, hi in zip(bounds, bounds[1:])]
    ids = np.stack([p[0] for p in parts]); raw = np.stack([p[1] for p in parts])
    m = search.merge_topk(ids, raw, k)
    assert np.array_equal(m[0], want[0]) and np.array_equal(m[1], want[1]) and np.array_equal(m[2], want[2])
