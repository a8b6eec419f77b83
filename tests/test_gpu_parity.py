"""GPU tier (-m gpu): the CUDA path, called through the C ABI, against the oracle.

Bit-exact for ids, raw scores (float stores: fl32 of the fp64-sequential dot), counts,
stored codes and MMR order.  The only tolerance in this file is the 1e-3 the north star
allows between the canonical fp16 score and the plain fp32 brute force on the original
fp32 embeddings."""
import json
import os

import numpy as np
import pytest

from helpers import clustered, queries_for
from oracle import encode, postprocess, search

pytestmark = pytest.mark.gpu

from compressed_rag_suite_b200.index import ShardIndex  # noqa: E402


def as_codes(raw_bytes, store, dp):
    dt = {"f16": np.float16, "bf16": np.uint16, "i8": np.int8, "b1": np.uint32}[store]
    return raw_bytes.view(dt).reshape(raw_bytes.shape[0], -1)


# ---------------------------------------------------------------- K0 ingest
@pytest.mark.parametrize("store", ["f16", "bf16", "i8", "b1"])
@pytest.mark.parametrize("dim", [384, 100, 64, 1000])
@pytest.mark.parametrize("n", [517, 6001])          # query-sized (shared-memory kernel) and bulk (streaming kernel) adds
def test_ingest_bit_exact(store, dim, n):
    rng = np.random.default_rng(dim)               # ragged n: not a multiple of the 32-row warp tile
    x = (rng.standard_normal((n, dim)) * np.exp(rng.uniform(-3, 3, (n, 1)))).astype(np.float32)
    x[7] = 0.0                                     # zero row stays zero
    x[8, :] = 0.0
    x[8, 3] = -2.5
    ix = ShardIndex(dim, dtype=store)
    ix.add(x[:200])
    ix.add(x[200:])                                # two adds, second one unaligned
    assert len(ix) == n
    got = as_codes(ix.fetch_rows(np.arange(n)), store, ix.dim_padded)
    want = encode.encode_rows(x, store)
    assert ix.dim_padded == encode.padded_dim(dim, store)
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8))


def test_ingest_ip_metric_keeps_values():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((100, 128)).astype(np.float32) * 3
    ix = ShardIndex(128, dtype="f16", metric="ip")
    ix.add(x)
    got = as_codes(ix.fetch_rows(np.arange(100)), "f16", 128)
    assert np.array_equal(got, encode.encode_rows(x, "f16", "ip"))


def test_ingest_from_device_tensor():
    import torch
    rng = np.random.default_rng(6)
    x = rng.standard_normal((1000, 384)).astype(np.float32)
    ix = ShardIndex(384)
    ix.add(torch.from_numpy(x).cuda())
    got = as_codes(ix.fetch_rows(np.arange(1000)), "f16", 384)
    assert np.array_equal(got, encode.encode_rows(x, "f16"))


# ---------------------------------------------------------------- K1/K2/K3 + finalize
def check_search(ix, x, q, store, k, min_similarity=-np.inf, row_base=0):
    dim = x.shape[1]
    codes = encode.encode_rows(x, store, ix.metric)
    qc = search.encode_queries(q, store, ix.metric)
    want = search.search(codes, qc, store, dim, k, min_similarity, row_base=row_base)
    got = ix.search(q, k, min_similarity)
    assert np.array_equal(got[2], want[2]), "counts differ"
    assert np.array_equal(got[0], want[0]), "ids differ"
    assert np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32)), "raw scores differ"
    return got


@pytest.mark.parametrize("store", ["f16", "bf16", "i8", "b1"])
@pytest.mark.parametrize("n,dim,k", [(20000, 384, 10), (5003, 384, 3), (777, 128, 16), (40000, 256, 100),
                                     (3000, 1024, 10), (9, 384, 5), (150, 384, 32)])
def test_search_bit_exact(store, n, dim, k):
    if store in ("f16", "bf16") and k > 16 and k > n:
        pytest.skip("covered elsewhere")
    x, centres = clustered(n, dim, seed=n + dim)
    q = queries_for(centres, x, 5, seed=k)
    ix = ShardIndex(dim, dtype=store)
    ix.add(x)
    kk = min(k, n)
    ix.set_option("force_path", 0)                      # stream scans (K1/K2/K3)
    check_search(ix, x, q, store, kk)
    assert ix.last_stats()["path"] == 0
    check_search(ix, x, q[:1], store, kk)
    ix.set_option("force_path", -1)                     # default dispatch: batches of >= 2 take the tensor cores
    check_search(ix, x, q, store, kk)                   # where the shape allows it


@pytest.mark.parametrize("store", ["f16", "i8", "b1"])
def test_search_threshold_and_count_lt_k(store):
    x, centres = clustered(30000, 384, seed=77)
    q = queries_for(centres, x, 6, seed=78)
    ix = ShardIndex(384, dtype=store)
    ix.add(x)
    for thr in (-0.1832, 0.293, 0.45, 0.9999, 1.5):
        got = check_search(ix, x, q, store, 10, min_similarity=thr)
    assert got[2].max() == 0                       # nothing reaches 1.5


def test_search_ties_lowest_id_and_duplicates():
    x, centres = clustered(8000, 384, seed=90, dup_frac=0.0)
    for dst in (11, 4000, 7999, 512, 513, 514):
        x[dst] = x[10]                             # seven identical rows
    q = np.stack([x[10], x[10] * 3.0, x[4000] + 1e-4 * x[1]])
    ix = ShardIndex(384)
    ix.add(x)
    ids, raw, cnt = check_search(ix, x, q, "f16", 10)
    assert list(ids[0][:7]) == [10, 11, 512, 513, 514, 4000, 7999]
    assert len(set(raw[0][:7].tolist())) == 1


def test_many_duplicates_take_the_exact_fallback():
    """More equal rows than the candidate list can hold: certification must fail and the
    exhaustive fp64 pass must still give the lowest ids."""
    x, centres = clustered(6000, 384, seed=91, dup_frac=0.0)
    x[100:180] = x[5]                              # 81 identical rows > M = 32
    q = np.stack([x[5], x[3000]])
    ix = ShardIndex(384)
    ix.add(x)
    ids, raw, cnt = check_search(ix, x, q, "f16", 10)
    assert list(ids[0]) == [5] + list(range(100, 109))
    assert ix.last_stats()["uncertified_total"] >= 1


def test_forced_exact_pass_equals_fast_pass():
    x, centres = clustered(12000, 384, seed=92)
    q = queries_for(centres, x, 4, seed=93)
    ix = ShardIndex(384)
    ix.add(x)
    fast = ix.search(q, 10)
    ix.set_option("force_exact", 1)
    slow = check_search(ix, x, q, "f16", 10)
    for a, b in zip(fast, slow):
        assert np.array_equal(a, b)


def test_row_base_and_device_buffers():
    import torch
    x, centres = clustered(9000, 384, seed=94)
    q = queries_for(centres, x, 7, seed=95)
    ix = ShardIndex(384, row_base=1_000_000)
    ix.add(x)
    host = check_search(ix, x, q, "f16", 10, row_base=1_000_000)
    ids, sc, cnt = ix.search(torch.from_numpy(q).cuda(), 10)
    torch.cuda.synchronize()
    assert np.array_equal(ids.cpu().numpy().view(np.uint32), host[0])
    assert np.array_equal(sc.cpu().numpy(), host[1])
    assert np.array_equal(cnt.cpu().numpy(), host[2])


def test_empty_index_and_bad_arguments():
    ix = ShardIndex(384)
    ids, raw, cnt = ix.search(np.zeros((2, 384), np.float32), 5)
    assert cnt.tolist() == [0, 0] and np.all(ids == 0xFFFFFFFF) and np.all(raw == -np.inf)
    with pytest.raises(ValueError):
        ix.search(np.zeros((1, 100), np.float32), 5)
    with pytest.raises(ValueError):
        ix.add(np.zeros((3, 7), np.float32))
    with pytest.raises(ValueError):
        ix.search(np.zeros((1, 384), np.float32), 0)
    with pytest.raises(ValueError):
        ShardIndex(384, dtype="f32")


def test_fp16_scores_within_1e3_of_fp32_bruteforce():
    """North-star tolerance: canonical fp16 scores vs plain fp32 brute force on the ORIGINAL
    fp32 embeddings: |delta| <= 1e-3 absolute."""
    x, centres = clustered(50000, 384, seed=96, dup_frac=0.0)
    q = queries_for(centres, x, 16, seed=97)
    ix = ShardIndex(384)
    ix.add(x)
    ids, raw, cnt = ix.search(q, 10)
    bi, bs = search.bruteforce_f32(x, q, 10)
    assert np.abs(raw - bs).max() <= 1e-3
    print(f"ids equal to fp32 brute force on the original embeddings: {np.mean(ids.astype(np.int64) == bi):.4f}")


@pytest.mark.parametrize("store", ["f16", "bf16"])
@pytest.mark.parametrize("nq", [1, 64])                 # stream scan / tensor-core path
def test_ids_equal_fp32_bruteforce_on_stored_rows_outside_near_ties(store, nq):
    """"Bit-exact with an exhaustive fp32 brute-force pass on the same embeddings": the embeddings the index
    holds are the stored (fp16 / bf16-rounded) rows.  A plain fp32 sgemm over the DECODED stored rows must give
    the same ids, except where two rows' fp32 scores differ by no more than the summation-order noise of an
    fp32 dot (2 * eps, eps = Dp * 2^-24 * |q| * |c|): there fp32 brute force itself has no defined order.
    Scores must agree within eps.  The disagreements are counted and printed."""
    n, dim, k = 60000, 384, 10
    x, centres = clustered(n, dim, seed=196, dup_frac=0.01)
    q = queries_for(centres, x, nq, seed=197)
    ix = ShardIndex(dim, dtype=store)
    ix.add(x)
    ids, raw, cnt = ix.search(q, k)
    dec = encode.decode_rows(encode.encode_rows(x, store), store).astype(np.float32)
    qd = encode.decode_rows(search.encode_queries(q, store), store).astype(np.float32)
    s32 = qd @ dec.T                                                        # the plain fp32 pass
    eps = dec.shape[1] * 2.0 ** -24 * 1.01 * 1.01
    differ = 0
    for i in range(nq):
        order = np.lexsort((np.arange(n), -s32[i]))[:k]
        assert np.abs(raw[i] - s32[i, ids[i].astype(np.int64)]).max() <= eps, "score differs from fp32 brute force by more than eps"
        kth = s32[i, order[-1]]
        for j in range(k):
            a, b = int(ids[i, j]), int(order[j])
            if a == b:
                continue
            differ += 1
            # a swap is only legitimate inside a near-tie band of the fp32 scores
            assert abs(float(s32[i, a]) - float(s32[i, b])) <= 2 * eps, (i, j, a, b, s32[i, a], s32[i, b])
            assert s32[i, a] >= kth - 2 * eps
    print(f"{store} nq={nq}: {differ} of {nq * k} positions differ from fp32 brute force, all inside the 2*eps near-tie band")


# ---------------------------------------------------------------- inner-product space (rag/retrieval.py:84-87 knows 'ip')
@pytest.mark.parametrize("store", ["f16", "bf16"])
@pytest.mark.parametrize("path", [0, 1])
def test_ip_search_bit_exact(store, path):
    """metric='ip': rows are stored as given (no normalisation), score = dot of the stored values.  Rows and
    queries of widely varying norm, thresholds in the dot domain, both kernels paths, single query and batch."""
    n, dim = 30000, 384
    rng = np.random.default_rng(300 + path)
    x, centres = clustered(n, dim, seed=301)
    x = (x * np.exp(rng.uniform(-2.0, 2.0, (n, 1)))).astype(np.float32)      # norms 0.13 .. 7.4
    x[17] = 0.0
    q = queries_for(centres, x, 20, seed=302)
    q = (q * np.exp(rng.uniform(-1.0, 1.0, (20, 1)))).astype(np.float32)
    ix = ShardIndex(dim, dtype=store, metric="ip")
    ix.add(x[:12345])
    ix.add(x[12345:])
    ix.set_option("force_path", path)
    for thr in (-np.inf, 0.5, 2.0, 1e9):
        got = check_search(ix, x, q, store, 10, min_similarity=thr)
        assert ix.last_stats()["path"] == path
    assert got[2].max() == 0
    check_search(ix, x, q[:1], store, 10)
    check_search(ix, x, q[:3], store, 100 if path == 0 else 50, min_similarity=-0.25)
    # the plain fp32 dot on the original rows agrees within the north star's tolerance scaled by the norms
    ids, raw, cnt = ix.search(q, 10)
    bi, bs = search.bruteforce_f32(x, q, 10, metric="ip")
    scale = np.linalg.norm(q, axis=1, keepdims=True) * np.linalg.norm(x, axis=1).max()
    assert (np.abs(raw - bs) <= 1e-3 * scale).all()


def test_ip_duplicates_force_the_exact_fallback_and_device_buffers():
    import torch
    n, dim = 9000, 384
    x, centres = clustered(n, dim, seed=310, dup_frac=0.0)
    x = (x * np.linspace(0.5, 4.0, n, dtype=np.float32)[:, None]).astype(np.float32)
    x[2000:2081] = x[8999]                            # 82 copies of the largest-norm row: more than a list holds
    q = np.concatenate([x[8999][None] * 0.3, queries_for(centres, x, 11, seed=311)]).astype(np.float32)
    for path in (0, 1):
        ix = ShardIndex(dim, dtype="f16", metric="ip")
        ix.add(x)
        ix.set_option("force_path", path)
        ids, raw, cnt = check_search(ix, x, q, "f16", 10)
        assert list(ids[0]) == list(range(2000, 2010))
        assert ix.last_stats()["uncertified_total"] >= 1
        d = ix.search(torch.from_numpy(q).cuda(), 10)
        torch.cuda.synchronize()
        assert np.array_equal(d[0].cpu().numpy().view(np.uint32), ids) and np.array_equal(d[1].cpu().numpy(), raw)
        ix.close()


def test_ip_with_integer_stores_is_rejected():
    """int8 / 1-bit codes are defined on unit rows (oracle/encode.py); un-normalised rows would saturate."""
    for store in ("i8", "b1"):
        with pytest.raises(ValueError):
            ShardIndex(384, dtype=store, metric="ip")


def test_save_load_roundtrip(tmp_path):
    x, centres = clustered(3000, 384, seed=98)
    q = queries_for(centres, x, 3, seed=99)
    for store in ("f16", "i8", "b1"):
        ix = ShardIndex(384, dtype=store)
        ix.add(x)
        want = ix.search(q, 8)
        p = str(tmp_path / f"{store}.crs")
        ix.save(p)
        jx = ShardIndex.load(p)
        assert (jx.dim, jx.dtype, len(jx)) == (384, store, 3000)
        got = jx.search(q, 8)
        for a, b in zip(want, got):
            assert np.array_equal(a, b)


# ---------------------------------------------------------------- K7 merge
@pytest.mark.parametrize("store", ["f16", "i8"])
@pytest.mark.parametrize("g,k", [(2, 10), (8, 10), (4, 100), (3, 1)])
def test_shard_merge_equals_single_index(store, g, k):
    import torch
    from compressed_rag_suite_b200.index import merge_topk
    n = 24000
    x, centres = clustered(n, 384, seed=100 + g)
    q = queries_for(centres, x, 9, seed=101)
    whole = ShardIndex(384, dtype=store)
    whole.add(x)
    want = whole.search(q, k)
    per = (n + g - 1) // g
    qd = torch.from_numpy(q).cuda()
    ids_l, sc_l = [], []
    for r in range(g):
        sh = ShardIndex(384, dtype=store, row_base=r * per)
        sh.add(x[r * per:(r + 1) * per])
        i, s, c = sh.search(qd, k)
        ids_l.append(i)
        sc_l.append(s)
    mi, ms, mc = merge_topk(torch.stack(ids_l), torch.stack(sc_l), k)
    torch.cuda.synchronize()
    assert np.array_equal(mi.cpu().numpy().view(np.uint32), want[0])
    assert np.array_equal(ms.cpu().numpy(), want[1])
    assert np.array_equal(mc.cpu().numpy(), want[2])
    # and against the oracle's merge of the oracle's shard results
    om = search.merge_topk(torch.stack(ids_l).cpu().numpy().view(np.uint32), torch.stack(sc_l).cpu().numpy(), k)
    assert np.array_equal(om[0], want[0])


# ---------------------------------------------------------------- K6 MMR
def test_mmr_dyadic_golden_bit_exact(golden_dir):
    """Expected orders were produced by the reference's own _apply_diversity."""
    cases = json.load(open(os.path.join(golden_dir, "mmr_dyadic_golden.json")))
    for c in cases:
        v = np.asarray(c["vectors_x64"], dtype=np.float32) / 64.0
        ix = ShardIndex(c["dim"], dtype="f16", metric="ip")       # ip: values stored as given
        ix.add(v)
        codes = ix.fetch_rows(np.arange(c["m"]))
        order = ix.mmr(codes, np.asarray(c["relevance"]), 1.0 - c["penalty"])
        assert order[0].tolist() == c["order"], (c["m"], c["mode"], c["penalty"])
        half = max(1, c["m"] // 2)
        assert ix.mmr(codes, np.asarray(c["relevance"]), 1.0 - c["penalty"], k_out=half)[0].tolist() == c["order"][:half]


@pytest.mark.parametrize("store", ["f16", "bf16", "i8", "b1"])
def test_mmr_matches_oracle_on_stored_vectors(store):
    x, centres = clustered(400, 384, seed=120, n_clusters=5)
    ix = ShardIndex(384, dtype=store)
    ix.add(x)
    rng = np.random.default_rng(121)
    for m, lam in [(3, 0.9), (20, 0.9), (100, 0.9), (64, 0.5), (7, 0.0)]:
        rows = rng.choice(400, size=m, replace=False)
        rel = np.sort(rng.uniform(0.3, 0.95, m))[::-1].copy()
        codes = ix.fetch_rows(rows)
        dec = encode.decode_rows(as_codes(codes, store, ix.dim_padded), store)
        if store == "b1":
            dec = dec[:, :384]
        want = postprocess.mmr_order([float(r) for r in rel], postprocess.pairwise_sims_f32(dec), lam,
                                     k_out=min(m, 10))
        got = ix.mmr(codes, rel, lam, k_out=min(m, 10))[0].tolist()
        assert got == want, (store, m, lam)


# ---------------------------------------------------------------- K4 tcgen05 batched path
@pytest.mark.parametrize("store", ["f16", "bf16", "i8"])
@pytest.mark.parametrize("n,dim,nq,k", [(30000, 384, 64, 10), (5003, 384, 130, 10), (777, 128, 8, 3),
                                        (20000, 256, 300, 20), (300, 64, 16, 10), (30000, 320, 140, 5),
                                        (255, 384, 9, 10), (257, 192, 128, 24)])
def test_gemm_path_bit_exact(store, n, dim, nq, k):
    x, centres = clustered(n, dim, seed=n + nq)
    q = queries_for(centres, x, nq, seed=k + 1)
    ix = ShardIndex(dim, dtype=store)
    ix.add(x)
    check_search(ix, x, q, store, k)
    assert ix.last_stats()["path"] == 1, "batched float search must take the tcgen05 path"


@pytest.mark.parametrize("store", ["f16", "bf16", "i8"])
@pytest.mark.parametrize("n,nq,k", [(60000, 300, 10), (150000, 1024, 10), (40000, 40, 100), (9000, 129, 24)])
def test_gemm_floor_sharing_keeps_results_exact(store, n, nq, k):
    """The slices of a query share their k-th best score while the contraction runs (on by default); the
    result must not depend on it, with duplicates, a threshold, a filter, and together with the sample pass."""
    dim = 384
    x, centres = clustered(n, dim, seed=n + nq)
    x[n // 2:n // 2 + 25] = x[3]                        # 26 equal rows spread over two slices
    x[n - 30:n - 10] = x[3]
    q = queries_for(centres, x, nq, seed=k + 7)
    q[2] = x[3]
    ix = ShardIndex(dim, dtype=store)
    ix.add(x)
    ix.set_option("share_floor", 0)
    plain = ix.search(q, k)
    plain_thr = ix.search(q, k, 0.3)
    ix.set_option("share_floor", 1)
    got = check_search(ix, x, q[:24], store, k)
    full = ix.search(q, k)
    assert ix.last_stats()["path"] == 1
    assert all(np.array_equal(u, v) for u, v in zip(full, plain))
    assert all(np.array_equal(u[:24], v) for u, v in zip(full, got))
    assert all(np.array_equal(u, v) for u, v in zip(ix.search(q, k, 0.3), plain_thr))
    ix.set_option("sample_rows", 2048)
    assert all(np.array_equal(u, v) for u, v in zip(ix.search(q, k), plain))
    ix.set_option("sample_rows", 0)
    # deferred warm-up: the first tiles of every slice only seed the floor and are computed again at the end
    for warm in (0, 1, 2, 3, 8):
        ix.set_option("gemm_warm", warm)
        for share in (1, 0):
            ix.set_option("share_floor", share)
            assert all(np.array_equal(u, v) for u, v in zip(ix.search(q, k), plain)), (warm, share)
            assert all(np.array_equal(u, v) for u, v in zip(ix.search(q, k, 0.3), plain_thr)), (warm, share)
    ix.set_option("gemm_warm", 2)
    ix.set_option("share_floor", 1)
    allow = np.random.default_rng(5).random(n) < 0.4
    a = ix.search(q, k, allow=allow)
    ix.set_option("share_floor", 0)
    b = ix.search(q, k, allow=allow)
    assert all(np.array_equal(u, v) for u, v in zip(a, b))


def test_gemm_path_equals_scan_path_and_threshold():
    x, centres = clustered(50000, 384, seed=130)
    q = queries_for(centres, x, 256, seed=131)
    ix = ShardIndex(384)
    ix.add(x)
    for thr in (-np.inf, 0.293, 0.6):
        ix.set_option("force_path", 1)
        a = ix.search(q, 10, thr)
        assert ix.last_stats()["path"] == 1
        ix.set_option("force_path", 0)
        b = ix.search(q, 10, thr)
        assert ix.last_stats()["path"] == 0
        for u, v in zip(a, b):
            assert np.array_equal(u, v)
    ix.set_option("force_path", 1)
    check_search(ix, x, q[:40], "f16", 10, min_similarity=0.293)
    one = ix.search(q[:1], 10)                       # forced GEMM path with a single query
    assert ix.last_stats()["path"] == 1
    ix.set_option("force_path", 0)
    ref = ix.search(q[:1], 10)
    assert all(np.array_equal(u, v) for u, v in zip(one, ref))


def test_int8_gemm_path_threshold_and_scan_equivalence():
    x, centres = clustered(40000, 384, seed=134)
    q = queries_for(centres, x, 200, seed=135)
    ix = ShardIndex(384, dtype="i8")
    ix.add(x)
    for thr in (-np.inf, 0.25, 0.5):
        ix.set_option("force_path", 1)
        a = ix.search(q, 10, thr)
        assert ix.last_stats()["path"] == 1
        ix.set_option("force_path", 0)
        b = ix.search(q, 10, thr)
        assert ix.last_stats()["path"] == 0
        for u, v in zip(a, b):
            assert np.array_equal(u, v)
    ix.set_option("force_path", -1)
    check_search(ix, x, q[:30], "i8", 20, min_similarity=0.25)


def test_gemm_path_duplicates_and_fallback():
    x, centres = clustered(9000, 384, seed=132, dup_frac=0.0)
    x[2000:2060] = x[7]                              # 61 identical rows inside one corpus tile range
    q = np.concatenate([np.stack([x[7], x[4000]]), queries_for(centres, x, 30, seed=133)])
    ix = ShardIndex(384)
    ix.add(x)
    ids, raw, cnt = check_search(ix, x, q, "f16", 10)
    assert ix.last_stats()["path"] == 1
    assert list(ids[0]) == [7] + list(range(2000, 2009))
    assert ix.last_stats()["uncertified_total"] >= 1


@pytest.mark.parametrize("store", ["f16", "bf16", "i8"])
def test_gemm_sampled_start_thresholds_keep_results_exact(store):
    """The contraction may start every slice from a per-query floor taken from a pass over the first rows
    (on by default from 512K rows x > 128 queries); forced on here at a small size, with duplicates of the
    best rows inside and outside the sample, a threshold and a filter."""
    n, dim, nq = 24000, 384, 132
    x, centres = clustered(n, dim, seed=170)
    x[100:130] = x[5]                                  # 31 copies inside the sample ...
    x[20000:20020] = x[5]                              # ... and 20 more far outside it
    q = queries_for(centres, x, nq, seed=171)
    q[2] = x[5]
    ix = ShardIndex(dim, dtype=store)
    ix.add(x)
    ix.set_option("sample_rows", 2048)
    for k in (10, 24):
        got = check_search(ix, x, q, store, k)
        assert ix.last_stats()["path"] == 1 and ix.last_stats()["kernel_launches"] >= 5
        ix.set_option("sample_rows", 0)
        plain = ix.search(q, k)
        ix.set_option("sample_rows", 2048)
        assert all(np.array_equal(u, v) for u, v in zip(got, plain))
    assert list(got[0][2][:24]) == [5] + list(range(100, 123))
    check_search(ix, x, q, store, 10, min_similarity=0.3)
    allow = np.random.default_rng(3).random(n) < 0.5
    rows = np.nonzero(allow)[0]
    want = search.search(encode.encode_rows(x, store)[rows], search.encode_queries(q, store), store, dim, 10)
    f = ix.search(q, 10, allow=allow)
    assert np.array_equal(f[2], want[2])
    assert np.array_equal(f[0], rows[want[0].astype(np.int64)].astype(np.uint32))


@pytest.mark.parametrize("store", ["f16", "bf16", "i8"])
@pytest.mark.parametrize("n,dim,nq,k", [(40000, 384, 24, 100), (9000, 384, 16, 50), (80000, 128, 130, 100),
                                        (2000, 256, 9, 112), (70, 384, 8, 64)])
def test_gemm_path_large_k_bit_exact(store, n, dim, nq, k):
    """k above the 32-key slice lists (BASELINE config 4: top-100 for a batch): certified or recomputed."""
    x, centres = clustered(n, dim, seed=n + k)
    q = queries_for(centres, x, nq, seed=k + 3)
    ix = ShardIndex(dim, dtype=store)
    ix.add(x)
    check_search(ix, x, q, store, k)
    assert ix.last_stats()["path"] == 1
    check_search(ix, x, q[:9], store, k, min_similarity=0.3)


def test_int8_gemm_large_k_overflowing_slice_takes_the_exact_fallback():
    x, centres = clustered(40000, 384, seed=150, dup_frac=0.0)
    x[5000:5080] = x[11]                              # 81 identical rows inside one corpus slice
    q = np.concatenate([x[11][None], queries_for(centres, x, 15, seed=151)])
    ix = ShardIndex(384, dtype="i8")
    ix.add(x)
    ids, raw, cnt = check_search(ix, x, q, "i8", 50)
    st = ix.last_stats()
    assert st["path"] == 1 and st["uncertified_total"] >= 1
    assert list(ids[0]) == [11] + list(range(5000, 5049))
    # device buffers: the conditional exact pass is enqueued without a host round trip
    import torch
    d = ix.search(torch.from_numpy(q).cuda(), 50)
    assert np.array_equal(d[0].cpu().numpy().view(np.uint32), ids) and np.array_equal(d[1].cpu().numpy(), raw)


# ---------------------------------------------------------------- K2/K3 shared-pass small batches
@pytest.mark.parametrize("store,dim", [("b1", 1024), ("b1", 384), ("b1", 2048), ("i8", 128), ("i8", 200)])
@pytest.mark.parametrize("nq,k", [(2, 10), (8, 100), (13, 100), (5, 10), (7, 32)])
def test_shared_pass_integer_batches_bit_exact(store, dim, nq, k):
    n = 30011
    x, centres = clustered(n, dim, seed=dim + nq)
    q = queries_for(centres, x, nq, seed=k)
    ix = ShardIndex(dim, dtype=store)
    ix.add(x)
    ix.set_option("force_path", 0)                      # i8 batches of >= 8 would otherwise take the tensor cores
    got = check_search(ix, x, q, store, k)
    assert ix.last_stats()["kernel_launches"] < 1 + nq + 1, "queries must share corpus passes"
    ix.set_option("multi_scan", 0)                      # one pass per query: same lists, same result
    one = ix.search(q, k)
    assert all(np.array_equal(u, v) for u, v in zip(got, one))
    ix.set_option("multi_scan", 8)
    thr = 0.2 if store == "i8" else 0.05
    check_search(ix, x, q, store, k, min_similarity=thr)
    allow = np.random.default_rng(nq).random(n) < 0.3
    rows = np.nonzero(allow)[0]
    want = search.search(encode.encode_rows(x, store)[rows], search.encode_queries(q, store), store, dim, k)
    f = ix.search(q, k, allow=allow)
    assert np.array_equal(f[2], want[2])
    for i in range(nq):
        c = want[2][i]
        assert np.array_equal(f[0][i, :c], rows[want[0][i, :c].astype(np.int64)].astype(np.uint32))


@pytest.mark.parametrize("store,dim", [("i8", 384), ("b1", 1024)])
def test_integer_scan_short_lists_certify_or_fall_back(store, dim):
    """k > 32 on the scan path: CTAs keep 32 keys; a CTA holding more of the top-k forces the exact pass."""
    n = 50000
    x, centres = clustered(n, dim, seed=160, dup_frac=0.0)
    x[7000:7060] = x[3]                               # 61 identical rows in consecutive tiles of one or two CTAs
    q = np.concatenate([x[3][None], queries_for(centres, x, 4, seed=161)])
    ix = ShardIndex(dim, dtype=store)
    ix.add(x)
    ix.set_option("force_path", 0)
    ids, raw, cnt = check_search(ix, x, q, store, 100)
    assert ix.last_stats()["uncertified_total"] >= 1
    assert list(ids[0][:61]) == [3] + list(range(7000, 7060))
    before = ix.last_stats()["uncertified_total"]
    ix.set_option("short_lists", 0)                     # full 128-key lists: same result, nothing to certify
    full = ix.search(q, 100)
    assert all(np.array_equal(u, v) for u, v in zip(full, (ids, raw, cnt)))
    assert ix.last_stats()["uncertified_total"] == before
    ix.set_option("short_lists", 1)
    import torch
    d = ix.search(torch.from_numpy(q).cuda(), 100, 0.1)           # device buffers + threshold
    w = ix.search(q, 100, 0.1)
    assert np.array_equal(d[0].cpu().numpy().view(np.uint32), w[0]) and np.array_equal(d[2].cpu().numpy(), w[2])
    check_search(ix, x, q, store, 100, min_similarity=0.1)


# ---------------------------------------------------------------- N3: row-bitmap filters
@pytest.mark.parametrize("store", ["f16", "bf16", "i8", "b1"])
def test_filtered_search_equals_oracle_on_allowed_rows(store):
    n = 20000
    x, centres = clustered(n, 384, seed=140)
    q = queries_for(centres, x, 12, seed=141)           # 12 queries: the tensor-core path for f16 / i8
    ix = ShardIndex(384, dtype=store)
    ix.add(x)
    rng = np.random.default_rng(142)
    codes = encode.encode_rows(x, store)
    qc = search.encode_queries(q, store)
    for frac, force in ((0.5, -1), (0.01, -1), (0.0003, -1), (0.2, 0)):
        allow = rng.random(n) < frac
        allow[17] = True
        rows = np.nonzero(allow)[0]
        want = search.search(codes[rows], qc, store, 384, 10)
        ix.set_option("force_path", force)
        got = ix.search(q, 10, allow=allow)
        assert ix.last_stats()["path"] == (1 if (store != "b1" and force != 0) else 0)
        assert np.array_equal(got[2], want[2])
        for i in range(len(q)):
            c = want[2][i]
            assert np.array_equal(got[0][i, :c], rows[want[0][i, :c].astype(np.int64)].astype(np.uint32))
            assert np.array_equal(got[1][i, :c], want[1][i, :c])
            assert np.all(got[0][i, c:] == 0xFFFFFFFF)
    with pytest.raises(ValueError):
        ix.search(q, 10, allow=np.ones(5, bool))
