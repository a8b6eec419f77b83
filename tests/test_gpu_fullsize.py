"""GPU tier (-m gpu), BASELINE.json's FULL per-GPU sizes: the oracle cannot brute-force these in
seconds, so parity is carried by size-independent properties, each of which the small-size
tests pin against the oracle first:

* exactness by exclusion (fp16, 10M x 384, 1024-query batch): the returned scores are the
  canonical ones (oracle arithmetic on the fetched rows) and NO other row can beat the k-th —
  every row whose independent fp32 score (torch matmul, used only as a filter) comes within
  1e-3 of the k-th is rescored canonically on the CPU;
* exhaustive integer check (int8, 12.5M x 384, top-100): fp32 matmul of int8 codes is exact
  (|dot| <= 384 * 127^2 < 2^24), so a blocked torch brute force gives the reference ids;
* path equivalence (tensor-core batch == one-pass-per-query scans == shared-pass scans),
  shard-and-merge == single index, threshold == prefix of the unthresholded list, planted
  duplicates -> lowest id first, order / counts invariants.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

from oracle import search  # noqa: E402

pytestmark = pytest.mark.gpu

import bench as hb  # noqa: E402
from compressed_rag_suite_b200.index import ShardIndex, merge_topk  # noqa: E402


def _build(n, dim, store, row_base=0, lo=0):
    dev = torch.device("cuda", 0)
    centres = hb.gen_centres(torch, dim, dev)
    ix = ShardIndex(dim, dtype=store, row_base=row_base, reserve_rows=n)
    hi = lo + n
    for blk in range(lo // hb.BLOCK_ROWS, (hi - 1) // hb.BLOCK_ROWS + 1):
        ix.add(hb.gen_block(torch, blk, lo, hi, dim, centres, dev))
    assert len(ix) == n
    return ix, centres


def _invariants(ids, sc, cnt, k):
    assert (cnt == k).all()
    s = sc.astype(np.float64)
    assert (np.diff(s, axis=1) <= 0).all(), "scores must be descending"
    tie = np.diff(s, axis=1) == 0
    assert (np.diff(ids.astype(np.int64), axis=1)[tie] > 0).all(), "equal scores must come in ascending id order"
    assert all(len(set(r)) == k for r in ids.tolist()), "no row may appear twice"


@pytest.fixture(scope="module")
def f16_10m():
    n, dim, nq = 10_000_000, 384, 1024
    ix, centres = _build(n, dim, "f16")
    q = hb.gen_queries(torch, nq, dim, centres, torch.device("cuda", 0), n)
    ids, sc, cnt = ix.search(q, 10)
    assert ix.last_stats()["path"] == 1
    yield ix, q, ids.cpu().numpy().view(np.uint32), sc.cpu().numpy(), cnt.cpu().numpy()
    ix.close()


def test_fp16_10m_batch_invariants_and_planted_rows(f16_10m):
    ix, q, ids, sc, cnt = f16_10m
    _invariants(ids, sc, cnt, 10)
    planted = np.arange(0, 1024, 100)                      # bench queries copied from rows (pos * 9973) % 2^20
    rows = (planted * 9973) % hb.BLOCK_ROWS
    for qi, r in zip(planted, rows):
        # rows r with r % 100 == 7 are copies of r - 1: the lower id must win the tie
        first = r - 1 if r % 100 == 7 else r
        assert ids[qi, 0] == first and sc[qi, 0] > 0.999
        if r % 100 in (6, 7):
            assert ids[qi, 1] == first + 1 and sc[qi, 1] == sc[qi, 0]


def test_fp16_10m_exact_by_exclusion(f16_10m):
    ix, q, ids, sc, cnt = f16_10m
    sample = np.r_[0, 100, np.arange(3, 1024, 97)]          # 13 queries incl. two planted ones
    qc = search.encode_queries(q[sample].cpu().numpy(), "f16")
    # (a) the reported scores are the canonical scores of the reported rows
    for j, qi in enumerate(sample):
        rows = ix.fetch_rows(ids[qi]).view(np.float16)
        want = search.raw_scores(rows, qc[j], "f16", 384)
        assert np.array_equal(want.view(np.uint32), sc[qi].view(np.uint32))
    # (b) nothing outside the list can beat the k-th: fp32 filter on the device, canonical rescoring on the CPU
    qd = torch.from_numpy(qc.astype(np.float32)).cuda()
    kth = torch.from_numpy(sc[sample, 9].copy()).cuda()
    suspects = [[] for _ in sample]
    step = 1 << 20
    for lo in range(0, len(ix), step):
        n = min(step, len(ix) - lo)
        rows_dev = torch.arange(lo, lo + n, dtype=torch.int32, device="cuda")
        blk = ix.fetch_rows_device(rows_dev).view(torch.float16)      # stored codes, [n, 384]
        approx = blk.float() @ qd.T                                   # [n, 13] fp32
        hit = (approx >= (kth - 1e-3)[None, :]).nonzero().cpu().numpy()
        for r, j in hit:
            suspects[j].append(lo + r)
    for j, qi in enumerate(sample):
        cand = np.asarray(sorted(suspects[j]), dtype=np.uint32)
        assert set(ids[qi].tolist()) <= set(cand.tolist())
        rows = ix.fetch_rows(cand).view(np.float16)
        exact = search.raw_scores(rows, qc[j], "f16", 384)
        order = np.lexsort((cand, -exact.astype(np.float64)))[:10]
        assert np.array_equal(cand[order], ids[qi]), f"query {qi}: a row outside the result beats the k-th"
        assert np.array_equal(exact[order].view(np.uint32), sc[qi].view(np.uint32))


def test_fp16_10m_paths_agree_and_threshold_is_a_prefix(f16_10m):
    ix, q, ids, sc, cnt = f16_10m
    sub = q[:24].contiguous()
    ix.set_option("force_path", 0)
    a = ix.search(sub, 10)                                            # 24 single-query scans
    assert ix.last_stats()["path"] == 0
    ix.set_option("force_path", -1)
    assert np.array_equal(a[0].cpu().numpy().view(np.uint32), ids[:24])
    assert np.array_equal(a[1].cpu().numpy().view(np.uint32), sc[:24].view(np.uint32))
    thr = 0.45
    t = ix.search(q, 10, thr)
    t_ids, t_sc, t_cnt = t[0].cpu().numpy().view(np.uint32), t[1].cpu().numpy(), t[2].cpu().numpy()
    want_cnt = (sc >= np.float32(thr)).sum(axis=1)
    assert np.array_equal(t_cnt, want_cnt)
    assert 0 < (want_cnt < 10).sum() < 1024, "the threshold must cut some lists and not others"
    for i in range(1024):
        c = want_cnt[i]
        assert np.array_equal(t_ids[i, :c], ids[i, :c]) and (t_ids[i, c:] == 0xFFFFFFFF).all()
    again = ix.search(q, 10)
    assert np.array_equal(again[0].cpu().numpy().view(np.uint32), ids), "deterministic"


def test_fp16_10m_shard_and_merge_equals_single_index(f16_10m):
    ix, q, ids, sc, cnt = f16_10m
    n, g = 10_000_000, 4
    per = n // g
    parts = []
    for r in range(g):
        s, _ = _build(per, 384, "f16", row_base=r * per, lo=r * per)
        parts.append(s.search(q, 10))
        s.close()
    m = merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]), 10)
    assert np.array_equal(m[0].cpu().numpy().view(np.uint32), ids)
    assert np.array_equal(m[1].cpu().numpy().view(np.uint32), sc.view(np.uint32))


def test_int8_12m_top100_equals_exhaustive_integer_bruteforce():
    """BASELINE config 4's per-GPU shard: 12.5M x 384 int8, top-100, single query and a batch."""
    n, dim, k = 12_500_000, 384, 100
    ix, centres = _build(n, dim, "i8")
    q = hb.gen_queries(torch, 16, dim, centres, torch.device("cuda", 0), n)
    one = ix.search(q[:1], k)                                         # stream scan, short lists + certification
    assert ix.last_stats()["path"] == 0
    batch = ix.search(q, k)                                           # tensor-core path, k above the slice lists
    assert ix.last_stats()["path"] == 1
    qc = torch.from_numpy(search.encode_queries(q.cpu().numpy(), "i8").astype(np.float32)).cuda()
    best_s = torch.full((16, k), -2.0 ** 30, device="cuda")
    best_i = torch.full((16, k), -1, dtype=torch.int64, device="cuda")
    step = 1 << 20
    for lo in range(0, n, step):
        m = min(step, n - lo)
        codes = ix.fetch_rows_device(torch.arange(lo, lo + m, dtype=torch.int32, device="cuda")).view(torch.int8)
        s = qc @ codes.float().T                                      # exact integers in fp32
        # top-k of (score desc, id asc): stable sort of the concatenation keeps the lower id first
        cs = torch.cat([best_s, s], dim=1)
        ci = torch.cat([best_i, torch.arange(lo, lo + m, device="cuda")[None, :].expand(16, m)], dim=1)
        o = torch.sort(cs, dim=1, descending=True, stable=True).indices[:, :k]
        best_s, best_i = torch.gather(cs, 1, o), torch.gather(ci, 1, o)
    want_i = best_i.cpu().numpy().astype(np.uint32)
    want_s = best_s.cpu().numpy().astype(np.int32)
    assert np.array_equal(batch[0].cpu().numpy().view(np.uint32), want_i)
    assert np.array_equal(batch[1].cpu().numpy(), want_s)
    assert np.array_equal(one[0].cpu().numpy().view(np.uint32), want_i[:1])
    assert np.array_equal(one[1].cpu().numpy(), want_s[:1])
    ix.close()


def test_binary_125m_planted_neighbours_and_shared_pass_equivalence():
    """BASELINE config 5's per-GPU shard: 125M x 1024-bit codes (16 GB), Hamming top-100."""
    import bench_configs as bc
    n, dim, k = 125_000_000, 1024, 100
    dev = torch.device("cuda", 0)
    centres = hb.gen_centres(torch, dim, dev)
    ix = ShardIndex(dim, dtype="b1", reserve_rows=n)
    for off in range(0, n, 1 << 19):
        rows = torch.arange(off, min(off + (1 << 19), n), device=dev, dtype=torch.int64)
        ix.add(bc.counter_rows(torch, rows, dim, centres))
    planted = torch.tensor([5, 1 << 19, 77_777_777, n - 1, 31_415_926, 124_999_000, 64_000_001, 99], device=dev)
    g = torch.Generator(device=dev); g.manual_seed(9)
    noise = torch.randn(8, dim, generator=g, device=dev)
    q = bc.counter_rows(torch, planted, dim, centres) + 0.25 * noise / noise.norm(dim=1, keepdim=True)
    q = (q / q.norm(dim=1, keepdim=True)).contiguous()
    shared = ix.search(q, k)                                          # 8 queries share corpus passes
    assert ix.last_stats()["kernel_launches"] <= 6
    ids = shared[0].cpu().numpy().view(np.uint32)
    sc = shared[1].cpu().numpy()
    assert np.array_equal(ids[:, 0], planted.cpu().numpy().astype(np.uint32)), "planted neighbours must come first"
    _invariants(ids, sc.astype(np.float32), shared[2].cpu().numpy(), k)
    ix.set_option("multi_scan", 0)
    ix.set_option("short_lists", 0)
    single = ix.search(q[:3].contiguous(), k)                          # one pass per query, full 128-key lists
    assert np.array_equal(single[0].cpu().numpy().view(np.uint32), ids[:3])
    assert np.array_equal(single[1].cpu().numpy(), sc[:3])
    # the canonical Hamming score of the reported rows, by the oracle on the fetched codes
    qc = search.encode_queries(q[:2].cpu().numpy(), "b1")
    for j in range(2):
        rows = ix.fetch_rows(ids[j]).view(np.uint32)
        assert np.array_equal(search.raw_scores(rows, qc[j], "b1", dim), sc[j])
    ix.close()
