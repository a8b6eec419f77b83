"""GPU tier (-m gpu): the peer-memory exchange + merge kernel (crs_exchange, csrc/exchange.cu).

On a one-GPU box the ranks of a sharded search are emulated in ONE process on ONE device: every rank has
its own shard index and its own exchange, the exchanges are wired with ``crs_exchange_set_peer_buffers``
(same-device pointers are valid "peers"), and the split form of the search is used — all pushes first, then
all merges — so that no kernel ever waits for a kernel that has not been launched.  The result must equal
the single-index search and the oracle's shard merge bit for bit.  The fused (one launch) form needs real
concurrency between the ranks and is covered by tools/multi_gpu_check.py under torchrun (>= 2 GPUs)."""
import numpy as np
import pytest

from helpers import clustered, queries_for

pytestmark = pytest.mark.gpu

from compressed_rag_suite_b200.index import ShardIndex  # noqa: E402
from compressed_rag_suite_b200.sharded import PeerExchange, shard_bounds  # noqa: E402


def _same(got, want):
    import torch
    torch.cuda.synchronize()
    assert np.array_equal(got[0].cpu().numpy().view(np.uint32), want[0])
    assert np.array_equal(got[1].cpu().numpy().view(np.uint32), want[1].view(np.uint32))
    assert np.array_equal(got[2].cpu().numpy(), want[2])


@pytest.mark.parametrize("store,world", [("f16", 2), ("i8", 3), ("b1", 2), ("bf16", 8)])
def test_emulated_ranks_push_then_merge_equal_single_index(store, world):
    import torch
    n, dim = 30011, (1024 if store == "b1" else 384)
    x, centres = clustered(n, dim, seed=400 + world)
    whole = ShardIndex(dim, dtype=store)
    whole.add(x)
    shards, exs = [], []
    for r in range(world):
        lo, hi = shard_bounds(n, world, r)
        sh = ShardIndex(dim, dtype=store, row_base=lo)
        sh.add(x[lo:hi])
        shards.append(sh)
        exs.append(PeerExchange(0, r, world, max_nq=300, max_k=100))
    PeerExchange.wire_local(exs)
    # consecutive steps of different shapes: the receive slots are double-buffered by step parity
    for nq, k, thr in [(5, 10, -np.inf), (300, 10, -np.inf), (1, 10, 0.3), (40, 100, -np.inf), (129, 24, 0.25), (5, 10, -np.inf)]:
        q = queries_for(centres, x, nq, seed=nq + k)
        want = whole.search(q, k, thr)
        qd = torch.from_numpy(q).cuda()
        for r in range(world):
            shards[r].search_push(exs[r], qd, k, thr)
        for r in range(world):
            _same(exs[r].merge(nq, k, whole.is_int), want)
    for e in exs:
        timed_out, step = e.status()
        assert not timed_out and step == 6
    # a merge whose peers never pushed gives up after ~2 s and reports it instead of hanging
    shards[0].search_push(exs[0], qd, 10)
    exs[0].merge(5, 10, whole.is_int)
    timed_out, step = exs[0].status()
    assert timed_out and step == 7


def test_world_one_exchange_is_the_plain_search():
    import torch
    x, centres = clustered(9000, 384, seed=410)
    q = queries_for(centres, x, 33, seed=411)
    ix = ShardIndex(384, row_base=77)
    ix.add(x)
    ex = PeerExchange(0, 0, 1, max_nq=64, max_k=16)
    want = ix.search(q, 10)
    _same(ix.search_sharded(ex, torch.from_numpy(q).cuda(), 10), want)
    got = ix.search_sharded(ex, q, 10)                       # host buffers: synchronous
    assert all(np.array_equal(u, v) for u, v in zip(got, want))
    with pytest.raises(ValueError):
        ix.search_sharded(ex, q, 32)                         # k above what the exchange was created for
    empty = ShardIndex(384)
    got = empty.search_sharded(ex, q, 10)                    # an empty shard still takes part in the step
    assert got[2].max() == 0 and np.all(got[0] == 0xFFFFFFFF)
    assert ex.status() == (False, 3)


@pytest.mark.parametrize("store", ["f16", "i8"])
def test_peer_shards_fetch_and_score_rows_equal_the_single_index(store):
    """Candidate rows by global id out of the owning shard (peer-mapped code buffers; here: three shards side by
    side on one device): the same bytes / canonical scores as the single index gives."""
    import torch
    n, dim, world = 9001, 384, 3
    x, centres = clustered(n, dim, seed=420)
    whole = ShardIndex(dim, dtype=store)
    whole.add(x)
    shards, exs = [], []
    for r in range(world):
        lo, hi = shard_bounds(n, world, r)
        sh = ShardIndex(dim, dtype=store, row_base=lo)
        sh.add(x[lo:hi])
        shards.append(sh)
        exs.append(PeerExchange(0, r, world, max_nq=8, max_k=8))
    PeerExchange.wire_local(exs)
    PeerExchange.register_shards_local(exs, shards)
    q = queries_for(centres, x, 6, seed=421)
    ids, raw, cnt = whole.search(q, 50)
    ids_t = torch.from_numpy(ids.view(np.int32)).cuda()
    ids_t[0, 3] = -1                                          # a pad id
    want_rows = whole.fetch_rows(ids.reshape(-1)).reshape(6, 50, -1)
    want_rows[0, 3] = 0
    for r in range(world):
        got = exs[r].fetch_rows(ids_t, whole.row_bytes)
        torch.cuda.synchronize()
        assert np.array_equal(got.cpu().numpy(), want_rows)
        sc = exs[r].score_rows(shards[r], torch.from_numpy(q).cuda(), ids_t)
        torch.cuda.synchronize()
        sc = sc.cpu().numpy()
        assert np.array_equal(sc[0, 4:].view(np.uint32), raw[0, 4:].view(np.uint32)) and np.array_equal(sc[1:].view(np.uint32), raw[1:].view(np.uint32))
        assert sc[0, 3] == (np.iinfo(np.int32).min if whole.is_int else -np.inf)
