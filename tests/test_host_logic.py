"""CPU tier: host-side logic of the Chroma-shaped collection that needs no GPU — the column-wise ``where`` /
``where_document`` evaluation must agree with the row-by-row definition (which mirrors Chroma's operators as
the reference can reach them, rag/indexing.py:129-130,174-175), and the multi-device id routing."""
import random

import numpy as np
import pytest

from compressed_rag_suite_b200 import collection as backend


def _random_metas(rng, n):
    metas = []
    for i in range(n):
        if rng.random() < 0.1:
            metas.append(None)
            continue
        m = {}
        if rng.random() < 0.9:
            m["page_number"] = rng.choice([0, 1, 2, 3, 7, 11, 2 ** 60 + 1])
        if rng.random() < 0.7:
            m["section"] = rng.choice(["s0", "s1", "intro", ""])
        if rng.random() < 0.5:
            m["tokens"] = rng.choice([0.5, 12.0, 100, 3])
        if rng.random() < 0.2:
            m["flag"] = rng.choice([True, False])
        metas.append(m)
    return metas


def _random_where(rng, depth=0):
    r = rng.random()
    if depth < 2 and r < 0.25:
        return {rng.choice(["$and", "$or"]): [_random_where(rng, depth + 1) for _ in range(rng.randint(1, 3))]}
    key = rng.choice(["page_number", "section", "tokens", "flag", "missing"])
    num = key in ("page_number", "tokens")
    ref = rng.choice([0, 3, 7, 2 ** 60 + 1, 12.0, 0.5]) if num else rng.choice(["s0", "intro", "", True, "zzz"])
    op = rng.choice(["$eq", "$ne", "$in", "$nin", "plain"] + (["$gt", "$gte", "$lt", "$lte"] if num else []))
    if op == "plain":
        return {key: ref}
    if op in ("$in", "$nin"):
        return {key: {op: [ref, rng.choice([1, "s1", 100])]}}
    return {key: {op: ref}}


def test_columnwise_where_equals_the_row_by_row_definition():
    rng = random.Random(5)
    metas = _random_metas(rng, 700)
    cols = backend._Columns(metas)
    for _ in range(400):
        w = _random_where(rng)
        want = np.array([backend._where_ok(m or {}, w) for m in metas])
        assert np.array_equal(cols.mask(w), want), w
    metas.extend(_random_metas(rng, 150))                # columns grow with the collection
    for _ in range(100):
        w = _random_where(rng)
        assert np.array_equal(cols.mask(w), np.array([backend._where_ok(m or {}, w) for m in metas])), w
    with pytest.raises(ValueError):
        cols.mask({"page_number": {"$regex": "x"}})
    with pytest.raises(TypeError):
        cols.mask({"section": {"$gt": 3}})               # a string against a number raises, as the row-by-row form does


def test_document_mask_equals_the_row_by_row_definition():
    rng = random.Random(6)
    docs = [rng.choice(["alpha beta", "omega", None, "", "beta omega alpha"]) for _ in range(300)]
    for cond in [{"$contains": "alpha"}, {"$not_contains": "omega"}, {"$and": [{"$contains": "beta"}, {"$not_contains": "omega"}]},
                 {"$or": [{"$contains": "omega"}, {"$contains": "zzz"}]}, None]:
        want = np.array([backend._doc_ok(d or "", cond) for d in docs])
        assert np.array_equal(backend._doc_mask(docs, cond), want), cond
    with pytest.raises(ValueError):
        backend._doc_mask(docs, {"$like": "x"})


def test_multi_device_id_routing_without_a_gpu():
    """multi.MultiDeviceIndex keeps (first global id, n, shard, first local row) segments; global ids are insertion
    indices.  The routing of ids and of a global allow-mask to the shards is plain numpy: checked here against a
    brute-force table, on an instance whose shards are stand-ins."""
    from compressed_rag_suite_b200.multi import MultiDeviceIndex

    class FakeShard:
        def __init__(self):
            self.n = 0

        def __len__(self):
            return self.n

    rng = np.random.default_rng(9)
    g = 3
    m = MultiDeviceIndex.__new__(MultiDeviceIndex)
    m.shards = [FakeShard() for _ in range(g)]
    m._seg, m._seg_start, m._count, m._next = [], [], 0, 0
    owner, local = [], []
    for n in [7, 1, 2, 50, 1, 1, 1, 13]:                      # the same dealing rule as MultiDeviceIndex.add
        if n < g:
            pieces = [(i, i + 1, (m._next + i) % g) for i in range(n)]
            m._next = (m._next + n) % g
        else:
            per = (n + g - 1) // g
            pieces = [(lo, min(lo + per, n), j) for j, lo in enumerate(range(0, n, per))]
        for lo, hi, j in pieces:
            sh = m.shards[j]
            m._seg_start.append(m._count + lo)
            m._seg.append((m._count + lo, hi - lo, j, sh.n))
            for r in range(hi - lo):
                owner.append(j)
                local.append(sh.n + r)
            sh.n += hi - lo
        m._count += n
    assert m._count == len(owner) == sum(len(s) for s in m.shards)
    ids = rng.integers(0, m._count, 200)
    sh, lo = m._locate(ids)
    assert np.array_equal(sh, np.array(owner)[ids]) and np.array_equal(lo, np.array(local)[ids])
    sh, lo = m._locate(np.array([-1, m._count, m._count + 5]))
    assert (sh == -1).all()
    allow = rng.random(m._count) < 0.4
    per_shard = m._local_allow(allow)
    for j in range(g):
        want = np.zeros(len(m.shards[j]), dtype=bool)
        for gid in range(m._count):
            if owner[gid] == j:
                want[local[gid]] = allow[gid]
        assert np.array_equal(per_shard[j], want)
    with pytest.raises(ValueError):
        m._local_allow(np.ones(3, dtype=bool))
    # every shard's ids grow with its local rows (ties inside a shard go to the lowest local row = first inserted)
    for j in range(g):
        gids = [gid for gid in range(m._count) if owner[gid] == j]
        assert [local[gid] for gid in gids] == list(range(len(gids)))
