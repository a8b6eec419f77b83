"""GPU tier (-m gpu): the drop-in boundary.  The product VectorStore / ContextRetriever
must return exactly what the reference's own classes returned for the same inputs
(tests/golden/retrieval_golden.*, produced by the unmodified reference code), and the
Chroma-shaped client must honour the guards the reference relies on."""
import json
import os

import numpy as np
import pytest

from oracle import encode, fake_chroma, postprocess

pytestmark = pytest.mark.gpu

from compressed_rag_suite_b200.rag import Chunk, ContextRetriever, VectorStore  # noqa: E402


class TableEmbedder:
    def __init__(self):
        self.table = {}
        self.calls = 0

    def embed(self, texts):
        self.calls += 1
        if isinstance(texts, str):
            texts = [texts]
        return np.stack([self.table[t] for t in texts]).astype(np.float32)


@pytest.fixture(scope="module")
def golden(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "retrieval_golden.json")))
    arr = np.load(os.path.join(golden_dir, "retrieval_golden.npz"))
    return g, arr["embeddings"], arr["queries"]


def build_store(g, x, name):
    chunks = [Chunk(text=t, chunk_id=f"chunk_{i}", start_char=0, end_char=len(t), **g["chunk_meta"][i])
              for i, t in enumerate(g["texts"])]
    vs = VectorStore({"collection_name": name})
    vs.create_index(chunks, x)
    return vs, chunks


def test_retrieve_equals_reference_golden(golden):
    g, x, queries = golden
    vs, _ = build_store(g, x, "golden_gpu")
    emb = TableEmbedder()
    for t, v in zip(g["query_texts"], queries):
        emb.table[t] = v
    assert len(g["cases"]) >= 80
    for case in g["cases"]:
        r = ContextRetriever(vs, emb, case["config"])
        got = r.retrieve(g["query_texts"][case["query"]])
        assert [c["chunk_id"] for c in got] == case["chunk_ids"], case
        assert [c["score"] for c in got] == case["scores"]                 # bit-exact Python floats
        assert [c["distance"] for c in got] == case["distances"]
        assert [c.get("rerank_score") for c in got] == case["rerank_scores"]
        assert [c["metadata"] for c in got] == case["metadatas"]
    assert set(emb.table) == set(g["query_texts"])      # chunk texts were never re-embedded


def test_retrieve_batch_equals_one_by_one(golden):
    g, x, queries = golden
    vs, _ = build_store(g, x, "golden_gpu_batch")
    emb = TableEmbedder()
    for t, v in zip(g["query_texts"], queries):
        emb.table[t] = v
    for cfg in ({"top_k": 3, "similarity_threshold": 0.3, "rerank": True, "diversity_penalty": 0.1},
                {"top_k": 5, "similarity_threshold": 0.75, "rerank": False, "diversity_penalty": 0.4}):
        r = ContextRetriever(vs, emb, cfg)
        single = [r.retrieve(t) for t in g["query_texts"]]
        calls = emb.calls
        batch = r.retrieve_batch(g["query_texts"])
        assert emb.calls == calls + 1
        assert batch == single
    assert r.get_context_string(g["query_texts"][0]) == "\n\n".join(c["text"] for c in single[0])


def test_product_pipeline_equals_oracle_pipeline_on_fresh_data():
    """Same inputs through (oracle fake-chroma, canonical f16) + OracleRetriever and through the product."""
    rng = np.random.default_rng(314)
    n, dim = 300, 384
    x = rng.standard_normal((n, dim)).astype(np.float32)
    texts = [f"alpha beta gamma {' '.join(rng.choice(['delta', 'eps', 'zeta', 'eta'], 4))} doc{i}" for i in range(n)]
    chunks = [Chunk(t, f"chunk_{i}", 0, len(t), page_number=i % 9) for i, t in enumerate(texts)]
    vs = VectorStore({"collection_name": "fresh"})
    vs.create_index(chunks, x)
    fake_chroma.PRECISION = "f16"
    oc = fake_chroma.Client().create_collection("fresh", {"hnsw:space": "cosine"})
    oc.add(ids=[c.chunk_id for c in chunks], embeddings=x, documents=texts,
           metadatas=[{"page_number": c.page_number} for c in chunks])

    class OStore:
        collection = oc

        def search(self, query_embedding, top_k=5, where=None, **_):
            return oc.query(np.asarray(query_embedding).reshape(1, -1), min(top_k, oc.count()), where)

    stored = encode.encode_rows(x, "f16").astype(np.float32)
    emb = TableEmbedder()
    qs = []
    for j in range(20):
        t = f"alpha eps question{j}"
        emb.table[t] = (x[rng.integers(0, n)] + 0.8 * rng.standard_normal(dim)).astype(np.float32)
        qs.append(t)
    for cfg in ({"top_k": 3, "similarity_threshold": 0.3, "rerank": True, "diversity_penalty": 0.1},
                {"top_k": 8, "similarity_threshold": 0.0, "rerank": True, "diversity_penalty": 0.6},
                {"top_k": 6, "similarity_threshold": 0.55, "rerank": False, "diversity_penalty": 0.2}):
        prod = ContextRetriever(vs, emb, cfg)
        orc = postprocess.OracleRetriever(OStore(), emb, cfg,
                                          lambda ids: np.stack([stored[int(i.split('_')[1])] for i in ids]))
        for t in qs:
            assert prod.retrieve(t) == orc.retrieve(t)


def test_vectorstore_contract():
    vs = VectorStore({"collection_name": "contract"})
    assert vs.collection is None and vs.get_stats() == {"status": "empty", "count": 0}
    with pytest.raises(ValueError):
        vs.search(np.zeros(4, np.float32))
    vs.create_index([], np.zeros((0, 4), np.float32))
    assert vs.collection is None
    with pytest.raises(ValueError):
        vs.create_index([Chunk("a", "chunk_0", 0, 1)], np.zeros((2, 4), np.float32))
    x = np.eye(4, dtype=np.float32)
    chunks = [Chunk(f"doc {i}", f"chunk_{i}", 0, 5, page_number=i, section=None, tokens=2) for i in range(4)]
    vs.create_index(chunks, x)
    vs.create_index(chunks[:2], x[:2])                  # same ids again: no-op, not an upsert
    assert vs.get_stats() == {"name": "contract", "count": 4, "metadata": {"hnsw:space": "cosine"}}
    out = vs.search(np.array([[1.0, 0.1, 0, 0]], np.float32), top_k=10)      # top_k clamped to count
    assert out["ids"] == [["chunk_0", "chunk_1", "chunk_2", "chunk_3"]]
    assert out["documents"][0][0] == "doc 0"
    assert out["metadatas"][0][0] == {"page_number": 0, "tokens": 2}
    assert out["distances"][0] == sorted(out["distances"][0])
    assert vs.search([1.0, 0.1, 0.0, 0.0], top_k=2)["ids"] == [["chunk_0", "chunk_1"]]     # plain list
    r = ContextRetriever(vs, None, {})
    assert (r.top_k, r.similarity_threshold, r.rerank, r.diversity_penalty, r.distance_metric) == \
        (3, 0.0, False, 0.0, "cosine")
    vs.delete_collection()
    assert vs.collection is None and vs.get_stats()["count"] == 0
    vs.reset_collection()


def test_persist_directory_roundtrip(tmp_path):
    rng = np.random.default_rng(8)
    x = rng.standard_normal((50, 384)).astype(np.float32)
    chunks = [Chunk(f"t{i}", f"chunk_{i}", 0, 2, page_number=i) for i in range(50)]
    cfg = {"collection_name": "persisted", "persist_directory": str(tmp_path / "vector_db")}
    a = VectorStore(cfg)
    a.create_index(chunks, x)
    want = a.search(x[7], top_k=5)
    b = VectorStore(cfg)                                # new "process": get_collection reload path
    assert b.collection is not None and b.get_stats()["count"] == 50
    assert b.search(x[7], top_k=5) == want
    b.create_index(chunks[:10], x[:10])                 # ids collide across processes: ignored
    assert b.get_stats()["count"] == 50


def test_where_filters_equal_oracle_chroma():
    rng = np.random.default_rng(2718)
    n, dim = 400, 384
    x = rng.standard_normal((n, dim)).astype(np.float32)
    texts = [f"{'alpha' if i % 3 else 'omega'} body text {i}" for i in range(n)]
    chunks = [Chunk(t, f"chunk_{i}", 0, len(t), page_number=i % 11, section=f"s{i % 4}" if i % 5 else None)
              for i, t in enumerate(texts)]
    vs = VectorStore({"collection_name": "filters"})
    vs.create_index(chunks, x)
    fake_chroma.PRECISION = "f16"
    oc = fake_chroma.Client().create_collection("filters", {"hnsw:space": "cosine"})
    metas = [VectorStore._chunk_metadata(c, ("page_number", "section", "tokens")) for c in chunks]
    oc.add(ids=[c.chunk_id for c in chunks], embeddings=x, documents=texts, metadatas=metas)
    q = rng.standard_normal(dim).astype(np.float32)
    cases = [({"page_number": 3}, None), ({"page_number": {"$gte": 9}}, None),
             ({"$and": [{"page_number": {"$lt": 4}}, {"section": {"$ne": "s1"}}]}, None),
             ({"$or": [{"section": "s2"}, {"page_number": {"$in": [0, 10]}}]}, None),
             (None, {"$contains": "omega"}), ({"page_number": {"$nin": [1, 2, 3]}}, {"$not_contains": "omega"}),
             ({"page_number": 99}, None)]
    for where, where_doc in cases:
        got = vs.search(q, top_k=7, where=where, where_document=where_doc)
        want = oc.query(q[None, :], 7, where, where_doc)
        for key in ("ids", "documents", "metadatas", "distances"):
            assert got[key] == want[key], (where, where_doc, key)
    emb = TableEmbedder()
    emb.table["find alpha"] = q
    r = ContextRetriever(vs, emb, {"top_k": 3, "rerank": True, "diversity_penalty": 0.2})
    out = r.retrieve("find alpha", filters={"page_number": 5})
    assert len(out) == 3 and all(c["metadata"]["page_number"] == 5 for c in out)
    assert r.retrieve_batch(["find alpha"], filters={"page_number": 5}) == [out]


def test_plain_c_host_runs_a_search(tmp_path):
    """examples/crs_example.c: a C99 program indexes 4096 rows and finds row 123 by its own vector."""
    import subprocess
    sys_path_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(sys_path_root, "compressed_rag_suite_b200")
    exe = os.path.join(str(tmp_path), "crs_example")
    b = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(sys_path_root, "include"),
                        os.path.join(sys_path_root, "examples", "crs_example.c"), "-o", exe, "-L" + libdir, "-lcrs",
                        "-Wl,-rpath," + libdir, "-lm"], capture_output=True, text=True)
    assert b.returncode == 0, b.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr
    assert "rank 0: row 123" in r.stdout


# ---------------------------------------------------------------- round 2: devices, device tensors, append-only persistence
def _texts_chunks(n):
    texts = [f"{'alpha' if i % 3 else 'omega'} body text {i}" for i in range(n)]
    chunks = [Chunk(t, f"chunk_{i}", 0, len(t), page_number=i % 11, section=f"s{i % 4}" if i % 5 else None)
              for i, t in enumerate(texts)]
    return texts, chunks


@pytest.mark.parametrize("dtype", ["f16", "i8"])
@pytest.mark.parametrize("devices", [[0, 0], [0, 0, 0]])
def test_vectorstore_over_several_devices_equals_one_device(dtype, devices, tmp_path):
    """``devices``: the collection is dealt out over several shards (here: side by side on the one GPU of the
    box) and every search runs on all of them from this one process — push to the peers' receive buffers,
    one merge.  Same dicts as the single-device store (rag/indexing.py:125-180 contract), with duplicates
    (ties -> first inserted row ACROSS shards), incremental adds, filters, MMR and a reload."""
    rng = np.random.default_rng(31)
    n, dim = 1500, 384
    x = rng.standard_normal((n, dim)).astype(np.float32)
    x[900] = x[3]
    x[1400] = x[3]                                       # equal rows on different shards
    texts, chunks = _texts_chunks(n)
    one = VectorStore({"collection_name": "one", "dtype": dtype})
    many = VectorStore({"collection_name": "many", "dtype": dtype, "devices": devices,
                        "persist_directory": str(tmp_path / "db")})
    for lo, hi in [(0, 700), (700, 701), (701, 703), (703, n)]:          # bulk, single-row and short adds
        one.create_index(chunks[lo:hi], x[lo:hi])
        many.create_index(chunks[lo:hi], x[lo:hi])
    assert many.get_stats()["count"] == n
    qs = [x[3], x[77] + 0.05 * rng.standard_normal(dim).astype(np.float32), rng.standard_normal(dim).astype(np.float32)]
    for q in qs:
        for where, where_doc in [(None, None), ({"page_number": {"$gte": 6}}, None), (None, {"$contains": "omega"})]:
            a = one.search(q, top_k=9, where=where, where_document=where_doc)
            b = many.search(q, top_k=9, where=where, where_document=where_doc)
            assert a == b, (where, where_doc)
    assert many.search(x[3], top_k=3)["ids"] == [["chunk_3", "chunk_900", "chunk_1400"]]
    emb = TableEmbedder()
    emb.table["q0"], emb.table["q1"] = qs[1], qs[2]
    cfg = {"top_k": 4, "rerank": True, "diversity_penalty": 0.2, "similarity_threshold": 0.1}
    ra, rb = ContextRetriever(one, emb, cfg), ContextRetriever(many, emb, cfg)
    assert ra.retrieve("q0") == rb.retrieve("q0")
    assert ra.retrieve_batch(["q0", "q1"]) == rb.retrieve_batch(["q0", "q1"])
    assert not any(t for t, _ in many.collection.index.exchange_status())
    again = VectorStore({"collection_name": "many", "dtype": dtype, "devices": devices,
                         "persist_directory": str(tmp_path / "db")})
    assert again.get_stats()["count"] == n
    for q in qs:
        assert again.search(q, top_k=9) == one.search(q, top_k=9)


def test_device_tensors_from_the_embedder_never_touch_the_host():
    """N4: ``SentenceTransformer.encode(convert_to_tensor=True)`` hands CUDA tensors to create_index / search /
    retrieve_batch (reference rag/embedding.py:65-71, rag/indexing.py:116): same results as numpy input."""
    import torch
    rng = np.random.default_rng(32)
    n, dim = 3000, 384
    x = rng.standard_normal((n, dim)).astype(np.float32)
    texts, chunks = _texts_chunks(n)
    host = VectorStore({"collection_name": "host"})
    host.create_index(chunks, x)
    devs = VectorStore({"collection_name": "dev"})
    xd = torch.from_numpy(x).cuda()
    devs.create_index(chunks[:2000], xd[:2000])
    devs.create_index(chunks[1990:], xd[1990:])          # 10 ids already present: dropped on the device
    assert devs.get_stats()["count"] == n
    q = rng.standard_normal((5, dim)).astype(np.float32)
    for i in range(5):
        assert devs.search(torch.from_numpy(q[i]).cuda(), top_k=7) == host.search(q[i], top_k=7)
        assert devs.search(torch.from_numpy(q[i]).cuda(), top_k=7, where={"page_number": 4}) == \
            host.search(q[i], top_k=7, where={"page_number": 4})

    class CudaEmbedder:
        def embed(self, texts):
            if isinstance(texts, str):
                return torch.from_numpy(q[int(texts[1:])]).cuda()
            return torch.from_numpy(np.stack([q[int(t[1:])] for t in texts])).cuda()

    class HostEmbedder:
        def embed(self, texts):
            if isinstance(texts, str):
                return q[int(texts[1:])]
            return np.stack([q[int(t[1:])] for t in texts])

    cfg = {"top_k": 3, "rerank": True, "diversity_penalty": 0.1}
    a = ContextRetriever(devs, CudaEmbedder(), cfg)
    b = ContextRetriever(host, HostEmbedder(), cfg)
    names = [f"q{i}" for i in range(5)]
    assert a.retrieve_batch(names) == b.retrieve_batch(names)
    assert [a.retrieve(t) for t in names] == [b.retrieve(t) for t in names]


def test_search_input_forms_follow_the_reference():
    """rag/indexing.py:156-168: an ndarray of any shape is ONE query; a list of lists is a batch."""
    x = np.eye(8, dtype=np.float32)
    chunks = [Chunk(f"d{i}", f"chunk_{i}", 0, 2) for i in range(8)]
    vs = VectorStore({"collection_name": "forms"})
    vs.create_index(chunks, x)
    assert vs.search(x[2].reshape(1, 8), top_k=1)["ids"] == [["chunk_2"]]
    assert vs.search(x[2].reshape(2, 4), top_k=1)["ids"] == [["chunk_2"]]           # flattened, as the reference does
    assert vs.search(iter(x[5].tolist()), top_k=1)["ids"] == [["chunk_5"]]
    out = vs.search([x[1].tolist(), x[6].tolist()], top_k=2)                        # list of lists: two queries
    assert [r[0] for r in out["ids"]] == ["chunk_1", "chunk_6"] and len(out["distances"]) == 2
    d = vs.search(x[0], top_k=8)["distances"][0]
    assert all(v == float(np.float32(v)) for v in d)                                # float32 distances, like Chroma
    with pytest.raises(ValueError):
        vs.search(np.zeros(7, np.float32), top_k=1)


def test_more_candidates_than_a_search_returns_is_a_clear_error():
    rng = np.random.default_rng(33)
    x = rng.standard_normal((400, 64)).astype(np.float32)
    chunks = [Chunk(f"d{i}", f"chunk_{i}", 0, 2) for i in range(400)]
    vs = VectorStore({"collection_name": "caps"})
    vs.create_index(chunks, x)
    assert len(vs.search(x[0], top_k=112)["ids"][0]) == 112
    with pytest.raises(ValueError, match="exceeds"):
        vs.search(x[0], top_k=113)
    r = ContextRetriever(vs, TableEmbedder(), {"top_k": 60, "rerank": True})
    r.embedding_model.table["q"] = x[0]
    with pytest.raises(ValueError, match="exceeds"):
        r.retrieve("q")                                                             # fetches 2 * 60


def test_persistence_is_append_only_and_survives_torn_writes(tmp_path):
    from compressed_rag_suite_b200 import collection as backend
    rng = np.random.default_rng(34)
    n, dim = 900, 384
    x = rng.standard_normal((n, dim)).astype(np.float32)
    texts, chunks = _texts_chunks(n)
    d = str(tmp_path / "db")
    cfg = {"collection_name": "col", "persist_directory": d}
    a = VectorStore(cfg)
    sizes = []
    for lo, hi in [(0, 300), (300, 301), (301, 900)]:
        a.create_index(chunks[lo:hi], x[lo:hi])
        sizes.append((os.path.getsize(os.path.join(d, "col.crs")), os.path.getsize(os.path.join(d, "col.rows.jsonl"))))
    assert [s[0] for s in sizes] == [64 + 300 * 768, 64 + 301 * 768, 64 + 900 * 768]       # the blob only ever grows by the new rows
    assert sizes[0][1] < sizes[1][1] < sizes[2][1]
    assert not [f for f in os.listdir(d) if f.endswith(".tmp")]
    want = a.search(x[5], top_k=6)
    # (1) a crash in the middle of appending a sidecar line, and garbage behind the last complete row of the blob
    with open(os.path.join(d, "col.rows.jsonl"), "ab") as f:
        f.write(b'["chunk_900", "torn li')
    with open(os.path.join(d, "col.crs"), "ab") as f:
        f.write(b"\x01" * 100)
    b = VectorStore(cfg)
    assert b.get_stats()["count"] == n and b.search(x[5], top_k=6) == want
    extra = [Chunk("late", "chunk_late", 0, 4, page_number=1)]
    b.create_index(extra, x[:1] * -1.0)                                  # appending after the repair works
    assert VectorStore(cfg).get_stats()["count"] == n + 1
    # (2) the sidecar lost its last rows (the blob was flushed, the lines were not): the common prefix is kept
    lines = open(os.path.join(d, "col.rows.jsonl"), "rb").read().splitlines(keepends=True)
    with open(os.path.join(d, "col.rows.jsonl"), "wb") as f:
        f.writelines(lines[:850])
    c = VectorStore(cfg)
    assert c.get_stats()["count"] == 850
    assert c.search(x[880], top_k=1)["ids"] != [["chunk_880"]]
    c.create_index(chunks[850:], x[850:])
    assert c.get_stats()["count"] == n and c.search(x[880], top_k=1)["ids"] == [["chunk_880"]]
    assert c.search(x[5], top_k=6) == want
    assert VectorStore(cfg).search(x[5], top_k=6) == want
    # (3) a header that lies about the row count is clamped to the rows present; a corrupt one is an error, not an empty store
    import struct
    blob = os.path.join(d, "col.crs")
    raw = bytearray(open(blob, "rb").read())
    struct.pack_into("<q", raw, 24, 10 ** 12)
    open(blob, "wb").write(bytes(raw))
    assert VectorStore(cfg).get_stats()["count"] == n
    struct.pack_into("<q", raw, 24, -5)
    open(blob, "wb").write(bytes(raw))
    with pytest.raises(backend.CorruptStoreError):
        VectorStore(cfg)
    # (4) a ChromaDB directory is not readable: the store starts empty (and logs why)
    os.makedirs(str(tmp_path / "chroma"), exist_ok=True)
    open(str(tmp_path / "chroma" / "chroma.sqlite3"), "wb").close()
    assert VectorStore({"collection_name": "col", "persist_directory": str(tmp_path / "chroma")}).collection is None


def test_config1_pipeline_equals_the_reference_ragpipeline(golden_dir):
    """BASELINE configs[0] shape: 14 page-sized chunks, top_k 3, threshold 0.3, rerank, MMR 0.1.  The expected outputs
    come from the reference's OWN RAGPipeline (tests/golden/make_golden.py: pipeline_c1_cases) with the same stand-in
    embedder; here its VectorStore / ContextRetriever are replaced by this repo's."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import c1_standins as st
    g = json.load(open(os.path.join(golden_dir, "pipeline_c1_golden.json")))
    chunks = [Chunk(text=t, chunk_id=cid, start_char=0, end_char=len(t), **m)
              for t, cid, m in zip(g["chunk_texts"], g["chunk_ids"], g["chunk_metas"])]

    class Embedder:                                     # EmbeddingModel stand-in: rag/embedding.py:47-73
        model = st.HashSentenceTransformer()

        def embed(self, texts, show_progress=False):
            return self.model.encode([texts] if isinstance(texts, str) else texts)

    emb = Embedder()
    vs = VectorStore(g["config"]["vector_store"])
    vs.create_index(chunks, emb.embed([c.text for c in chunks]))
    r = ContextRetriever(vs, emb, g["config"]["retrieval"])
    assert len(g["cases"]) >= 15
    for case in g["cases"]:
        got = r.retrieve(case["query"])
        assert [c["chunk_id"] for c in got] == case["chunk_ids"], case["query"]
        assert [c["score"] for c in got] == case["scores"]
        assert [c["distance"] for c in got] == case["distances"]
        assert [c.get("rerank_score") for c in got] == case["rerank_scores"]
        assert [c["metadata"] for c in got] == case["metadatas"]
        assert r.get_context_string(case["query"]) == case["context"]
    batch = r.retrieve_batch([c["query"] for c in g["cases"]])
    assert [[c["chunk_id"] for c in one] for one in batch] == [c["chunk_ids"] for c in g["cases"]]
