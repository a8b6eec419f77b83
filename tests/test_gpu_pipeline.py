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
