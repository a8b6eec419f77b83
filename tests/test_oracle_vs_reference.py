"""CPU tier, build container only: the oracle restatement against the reference's own
UNMODIFIED rag/indexing.py + rag/retrieval.py (imported from /root/reference).  Skipped
where the reference tree is absent (the GPU box)."""
import json
import os

import numpy as np
import pytest

from oracle import encode, postprocess
from oracle.reference_loader import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


class TableEmbedder:
    def __init__(self):
        self.table = {}

    def embed(self, texts):
        if isinstance(texts, str):
            texts = [texts]
        return np.stack([self.table[t] for t in texts]).astype(np.float32)


@pytest.fixture(scope="module")
def ref():
    return load_reference()


def test_golden_files_are_current(ref, golden_dir):
    """Regenerating a golden case from the live reference gives the committed values."""
    ri, rr, rc, fc = ref
    fc.PRECISION = "f16"
    g = json.load(open(os.path.join(golden_dir, "retrieval_golden.json")))
    arr = np.load(os.path.join(golden_dir, "retrieval_golden.npz"))
    x, queries = arr["embeddings"], arr["queries"]
    stored = encode.encode_rows(x, "f16")[:, :x.shape[1]].astype(np.float32)
    emb = TableEmbedder()
    for t, v in zip(g["texts"], stored):
        emb.table[t] = v
    for t, v in zip(g["query_texts"], queries):
        emb.table[t] = v
    chunks = [rc.Chunk(text=t, chunk_id=f"chunk_{i}", start_char=0, end_char=len(t), **g["chunk_meta"][i])
              for i, t in enumerate(g["texts"])]
    stores = {}
    for n, case in enumerate(g["cases"][::5]):
        key = json.dumps(case["config"], sort_keys=True)
        if key not in stores:
            vs = ri.VectorStore({"collection_name": f"regen_{len(stores)}_{os.getpid()}"})
            vs.create_index(chunks, x)
            stores[key] = vs
        r = rr.ContextRetriever(stores[key], emb, case["config"])
        got = r.retrieve(g["query_texts"][case["query"]])
        assert [c["chunk_id"] for c in got] == case["chunk_ids"]
        assert [c["score"] for c in got] == case["scores"]
        assert [c["distance"] for c in got] == case["distances"]


def test_distance_transform_matches_reference(ref):
    _, rr, _, _ = ref

    class S:
        pass

    rng = np.random.default_rng(0)
    for metric in ("cosine", "l2", "ip", "other"):
        s = S()
        s.distance_metric = metric
        for d in np.concatenate([rng.uniform(-1, 3, 300), [0.0, 2.0, 1.0]]):
            assert postprocess.distance_to_similarity(float(d), metric) == \
                rr.ContextRetriever._distance_to_similarity(s, float(d))


def test_mmr_restatement_matches_reference_on_random_dyadic(ref):
    _, rr, _, _ = ref
    rng = np.random.default_rng(11)

    class S:
        pass

    for trial in range(60):
        m = int(rng.integers(2, 30))
        dim = int(rng.choice([16, 64, 96]))
        v = rng.integers(-16, 17, size=(m, dim)).astype(np.float32) / 32.0
        v[:, 0] += (v == 0).all(axis=1)
        rel = [float(r) for r in rng.uniform(0, 1, m)]
        emb = TableEmbedder()
        for i in range(m):
            emb.table[f"t{i}"] = v[i]
        s = S()
        s.diversity_penalty = float(rng.choice([0.1, 0.5, 0.9, 1.0, 0.01]))
        s.embedding_model = emb
        out = rr.ContextRetriever._apply_diversity(s, [{"text": f"t{i}", "score": rel[i], "chunk_id": i} for i in range(m)])
        want = postprocess.mmr_order(rel, postprocess.pairwise_sims_f32(v), 1.0 - s.diversity_penalty)
        assert [c["chunk_id"] for c in out] == want


def test_vectorstore_contract_on_fake_chroma(ref):
    """Pins the dict shapes / guards the product VectorStore must reproduce (rag/indexing.py)."""
    ri, _, rc, fc = ref
    fc.PRECISION = "f32"
    vs = ri.VectorStore({"collection_name": f"contract_{os.getpid()}"})
    assert vs.collection is None and vs.get_stats() == {"status": "empty", "count": 0}
    with pytest.raises(ValueError):
        vs.search(np.zeros(4, np.float32))
    vs.create_index([], np.zeros((0, 4), np.float32))
    assert vs.collection is None
    with pytest.raises(ValueError):
        vs.create_index([rc.Chunk("a", "chunk_0", 0, 1)], np.zeros((2, 4), np.float32))
    x = np.eye(4, dtype=np.float32)
    chunks = [rc.Chunk(f"doc {i}", f"chunk_{i}", 0, 5, page_number=i, section=None, tokens=2) for i in range(4)]
    vs.create_index(chunks, x)
    vs.create_index(chunks[:2], x[:2])                  # same ids again: no-op
    assert vs.get_stats()["count"] == 4
    out = vs.search(np.array([[1.0, 0.1, 0, 0]], np.float32), top_k=10)
    assert out["ids"] == [["chunk_0", "chunk_1", "chunk_2", "chunk_3"]]
    assert out["metadatas"][0][0] == {"page_number": 0, "tokens": 2}
    assert out["distances"][0] == sorted(out["distances"][0])


def test_config1_golden_is_what_the_reference_ragpipeline_returns(ref, golden_dir):
    """tests/golden/pipeline_c1_golden.json: re-run the reference's OWN RAGPipeline (index_documents on the 14 synthetic
    pages, retrieve on the queries) with the stand-in embedder / tokenizer and compare with the committed file — the
    chunks its chunker makes and the contexts it returns."""
    import sys
    import types
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import c1_standins as st
    ri, rr, rc, fc = ref
    fc.PRECISION = "f16"
    import rag.chunking as ref_chunking
    import rag.embedding as ref_embedding
    import rag.pipeline as ref_pipeline
    g = json.load(open(os.path.join(golden_dir, "pipeline_c1_golden.json")))
    stored_of = {}

    class PipelineEmbedder(st.HashSentenceTransformer):
        def encode(self, texts, **kw):
            out = super().encode(texts, **kw)
            for i, t in enumerate([texts] if isinstance(texts, str) else list(texts)):
                if t in stored_of:
                    out[i] = stored_of[t]
            return out

    ref_embedding.SentenceTransformer = PipelineEmbedder
    ref_chunking.nltk = types.SimpleNamespace(data=types.SimpleNamespace(load=lambda _p: st.RegexPunkt(), find=lambda _p: True),
                                              download=lambda *a, **k: True)
    cfg = json.loads(json.dumps(g["config"]))
    cfg["vector_store"]["collection_name"] = f"c1_regen_{os.getpid()}"
    pipe = ref_pipeline.RAGPipeline(cfg)
    pipe.setup(model_interface=types.SimpleNamespace(model_type="instruct"))
    pipe.index_documents(st.make_pages(g["n_pages"]), show_progress=False)
    col = pipe.vector_store.collection
    assert list(col._ids) == g["chunk_ids"] and list(col._docs) == g["chunk_texts"] and list(col._metas) == g["chunk_metas"]
    x = np.stack(col._emb).astype(np.float32)
    stored = encode.encode_rows(x, "f16", "cosine")[:, :x.shape[1]].astype(np.float32)
    for t, v in zip(g["chunk_texts"], stored):
        stored_of[t] = v
    assert len(g["cases"]) >= 15
    for case in g["cases"]:
        got = pipe.retrieve(case["query"])
        assert [c["chunk_id"] for c in got] == case["chunk_ids"]
        assert [c["score"] for c in got] == case["scores"]
        assert [c["distance"] for c in got] == case["distances"]
        assert [c.get("rerank_score") for c in got] == case["rerank_scores"]
        assert pipe.retriever.get_context_string(case["query"]) == case["context"]
