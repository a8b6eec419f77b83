"""Generate tests/golden/*.npz + *.json from the reference's OWN code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference's unmodified ``rag/indexing.py`` + ``rag/retrieval.py`` are imported through
``oracle/reference_loader.py`` (third-party imports stubbed; the Chroma arithmetic is
``oracle/fake_chroma.py`` in canonical f16 precision, because chromadb itself cannot be
installed offline).  Inputs are stored next to the outputs, so the fixtures do not depend
on a random generator's stream.

Fixtures
  retrieval_golden.npz/.json  end-to-end ``ContextRetriever.retrieve`` outputs for several
                              configs.  For MMR the stub embedder hands back the stored
                              fp16 rows (as fp32), and a case is only kept when the
                              reference (BLAS-order fp32 sims) and the canonical restatement
                              (exactly rounded sims) agree — i.e. no MMR decision sits on an
                              fp32 summation-order tie.
  mmr_dyadic_golden.json      ``_apply_diversity`` on dyadic vectors, where every fp32 sum is
                              exact in any order, so the reference output is the bit-exact
                              expectation for the CUDA MMR kernel (incl. the numpy-2 scalar
                              typing paths).
  transform_golden.json       ``_distance_to_similarity`` and ``_rerank`` outputs.
  pipeline_c1_golden.json     BASELINE config 1 shape: the reference's own ``RAGPipeline`` (rag/pipeline.py:85-163)
                              — DocumentProcessor.process_string, TextChunker (semantic, code defaults 512/50),
                              EmbeddingModel, VectorStore, ContextRetriever with config.json's retrieval keys
                              (top_k 3, threshold 0.3, rerank, MMR 0.1) — over 14 synthetic "pages" and 20 queries,
                              with stand-ins only for what cannot exist offline (tests/c1_standins.py: a hash-seeded
                              SentenceTransformer, a regex punkt tokenizer).  Holds the chunks the reference's
                              chunker produced and what ``RAGPipeline.retrieve`` returned.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import encode, postprocess  # noqa: E402
from oracle.reference_loader import load_reference  # noqa: E402

WORDS = ("model compression quantization pruning distillation low rank factorization survey "
         "llm inference memory latency accuracy benchmark dataset metric training sparse dense "
         "weights activations calibration perplexity").split()


class TableEmbedder:
    def __init__(self):
        self.table = {}

    def embed(self, texts):
        if isinstance(texts, str):
            texts = [texts]
        return np.stack([self.table[t] for t in texts]).astype(np.float32)


def make_corpus(rng, n, dim, n_clusters=6):
    centres = rng.standard_normal((n_clusters, dim)).astype(np.float32)
    centres /= np.linalg.norm(centres, axis=1, keepdims=True)
    x = np.empty((n, dim), dtype=np.float32)
    for i in range(n):
        noise = rng.standard_normal(dim).astype(np.float32)
        noise /= np.linalg.norm(noise)
        v = 0.6 * centres[i % n_clusters] + 0.8 * noise
        x[i] = v / np.linalg.norm(v)
    return x, centres


def retrieval_cases():
    ri, rr, rc, fc = load_reference()
    fc.PRECISION = "f16"
    rng = np.random.default_rng(20240607)
    n, dim = 96, 384
    x, centres = make_corpus(rng, n, dim)
    x[17] = x[5]                       # exact duplicates: ties -> lowest id
    x[60] = x[5]
    texts = []
    for i in range(n):
        w = rng.choice(WORDS, size=12)
        texts.append(" ".join(w) + f" passage{i}")
    queries, qtexts = [], []
    for j in range(12):
        noise = rng.standard_normal(dim).astype(np.float32)
        noise /= np.linalg.norm(noise)
        q = 0.7 * centres[j % len(centres)] + 0.7 * noise
        if j == 3:
            q = x[5].copy()            # query equal to a (triplicated) corpus row
        if j == 7:
            q = 3.5 * q                # un-normalised query: cosine must not care
        queries.append(q.astype(np.float32))
        qtexts.append(" ".join(rng.choice(WORDS, size=5)) + f" question{j}")
    queries = np.stack(queries)

    stored = encode.encode_rows(x, "f16", "cosine")[:, :dim].astype(np.float32)
    emb = TableEmbedder()
    for t, v in zip(texts, stored):
        emb.table[t] = v               # MMR "re-embedding" returns the stored rows
    for t, v in zip(qtexts, queries):
        emb.table[t] = v

    chunks = [rc.Chunk(text=t, chunk_id=f"chunk_{i}", start_char=0, end_char=len(t),
                       page_number=(i % 7) + 1, section=None if i % 3 else f"sec{i % 5}", tokens=12)
              for i, t in enumerate(texts)]
    configs = [
        {"top_k": 3, "similarity_threshold": 0.3, "rerank": True, "diversity_penalty": 0.1},   # config.json:20-25
        {"top_k": 3},                                                                          # code defaults
        {"top_k": 5, "similarity_threshold": 0.0, "rerank": False, "diversity_penalty": 0.5},
        {"top_k": 10, "similarity_threshold": 0.3, "rerank": True, "diversity_penalty": 0.3},
        {"top_k": 4, "similarity_threshold": 0.75, "rerank": True, "diversity_penalty": 0.1},  # count < k
        {"top_k": 3, "similarity_threshold": 0.995, "rerank": False, "diversity_penalty": 0.0},
        {"top_k": 1, "similarity_threshold": 0.0, "rerank": True, "diversity_penalty": 0.9},
    ]
    by_id = {c.chunk_id: stored[i] for i, c in enumerate(chunks)}
    cases = []
    for ci, cfg in enumerate(configs):
        vs = ri.VectorStore({"collection_name": f"golden_{ci}"})
        vs.create_index(chunks, x)
        ref = rr.ContextRetriever(vs, emb, cfg)
        orc = postprocess.OracleRetriever(vs, emb, cfg, lambda ids: np.stack([by_id[i] for i in ids]))
        for qi, qt in enumerate(qtexts):
            got = ref.retrieve(qt)
            want = orc.retrieve(qt)
            agree = [g["chunk_id"] for g in got] == [w["chunk_id"] for w in want]
            if not agree:
                print(f"  skip cfg{ci} q{qi}: reference MMR sits on an fp32 summation-order tie")
                continue
            assert got == want, "oracle restatement differs from the reference"
            cases.append({"config": cfg, "query": qi,
                          "chunk_ids": [g["chunk_id"] for g in got],
                          "scores": [g["score"] for g in got],
                          "distances": [g["distance"] for g in got],
                          "rerank_scores": [g.get("rerank_score") for g in got],
                          "metadatas": [g["metadata"] for g in got]})
        stats = vs.get_stats()
        assert stats["count"] == n and stats["metadata"] == {"hnsw:space": "cosine"}
    np.savez_compressed(os.path.join(HERE, "retrieval_golden.npz"), embeddings=x, queries=queries)
    with open(os.path.join(HERE, "retrieval_golden.json"), "w") as f:
        json.dump({"texts": texts, "query_texts": qtexts, "cases": cases,
                   "chunk_meta": [{"page_number": c.page_number, "section": c.section, "tokens": c.tokens} for c in chunks]},
                  f, indent=0)
    print(f"retrieval_golden: {len(cases)} cases")


def mmr_dyadic_cases():
    ri, rr, rc, fc = load_reference()
    rng = np.random.default_rng(7)
    cases = []

    class Stub:                         # just enough `self` for ContextRetriever._apply_diversity
        def __init__(self, penalty, emb):
            self.diversity_penalty = penalty
            self.embedding_model = emb

    for m, dim, penalty, mode in [(2, 64, 0.1, "mixed"), (5, 64, 0.1, "mixed"), (16, 64, 0.5, "mixed"),
                                  (40, 64, 0.1, "mixed"), (12, 128, 1.0, "mixed"), (12, 128, 0.0001, "mixed"),
                                  (9, 64, 0.3, "negative"), (24, 64, 0.7, "positive"), (7, 64, 0.25, "dup"),
                                  (100, 64, 0.1, "mixed"), (33, 192, 0.45, "mixed")]:
        for rep in range(3):
            v = rng.integers(-32, 33, size=(m, dim)).astype(np.float32) / 64.0     # dyadic, exact in fp16
            if mode == "negative":      # all pairwise sims <= 0: max_sim stays the Python 0.0 (fp64 path)
                v = np.zeros((m, dim), dtype=np.float32)
                for i in range(m):
                    v[i, i * 2] = 0.5
                    v[i, i * 2 + 1] = rng.integers(1, 8) / 64.0
                v[1:, 0] = -0.25
            if mode == "positive":
                v = np.abs(v) + 1.0 / 64.0
            if mode == "dup":
                v[3] = v[1]
                v[5] = v[1]
            v[:, 0] += (v == 0).all(axis=1) * 0.5                                   # no zero rows
            rel = [float(r) for r in rng.uniform(0.2, 1.0, size=m)]
            if mode == "dup":
                rel[3] = rel[1]
                rel[5] = rel[1]
            emb = TableEmbedder()
            chunks = []
            for i in range(m):
                t = f"t{i}"
                emb.table[t] = v[i]
                chunks.append({"text": t, "score": rel[i], "chunk_id": f"c{i}"})
            out = rr.ContextRetriever._apply_diversity(Stub(penalty, emb), list(chunks))
            order = [int(c["chunk_id"][1:]) for c in out]
            want = postprocess.mmr_order(rel, postprocess.pairwise_sims_f32(v), 1.0 - penalty)
            assert order == want, (m, dim, penalty, mode, order, want)
            cases.append({"m": m, "dim": dim, "penalty": penalty, "mode": mode,
                          "vectors_x64": (v * 64).astype(np.int32).tolist(), "relevance": rel, "order": order})
    with open(os.path.join(HERE, "mmr_dyadic_golden.json"), "w") as f:
        json.dump(cases, f)
    print(f"mmr_dyadic_golden: {len(cases)} cases")


def transform_cases():
    ri, rr, rc, fc = load_reference()

    class Stub:
        def __init__(self, metric):
            self.distance_metric = metric

    out = {"distance_to_similarity": [], "rerank": []}
    ds = [-0.5, 0.0, 1e-9, 0.004, 0.1, 0.3, 0.64, 0.999999, 1.0, 1.1832, 1.5, 2.0, 2.5, 1 - 0.3600001]
    for metric in ("cosine", "l2", "ip", "weird"):
        for d in ds:
            out["distance_to_similarity"].append(
                {"metric": metric, "distance": d,
                 "score": rr.ContextRetriever._distance_to_similarity(Stub(metric), d)})
    rng = np.random.default_rng(3)
    for rep in range(6):
        query = " ".join(rng.choice(WORDS, size=rng.integers(1, 7))) if rep else "   "
        chunks = [{"text": " ".join(rng.choice(WORDS, size=10)).title(), "score": float(s), "chunk_id": f"c{i}"}
                  for i, s in enumerate(np.round(rng.uniform(0.3, 0.9, size=8), 2))]   # rounded: forces ties
        got = rr.ContextRetriever._rerank(Stub("cosine"), query, [dict(c) for c in chunks], 4)
        out["rerank"].append({"query": query, "chunks": chunks, "top_k": 4,
                              "order": [g["chunk_id"] for g in got],
                              "rerank_scores": [g["rerank_score"] for g in got]})
    with open(os.path.join(HERE, "transform_golden.json"), "w") as f:
        json.dump(out, f)
    print("transform_golden: ok")


def pipeline_c1_cases():
    """BASELINE configs[0] through the reference's RAGPipeline (see the module docstring)."""
    import types
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import c1_standins as st
    ri, rr, rc, fc = load_reference()
    fc.PRECISION = "f16"
    import rag.embedding as ref_embedding
    import rag.chunking as ref_chunking
    import rag.pipeline as ref_pipeline
    stored_of = {}

    class PipelineEmbedder(st.HashSentenceTransformer):
        """Indexing and queries get the hash vectors; the MMR step's "re-embedding" of chunk texts gets the rows the
        store holds for them (canonical fp16 codes as fp32), exactly like the retrieval golden does."""

        def encode(self, texts, **kw):
            out = super().encode(texts, **kw)
            for i, t in enumerate([texts] if isinstance(texts, str) else list(texts)):
                if t in stored_of:
                    out[i] = stored_of[t]
            return out

    ref_embedding.SentenceTransformer = PipelineEmbedder
    ref_chunking.nltk = types.SimpleNamespace(data=types.SimpleNamespace(load=lambda _p: st.RegexPunkt(), find=lambda _p: True),
                                              download=lambda *a, **k: True)
    cfg = {"chunking": {"strategy": "semantic"},                                   # code defaults 512 / 50 / 100
           "embedding": {"device": "cpu"},
           "vector_store": {"collection_name": "c1_golden"},
           "retrieval": {"top_k": 3, "similarity_threshold": 0.3, "rerank": True, "diversity_penalty": 0.1}}   # config.json:20-25
    pipe = ref_pipeline.RAGPipeline(cfg)
    pipe.setup(model_interface=types.SimpleNamespace(model_type="instruct"))     # the generator is never called
    pages = st.make_pages()
    # index once to learn the chunk texts, then register their stored rows for the MMR step
    pipe.index_documents(pages, show_progress=False)
    col = pipe.vector_store.collection
    texts = list(col._docs)
    x = np.stack(col._emb).astype(np.float32)
    stored = encode.encode_rows(x, "f16", "cosine")[:, :x.shape[1]].astype(np.float32)
    for t, v in zip(texts, stored):
        stored_of[t] = v
    queries = st.make_queries()
    by_id = {cid: stored[i] for i, cid in enumerate(col._ids)}
    orc = postprocess.OracleRetriever(pipe.vector_store, pipe.embedding_model, cfg["retrieval"],
                                      lambda ids: np.stack([by_id[i] for i in ids]))
    cases = []
    for q in queries:
        got = pipe.retrieve(q)
        want = orc.retrieve(q)
        if [g["chunk_id"] for g in got] != [w["chunk_id"] for w in want]:
            print(f"  skip {q!r}: reference MMR sits on an fp32 summation-order tie")
            continue
        assert got == want, "oracle restatement differs from the reference pipeline"
        cases.append({"query": q, "chunk_ids": [g["chunk_id"] for g in got], "scores": [g["score"] for g in got],
                      "distances": [g["distance"] for g in got], "rerank_scores": [g.get("rerank_score") for g in got],
                      "metadatas": [g["metadata"] for g in got], "context": pipe.retriever.get_context_string(q)})
    with open(os.path.join(HERE, "pipeline_c1_golden.json"), "w") as f:
        json.dump({"config": cfg, "n_pages": len(pages), "chunk_ids": list(col._ids), "chunk_texts": texts,
                   "chunk_metas": list(col._metas), "cases": cases}, f, indent=0)
    print(f"pipeline_c1_golden: {len(col._ids)} chunks from {len(pages)} pages, {len(cases)} of {len(queries)} queries kept")


if __name__ == "__main__":
    pipeline_c1_cases()
    retrieval_cases()
    mmr_dyadic_cases()
    transform_cases()
