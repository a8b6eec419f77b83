"""Oracle: restatement of ``rag/retrieval.py`` post-processing (a6-a12 in SURVEY.md §8a).

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.

Each function cites the reference lines it follows.  ``tests/test_oracle_vs_reference.py``
checks every one of them against the reference's own unmodified module (imported
through ``oracle/reference_loader.py``) and ``tests/golden/`` holds vectors
generated from that module, so this half of the oracle is pinned.

numpy-2 typing (NEP 50) is part of the reference's behaviour inside MMR: ``sim``
is an ``np.float32`` scalar, ``relevance`` a Python float, so
``lambda*rel - (1-lambda)*max_sim`` is evaluated in fp32 once ``max_sim`` has
become an ``np.float32`` (any positive sim) and in fp64 while it is still the
Python ``0.0``; mixed comparisons round the Python float to fp32.  The code
below keeps the same scalar types so it inherits exactly that behaviour.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from .encode import decode_rows


def distance_to_similarity(distance: float, metric: str = "cosine") -> float:
    """``rag/retrieval.py:55-91`` verbatim arithmetic (Python floats)."""
    if metric == "cosine":
        distance = max(0.0, min(2.0, distance))                 # :75
        cosine_sim = 1.0 - (distance * distance / 2.0)          # :76
        return max(0.0, min(1.0, cosine_sim))                   # :77
    if metric == "l2":
        return 1.0 / (1.0 + distance)                           # :82
    if metric == "ip":
        return max(0.0, min(1.0, (distance + 2.0) / 2.0))       # :87
    return max(0.0, 1.0 - (distance / 2.0))                     # :91


def min_cosine_for_threshold(threshold: float) -> float:
    """Push-down bound: smallest cosine whose transformed score can reach
    ``threshold`` under the cosine branch (:75-77): score = 1 - (1-cos)^2/2 for
    cos in [-1, 1].  Conservative (slightly low); the exact Python filter at
    ``rag/retrieval.py:143`` is still applied afterwards."""
    if threshold <= 0.0:
        return -np.inf                      # score is clamped to >= 0: filter is a no-op
    if threshold > 1.0:
        return np.inf                       # nothing can pass
    return 1.0 - np.sqrt(2.0 * (1.0 - threshold)) - 1e-6


def rerank(query: str, chunks: List[Dict], top_k: int) -> List[Dict]:
    """``rag/retrieval.py:190-217``: 0.7*score + 0.3*token-overlap, stable sort desc."""
    query_tokens = set(query.lower().split())                   # :202
    for chunk in chunks:
        chunk_tokens = set(chunk["text"].lower().split())       # :205
        overlap = len(query_tokens & chunk_tokens)              # :208
        overlap_score = overlap / max(len(query_tokens), 1)     # :209
        chunk["rerank_score"] = chunk["score"] * 0.7 + overlap_score * 0.3   # :213
    chunks.sort(key=lambda x: x.get("rerank_score", x["score"]), reverse=True)   # :216
    return chunks[:top_k]


def pairwise_sims_f32(vecs: np.ndarray) -> np.ndarray:
    """Canonical candidate-candidate cosine used by MMR, fp32 [m, m].

    Follows ``rag/retrieval.py:258-260`` (``np.dot(a,b) / (norm(a)*norm(b))`` on
    fp32 arrays, i.e. fp32 sqrt / mul / div) with one difference that makes it
    reproducible: the two inner products are the exactly rounded
    ``fl32(sum in fp64, sequential)`` instead of BLAS-order fp32 sums.  For inputs
    whose fp32 partial sums are exact (tests use dyadic vectors) this equals the
    reference bit for bit; otherwise it differs by summation-order noise only."""
    v = np.asarray(vecs, dtype=np.float64)
    m = v.shape[0]
    dots = np.zeros((m, m), dtype=np.float64)
    for j in range(v.shape[1]):
        dots += np.outer(v[:, j], v[:, j])
    dots32 = dots.astype(np.float32)
    norms = np.sqrt(np.diag(dots32).astype(np.float32))         # fp32 sqrt, correctly rounded
    with np.errstate(divide="ignore", invalid="ignore"):
        den = (norms[:, None] * norms[None, :]).astype(np.float32)
        return (dots32 / den).astype(np.float32)


def mmr_order(relevance: Sequence[float], sims: np.ndarray, lambda_param: float,
              k_out: Optional[int] = None) -> List[int]:
    """Greedy MMR order over positions 0..m-1 (``rag/retrieval.py:241-275``).

    ``relevance`` are the chunks' ``score`` values (Python floats, :252), ``sims``
    the fp32 candidate-candidate cosines.  The reference always runs to
    ``k_out = m`` (it never drops items, :246); ``k_out < m`` returns the prefix of
    the same greedy order (the optional ``fetch_k`` extension)."""
    m = len(relevance)
    if k_out is None:
        k_out = m
    sims = np.asarray(sims, dtype=np.float32)
    selected = [0]                                              # :242
    remaining = list(range(1, m))                               # :244
    while len(selected) < min(k_out, m) and remaining:          # :246
        best_idx = None
        best_score = -float("inf")                              # :248
        for idx in remaining:                                   # :250
            rel = relevance[idx]                                # :252
            max_sim = 0.0                                       # :255
            for s in selected:                                  # :256
                sim = sims[idx, s]                              # np.float32 scalar, as at :258
                max_sim = max(max_sim, sim)                     # :261
            mmr_score = lambda_param * rel - (1 - lambda_param) * max_sim   # :264
            if mmr_score > best_score:                          # :266
                best_score = mmr_score
                best_idx = idx
        if best_idx is not None:                                # :270
            selected.append(best_idx)
            remaining.remove(best_idx)
        else:
            break
    return selected


class OracleRetriever:
    """Restatement of ``ContextRetriever`` (``rag/retrieval.py:19-188``) on top of any
    object with the reference ``VectorStore.search`` signature.  ``row_vectors(chunk_ids)``
    returns the stored vectors MMR compares (instead of re-embedding the texts, :238-239)."""

    def __init__(self, vector_store, embedding_model, config: dict, row_vectors):
        self.vector_store = vector_store
        self.embedding_model = embedding_model
        self.top_k = config.get("top_k", 3)                               # :36
        self.similarity_threshold = config.get("similarity_threshold", 0.0)   # :37
        self.rerank = config.get("rerank", False)                         # :38
        self.diversity_penalty = config.get("diversity_penalty", 0.0)     # :39
        self.distance_metric = "cosine"
        try:                                                              # :45-53
            if self.vector_store.collection:
                self.distance_metric = self.vector_store.collection.metadata.get("hnsw:space", "cosine")
        except Exception:
            pass
        self._row_vectors = row_vectors

    def retrieve(self, query: str, top_k: Optional[int] = None, filters: Optional[dict] = None):
        k = top_k or self.top_k                                           # :110
        q = self.embedding_model.embed(query)                             # :114
        results = self.vector_store.search(query_embedding=q,
                                           top_k=k * 2 if self.rerank else k,
                                           where=filters)                 # :117-121
        if not results["ids"][0]:                                         # :124
            return []
        kept = []
        for i in range(len(results["ids"][0])):                           # :130
            distance = results["distances"][0][i]
            score = distance_to_similarity(distance, self.distance_metric)
            item = {"text": results["documents"][0][i], "score": score, "distance": distance,
                    "metadata": results["metadatas"][0][i] if results["metadatas"] else {},
                    "chunk_id": results["ids"][0][i]}                     # :134-140
            if item["score"] >= self.similarity_threshold:                # :143
                kept.append(item)
        if not kept:                                                      # :146
            return []
        if self.rerank and len(kept) > k:                                 # :151
            kept = rerank(query, kept, k)
        else:
            kept = kept[:k]                                               # :154
        if self.diversity_penalty > 0 and len(kept) > 1:                  # :157
            lam = 1.0 - self.diversity_penalty                            # :235
            vecs = self._row_vectors([c["chunk_id"] for c in kept])
            order = mmr_order([c["score"] for c in kept], pairwise_sims_f32(vecs), lam)
            kept = [kept[i] for i in order]
        return kept
