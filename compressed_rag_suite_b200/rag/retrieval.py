"""ContextRetriever — same public surface as the reference's ``rag/retrieval.py`` (:13-277).

Pipeline kept in the reference's order (:110-160): embed the query, fetch ``2k`` hits when
``rerank`` else ``k``, convert each Chroma distance with ``_distance_to_similarity``
(:55-91, arithmetic verbatim), keep ``score >= similarity_threshold`` (:143), lexical
rerank when more than ``k`` survive else truncate (:151-154), then MMR when
``diversity_penalty > 0`` and more than one chunk is left (:157-158).  Result dicts carry
``text, score, distance, metadata, chunk_id`` (+ ``rerank_score``).

What runs where:
  * the N x D scoring, threshold and top-k: GPU (VectorStore.search -> libcrs); the
    threshold is also pushed down as a cosine-domain bound, the Python filter stays;
  * lexical rerank (:190-217): host, string work;
  * MMR (:219-277): GPU kernel K6 over the STORED vectors of the surviving chunks
    instead of re-embedding their texts.
Additive: ``retrieve_batch`` (many queries, one batched search + one MMR launch) and the
optional config key ``fetch_k`` (candidates fetched before MMR, default = reference).
"""
from __future__ import annotations

import logging
import math
from typing import Dict, List, Optional

import numpy as np

logger = logging.getLogger(__name__)


class ContextRetriever:
    def __init__(self, vector_store, embedding_model, config: dict):
        self.vector_store = vector_store
        self.embedding_model = embedding_model
        self.top_k = config.get("top_k", 3)
        self.similarity_threshold = config.get("similarity_threshold", 0.0)
        self.rerank = config.get("rerank", False)
        self.diversity_penalty = config.get("diversity_penalty", 0.0)
        self.fetch_k = config.get("fetch_k", None)
        self.distance_metric = self._get_distance_metric()
        logger.info(f"Using distance metric: {self.distance_metric}")

    def _get_distance_metric(self) -> str:
        try:
            if self.vector_store.collection:
                return self.vector_store.collection.metadata.get("hnsw:space", "cosine")
        except Exception:
            pass
        return "cosine"

    def _distance_to_similarity(self, distance: float) -> float:
        metric = self.distance_metric
        if metric == "cosine":
            d = max(0.0, min(2.0, distance))
            return max(0.0, min(1.0, 1.0 - (d * d / 2.0)))
        if metric == "l2":
            return 1.0 / (1.0 + distance)
        if metric == "ip":
            return max(0.0, min(1.0, (distance + 2.0) / 2.0))
        logger.warning(f"Unknown distance metric: {metric}, using default conversion")
        return max(0.0, 1.0 - (distance / 2.0))

    # cosine-domain lower bound equivalent to `score >= similarity_threshold`, slightly loose
    def _pushdown_similarity(self) -> Optional[float]:
        if self.distance_metric != "cosine" or not getattr(self.vector_store, "supports_min_similarity", False):
            return None
        t = self.similarity_threshold
        if t <= 0.0:
            return None
        if t > 1.0:
            return math.inf
        return 1.0 - math.sqrt(2.0 * (1.0 - t)) - 1e-6

    def _n_fetch(self, k: int) -> int:
        if self.fetch_k:
            return max(int(self.fetch_k), k)
        return k * 2 if self.rerank else k

    def _format(self, results: dict, col: int = 0) -> List[Dict]:
        kept = []
        ids = results["ids"][col]
        for i in range(len(ids)):
            distance = results["distances"][col][i]
            item = {
                "text": results["documents"][col][i],
                "score": self._distance_to_similarity(distance),
                "distance": distance,
                "metadata": results["metadatas"][col][i] if results["metadatas"] else {},
                "chunk_id": ids[i],
            }
            if item["score"] >= self.similarity_threshold:
                kept.append(item)
        return kept

    def _post(self, query: str, kept: List[Dict], k: int) -> List[Dict]:
        if self.rerank and len(kept) > k:
            return self._rerank(query, kept, k)
        return kept[:k]

    def retrieve(self, query: str, top_k: Optional[int] = None, filters: Optional[dict] = None) -> List[Dict]:
        k = top_k or self.top_k
        try:
            query_embedding = self.embedding_model.embed(query)
            extra = {}
            bound = self._pushdown_similarity()
            if bound is not None:
                extra["min_similarity"] = bound
            results = self.vector_store.search(query_embedding=query_embedding, top_k=self._n_fetch(k),
                                               where=filters, **extra)
            if not results["ids"][0]:
                logger.warning("No results found for query")
                return []
            kept = self._format(results)
            if not kept:
                logger.warning(f"No chunks passed similarity threshold of {self.similarity_threshold}")
                return []
            kept = self._post(query, kept, k)
            if self.diversity_penalty > 0 and len(kept) > 1:
                kept = self._apply_diversity(kept)
            return kept
        except Exception as e:
            logger.error(f"Retrieval failed: {e}")
            raise

    def retrieve_batch(self, queries: List[str], top_k: Optional[int] = None,
                       filters: Optional[dict] = None) -> List[List[Dict]]:
        """Many queries at once: one embedder call, one batched GPU search, one MMR launch."""
        k = top_k or self.top_k
        if not queries:
            return []
        try:
            emb = self.embedding_model.embed(list(queries))
            if not (type(emb).__module__.startswith("torch") and getattr(emb, "is_cuda", False)):
                emb = np.asarray(emb, dtype=np.float32)               # a CUDA tensor from the embedder stays on the device
            col = self.vector_store.collection
            if col is None:
                raise ValueError("No collection available. Create index first.")
            if col.count() == 0:
                return [[] for _ in queries]
            extra = {}
            bound = self._pushdown_similarity()
            if bound is not None:
                extra["min_similarity"] = bound
            results = col.query(query_embeddings=emb, n_results=min(self._n_fetch(k), col.count()),
                                where=filters, **extra)
            lists = [self._post(q, self._format(results, i), k) for i, q in enumerate(queries)]
            if self.diversity_penalty > 0:
                lists = self._apply_diversity_batch(lists)
            return lists
        except Exception as e:
            logger.error(f"Batch retrieval failed: {e}")
            raise

    def get_context_string(self, query: str, top_k: Optional[int] = None, separator: str = "\n\n") -> str:
        chunks = self.retrieve(query, top_k=top_k)
        return separator.join(c["text"] for c in chunks) if chunks else ""

    def _rerank(self, query: str, chunks: List[Dict], top_k: int) -> List[Dict]:
        q_tokens = set(query.lower().split())
        denom = max(len(q_tokens), 1)
        for chunk in chunks:
            overlap = len(q_tokens & set(chunk["text"].lower().split()))
            chunk["rerank_score"] = chunk["score"] * 0.7 + (overlap / denom) * 0.3
        chunks.sort(key=lambda c: c.get("rerank_score", c["score"]), reverse=True)
        return chunks[:top_k]

    # ------------------------------------------------------------------ MMR on the GPU
    def _gpu_collection(self):
        col = getattr(self.vector_store, "collection", None)
        if col is None or not hasattr(col, "stored_vectors") or col.index is None:
            raise TypeError("MMR needs the GPU VectorStore of compressed_rag_suite_b200 (stored vectors); "
                            "there is no CPU path")
        return col

    def _apply_diversity(self, chunks: List[Dict]) -> List[Dict]:
        if len(chunks) <= 1:
            return chunks
        return self._apply_diversity_batch([chunks])[0]

    def _apply_diversity_batch(self, lists: List[List[Dict]]) -> List[List[Dict]]:
        todo = [i for i, l in enumerate(lists) if len(l) > 1]
        if not todo:
            return lists
        col = self._gpu_collection()
        lam = 1.0 - self.diversity_penalty
        m = max(len(lists[i]) for i in todo)
        rb = col.index.row_bytes
        vecs = np.zeros((len(todo), m, rb), dtype=np.uint8)
        # padding candidates (zero vector, relevance -inf) can never be picked: their mmr
        # score is -inf and the selection needs a strict improvement over -inf
        rel = np.full((len(todo), m), -np.inf, dtype=np.float64)
        for j, i in enumerate(todo):
            ids = [c["chunk_id"] for c in lists[i]]
            vecs[j, :len(ids)] = col.stored_vectors(ids)
            rel[j, :len(ids)] = [c["score"] for c in lists[i]]
        order = col.index.mmr(vecs, rel, lam)
        out = list(lists)
        for j, i in enumerate(todo):
            out[i] = [lists[i][p] for p in order[j] if p >= 0]
        return out
