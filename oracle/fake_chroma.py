"""Oracle: an in-memory stand-in for the ``chromadb`` surface the reference uses.

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.

``chromadb==1.3.0`` (``requirements.txt:19`` of the reference) is not vendored
and cannot be installed offline, so the engine half of the hot path is restated
here from Chroma's documented behaviour.  Only the calls the reference makes are
provided (``rag/indexing.py:33,36,50,81-84,114-119,171-176,186,206-207``;
``rag/retrieval.py:49-50``):

``Client(settings)``, ``PersistentClient(path)``, ``get_collection(name)``,
``create_collection(name, metadata)``, ``delete_collection(name)``;
``Collection.add(ids, embeddings, documents, metadatas)``,
``Collection.query(query_embeddings, n_results, where, where_document)``,
``Collection.count()``, ``Collection.metadata``.

Search is EXHAUSTIVE (Chroma's HNSW is approximate; at the reference's N ~ 14 it
is exact in practice).  Distances follow Chroma's convention for the collection's
``hnsw:space``: cosine ``1 - cos``, ip ``1 - dot``, l2 squared L2; results are
ascending distance, ties -> first inserted.  ``add`` of an id that already exists
is a no-op, not an upsert.

``precision`` selects the arithmetic:
* ``"f32"`` — what a real Chroma does: fp32 normalise + fp32 dot (numpy).
* ``"f16" | "bf16" | "i8" | "b1"`` — the canonical stored-code arithmetic of
  ``oracle/search.py``; this is the comparand for bit-exact parity with the CUDA
  backend configured with the same ``dtype``.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import numpy as np

from . import encode as _enc
from . import search as _srch

PRECISION = "f32"          # module-level default, tests set it per run


class Settings:            # chromadb.config.Settings(anonymized_telemetry=False)
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _match_where(meta: dict, where: Optional[dict]) -> bool:
    if not where:
        return True
    for key, cond in where.items():
        if key == "$and":
            if not all(_match_where(meta, w) for w in cond):
                return False
        elif key == "$or":
            if not any(_match_where(meta, w) for w in cond):
                return False
        else:
            have = key in meta
            val = meta.get(key)
            if not isinstance(cond, dict):
                cond = {"$eq": cond}
            for op, ref in cond.items():
                if op == "$eq":
                    ok = have and val == ref
                elif op == "$ne":
                    ok = (not have) or val != ref
                elif op == "$gt":
                    ok = have and val > ref
                elif op == "$gte":
                    ok = have and val >= ref
                elif op == "$lt":
                    ok = have and val < ref
                elif op == "$lte":
                    ok = have and val <= ref
                elif op == "$in":
                    ok = have and val in ref
                elif op == "$nin":
                    ok = (not have) or val not in ref
                else:
                    raise ValueError(f"unsupported where operator {op}")
                if not ok:
                    return False
    return True


def _match_doc(doc: str, cond: Optional[dict]) -> bool:
    if not cond:
        return True
    for op, ref in cond.items():
        if op == "$contains":
            if ref not in doc:
                return False
        elif op == "$not_contains":
            if ref in doc:
                return False
        elif op == "$and":
            if not all(_match_doc(doc, c) for c in ref):
                return False
        elif op == "$or":
            if not any(_match_doc(doc, c) for c in ref):
                return False
        else:
            raise ValueError(f"unsupported where_document operator {op}")
    return True


class Collection:
    def __init__(self, name: str, metadata: Optional[dict], precision: str):
        self.name = name
        self.metadata = dict(metadata) if metadata else None
        self._precision = precision
        self._ids: List[str] = []
        self._id_set = set()
        self._docs: List[str] = []
        self._metas: List[dict] = []
        self._emb: List[np.ndarray] = []          # original fp32 rows

    @property
    def _space(self) -> str:
        return (self.metadata or {}).get("hnsw:space", "l2")

    def count(self) -> int:
        return len(self._ids)

    def add(self, ids, embeddings, documents=None, metadatas=None):
        emb = np.asarray(embeddings, dtype=np.float32)
        if emb.ndim != 2 or emb.shape[0] != len(ids):
            raise ValueError("embeddings must be [len(ids), dim]")
        if self._emb and emb.shape[1] != self._emb[0].shape[0]:
            raise ValueError("embedding dimension mismatch")
        for i, cid in enumerate(ids):
            if cid in self._id_set:            # existing id: ignored, not upserted
                continue
            self._id_set.add(cid)
            self._ids.append(cid)
            self._docs.append(documents[i] if documents is not None else None)
            self._metas.append(dict(metadatas[i]) if metadatas is not None else None)
            self._emb.append(emb[i].copy())

    # -- distance arithmetic -------------------------------------------------
    def _distances(self, q: np.ndarray, rows: np.ndarray) -> np.ndarray:
        """One query against the selected rows -> python-float-compatible distances."""
        x = np.stack([self._emb[i] for i in rows]).astype(np.float32)
        space, prec = self._space, self._precision
        if space == "l2":
            if prec != "f32":
                raise ValueError("canonical stores support cosine / ip only")
            d = x - q[None, :]
            return np.einsum("ij,ij->i", d, d).astype(np.float32).astype(np.float64)
        metric = "cosine" if space == "cosine" else "ip"
        if prec == "f32":
            if metric == "cosine":
                xn = x / np.maximum(np.linalg.norm(x, axis=1, keepdims=True), np.float32(1e-30))
                qn = q / max(np.linalg.norm(q), np.float32(1e-30))
            else:
                xn, qn = x, q
            sims = (xn @ qn).astype(np.float32)
        else:
            dim = x.shape[1]
            codes = _enc.encode_rows(x, prec, metric)
            qc = _enc.encode_rows(q[None, :], prec, metric)[0]
            raw = _srch.raw_scores(codes, qc, prec, dim)
            sims = _srch.similarity_from_raw(raw, prec, dim)
            self._last_raw = raw
        # Chroma (hnswlib) computes the distance in float32 (1.0f - dot) and hands that float32 to Python
        return (np.float32(1.0) - sims.astype(np.float32)).astype(np.float64)

    def query(self, query_embeddings, n_results=10, where=None, where_document=None,
              include=None) -> Dict[str, Any]:
        qs = np.asarray(query_embeddings, dtype=np.float32)
        if qs.ndim == 1:
            qs = qs[None, :]
        rows = np.array([i for i in range(len(self._ids))
                         if _match_where(self._metas[i] or {}, where)
                         and _match_doc(self._docs[i] or "", where_document)], dtype=np.int64)
        out = {"ids": [], "documents": [], "metadatas": [], "distances": [],
               "embeddings": None, "uris": None, "data": None,
               "included": ["metadatas", "documents", "distances"]}
        for q in qs:
            if len(rows) == 0:
                sel, dist = [], []
            else:
                d = self._distances(q, rows)
                if self._precision in ("i8", "b1"):
                    # integer stores rank on the exact raw score, not on the rounded float
                    order = np.lexsort((rows, -self._last_raw.astype(np.float64)))[:n_results]
                else:
                    order = np.lexsort((rows, d))[:n_results]
                sel, dist = rows[order].tolist(), d[order].tolist()
            out["ids"].append([self._ids[i] for i in sel])
            out["documents"].append([self._docs[i] for i in sel])
            out["metadatas"].append([self._metas[i] for i in sel])
            out["distances"].append([float(v) for v in dist])
        return out


_PERSISTED: Dict[str, Dict[str, Collection]] = {}      # path -> collections (process-local)


class _ClientBase:
    def __init__(self, store: Dict[str, Collection]):
        self._collections = store

    def get_collection(self, name: str) -> Collection:
        if name not in self._collections:
            raise ValueError(f"Collection {name} does not exist.")
        return self._collections[name]

    def create_collection(self, name: str, metadata: Optional[dict] = None) -> Collection:
        if name in self._collections:
            raise ValueError(f"Collection {name} already exists.")
        col = Collection(name, metadata, PRECISION)
        self._collections[name] = col
        return col

    def delete_collection(self, name: str) -> None:
        if name not in self._collections:
            raise ValueError(f"Collection {name} does not exist.")
        del self._collections[name]


class Client(_ClientBase):
    def __init__(self, settings: Optional[Settings] = None):
        super().__init__({})


class PersistentClient(_ClientBase):
    def __init__(self, path: str = "./chroma"):
        super().__init__(_PERSISTED.setdefault(str(path), {}))
