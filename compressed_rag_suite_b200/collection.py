"""The Chroma-shaped surface the reference talks to, backed by libcrs on one B200.

``rag/indexing.py`` of the reference uses exactly these calls (SURVEY.md §1 L1★):
``Client(settings)`` / ``PersistentClient(path)``, ``get_collection``,
``create_collection(name, metadata)``, ``delete_collection``; ``Collection.add``,
``.query``, ``.count()``, ``.metadata``.  Providing them means either the mirror
``compressed_rag_suite_b200.rag.indexing.VectorStore`` or the reference's own
unmodified ``rag/indexing.py`` (with ``sys.modules["chromadb"]`` pointed here, see
INTEGRATION.md) runs on the exact GPU search.

Behaviour kept from Chroma: ``add`` of an existing id is a no-op, not an upsert;
``query`` returns ascending distance in Chroma's convention (cosine ``1 - cos``,
ip ``1 - dot``); ties go to the first inserted row; ``get_collection`` of a
missing name raises.  The embeddings live in HBM as fp16 / bf16 / int8 / 1-bit
rows (``dtype``); ids, documents and metadatas stay on the host.
"""
from __future__ import annotations

import json
import math
import os
from typing import Any, Dict, List, Optional

import numpy as np

from .index import ShardIndex

DEFAULT_DTYPE = "f16"
DEFAULT_DEVICE = 0


class Settings:
    """chromadb.config.Settings stand-in (the reference passes anonymized_telemetry=False)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def _where_ok(meta: dict, where: Optional[dict]) -> bool:
    if not where:
        return True
    for key, cond in where.items():
        if key == "$and":
            if not all(_where_ok(meta, w) for w in cond):
                return False
            continue
        if key == "$or":
            if not any(_where_ok(meta, w) for w in cond):
                return False
            continue
        present = key in meta
        val = meta.get(key)
        ops = cond if isinstance(cond, dict) else {"$eq": cond}
        for op, ref in ops.items():
            if op == "$eq":
                good = present and val == ref
            elif op == "$ne":
                good = (not present) or val != ref
            elif op in ("$gt", "$gte", "$lt", "$lte"):
                if not present:
                    good = False
                elif op == "$gt":
                    good = val > ref
                elif op == "$gte":
                    good = val >= ref
                elif op == "$lt":
                    good = val < ref
                else:
                    good = val <= ref
            elif op == "$in":
                good = present and val in ref
            elif op == "$nin":
                good = (not present) or val not in ref
            else:
                raise ValueError(f"unsupported where operator {op!r}")
            if not good:
                return False
    return True


def _doc_ok(doc: str, cond: Optional[dict]) -> bool:
    if not cond:
        return True
    for op, ref in cond.items():
        if op == "$contains":
            good = ref in doc
        elif op == "$not_contains":
            good = ref not in doc
        elif op == "$and":
            good = all(_doc_ok(doc, c) for c in ref)
        elif op == "$or":
            good = any(_doc_ok(doc, c) for c in ref)
        else:
            raise ValueError(f"unsupported where_document operator {op!r}")
        if not good:
            return False
    return True


class Collection:
    def __init__(self, name: str, metadata: Optional[dict] = None, dtype: str = DEFAULT_DTYPE,
                 device: int = DEFAULT_DEVICE, directory: Optional[str] = None):
        self.name = name
        self.metadata = dict(metadata) if metadata else None
        self._dtype = dtype
        self._device = device
        self._dir = directory
        self._index: Optional[ShardIndex] = None
        self._ids: List[str] = []
        self._row_of: Dict[str, int] = {}
        self._docs: List[Optional[str]] = []
        self._metas: List[Optional[dict]] = []

    # ------------------------------------------------------------------ basics
    @property
    def space(self) -> str:
        return (self.metadata or {}).get("hnsw:space", "l2")

    def _metric(self) -> str:
        sp = self.space
        if sp in ("cosine", "ip"):
            return sp
        raise ValueError(f"hnsw:space={sp!r} is not supported by the GPU backend (cosine and ip are)")

    def count(self) -> int:
        return len(self._ids)

    @property
    def index(self) -> Optional[ShardIndex]:
        return self._index

    # ------------------------------------------------------------------ add
    def add(self, ids, embeddings, documents=None, metadatas=None) -> None:
        emb = np.asarray(embeddings, dtype=np.float32)
        if emb.ndim != 2 or emb.shape[0] != len(ids):
            raise ValueError(f"embeddings must be [len(ids), dim], got {emb.shape} for {len(ids)} ids")
        if documents is not None and len(documents) != len(ids):
            raise ValueError("documents and ids differ in length")
        if metadatas is not None and len(metadatas) != len(ids):
            raise ValueError("metadatas and ids differ in length")
        if self._index is None:
            self._index = ShardIndex(emb.shape[1], dtype=self._dtype, metric=self._metric(), device=self._device)
        elif emb.shape[1] != self._index.dim:
            raise ValueError(f"embedding dimension {emb.shape[1]} does not match collection dimension {self._index.dim}")
        fresh = []
        seen = set()
        for i, cid in enumerate(ids):
            if cid in self._row_of or cid in seen:       # existing id: ignored (Chroma add is not an upsert)
                continue
            seen.add(cid)
            fresh.append(i)
        if not fresh:
            return
        self._index.add(emb[fresh] if len(fresh) != len(ids) else emb)
        for i in fresh:
            self._row_of[ids[i]] = len(self._ids)
            self._ids.append(ids[i])
            self._docs.append(documents[i] if documents is not None else None)
            self._metas.append(dict(metadatas[i]) if metadatas is not None and metadatas[i] is not None else None)
        if self._dir:
            self.persist()

    # ------------------------------------------------------------------ query
    def query(self, query_embeddings, n_results: int = 10, where: Optional[dict] = None,
              where_document: Optional[dict] = None, include=None,
              min_similarity: float = -math.inf) -> Dict[str, Any]:
        q = np.asarray(query_embeddings, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        out = {"ids": [], "documents": [], "metadatas": [], "distances": [],
               "embeddings": None, "uris": None, "data": None,
               "included": ["metadatas", "documents", "distances"]}
        n = self.count()
        if n == 0 or n_results <= 0:
            for _ in range(q.shape[0]):
                for key in ("ids", "documents", "metadatas", "distances"):
                    out[key].append([])
            return out
        allow = None
        if where or where_document:
            # predicates are dict / string work: evaluated here, pushed into the scans as a row bitmap
            allow = np.fromiter((_where_ok(self._metas[r] or {}, where) and _doc_ok(self._docs[r] or "", where_document)
                                 for r in range(n)), dtype=bool, count=n)
            if not allow.any():
                for _ in range(q.shape[0]):
                    for key in ("ids", "documents", "metadatas", "distances"):
                        out[key].append([])
                return out
        k = min(int(n_results), n)
        ids, raw, counts = self._index.search(q, k, min_similarity, allow=allow)
        sims = self._index.similarity(raw)
        for i in range(q.shape[0]):
            c = int(counts[i])
            rows = [int(r) for r in ids[i, :c]]
            out["ids"].append([self._ids[r] for r in rows])
            out["documents"].append([self._docs[r] for r in rows])
            out["metadatas"].append([self._metas[r] for r in rows])
            out["distances"].append([1.0 - float(s) for s in sims[i, :c]])
        return out

    def stored_vectors(self, ids: List[str]) -> np.ndarray:
        """Stored codes [len(ids), row_bytes] (uint8) of the given chunk ids — what MMR compares."""
        rows = np.array([self._row_of[c] for c in ids], dtype=np.uint32)
        return self._index.fetch_rows(rows)

    # ------------------------------------------------------------------ persistence (N2)
    def _paths(self):
        base = os.path.join(self._dir, self.name)
        return base + ".crs", base + ".json"

    def persist(self) -> None:
        os.makedirs(self._dir, exist_ok=True)
        blob, side = self._paths()
        if self._index is not None:
            self._index.save(blob)
        with open(side, "w") as f:
            json.dump({"name": self.name, "metadata": self.metadata, "dtype": self._dtype,
                       "ids": self._ids, "documents": self._docs, "metadatas": self._metas}, f)

    @classmethod
    def load(cls, name: str, directory: str, device: int = DEFAULT_DEVICE) -> "Collection":
        base = os.path.join(directory, name)
        with open(base + ".json") as f:
            side = json.load(f)
        col = cls(name, side["metadata"], dtype=side.get("dtype", DEFAULT_DTYPE), device=device, directory=directory)
        col._ids = list(side["ids"])
        col._docs = list(side["documents"])
        col._metas = list(side["metadatas"])
        col._row_of = {c: i for i, c in enumerate(col._ids)}
        if os.path.exists(base + ".crs"):
            col._index = ShardIndex.load(base + ".crs", device=device)
            if len(col._index) != len(col._ids):
                raise ValueError("index blob and sidecar disagree on the row count")
        return col

    def drop_files(self) -> None:
        if self._dir:
            for p in self._paths():
                if os.path.exists(p):
                    os.remove(p)

    def close(self) -> None:
        if self._index is not None:
            self._index.close()
            self._index = None


class _ClientBase:
    def __init__(self, directory: Optional[str], dtype: str, device: int):
        self._dir = directory
        self._dtype = dtype
        self._device = device
        self._open: Dict[str, Collection] = {}

    def get_collection(self, name: str) -> Collection:
        if name in self._open:
            return self._open[name]
        if self._dir and os.path.exists(os.path.join(self._dir, name + ".json")):
            col = Collection.load(name, self._dir, self._device)
            self._open[name] = col
            return col
        raise ValueError(f"Collection {name} does not exist.")

    def create_collection(self, name: str, metadata: Optional[dict] = None) -> Collection:
        exists = name in self._open or (self._dir and os.path.exists(os.path.join(self._dir, name + ".json")))
        if exists:
            raise ValueError(f"Collection {name} already exists.")
        col = Collection(name, metadata, dtype=self._dtype, device=self._device, directory=self._dir)
        self._open[name] = col
        if self._dir:
            col.persist()
        return col

    def delete_collection(self, name: str) -> None:
        col = self.get_collection(name)
        col.drop_files()
        col.close()
        del self._open[name]


class Client(_ClientBase):
    """In-memory client (chromadb.Client)."""

    def __init__(self, settings: Optional[Settings] = None, dtype: str = DEFAULT_DTYPE, device: int = DEFAULT_DEVICE):
        super().__init__(None, dtype, device)


class PersistentClient(_ClientBase):
    """Directory-backed client (chromadb.PersistentClient): raw code blob + JSON sidecar per collection."""

    def __init__(self, path: str = "./chroma", dtype: str = DEFAULT_DTYPE, device: int = DEFAULT_DEVICE):
        super().__init__(str(path), dtype, device)
