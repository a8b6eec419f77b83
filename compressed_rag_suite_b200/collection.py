"""The Chroma-shaped surface the reference talks to, backed by libcrs on one or several B200s.

``rag/indexing.py`` of the reference uses exactly these calls (SURVEY.md §1 L1★):
``Client(settings)`` / ``PersistentClient(path)``, ``get_collection``,
``create_collection(name, metadata)``, ``delete_collection``; ``Collection.add``,
``.query``, ``.count()``, ``.metadata``.  Providing them means either the mirror
``compressed_rag_suite_b200.rag.indexing.VectorStore`` or the reference's own
unmodified ``rag/indexing.py`` (with ``sys.modules["chromadb"]`` pointed here, see
INTEGRATION.md) runs on the exact GPU search.

Behaviour kept from Chroma: ``add`` of an id the collection already holds is a no-op, not an
upsert; ``query`` returns ascending distance in Chroma's convention (cosine ``1 - cos``,
ip ``1 - dot``), computed in float32 like Chroma's; ties go to the first inserted row;
``get_collection`` of a missing name raises.  Deviations: an id repeated INSIDE one ``add``
batch keeps its first occurrence (Chroma raises ``DuplicateIDError``); ``l2`` space is not
offered (the reference never asks for it, rag/indexing.py:83).

The embeddings live in HBM as fp16 / bf16 / int8 / 1-bit rows (``dtype``), on one GPU or dealt
out over several (``devices``, multi.MultiDeviceIndex); ids, documents and metadatas stay on the
host.  ``embeddings`` and ``query_embeddings`` may be torch CUDA tensors (SentenceTransformer
``convert_to_tensor=True``, reference rag/embedding.py:65-71): they go to the ingest / search
kernels without a host round trip.

Persistence (``persist_directory``) is append-only and crash-safe:
  <name>.meta.json   collection header (name, metadata, dtype, shard layout) — rewritten atomically
  <name>.rows.jsonl  one JSON line per row: [id, document, metadata], appended and fsync'ed
  <name>.crs[.dN]    raw code blob(s): rows appended, then the header's row count advanced (libcrs)
An ``add`` costs O(new rows).  On load the common prefix of blob and sidecar is kept: a torn last line,
a torn row, or rows whose counterpart did not reach the disk are dropped, never misread.  A directory
written by ChromaDB itself (chroma.sqlite3) cannot be read: the store starts empty and says so.
"""
from __future__ import annotations

import json
import logging
import math
import os
from typing import Any, Dict, List, Optional

import numpy as np

from .index import ShardIndex

logger = logging.getLogger(__name__)

DEFAULT_DTYPE = "f16"
DEFAULT_DEVICE = 0
# candidates one search can return: the kernels keep <= 128 keys per list and float stores need head-room
# above k for the certification (crs_index_search)
MAX_RESULTS = {"f16": 112, "bf16": 112, "i8": 128, "b1": 128}
FORMAT = 2


class CollectionNotFound(ValueError):
    pass


class CorruptStoreError(RuntimeError):
    pass


class Settings:
    """chromadb.config.Settings stand-in (the reference passes anonymized_telemetry=False)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def _is_torch_cuda(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "is_cuda") and x.is_cuda


# ---------------------------------------------------------------------------- predicates (row by row: the definition)
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


class _Columns:
    """Metadata as columns, so a ``where`` predicate is a handful of numpy operations over all rows instead of
    a Python loop over dicts.  Built lazily per key and extended when rows are added."""

    def __init__(self, metas: List[Optional[dict]]):
        self._metas = metas
        self._cols: Dict[str, tuple] = {}

    def _column(self, key: str):
        n = len(self._metas)
        have = self._cols.get(key)
        start = 0
        if have is not None and have[0].shape[0] == n:
            return have
        if have is not None:
            start = have[0].shape[0]
        present = np.zeros(n, dtype=bool)
        is_num = np.zeros(n, dtype=bool)
        num = np.full(n, np.nan, dtype=np.float64)
        obj = np.empty(n, dtype=object)
        if have is not None:
            present[:start], is_num[:start], num[:start], obj[:start] = have
        for r in range(start, n):
            m = self._metas[r]
            if m and key in m:
                v = m[key]
                present[r] = True
                obj[r] = v
                if isinstance(v, (int, float)) and not isinstance(v, bool) and (isinstance(v, float) or abs(v) < (1 << 53)):
                    is_num[r] = True                        # exactly representable as float64; anything else is compared the Python way
                    num[r] = float(v)
        self._cols[key] = (present, is_num, num, obj)
        return self._cols[key]

    @staticmethod
    def _exact_int(ref) -> bool:
        return isinstance(ref, int) and not isinstance(ref, bool) and abs(ref) < (1 << 53)

    def _cmp(self, key: str, op: str, ref) -> np.ndarray:
        present, is_num, num, obj = self._column(key)
        n = present.shape[0]
        numeric_ref = isinstance(ref, (int, float)) and not isinstance(ref, bool)
        if op in ("$eq", "$ne"):
            if numeric_ref and (isinstance(ref, float) or self._exact_int(ref)):
                eq = is_num & (num == float(ref))
                other = present & ~is_num                      # strings / bools compared the Python way
                if other.any():
                    idx = np.nonzero(other)[0]
                    eq[idx] = [obj[i] == ref for i in idx]
            else:
                eq = np.zeros(n, dtype=bool)
                idx = np.nonzero(present)[0]
                eq[idx] = [obj[i] == ref for i in idx]
            return eq if op == "$eq" else ~eq                  # $ne: absent keys pass
        if op in ("$gt", "$gte", "$lt", "$lte"):
            fn = {"$gt": np.greater, "$gte": np.greater_equal, "$lt": np.less, "$lte": np.less_equal}[op]
            if numeric_ref and (isinstance(ref, float) or self._exact_int(ref)):
                out = is_num & fn(num, float(ref))
                other = present & ~is_num                      # huge ints, strings, bools: the Python way (a string
                if other.any():                                # against a number raises there, so it does here)
                    idx = np.nonzero(other)[0]
                    out[idx] = [_where_ok({key: obj[i]}, {key: {op: ref}}) for i in idx]
                return out
            out = np.zeros(n, dtype=bool)
            idx = np.nonzero(present)[0]
            out[idx] = [_where_ok({key: obj[i]}, {key: {op: ref}}) for i in idx]
            return out
        if op in ("$in", "$nin"):
            inn = np.zeros(n, dtype=bool)
            idx = np.nonzero(present)[0]
            inn[idx] = [obj[i] in ref for i in idx]
            return inn if op == "$in" else ~inn
        raise ValueError(f"unsupported where operator {op!r}")

    def mask(self, where: Optional[dict]) -> np.ndarray:
        n = len(self._metas)
        out = np.ones(n, dtype=bool)
        if not where:
            return out
        for key, cond in where.items():
            if key == "$and":
                for w in cond:
                    out &= self.mask(w)
            elif key == "$or":
                acc = np.zeros(n, dtype=bool)
                for w in cond:
                    acc |= self.mask(w)
                out &= acc
            else:
                ops = cond if isinstance(cond, dict) else {"$eq": cond}
                for op, ref in ops.items():
                    out &= self._cmp(key, op, ref)
        return out


def _doc_mask(docs: List[Optional[str]], cond: Optional[dict]) -> np.ndarray:
    n = len(docs)
    if not cond:
        return np.ones(n, dtype=bool)
    out = np.ones(n, dtype=bool)
    for op, ref in cond.items():
        if op == "$contains":
            out &= np.fromiter((ref in (d or "") for d in docs), dtype=bool, count=n)
        elif op == "$not_contains":
            out &= np.fromiter((ref not in (d or "") for d in docs), dtype=bool, count=n)
        elif op == "$and":
            for c in ref:
                out &= _doc_mask(docs, c)
        elif op == "$or":
            acc = np.zeros(n, dtype=bool)
            for c in ref:
                acc |= _doc_mask(docs, c)
            out &= acc
        else:
            raise ValueError(f"unsupported where_document operator {op!r}")
    return out


def _fsync_dir(path: str) -> None:
    try:
        fd = os.open(path, os.O_RDONLY)
        try:
            os.fsync(fd)
        finally:
            os.close(fd)
    except OSError:
        pass


def _write_atomic(path: str, text: str) -> None:
    tmp = path + ".tmp"
    with open(tmp, "w") as f:
        f.write(text)
        f.flush()
        os.fsync(f.fileno())
    os.replace(tmp, path)
    _fsync_dir(os.path.dirname(path) or ".")


class Collection:
    def __init__(self, name: str, metadata: Optional[dict] = None, dtype: str = DEFAULT_DTYPE,
                 device: int = DEFAULT_DEVICE, directory: Optional[str] = None, devices: Optional[List[int]] = None):
        self.name = name
        self.metadata = dict(metadata) if metadata else None
        self._dtype = dtype
        self._device = device
        self._devices = [int(d) for d in devices] if devices and len(devices) > 1 else None
        self._dir = directory
        self._index = None                       # ShardIndex or multi.MultiDeviceIndex
        self._ids: List[str] = []
        self._row_of: Dict[str, int] = {}
        self._docs: List[Optional[str]] = []
        self._metas: List[Optional[dict]] = []
        self._columns = _Columns(self._metas)
        self._allow_cache: Dict[str, tuple] = {}
        self._persisted_rows = 0

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
    def index(self):
        return self._index

    def _new_index(self, dim: int):
        if self._devices:
            from .multi import MultiDeviceIndex
            return MultiDeviceIndex(dim, dtype=self._dtype, metric=self._metric(), devices=self._devices)
        return ShardIndex(dim, dtype=self._dtype, metric=self._metric(), device=self._device)

    # ------------------------------------------------------------------ add
    def add(self, ids, embeddings, documents=None, metadatas=None) -> None:
        on_device = _is_torch_cuda(embeddings)
        if on_device:
            import torch
            emb = embeddings if embeddings.dtype == torch.float32 else embeddings.float()
        else:
            emb = np.asarray(embeddings, dtype=np.float32)
        if emb.ndim != 2 or emb.shape[0] != len(ids):
            raise ValueError(f"embeddings must be [len(ids), dim], got {tuple(emb.shape)} for {len(ids)} ids")
        if documents is not None and len(documents) != len(ids):
            raise ValueError("documents and ids differ in length")
        if metadatas is not None and len(metadatas) != len(ids):
            raise ValueError("metadatas and ids differ in length")
        if self._index is None:
            self._index = self._new_index(int(emb.shape[1]))
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
        if len(fresh) != len(ids):
            emb = emb[torch.as_tensor(fresh, device=emb.device)] if on_device else emb[fresh]
        self._index.add(emb.contiguous() if on_device else emb)
        for i in fresh:
            self._row_of[ids[i]] = len(self._ids)
            self._ids.append(ids[i])
            self._docs.append(documents[i] if documents is not None else None)
            self._metas.append(dict(metadatas[i]) if metadatas is not None and metadatas[i] is not None else None)
        self._allow_cache.clear()
        if self._dir:
            self.persist()

    # ------------------------------------------------------------------ query
    def _allow_mask(self, where, where_document) -> np.ndarray:
        """Row mask of a where / where_document pair: evaluated over columns and cached per predicate
        (the cache is dropped when rows are added)."""
        key = json.dumps([where, where_document], sort_keys=True, default=str)
        hit = self._allow_cache.get(key)
        if hit is not None:
            return hit[0]
        mask = self._columns.mask(where) & _doc_mask(self._docs, where_document)
        if isinstance(self._index, ShardIndex) and mask.any():
            mask = self._index.pack_allow(mask)           # the uint32 bitmap the kernels read; packed once per predicate
        if len(self._allow_cache) >= 64:
            self._allow_cache.pop(next(iter(self._allow_cache)))
        self._allow_cache[key] = (mask,)
        return mask

    def query(self, query_embeddings, n_results: int = 10, where: Optional[dict] = None,
              where_document: Optional[dict] = None, include=None,
              min_similarity: float = -math.inf) -> Dict[str, Any]:
        on_device = _is_torch_cuda(query_embeddings)
        if on_device:
            q = query_embeddings if query_embeddings.dim() == 2 else query_embeddings[None, :]
            if str(q.dtype) != "torch.float32":
                q = q.float()
        else:
            q = np.asarray(query_embeddings, dtype=np.float32)
            if q.ndim == 1:
                q = q[None, :]
        nq = int(q.shape[0])
        out = {"ids": [], "documents": [], "metadatas": [], "distances": [],
               "embeddings": None, "uris": None, "data": None,
               "included": ["metadatas", "documents", "distances"]}
        n = self.count()

        def empty():
            for _ in range(nq):
                for key in ("ids", "documents", "metadatas", "distances"):
                    out[key].append([])
            return out

        if n == 0 or n_results <= 0:
            return empty()
        allow = None
        if where or where_document:
            # predicates are dict / string work: evaluated on the host, pushed into the kernels as a row bitmap
            allow = self._allow_mask(where, where_document)
            if not allow.any():
                return empty()
        k = min(int(n_results), n)
        cap = MAX_RESULTS[self._dtype]
        if k > cap:
            raise ValueError(f"n_results={k} exceeds what one exact GPU search returns for a {self._dtype} store ({cap}); "
                             f"ask for at most {cap} candidates (ContextRetriever fetches 2 * top_k when rerank is on)")
        ids, raw, counts = self._index.search(q, k, min_similarity, allow=allow)
        if on_device:
            import torch
            torch.cuda.current_stream(ids.device).synchronize()
            ids, raw, counts = ids.cpu().numpy().view(np.uint32), raw.cpu().numpy(), counts.cpu().numpy()
        sims = self._index.similarity(raw)
        one = np.float32(1.0)
        for i in range(nq):
            c = int(counts[i])
            rows = [int(r) for r in ids[i, :c]]
            out["ids"].append([self._ids[r] for r in rows])
            out["documents"].append([self._docs[r] for r in rows])
            out["metadatas"].append([self._metas[r] for r in rows])
            out["distances"].append([float(one - s) for s in sims[i, :c]])      # float32 arithmetic, like Chroma
        return out

    def stored_vectors(self, ids: List[str]) -> np.ndarray:
        """Stored codes [len(ids), row_bytes] (uint8) of the given chunk ids — what MMR compares."""
        rows = np.array([self._row_of[c] for c in ids], dtype=np.uint32)
        return self._index.fetch_rows(rows)

    # ------------------------------------------------------------------ persistence (N2)
    def _paths(self):
        base = os.path.join(self._dir, self.name)
        return base + ".crs", base + ".meta.json", base + ".rows.jsonl"

    def _header(self) -> dict:
        h = {"format": FORMAT, "name": self.name, "metadata": self.metadata, "dtype": self._dtype}
        if self._devices and self._index is not None:
            h["layout"] = self._index.layout()
        elif self._devices:
            h["layout"] = {"devices": len(self._devices), "segments": [], "count": 0}
        return h

    def persist(self) -> None:
        """Bring the files up to date with the collection: O(rows added since the last call).
        Order: code blob (rows, then its header) -> sidecar lines -> collection header; every step is flushed
        to disk before the next, and load() keeps the common prefix, so a crash at any point loses at most
        the rows of the interrupted add."""
        os.makedirs(self._dir, exist_ok=True)
        blob, meta, rows = self._paths()
        n = len(self._ids)
        if self._index is not None and n > self._persisted_rows:
            multi = self._devices is not None
            if self._persisted_rows == 0 and not multi:
                self._index.save(blob)
            else:
                self._index.append_to(blob)
            with open(rows, "ab") as f:
                for r in range(self._persisted_rows, n):
                    f.write((json.dumps([self._ids[r], self._docs[r], self._metas[r]], ensure_ascii=False) + "\n").encode("utf-8"))
                f.flush()
                os.fsync(f.fileno())
            self._persisted_rows = n
        _write_atomic(meta, json.dumps(self._header()))

    @classmethod
    def load(cls, name: str, directory: str, device: int = DEFAULT_DEVICE,
             devices: Optional[List[int]] = None) -> "Collection":
        base = os.path.join(directory, name)
        legacy = base + ".json"
        try:
            if os.path.exists(base + ".meta.json"):
                with open(base + ".meta.json") as f:
                    head = json.load(f)
                col = cls(name, head["metadata"], dtype=head.get("dtype", DEFAULT_DTYPE), device=device, directory=directory,
                          devices=devices if head.get("layout") else None)
                valid_bytes = 0
                if os.path.exists(base + ".rows.jsonl"):
                    with open(base + ".rows.jsonl", "rb") as f:
                        for line in f:
                            if not line.endswith(b"\n"):
                                break                                  # torn last line
                            try:
                                cid, doc, meta = json.loads(line.decode("utf-8"))
                            except Exception:
                                break
                            col._ids.append(cid)
                            col._docs.append(doc)
                            col._metas.append(meta)
                            valid_bytes += len(line)
                if head.get("layout"):
                    from .multi import MultiDeviceIndex
                    lay = head["layout"]
                    if lay.get("segments"):
                        devs = devices if devices and len(devices) == int(lay["devices"]) else [device] * int(lay["devices"])
                        col._devices = [int(d) for d in devs]
                        col._index = MultiDeviceIndex.load(base + ".crs", lay, col._devices)
                elif os.path.exists(base + ".crs"):
                    col._index = ShardIndex.load(base + ".crs", device=device)
                rows_path = base + ".rows.jsonl"
                if os.path.exists(rows_path) and os.path.getsize(rows_path) != valid_bytes:
                    with open(rows_path, "r+b") as f:                 # drop a torn last line so that appends start on a line boundary
                        f.truncate(valid_bytes)
                        f.flush()
                        os.fsync(f.fileno())
                n_idx = len(col._index) if col._index is not None else 0
                n = min(n_idx, len(col._ids))
                if n < len(col._ids) or n < n_idx:
                    logger.warning(f"collection {name!r}: blob holds {n_idx} rows, sidecar {len(col._ids)}; keeping the first {n}")
                    if n < len(col._ids):                             # rewrite the sidecar prefix, atomically
                        del col._ids[n:], col._docs[n:], col._metas[n:]
                        _write_atomic(rows_path, "".join(
                            json.dumps([col._ids[r], col._docs[r], col._metas[r]], ensure_ascii=False) + "\n" for r in range(n)))
                    if col._index is not None and n < n_idx:
                        if isinstance(col._index, ShardIndex):
                            col._index.truncate(n)
                            col._index.save(base + ".crs")            # atomic rewrite: the blob's header must not promise more rows
                        else:
                            raise CorruptStoreError(f"collection {name!r}: shard blobs and sidecar disagree ({n_idx} vs {len(col._ids)} rows)")
            elif os.path.exists(legacy):                               # round-1 layout: everything in one JSON
                with open(legacy) as f:
                    side = json.load(f)
                col = cls(name, side["metadata"], dtype=side.get("dtype", DEFAULT_DTYPE), device=device, directory=directory)
                col._ids, col._docs, col._metas = list(side["ids"]), list(side["documents"]), list(side["metadatas"])
                if os.path.exists(base + ".crs"):
                    col._index = ShardIndex.load(base + ".crs", device=device)
                    if len(col._index) != len(col._ids):
                        raise CorruptStoreError("index blob and sidecar disagree on the row count")
            else:
                raise CollectionNotFound(f"Collection {name} does not exist.")
        except (CollectionNotFound, CorruptStoreError):
            raise
        except Exception as e:
            raise CorruptStoreError(f"collection {name!r} in {directory!r} cannot be read: {e}") from e
        col._columns = _Columns(col._metas)
        col._row_of = {c: i for i, c in enumerate(col._ids)}
        col._persisted_rows = len(col._ids) if not os.path.exists(legacy) or os.path.exists(base + ".meta.json") else 0
        return col

    def drop_files(self) -> None:
        if self._dir:
            base = os.path.join(self._dir, self.name)
            for p in os.listdir(self._dir):
                full = os.path.join(self._dir, p)
                if full in (base + ".json", base + ".meta.json", base + ".rows.jsonl") or full.startswith(base + ".crs"):
                    os.remove(full)

    def close(self) -> None:
        if self._index is not None:
            self._index.close()
            self._index = None


class _ClientBase:
    def __init__(self, directory: Optional[str], dtype: str, device: int, devices: Optional[List[int]] = None):
        self._dir = directory
        self._dtype = dtype
        self._device = device
        self._devices = devices
        self._open: Dict[str, Collection] = {}
        if directory and os.path.exists(os.path.join(directory, "chroma.sqlite3")):
            logger.warning(f"{directory!r} holds a ChromaDB store (chroma.sqlite3): this backend cannot read it; "
                           "collections start empty until they are indexed again")

    def _exists_on_disk(self, name: str) -> bool:
        return bool(self._dir) and (os.path.exists(os.path.join(self._dir, name + ".meta.json")) or
                                    os.path.exists(os.path.join(self._dir, name + ".json")))

    def get_collection(self, name: str) -> Collection:
        if name in self._open:
            return self._open[name]
        if self._exists_on_disk(name):
            col = Collection.load(name, self._dir, self._device, self._devices)      # CorruptStoreError surfaces
            self._open[name] = col
            return col
        raise CollectionNotFound(f"Collection {name} does not exist.")

    def create_collection(self, name: str, metadata: Optional[dict] = None) -> Collection:
        if name in self._open or self._exists_on_disk(name):
            raise ValueError(f"Collection {name} already exists.")
        col = Collection(name, metadata, dtype=self._dtype, device=self._device, directory=self._dir, devices=self._devices)
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

    def __init__(self, settings: Optional[Settings] = None, dtype: str = DEFAULT_DTYPE, device: int = DEFAULT_DEVICE,
                 devices: Optional[List[int]] = None):
        super().__init__(None, dtype, device, devices)


class PersistentClient(_ClientBase):
    """Directory-backed client (chromadb.PersistentClient): append-only code blob + JSON-lines sidecar."""

    def __init__(self, path: str = "./chroma", dtype: str = DEFAULT_DTYPE, device: int = DEFAULT_DEVICE,
                 devices: Optional[List[int]] = None):
        super().__init__(str(path), dtype, device, devices)
