"""VectorStore — same public surface as the reference's ``rag/indexing.py`` (:14-211),
with the ChromaDB engine replaced by the exact GPU search of libcrs.

Kept from the reference: constructor keys ``collection_name`` (default
``rag_documents``) and ``persist_directory`` (:27-28); ``collection`` is ``None`` until
the first add (:43, :79-84) and afterwards exposes ``.metadata`` (with
``hnsw:space = cosine``) and ``.count()``; ``create_index`` warns and returns on empty
input (:71-73) and raises ``ValueError`` on a length mismatch (:75-76); metadata is built
from ``page_number / section / tokens`` skipping ``None`` and stringifying anything that
is not str/int/float (:95-109); ``search`` raises ``ValueError`` without a collection
(:144-145), returns four ``[[]]`` on an empty one (:147-149), clamps ``top_k`` to the
collection size (:152-153), accepts a 1-D or 2-D array, list or iterable (:156-168) and
returns ``ids / documents / metadatas / distances`` as lists of one list in ascending
Chroma distance; every failure is logged and re-raised (:121-123, :178-180).

Optional additive keys (defaults reproduce the reference): ``dtype`` (f16 | bf16 | i8 |
b1, default f16), ``device`` (CUDA ordinal, default 0), ``devices`` (list of CUDA ordinals: the
collection is dealt out over these GPUs and searched on all of them from this one process, see
multi.MultiDeviceIndex).  ``embeddings`` / ``query_embedding`` may be torch CUDA tensors
(``SentenceTransformer.encode(convert_to_tensor=True)``, reference rag/embedding.py:65-71): they reach
the ingest / search kernels without a host round trip.  One search returns at most 112 (f16 / bf16) or
128 (i8 / b1) candidates (``collection.MAX_RESULTS``); more raises ``ValueError``.
"""
from __future__ import annotations

import logging
from typing import Any, Dict, List, Optional

import numpy as np

from .. import collection as _backend

logger = logging.getLogger(__name__)

_META_DEFAULT = ("page_number", "section", "tokens")


class VectorStore:
    def __init__(self, config: dict):
        self.collection_name = config.get("collection_name", "rag_documents")
        self.persist_directory = config.get("persist_directory", None)
        dtype = config.get("dtype", _backend.DEFAULT_DTYPE)
        devices = config.get("devices", None)
        device = config.get("device", devices[0] if devices else _backend.DEFAULT_DEVICE)
        try:
            if self.persist_directory:
                self.client = _backend.PersistentClient(path=self.persist_directory, dtype=dtype, device=device, devices=devices)
                logger.info(f"Using persistent storage: {self.persist_directory}")
            else:
                self.client = _backend.Client(_backend.Settings(anonymized_telemetry=False), dtype=dtype, device=device,
                                              devices=devices)
                logger.info("Using in-memory storage")
        except Exception as e:
            logger.error(f"Failed to initialize vector store client: {e}")
            raise
        self.collection = None
        self._initialize_collection()

    def _initialize_collection(self):
        try:
            self.collection = self.client.get_collection(self.collection_name)
        except _backend.CollectionNotFound:
            logger.info(f"Collection '{self.collection_name}' will be created on first add")
            return
        except Exception as e:
            # the reference treats every failure here as "no collection yet" (rag/indexing.py:53-55); a store that
            # EXISTS but cannot be read must not be silently replaced by an empty one
            logger.error(f"Collection '{self.collection_name}' exists but cannot be loaded: {e}")
            raise
        logger.info(f"Loaded existing collection: {self.collection_name}")
        logger.info(f"Collection size: {self.collection.count()}")

    @staticmethod
    def _chunk_metadata(chunk, fields) -> dict:
        meta = {}
        for name in fields:
            value = getattr(chunk, name, None)
            if value is None:
                continue
            meta[name] = value if isinstance(value, (str, int, float)) else str(value)
        return meta

    def create_index(self, chunks: List[Any], embeddings: np.ndarray,
                     metadata_fields: Optional[List[str]] = None):
        if len(chunks) == 0:
            logger.warning("No chunks provided for indexing")
            return
        if len(chunks) != len(embeddings):
            raise ValueError(f"Chunk count ({len(chunks)}) doesn't match embedding count ({len(embeddings)})")
        if self.collection is None:
            try:
                self.collection = self.client.create_collection(name=self.collection_name,
                                                                metadata={"hnsw:space": "cosine"})
                logger.info(f"Created new collection: {self.collection_name}")
            except Exception as e:
                logger.error(f"Failed to create collection: {e}")
                raise
        fields = _META_DEFAULT if metadata_fields is None else metadata_fields
        try:
            logger.info(f"Adding {len(chunks)} chunks to index...")
            on_device = type(embeddings).__module__.startswith("torch") and getattr(embeddings, "is_cuda", False)
            self.collection.add(ids=[c.chunk_id for c in chunks],
                                # no .tolist() boxing; a CUDA tensor from the embedder stays on the device
                                embeddings=embeddings if on_device else np.asarray(embeddings, dtype=np.float32),
                                documents=[c.text for c in chunks],
                                metadatas=[self._chunk_metadata(c, fields) for c in chunks])
            logger.info(f"Index created successfully! Total documents: {self.collection.count()}")
        except Exception as e:
            logger.error(f"Failed to add documents to collection: {e}")
            raise

    supports_min_similarity = True      # additive: lets the retriever push its threshold into the scan

    def search(self, query_embedding, top_k: int = 5, where: Optional[dict] = None,
               where_document: Optional[dict] = None, min_similarity: Optional[float] = None) -> Dict[str, Any]:
        if self.collection is None:
            raise ValueError("No collection available. Create index first.")
        size = self.collection.count()
        if size == 0:
            logger.warning("Collection is empty. No results to return.")
            return {"ids": [[]], "documents": [[]], "metadatas": [[]], "distances": [[]]}
        top_k = min(top_k, size)
        # the reference's conversion (:156-168): an ndarray of any shape is ONE query (flattened); a list whose first
        # element is a list is passed through as a batch of queries; anything else is one query
        if isinstance(query_embedding, np.ndarray):
            batch = query_embedding.reshape(1, -1)
        elif type(query_embedding).__module__.startswith("torch") and getattr(query_embedding, "is_cuda", False):
            batch = query_embedding.reshape(1, -1)                    # stays on the device
        else:
            as_list = query_embedding if isinstance(query_embedding, list) else list(query_embedding)
            if len(as_list) and isinstance(as_list[0], list):
                batch = np.asarray(as_list, dtype=np.float32)
            else:
                batch = np.asarray(as_list, dtype=np.float32).reshape(1, -1)
        try:
            extra = {} if min_similarity is None else {"min_similarity": float(min_similarity)}
            return self.collection.query(query_embeddings=batch, n_results=top_k,
                                         where=where, where_document=where_document, **extra)
        except Exception as e:
            logger.error(f"Search failed: {e}")
            raise

    def delete_collection(self):
        if self.collection:
            try:
                self.client.delete_collection(self.collection_name)
                self.collection = None
                logger.info(f"Deleted collection: {self.collection_name}")
            except Exception as e:
                logger.error(f"Failed to delete collection: {e}")
                raise

    def reset_collection(self):
        self.delete_collection()
        self._initialize_collection()

    def get_stats(self) -> Dict[str, Any]:
        if self.collection is None:
            return {"status": "empty", "count": 0}
        try:
            return {"name": self.collection_name, "count": self.collection.count(),
                    "metadata": self.collection.metadata}
        except Exception as e:
            logger.error(f"Failed to get stats: {e}")
            return {"status": "error", "error": str(e)}
