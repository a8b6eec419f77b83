"""compressed_rag_suite_b200 — B200-native exact similarity search behind the
``rag/indexing.py`` / ``rag/retrieval.py`` API of zahraamselim/compressed-rag-suite.

Layout (only what the hot path needs):
  csrc/          hand-written sm_100a CUDA kernels + the C ABI (include/crs.h) -> libcrs.so
  _native.py     ctypes binding of libcrs.so (no CPU fallback)
  index.py       ShardIndex: one GPU shard;  merge_topk: K7
  collection.py  Chroma-shaped Client / Collection on top of ShardIndex
  rag/           drop-in VectorStore / ContextRetriever with the reference's signatures
  sharded.py     row-sharded multi-GPU search (one process per GPU, one allgather)
"""
__version__ = "0.1.0"
