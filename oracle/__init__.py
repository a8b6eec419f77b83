"""CPU oracle for the retrieval hot path — TEST INFRASTRUCTURE ONLY.

This package is a numpy restatement of the arithmetic that the reference
(zahraamselim/compressed-rag-suite) hands to ChromaDB from ``rag/indexing.py``
and post-processes in ``rag/retrieval.py``.  It exists to check the CUDA path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  Nothing under
``compressed_rag_suite_b200/`` imports it, and the product raises when the CUDA
library is missing instead of falling back to this code.

Parity status:

* ``rag/retrieval.py`` post-processing (distance→score transform, threshold,
  lexical rerank, MMR): PINNED — the reference's own unmodified module is
  imported from ``/root/reference`` in the build container (third-party imports
  stubbed, ``oracle/reference_loader.py``) and the restatement is checked against
  it; golden vectors generated that way are committed under ``tests/golden/``.
* The distance arithmetic itself lives in ``chromadb==1.3.0``
  (``requirements.txt:19``), which is not vendored in the reference tree and is
  not installable offline; the reference has no tests or golden vectors for it.
  That part is restated from Chroma's documented distance definitions
  (cosine ``1 - cos``, ip ``1 - dot``, l2 squared) and is **parity unpinned**
  against a real Chroma; it is cross-checked against an independent fp32 brute
  force (numpy sgemm, sklearn brute kNN).
"""
