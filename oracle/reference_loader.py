"""Oracle: import the reference's OWN ``rag/indexing.py`` + ``rag/retrieval.py``.

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.

Works only where ``/root/reference`` exists (the build container); the GPU box
does not have it, so only ``-m "not gpu"`` tests and the golden-vector generator
(``tests/golden/make_golden.py``) call this.  Third-party modules that are not
installable offline are stubbed in ``sys.modules`` before the import:

* ``chromadb`` / ``chromadb.config``  -> ``oracle.fake_chroma``
* ``sentence_transformers``, ``nltk``, ``PyPDF2`` -> empty stubs (never called:
  tests pass their own embedder object to ``ContextRetriever``)

The reference modules themselves run UNMODIFIED.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CRS_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "rag", "retrieval.py"))


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference():
    """-> (ref_indexing, ref_retrieval, ref_chunking, fake_chroma) modules.

    The reference package is imported under its own name ``rag`` from a private
    sys.path entry, so it cannot collide with ``compressed_rag_suite_b200.rag``."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    from . import fake_chroma

    chroma = _stub("chromadb", Client=fake_chroma.Client,
                   PersistentClient=fake_chroma.PersistentClient)
    cfg = _stub("chromadb.config", Settings=fake_chroma.Settings)
    chroma.config = cfg

    class _NoSentenceTransformer:                      # rag/embedding.py:5 import only
        def __init__(self, *a, **k):
            raise RuntimeError("sentence-transformers is not available offline")

    _stub("sentence_transformers", SentenceTransformer=_NoSentenceTransformer)

    class _Data:                                       # rag/chunking.py:11-21 import-time probe
        @staticmethod
        def find(_):
            return True

        @staticmethod
        def load(_):
            raise RuntimeError("nltk punkt is not available offline")

    _stub("nltk", data=_Data, download=lambda *a, **k: True)
    if "PyPDF2" not in sys.modules:
        _stub("PyPDF2")

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    for name in [m for m in sys.modules if m == "rag" or m.startswith("rag.")]:
        del sys.modules[name]
    ref_indexing = importlib.import_module("rag.indexing")
    ref_retrieval = importlib.import_module("rag.retrieval")
    ref_chunking = importlib.import_module("rag.chunking")
    assert os.path.realpath(ref_retrieval.__file__).startswith(os.path.realpath(REFERENCE_ROOT))
    return ref_indexing, ref_retrieval, ref_chunking, fake_chroma
