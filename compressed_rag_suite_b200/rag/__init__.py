"""Drop-in replacements for the hot-path classes of the reference's ``rag`` package."""
from .chunking import Chunk
from .indexing import VectorStore
from .retrieval import ContextRetriever

__all__ = ["Chunk", "VectorStore", "ContextRetriever"]
