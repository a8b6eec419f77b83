"""``Chunk`` record — the only part of the reference's ``rag/chunking.py`` that crosses the
hot-path boundary (reference rag/chunking.py:24-33).  ``VectorStore.create_index`` reads
``chunk_id``, ``text`` and the metadata fields by attribute, so any object with these
attributes (including the reference's own dataclass) is accepted."""
from dataclasses import dataclass
from typing import Optional


@dataclass
class Chunk:
    text: str
    chunk_id: str
    start_char: int
    end_char: int
    page_number: Optional[int] = None
    section: Optional[str] = None
    tokens: Optional[int] = None
