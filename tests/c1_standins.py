"""Stand-ins for what BASELINE config 1 needs and cannot exist offline — MiniLM weights and nltk's punkt model —
shared by the golden generator (the reference's own RAGPipeline, CPU, build container only) and by the GPU run of
the same shape (tests/test_gpu_pipeline.py, tools/bench_configs.py c1).  Deterministic: no random stream, only hashes."""
import hashlib
import re

import numpy as np

DIM = 384
TOPICS = ["quantization", "pruning", "distillation", "low rank factorization", "sparse attention", "kv cache",
          "speculative decoding", "mixture of experts", "activation outliers", "calibration data", "perplexity",
          "throughput", "memory footprint", "hardware support"]
FILLER = ("large language models are compressed to reduce memory and latency while keeping accuracy the survey "
          "compares methods on benchmarks and reports metrics for inference training and deployment").split()


def _unit(seed: str) -> np.ndarray:
    h = hashlib.sha256(seed.encode("utf-8")).digest()
    rng = np.random.default_rng(int.from_bytes(h[:8], "little"))
    v = rng.standard_normal(DIM).astype(np.float32)
    return v / np.linalg.norm(v)


def embed_text(text: str) -> np.ndarray:
    """Hash-seeded unit vector: a mix of the topics the text mentions plus a text-specific component, so that
    related chunks and queries land near each other like sentence embeddings do."""
    low = text.lower()
    v = 0.9 * _unit("text:" + text)
    for t in TOPICS:
        c = low.count(t)
        if c:
            v = v + min(c, 4) * 0.55 * _unit("topic:" + t)
    v = v.astype(np.float32)
    return v / np.linalg.norm(v)


class HashSentenceTransformer:
    """SentenceTransformer stand-in (rag/embedding.py:31-33,65-71)."""

    def __init__(self, *a, **k):
        pass

    def get_sentence_embedding_dimension(self):
        return DIM

    def encode(self, texts, **kw):
        one = isinstance(texts, str)
        out = np.stack([embed_text(t) for t in ([texts] if one else list(texts))]).astype(np.float32)
        return out


class RegexPunkt:
    """nltk punkt stand-in (rag/chunking.py:58): split after sentence punctuation."""

    def tokenize(self, text):
        return [s for s in re.split(r"(?<=[.!?])\s+", text) if s]


def make_pages(n_pages: int = 14):
    """~2.5-3.5 kchar "pages" (the real run has one chunk per PDF page)."""
    pages = []
    for p in range(n_pages):
        sents = []
        for s in range(34 + (p * 7) % 11):
            h = hashlib.sha256(f"page{p}:sent{s}".encode()).digest()
            words = [FILLER[b % len(FILLER)] for b in h[:10]]
            topic = TOPICS[(p + (h[10] % 3 == 0) * (h[11] % len(TOPICS))) % len(TOPICS)]
            words.insert(h[12] % 8, topic)
            sents.append(" ".join(words).capitalize() + ".")
        pages.append(" ".join(sents) + f"\n\nSection {p + 1} closes with remarks on {TOPICS[p % len(TOPICS)]}.")
    return pages


def make_queries(n: int = 20):
    qs = []
    for i in range(n):
        t = TOPICS[(i * 5) % len(TOPICS)]
        u = TOPICS[(i * 3 + 1) % len(TOPICS)]
        qs.append([f"How does {t} affect accuracy?", f"What is the trade-off between {t} and {u}?",
                   f"Which methods use {t} for inference?", f"Explain {t}."][i % 4])
    return qs
