"""Oracle-side CPU baseline: the exact search as the reference's CPU stack would do it
without ChromaDB's HNSW — one fp32 BLAS matmul over L2-normalised rows, then top-k by
(score desc, id asc).  TEST / BENCH INFRASTRUCTURE (see oracle/__init__.py): only
bench.py's ``cpu_baseline`` leg and ``--impl reference`` time this; nothing ships it.

It follows the arithmetic Chroma performs for ``collection.query`` in cosine space
(reference call site rag/indexing.py:171-176; distance = 1 - cos) exhaustively, i.e. it is
the "fp32 brute-force pass" the north star names as the correctness reference.
"""
from __future__ import annotations

import os
import time

import numpy as np


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def make_slice(rows: int, dim: int, seed: int = 1234, n_clusters: int = 4096):
    """Clustered unit rows + queries with the same recipe as bench.py's device generator
    (numpy stream, so values differ; the distribution is the same)."""
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((n_clusters, dim), dtype=np.float32)
    centres /= np.linalg.norm(centres, axis=1, keepdims=True)
    x = rng.standard_normal((rows, dim), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    x *= np.float32(0.8)
    x += np.float32(0.6) * centres[np.arange(rows) % n_clusters]
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x, centres


def make_queries(centres: np.ndarray, nq: int, seed: int = 4321):
    rng = np.random.default_rng(seed)
    dim = centres.shape[1]
    q = rng.standard_normal((nq, dim), dtype=np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q = np.float32(0.6) * centres[rng.integers(0, centres.shape[0], nq)] + np.float32(0.8) * q
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return q


def search_f32(x: np.ndarray, q: np.ndarray, k: int, min_similarity: float = -np.inf, block: int = 262144):
    """Exact top-k of every query over unit rows x: blocked sgemm + argpartition + ordered
    merge.  -> (ids int64 [nq,k], scores f32 [nq,k])."""
    nq = q.shape[0]
    best_s = np.full((nq, k), -np.inf, dtype=np.float32)
    best_i = np.full((nq, k), -1, dtype=np.int64)
    for lo in range(0, x.shape[0], block):
        s = q @ x[lo:lo + block].T                                  # [nq, block] fp32 (BLAS)
        kk = min(k, s.shape[1])
        part = np.argpartition(-s, kk - 1, axis=1)[:, :kk]
        cs = np.take_along_axis(s, part, axis=1)
        all_s = np.concatenate([best_s, cs], axis=1)
        all_i = np.concatenate([best_i, part + lo], axis=1)
        order = np.lexsort((all_i, -all_s), axis=1)[:, :k]
        best_s = np.take_along_axis(all_s, order, axis=1)
        best_i = np.take_along_axis(all_i, order, axis=1)
    if np.isfinite(min_similarity):
        drop = best_s < np.float32(min_similarity)
        best_i[drop] = -1
    return best_i, best_s


def time_search(x, q, k, steps: int, warmup: int, min_similarity: float = -np.inf):
    for _ in range(warmup):
        search_f32(x, q, k, min_similarity)
    t0 = time.perf_counter()
    for _ in range(steps):
        search_f32(x, q, k, min_similarity)
    return (time.perf_counter() - t0) / max(steps, 1)
