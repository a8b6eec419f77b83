"""Oracle: exhaustive scoring + threshold + top-k (the arithmetic Chroma does for
``collection.query`` at the reference call site ``rag/indexing.py:171-176``).

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.

Canonical result definition (what "bit-exact" means for the CUDA path):

* float stores (f16/bf16): ``score(q, c) = fl32( sum_j c_j * q_j )`` where c, q
  are the *stored* (rounded) values, the sum runs **sequentially in fp64** over
  j = 0..Dp-1.  Each product of two 11-bit (8-bit for bf16) significands is
  exact in fp64, so FMA and mul+add agree and CPU/GPU give identical bits.
* i8: ``dot = sum_j c_j * q_j`` in int32 (exact); float score
  ``fl32(fl32(dot) * fl32((a/127)^2))``.
* b1: Hamming distance ``h = popcount(c xor q)`` over the padded row (exact);
  raw score is the inner product of the +-1 sign vectors ``D - 2h`` (int32),
  float similarity ``fl32((D - 2h) / D)`` = their cosine.
* every store therefore has a 32-bit *raw score* (f32 or i32) = inner product
  of the stored codes, larger is better.  Order: raw score descending, ties ->
  lowest row id.  The threshold keeps rows whose float similarity is
  ``>= min_similarity`` (fp32 compare).

``bruteforce_f32`` is the plain ``np.float32`` matmul pass the north star names;
it differs from the canonical score by fp32 summation-order noise only and is
used to report agreement / tolerance, never as the bit-exact comparand.
"""
from __future__ import annotations

import numpy as np

from .encode import decode_rows, encode_rows


def exact_scores_seq(codes: np.ndarray, qcodes: np.ndarray, store: str) -> np.ndarray:
    """Canonical fp64-sequential dot of every stored row with one stored query.

    codes [n, Dp] / qcodes [Dp] in store dtype -> fp64 [n] (before the fl32 rounding)."""
    c = decode_rows(codes, store)
    q = decode_rows(qcodes[None, :], store)[0]
    acc = np.zeros(c.shape[0], dtype=np.float64)
    for j in range(c.shape[1]):
        acc += c[:, j] * q[j]
    return acc


def int_dots(codes: np.ndarray, qcodes: np.ndarray) -> np.ndarray:
    """int8 codes [n, Dp] . [Dp] -> int32 [n] (exact)."""
    return codes.astype(np.int32) @ qcodes.astype(np.int32)


def hamming(words: np.ndarray, qwords: np.ndarray) -> np.ndarray:
    """uint32 words [n, W] vs [W] -> int32 Hamming distance [n] (exact)."""
    x = np.bitwise_xor(words, qwords[None, :])
    return np.bitwise_count(x).sum(axis=1, dtype=np.int64).astype(np.int32)


def i8_similarity(dots: np.ndarray, i8_scale: float = 1.0) -> np.ndarray:
    s2 = np.float32((float(i8_scale) / 127.0) ** 2)
    return dots.astype(np.float32) * s2


def b1_similarity(h: np.ndarray, dim: int) -> np.ndarray:
    return (1.0 - 2.0 * h.astype(np.float64) / float(dim)).astype(np.float32)


def raw_scores(codes: np.ndarray, qcodes: np.ndarray, store: str, dim: int) -> np.ndarray:
    """Raw 32-bit score of every stored row against one stored query:
    f32 (f16/bf16 stores) or i32 (i8: dot; b1: D - 2*hamming)."""
    if store in ("f16", "bf16"):
        return exact_scores_seq(codes, qcodes, store).astype(np.float32)
    if store == "i8":
        return int_dots(codes, qcodes)
    if store == "b1":
        return (np.int32(dim) - 2 * hamming(codes, qcodes)).astype(np.int32)
    raise ValueError(store)


def similarity_from_raw(raw: np.ndarray, store: str, dim: int, i8_scale: float = 1.0) -> np.ndarray:
    """Raw score -> fp32 cosine-domain similarity (what the host turns into a Chroma distance)."""
    if store in ("f16", "bf16"):
        return raw.astype(np.float32)
    if store == "i8":
        return i8_similarity(raw, i8_scale)
    if store == "b1":
        return (raw.astype(np.float64) / float(dim)).astype(np.float32)
    raise ValueError(store)


def select_topk(raw: np.ndarray, sims: np.ndarray, k: int,
                min_similarity: float = -np.inf, row_base: int = 0):
    """Top-k by (raw desc, id asc) among rows with sims >= min_similarity.

    -> (ids uint32 [m], raw [m]) with m <= k."""
    n = raw.shape[0]
    ids = np.arange(n, dtype=np.int64)
    if np.isfinite(min_similarity):
        ids = ids[sims >= np.float32(min_similarity)]
    key = raw[ids].astype(np.float64)                    # f32 and i32 are exact in fp64
    order = np.lexsort((ids, -key))                      # last key is primary
    sel = ids[order[:k]]
    return (sel + row_base).astype(np.uint32), raw[sel]


def pad_id():
    return np.uint32(0xFFFFFFFF)


def pad_raw(dtype):
    return -np.inf if np.dtype(dtype) == np.float32 else np.iinfo(np.int32).min


def search(codes: np.ndarray, qcodes: np.ndarray, store: str, dim: int, k: int,
           min_similarity: float = -np.inf, i8_scale: float = 1.0, row_base: int = 0):
    """Batched canonical search.  qcodes [nq, Dp] -> (ids [nq,k] u32 padded with
    0xFFFFFFFF, raw [nq,k] f32|i32 padded with -inf|INT_MIN, counts [nq] i32)."""
    if qcodes.ndim == 1:
        qcodes = qcodes[None, :]
    nq = qcodes.shape[0]
    rdt = np.float32 if store in ("f16", "bf16") else np.int32
    out_ids = np.full((nq, k), 0xFFFFFFFF, dtype=np.uint32)
    out_raw = np.full((nq, k), pad_raw(rdt), dtype=rdt)
    counts = np.zeros(nq, dtype=np.int32)
    for i in range(nq):
        raw = raw_scores(codes, qcodes[i], store, dim)
        sims = similarity_from_raw(raw, store, dim, i8_scale)
        ids, r = select_topk(raw, sims, k, min_similarity, row_base)
        counts[i] = len(ids)
        out_ids[i, :len(ids)] = ids
        out_raw[i, :len(ids)] = r
    return out_ids, out_raw, counts


def search_f16_shortlist(codes: np.ndarray, qcodes: np.ndarray, k: int,
                         min_similarity: float = -np.inf, row_base: int = 0,
                         slack: float = 1e-9):
    """Same result as ``search(..., store='f16')`` for large n: a BLAS fp64 matmul
    (error <= Dp * 2^-53 << slack) shortlists every row that can reach the top-k,
    then the canonical sequential score is computed on the shortlist only."""
    if qcodes.ndim == 1:
        qcodes = qcodes[None, :]
    c64 = codes.astype(np.float64)
    nq = qcodes.shape[0]
    out_ids = np.full((nq, k), 0xFFFFFFFF, dtype=np.uint32)
    out_sims = np.full((nq, k), -np.inf, dtype=np.float32)
    counts = np.zeros(nq, dtype=np.int32)
    for i in range(nq):
        approx = c64 @ qcodes[i].astype(np.float64)
        kk = min(k, approx.shape[0])
        kth = np.partition(approx, approx.shape[0] - kk)[approx.shape[0] - kk]
        # fl32 rounding can merge scores within half an fp32 ulp; keep those too
        cut = kth - slack - 2.0 ** -23 * max(1.0, abs(kth))
        short = np.nonzero(approx >= cut)[0]
        s = exact_scores_seq(codes[short], qcodes[i], "f16").astype(np.float32)
        ids, ss = select_topk(s, s, k, min_similarity, 0)
        ids = short[ids.astype(np.int64)]
        counts[i] = len(ids)
        out_ids[i, :len(ids)] = (ids + row_base).astype(np.uint32)
        out_sims[i, :len(ids)] = ss
    return out_ids, out_sims, counts


def bruteforce_f32(x: np.ndarray, q: np.ndarray, k: int, metric: str = "cosine"):
    """The plain fp32 brute-force pass on the ORIGINAL fp32 embeddings: normalise
    in fp32, one sgemm, (score desc, id asc).  -> (ids [nq,k] int64, scores fp32)."""
    x = np.asarray(x, dtype=np.float32)
    q = np.atleast_2d(np.asarray(q, dtype=np.float32))
    if metric == "cosine":
        x = x / np.maximum(np.linalg.norm(x, axis=1, keepdims=True), np.float32(1e-30))
        q = q / np.maximum(np.linalg.norm(q, axis=1, keepdims=True), np.float32(1e-30))
    s = q @ x.T
    ids = np.empty((q.shape[0], min(k, x.shape[0])), dtype=np.int64)
    for i in range(q.shape[0]):
        ids[i] = np.lexsort((np.arange(x.shape[0]), -s[i]))[:k]
    return ids, np.take_along_axis(s, ids, axis=1)


def merge_topk(ids_lists: np.ndarray, raw_lists: np.ndarray, k: int):
    """Cross-shard merge (K7): ids/raw [G, nq, k_in] (padding id 0xFFFFFFFF) ->
    ([nq,k] ids, [nq,k] raw, [nq] counts) by (raw desc, id asc)."""
    g, nq, kin = ids_lists.shape
    out_ids = np.full((nq, k), 0xFFFFFFFF, dtype=np.uint32)
    out_raw = np.full((nq, k), pad_raw(raw_lists.dtype), dtype=raw_lists.dtype)
    counts = np.zeros(nq, dtype=np.int32)
    for i in range(nq):
        ids = ids_lists[:, i, :].reshape(-1)
        raw = raw_lists[:, i, :].reshape(-1)
        valid = ids != 0xFFFFFFFF
        ids, raw = ids[valid], raw[valid]
        order = np.lexsort((ids, -raw.astype(np.float64)))[:k]
        counts[i] = len(order)
        out_ids[i, :len(order)] = ids[order]
        out_raw[i, :len(order)] = raw[order]
    return out_ids, out_raw, counts


def encode_queries(q: np.ndarray, store: str, metric: str = "cosine", i8_scale: float = 1.0):
    """Queries go through the same canonical encoder as corpus rows."""
    return encode_rows(np.atleast_2d(np.asarray(q, dtype=np.float32)), store, metric, i8_scale)
