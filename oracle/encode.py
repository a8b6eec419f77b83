"""Oracle: canonical ingest arithmetic (normalise + cast / quantise / sign-pack).

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.

The reference stores fp32 MiniLM embeddings in a Chroma collection created with
``metadata={"hnsw:space": "cosine"}`` (``rag/indexing.py:81-84,114-119``); Chroma
normalises for cosine space.  The reference has no embedding quantisation at all
(SURVEY.md §8c "Quantisation definitions"), so the stored formats are frozen
here first and the CUDA ingest kernel (``csrc/ingest.cu``) must reproduce them
bit for bit:

* normalisation (cosine only): ``n2 = sum_j x_j^2`` accumulated **sequentially
  in fp64** (j = 0..D-1), ``y_j = x_j / sqrt(n2)`` in fp64; an all-zero row stays
  zero.  Inner-product space stores ``y = x`` unchanged.
* ``f16`` / ``bf16``: ``y`` rounded once (round-to-nearest-even) from fp64.
* ``i8``: ``code = clip(rint(y * 127 / a), -127, 127)`` with one global scale
  ``a`` (default 1.0: unit-norm rows have ``|y| <= 1``); ``rint`` is
  round-half-to-even.  Float score of two codes = ``dot * (a/127)^2``.
* ``b1``: bit j = ``y_j > 0``, packed little-endian into uint32 words
  (bit ``j % 32`` of word ``j // 32``).

Every function is vectorised over rows but keeps the per-row operation order
stated above, so CPU and GPU results are identical IEEE-754 values.
"""
from __future__ import annotations

import numpy as np

STORE_DTYPES = ("f16", "bf16", "i8", "b1")


_FLOAT_NCH = (1, 2, 3, 4, 6, 8, 12, 16)
_INT_NCH = (1, 2, 3, 4, 6, 8)


def padded_dim(dim: int, store: str) -> int:
    """Stored row width in elements.  Rows are zero-padded so that a row is one of the
    widths the scan kernels are instantiated for: 128-byte units x {1,2,3,4,6,8[,12,16]}
    (f16/bf16: 64 elements per unit, i8: 128, b1: 1024 bits)."""
    unit = 64 if store in ("f16", "bf16") else (128 if store == "i8" else 1024)
    for nch in (_FLOAT_NCH if store in ("f16", "bf16") else _INT_NCH):
        if nch * unit >= dim:
            return nch * unit
    raise ValueError(f"dim {dim} too large for store {store}")


def normalise_rows(x: np.ndarray, metric: str = "cosine") -> np.ndarray:
    """fp32 [n, D] -> fp64 [n, D] canonical unit rows (cosine) or a plain upcast (ip)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim == 1:
        x = x[None, :]
    x64 = x.astype(np.float64)
    if metric == "ip":
        return x64
    if metric != "cosine":
        raise ValueError(f"unsupported metric {metric!r}")
    n2 = np.zeros(x64.shape[0], dtype=np.float64)
    for j in range(x64.shape[1]):          # sequential order is part of the definition
        n2 += x64[:, j] * x64[:, j]        # product of two fp32 values is exact in fp64
    norm = np.sqrt(n2)
    safe = np.where(norm > 0.0, norm, 1.0)
    y = x64 / safe[:, None]
    y[norm == 0.0] = 0.0
    return y


def f64_to_bf16_bits(y: np.ndarray) -> np.ndarray:
    """Correctly rounded (single RNE) fp64 -> bf16, returned as uint16 bit patterns.

    Goes through fp32 with round-to-odd so the second rounding cannot double-round
    (the same construction CUDA's ``__double2bfloat16`` uses)."""
    y = np.asarray(y, dtype=np.float64)
    f = y.astype(np.float32)
    inexact = f.astype(np.float64) != y
    # truncate toward zero where the RNE conversion rounded away from zero
    away = inexact & (np.abs(f.astype(np.float64)) > np.abs(y))
    bits = f.view(np.uint32).copy()
    bits[away] -= 1                      # one ulp toward zero (same sign, magnitude bits)
    bits[inexact] |= 1                   # sticky -> odd
    rnd = ((bits >> 16) & 1) + np.uint32(0x7FFF)
    out = ((bits + rnd) >> 16).astype(np.uint16)
    return out


def bf16_bits_to_f64(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << 16).view(np.float32).astype(np.float64)


def encode_rows(x: np.ndarray, store: str = "f16", metric: str = "cosine",
                i8_scale: float = 1.0) -> np.ndarray:
    """fp32 [n, D] -> stored codes [n, Dp] (Dp = padded_dim) in the store dtype.

    f16 -> np.float16, bf16 -> np.uint16 bit patterns, i8 -> np.int8,
    b1 -> np.uint32 words [n, Dp/32].
    """
    y = normalise_rows(x, metric)
    n, d = y.shape
    dp = padded_dim(d, store)
    if store == "f16":
        out = np.zeros((n, dp), dtype=np.float16)
        out[:, :d] = y.astype(np.float16)
        return out
    if store == "bf16":
        out = np.zeros((n, dp), dtype=np.uint16)
        out[:, :d] = f64_to_bf16_bits(y)
        return out
    if store == "i8":
        out = np.zeros((n, dp), dtype=np.int8)
        q = np.rint(y * (127.0 / float(i8_scale)))
        out[:, :d] = np.clip(q, -127.0, 127.0).astype(np.int8)
        return out
    if store == "b1":
        bits = np.zeros((n, dp), dtype=np.uint8)
        bits[:, :d] = (y > 0.0)
        words = bits.reshape(n, dp // 32, 32).astype(np.uint32)
        shifts = np.arange(32, dtype=np.uint32)
        return (words << shifts).sum(axis=2, dtype=np.uint64).astype(np.uint32)
    raise ValueError(f"unsupported store dtype {store!r}")


def decode_rows(codes: np.ndarray, store: str) -> np.ndarray:
    """stored codes -> fp64 values (i8: integer codes; b1: +1 / -1 per bit, pad bits -> -1)."""
    if store == "f16":
        return codes.astype(np.float64)
    if store == "bf16":
        return bf16_bits_to_f64(codes)
    if store == "i8":
        return codes.astype(np.float64)
    if store == "b1":
        n, w = codes.shape
        shifts = np.arange(32, dtype=np.uint32)
        bits = ((codes[:, :, None] >> shifts) & 1).reshape(n, w * 32)
        return bits.astype(np.float64) * 2.0 - 1.0
    raise ValueError(store)
