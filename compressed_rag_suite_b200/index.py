"""ShardIndex — Python handle on one ``crs_index`` (one GPU shard of the corpus).

Thin plumbing over the C ABI: numpy arrays go through the host-buffer form of each
call (the library copies and synchronises), torch CUDA tensors through the
device-buffer form (the call only enqueues on torch's current stream).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Tuple

import numpy as np

from . import _native as N


def _is_torch_cuda(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "is_cuda") and x.is_cuda


class ShardIndex:
    def __init__(self, dim: int, dtype: str = "f16", metric: str = "cosine", device: int = 0,
                 row_base: int = 0, reserve_rows: int = 0, _handle=None):
        self._lib = N.lib()
        self._h = C.c_void_p()
        if _handle is not None:
            self._h = _handle
        else:
            if dtype not in ("f16", "bf16", "i8", "b1"):
                raise ValueError(f"dtype must be f16, bf16, i8 or b1, got {dtype!r}")
            if metric not in N.METRIC_CODES:
                raise ValueError(f"metric must be cosine or ip, got {metric!r}")
            N.check(self._lib.crs_index_create(C.byref(self._h), int(dim), N.DTYPE_CODES[dtype],
                                               N.METRIC_CODES[metric], int(device), int(row_base),
                                               int(reserve_rows)))
        d, dp, rb, st, me = C.c_int32(), C.c_int32(), C.c_int64(), C.c_int32(), C.c_int32()
        N.check(self._lib.crs_index_info(self._h, C.byref(d), C.byref(dp), C.byref(rb), C.byref(st), C.byref(me)))
        self.dim, self.dim_padded, self.row_bytes = d.value, dp.value, rb.value
        self.dtype = {v: k for k, v in N.DTYPE_CODES.items()}[st.value]
        self.metric = {v: k for k, v in N.METRIC_CODES.items()}[me.value]
        self.device = int(device)
        self.row_base = int(row_base)
        sc = C.c_double()
        N.check(self._lib.crs_index_similarity_scale(self._h, C.byref(sc)))
        self.similarity_scale = sc.value
        self.is_int = self.dtype in ("i8", "b1")

    # ------------------------------------------------------------------ lifecycle
    def close(self) -> None:
        if self._h:
            self._lib.crs_index_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        n = C.c_int64()
        N.check(self._lib.crs_index_count(self._h, C.byref(n)))
        return n.value

    def _use_torch_stream(self) -> None:
        import torch
        N.check(self._lib.crs_index_set_stream(self._h, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))

    def set_option(self, name: str, value: int) -> None:
        N.check(self._lib.crs_index_set_option(self._h, name.encode(), int(value)))

    def last_stats(self) -> dict:
        s = N.SearchStats()
        N.check(self._lib.crs_index_last_stats(self._h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in s._fields_}

    def last_kernel_ms(self) -> float:
        """Device time of the dominant kernel(s) of the last search (needs set_option('profiling', 1))."""
        ms = C.c_float()
        N.check(self._lib.crs_index_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    def kernel_ms_history(self) -> list:
        """Dominant-kernel device times of up to the last 32 searches, oldest first."""
        buf = (C.c_float * 32)()
        n = C.c_int()
        N.check(self._lib.crs_index_kernel_ms_history(self._h, buf, 32, C.byref(n)))
        return [buf[i] for i in range(n.value)]

    # ------------------------------------------------------------------ ingest
    def add(self, rows) -> None:
        """rows: float32 [n, dim] numpy array (host) or torch CUDA tensor (device)."""
        if _is_torch_cuda(rows):
            import torch
            if rows.dtype != torch.float32 or rows.dim() != 2 or rows.shape[1] != self.dim:
                raise ValueError(f"rows must be float32 [n, {self.dim}]")
            rows = rows.contiguous()
            self._use_torch_stream()
            N.check(self._lib.crs_index_add(self._h, C.c_void_p(rows.data_ptr()), rows.shape[0], N.CRS_F32))
            return
        a = np.ascontiguousarray(rows, dtype=np.float32)
        if a.ndim != 2 or a.shape[1] != self.dim:
            raise ValueError(f"rows must be float32 [n, {self.dim}], got {a.shape}")
        N.check(self._lib.crs_index_add(self._h, a.ctypes.data_as(C.c_void_p), a.shape[0], N.CRS_F32))

    # ------------------------------------------------------------------ search
    def search_into(self, queries, k: int, min_similarity, ids, scores, counts) -> None:
        """Device-buffer search writing into caller-provided CUDA tensors (ids int32 [nq,k], scores
        f32|i32 [nq,k] — any 32-bit dtype view —, counts int32 [nq]); enqueued on the current stream."""
        q = queries.contiguous()
        if q.dim() == 1:
            q = q[None, :]
        self._use_torch_stream()
        N.check(self._lib.crs_index_search(self._h, C.c_void_p(q.data_ptr()), q.shape[0], int(k), float(min_similarity),
                                           C.c_void_p(ids.data_ptr()), C.c_void_p(scores.data_ptr()),
                                           C.c_void_p(counts.data_ptr())))

    def search(self, queries, k: int, min_similarity: float = -math.inf, allow=None, out=None):
        """-> (ids uint32 [nq,k], raw scores f32|i32 [nq,k], counts i32 [nq]).

        out: optional preallocated host result arrays (ids, scores, counts) for the numpy form —
        e.g. views of pinned memory, so the device->host copies of the call are asynchronous.

        numpy in -> numpy out (synchronous); torch CUDA in -> torch CUDA out (enqueued on
        the current stream).  ids are global row ids, padded with 0xFFFFFFFF."""
        if _is_torch_cuda(queries):
            import torch
            q = queries
            if q.dim() == 1:
                q = q[None, :]
            if q.dtype != torch.float32 or q.shape[1] != self.dim:
                raise ValueError(f"queries must be float32 [nq, {self.dim}]")
            q = q.contiguous()
            nq = q.shape[0]
            dev = q.device
            # torch has no uint32 arithmetic: ids travel as int32 bit patterns
            ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
            sc = torch.empty((nq, k), dtype=torch.int32 if self.is_int else torch.float32, device=dev)
            cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
            self._use_torch_stream()
            if allow is None:
                N.check(self._lib.crs_index_search(self._h, C.c_void_p(q.data_ptr()), nq, int(k), float(min_similarity),
                                                   C.c_void_p(ids.data_ptr()), C.c_void_p(sc.data_ptr()),
                                                   C.c_void_p(cnt.data_ptr())))
            else:
                bits = allow if (isinstance(allow, np.ndarray) and allow.dtype == np.uint32) else self.pack_allow(allow)
                N.check(self._lib.crs_index_search_filtered(self._h, C.c_void_p(q.data_ptr()), nq, int(k),
                                                            float(min_similarity), bits.ctypes.data_as(C.c_void_p),
                                                            C.c_void_p(ids.data_ptr()), C.c_void_p(sc.data_ptr()),
                                                            C.c_void_p(cnt.data_ptr())))
            return ids, sc, cnt
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be float32 [nq, {self.dim}], got {q.shape}")
        nq = q.shape[0]
        if out is not None:
            ids, sc, cnt = out
            if (ids.shape != (nq, k) or sc.shape != (nq, k) or cnt.shape != (nq,) or ids.dtype.itemsize != 4
                    or sc.dtype.itemsize != 4 or cnt.dtype != np.int32
                    or not (ids.flags.c_contiguous and sc.flags.c_contiguous and cnt.flags.c_contiguous)):
                raise ValueError("out must be C-contiguous (ids [nq,k] 32-bit, scores [nq,k] 32-bit, counts [nq] int32)")
        else:
            ids = np.empty((nq, k), dtype=np.uint32)
            sc = np.empty((nq, k), dtype=np.int32 if self.is_int else np.float32)
            cnt = np.empty((nq,), dtype=np.int32)
        if allow is None:
            N.check(self._lib.crs_index_search(self._h, q.ctypes.data_as(C.c_void_p), nq, int(k), float(min_similarity),
                                               ids.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p),
                                               cnt.ctypes.data_as(C.c_void_p)))
        else:
            # a ready-made bitmap (uint32 words, e.g. a cached predicate) or a bool mask over the rows
            bits = allow if (isinstance(allow, np.ndarray) and allow.dtype == np.uint32) else self.pack_allow(allow)
            if bits.shape[0] != (len(self) + 31) // 32:
                raise ValueError("allow bitmap has the wrong number of words")
            N.check(self._lib.crs_index_search_filtered(self._h, q.ctypes.data_as(C.c_void_p), nq, int(k),
                                                        float(min_similarity), bits.ctypes.data_as(C.c_void_p),
                                                        ids.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p),
                                                        cnt.ctypes.data_as(C.c_void_p)))
        return ids, sc, cnt

    # ------------------------------------------------------------------ sharded search over peer memory
    def search_sharded(self, exchange, queries, k: int, min_similarity: float = -math.inf, out=None):
        """Local search + exchange + merge in the library (crs_index_search_sharded): every rank gets the
        GLOBAL top-k.  torch CUDA in -> torch CUDA out (enqueued); numpy in -> numpy out (synchronous).
        out: optional preallocated (ids, scores, counts) of the matching kind."""
        if _is_torch_cuda(queries):
            import torch
            q = queries if queries.dim() == 2 else queries[None, :]
            if q.dtype != torch.float32 or q.shape[1] != self.dim:
                raise ValueError(f"queries must be float32 [nq, {self.dim}]")
            q = q.contiguous()
            nq = q.shape[0]
            if out is not None:
                ids, sc, cnt = out
            else:
                ids = torch.empty((nq, k), dtype=torch.int32, device=q.device)
                sc = torch.empty((nq, k), dtype=torch.int32 if self.is_int else torch.float32, device=q.device)
                cnt = torch.empty((nq,), dtype=torch.int32, device=q.device)
            self._use_torch_stream()
            N.check(self._lib.crs_index_search_sharded(self._h, exchange._h, C.c_void_p(q.data_ptr()), nq, int(k),
                                                       float(min_similarity), C.c_void_p(ids.data_ptr()),
                                                       C.c_void_p(sc.data_ptr()), C.c_void_p(cnt.data_ptr())))
            return ids, sc, cnt
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be float32 [nq, {self.dim}], got {q.shape}")
        nq = q.shape[0]
        if out is not None:
            ids, sc, cnt = out
        else:
            ids = np.empty((nq, k), dtype=np.uint32)
            sc = np.empty((nq, k), dtype=np.int32 if self.is_int else np.float32)
            cnt = np.empty((nq,), dtype=np.int32)
        N.check(self._lib.crs_index_search_sharded(self._h, exchange._h, q.ctypes.data_as(C.c_void_p), nq, int(k),
                                                   float(min_similarity), ids.ctypes.data_as(C.c_void_p),
                                                   sc.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p)))
        return ids, sc, cnt

    def search_push(self, exchange, queries, k: int, min_similarity: float = -math.inf, allow=None) -> None:
        """First half of a sharded search for a host that drives several GPUs from one thread: local
        search + stores into every peer's receive buffer; waits for nobody (torch CUDA queries).
        allow: optional bool mask / uint32 bitmap over this shard's local rows."""
        q = queries if queries.dim() == 2 else queries[None, :]
        q = q.contiguous()
        bits = None
        if allow is not None:
            bits = allow if (isinstance(allow, np.ndarray) and allow.dtype == np.uint32) else self.pack_allow(allow)
        self._use_torch_stream()
        N.check(self._lib.crs_index_search_push(self._h, exchange._h, C.c_void_p(q.data_ptr()), q.shape[0], int(k),
                                                float(min_similarity),
                                                bits.ctypes.data_as(C.c_void_p) if bits is not None else None))

    def map_ids(self, first_row: int, n: int, first_global_id: int) -> None:
        """Local rows [first_row, first_row + n) report the global ids first_global_id.. in searches."""
        N.check(self._lib.crs_index_map_ids(self._h, int(first_row), int(n), int(first_global_id)))

    def capture_search(self, nq: int, k: int, min_similarity: float = -math.inf):
        """CUDA-graph form of the device-buffer search for latency-bound callers (a single query over a
        small shard is ~5 short kernels: their launch cost, not their run time, bounds the call).
        Returns a GraphedSearch: copy queries into ``.queries`` ([nq, dim] float32 CUDA), call
        ``.replay()``, read ``.ids / .scores / .counts``.  The index must not grow while the graph
        is in use (the captured launches hold its row count and buffer addresses)."""
        import torch
        dev = torch.device("cuda", self.device)
        # warm-up / capture queries: random, NOT zeros (an all-zero query ties every row, which sends every query
        # through the exhaustive fallback: seconds of warm-up on a large shard)
        gen = torch.Generator(device=dev)
        gen.manual_seed(7)
        q = torch.randn((nq, self.dim), dtype=torch.float32, device=dev, generator=gen)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                      # warm-up: sizes every scratch buffer the search needs
            for _ in range(2):
                self.search(q, k, min_similarity)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.set_option("profiling", 0)                    # event pairs cannot be read back from a graph
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            ids, sc, cnt = self.search(q, k, min_similarity)
        self._use_torch_stream()                           # the capture stream is gone: back to the caller's stream
        return GraphedSearch(graph, q, ids, sc, cnt, len(self))

    def pack_allow(self, allow) -> np.ndarray:
        """bool mask over the local rows -> uint32 bitmap (bit r%32 of word r/32)."""
        a = np.asarray(allow, dtype=bool).reshape(-1)
        n = len(self)
        if a.shape[0] != n:
            raise ValueError(f"allow mask has {a.shape[0]} entries, index has {n} rows")
        pad = (-n) % 32
        if pad:
            a = np.concatenate([a, np.zeros(pad, dtype=bool)])
        return np.packbits(a.reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).reshape(-1).copy()

    def similarity(self, raw):
        """raw scores -> float32 cosine-domain similarity (numpy)."""
        raw = np.asarray(raw)
        if self.dtype == "i8":
            return raw.astype(np.float32) * np.float32(self.similarity_scale)
        if self.dtype == "b1":
            return (raw.astype(np.float64) / float(self.dim)).astype(np.float32)
        return raw.astype(np.float32)

    # ------------------------------------------------------------------ candidate vectors / MMR
    def code_dtype(self):
        return {"f16": np.float16, "bf16": np.uint16, "i8": np.int8, "b1": np.uint32}[self.dtype]

    def fetch_rows(self, ids, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Stored codes of the given global row ids -> [n, row_bytes] uint8 (host).
        Rows owned by other shards are left as they are in ``out`` (zeros by default)."""
        ids = np.ascontiguousarray(ids, dtype=np.uint32).reshape(-1)
        if out is None:
            out = np.zeros((ids.shape[0], self.row_bytes), dtype=np.uint8)
        N.check(self._lib.crs_index_fetch_rows(self._h, ids.ctypes.data_as(C.c_void_p), ids.shape[0],
                                               out.ctypes.data_as(C.c_void_p)))
        return out

    def fetch_rows_device(self, ids, out=None):
        """Device form of fetch_rows: ids int32 CUDA tensor (uint32 bit patterns) of any shape ->
        uint8 CUDA tensor ids.shape + (row_bytes,); rows of other shards stay as they are in `out`
        (zeros by default).  Enqueued on the current stream."""
        import torch
        ids = ids.contiguous()
        if out is None:
            out = torch.zeros(tuple(ids.shape) + (self.row_bytes,), dtype=torch.uint8, device=ids.device)
        self._use_torch_stream()
        N.check(self._lib.crs_index_fetch_rows(self._h, C.c_void_p(ids.data_ptr()), ids.numel(),
                                               C.c_void_p(out.data_ptr())))
        return out

    def score_rows(self, queries, ids):
        """K8: canonical score of every (query q, row ids[q, j]) pair on this shard.
        queries [nq, dim] float32, ids [nq, m] (numpy uint32 or torch CUDA int32 bit patterns)
        -> [nq, m] float32 | int32; rows this shard does not hold -> -inf / INT32_MIN."""
        if _is_torch_cuda(queries):
            import torch
            q = queries.contiguous()
            ids = ids.contiguous()
            nq, m = ids.shape
            out = torch.empty((nq, m), dtype=torch.int32 if self.is_int else torch.float32, device=q.device)
            self._use_torch_stream()
            N.check(self._lib.crs_index_score_rows(self._h, C.c_void_p(q.data_ptr()), nq, C.c_void_p(ids.data_ptr()), m,
                                                   C.c_void_p(out.data_ptr())))
            return out
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        i = np.ascontiguousarray(ids, dtype=np.uint32)
        if i.ndim == 1:
            i = i[None, :]
        if q.shape[1] != self.dim or i.shape[0] != q.shape[0]:
            raise ValueError(f"queries must be [nq, {self.dim}] and ids [nq, m]")
        nq, m = i.shape
        out = np.empty((nq, m), dtype=np.int32 if self.is_int else np.float32)
        N.check(self._lib.crs_index_score_rows(self._h, q.ctypes.data_as(C.c_void_p), nq, i.ctypes.data_as(C.c_void_p), m,
                                               out.ctypes.data_as(C.c_void_p)))
        return out

    def score_vectors(self, queries, rows):
        """K8 on caller-supplied candidate vectors: queries [nq, dim], rows [nq, m, dim] float32
        (both numpy or both torch CUDA) -> canonical scores [nq, m] of this index's store dtype."""
        if _is_torch_cuda(queries):
            import torch
            q = queries.contiguous()
            r = rows.contiguous()
            nq, m = r.shape[0], r.shape[1]
            if r.dtype != torch.float32 or r.shape[2] != self.dim or q.shape[0] != nq:
                raise ValueError(f"rows must be float32 [nq, m, {self.dim}]")
            out = torch.empty((nq, m), dtype=torch.int32 if self.is_int else torch.float32, device=q.device)
            self._use_torch_stream()
            N.check(self._lib.crs_index_score_vectors(self._h, C.c_void_p(q.data_ptr()), nq, C.c_void_p(r.data_ptr()), m,
                                                      C.c_void_p(out.data_ptr())))
            return out
        q = np.ascontiguousarray(queries, dtype=np.float32)
        r = np.ascontiguousarray(rows, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if r.ndim != 3 or r.shape[2] != self.dim or r.shape[0] != q.shape[0]:
            raise ValueError(f"rows must be float32 [nq, m, {self.dim}]")
        nq, m = r.shape[0], r.shape[1]
        out = np.empty((nq, m), dtype=np.int32 if self.is_int else np.float32)
        N.check(self._lib.crs_index_score_vectors(self._h, q.ctypes.data_as(C.c_void_p), nq, r.ctypes.data_as(C.c_void_p), m,
                                                  out.ctypes.data_as(C.c_void_p)))
        return out

    def mmr_device(self, vecs, relevance, lam: float, k_out: int):
        """Device form of mmr: vecs uint8 CUDA [nq, m, row_bytes], relevance float64 CUDA [nq, m]
        -> int32 CUDA [nq, k_out]; enqueued on the current stream."""
        import torch
        v = vecs.contiguous()
        r = relevance.contiguous()
        nq, m = r.shape
        out = torch.empty((nq, int(k_out)), dtype=torch.int32, device=v.device)
        self._use_torch_stream()
        N.check(self._lib.crs_mmr(self._h, C.c_void_p(v.data_ptr()), C.c_void_p(r.data_ptr()), nq, m, int(k_out),
                                  float(lam), C.c_void_p(out.data_ptr())))
        return out

    def mmr_select(self, vecs, ids, raw, counts, lam: float, k_out: int):
        """Fused search -> MMR finish (torch CUDA tensors): candidate codes [nq, m, row_bytes], the search
        output (ids int32 bit patterns [nq, m], raw scores [nq, m], counts [nq]) -> (ids [nq,k_out] (pad -1),
        similarity f32 [nq,k_out], reference score f64 [nq,k_out], counts [nq]) in greedy MMR order."""
        import torch
        v, i, r, c = vecs.contiguous(), ids.contiguous(), raw.contiguous(), counts.contiguous()
        nq, m = i.shape
        dev = i.device
        o_ids = torch.empty((nq, int(k_out)), dtype=torch.int32, device=dev)
        o_sim = torch.empty((nq, int(k_out)), dtype=torch.float32, device=dev)
        o_rel = torch.empty((nq, int(k_out)), dtype=torch.float64, device=dev)
        o_cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
        self._use_torch_stream()
        N.check(self._lib.crs_mmr_select(self._h, C.c_void_p(v.data_ptr()), C.c_void_p(i.data_ptr()), C.c_void_p(r.data_ptr()),
                                         C.c_void_p(c.data_ptr()), nq, m, int(k_out), float(lam), C.c_void_p(o_ids.data_ptr()),
                                         C.c_void_p(o_sim.data_ptr()), C.c_void_p(o_rel.data_ptr()), C.c_void_p(o_cnt.data_ptr())))
        return o_ids, o_sim, o_rel, o_cnt

    def mmr(self, vecs: np.ndarray, relevance, lam: float, k_out: Optional[int] = None) -> np.ndarray:
        """Greedy MMR order.  vecs: [nq, m, row_bytes] uint8 stored codes (or [m, row_bytes]);
        relevance: [nq, m] float64.  -> int32 [nq, k_out] positions."""
        v = np.ascontiguousarray(vecs, dtype=np.uint8)
        if v.ndim == 2:
            v = v[None]
        r = np.ascontiguousarray(relevance, dtype=np.float64).reshape(v.shape[0], -1)
        nq, m = r.shape
        if v.shape[1] != m or v.shape[2] != self.row_bytes:
            raise ValueError("vecs must be [nq, m, row_bytes]")
        k_out = m if k_out is None else int(k_out)
        out = np.empty((nq, k_out), dtype=np.int32)
        N.check(self._lib.crs_mmr(self._h, v.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p), nq, m, k_out,
                                  float(lam), out.ctypes.data_as(C.c_void_p)))
        return out

    # ------------------------------------------------------------------ persistence
    def save(self, path: str) -> None:
        """Whole index -> path, atomically (temp file + rename)."""
        N.check(self._lib.crs_index_save(self._h, path.encode()))

    def append_to(self, path: str) -> None:
        """Append the rows `path` does not hold yet, then advance its header (crs_index_append)."""
        N.check(self._lib.crs_index_append(self._h, path.encode()))

    def truncate(self, new_count: int) -> None:
        N.check(self._lib.crs_index_truncate(self._h, int(new_count)))

    @classmethod
    def load(cls, path: str, device: int = 0, row_base: int = 0) -> "ShardIndex":
        lib = N.lib()
        h = C.c_void_p()
        N.check(lib.crs_index_load(C.byref(h), path.encode(), int(device), int(row_base)))
        return cls(0, device=device, row_base=row_base, _handle=h)


class GraphedSearch:
    """One captured search (see ShardIndex.capture_search)."""

    def __init__(self, graph, queries, ids, scores, counts, rows):
        self.graph, self.queries, self.ids, self.scores, self.counts, self.rows = graph, queries, ids, scores, counts, rows

    def replay(self):
        self.graph.replay()
        return self.ids, self.scores, self.counts


def select_topk(ids, scores, k_out: int):
    """Order UNSORTED candidates (torch CUDA: ids int32 bit patterns [nq, m], scores f32|i32 [nq, m])
    by (score desc, id asc) -> (ids [nq,k_out], scores [nq,k_out], counts [nq]); pad ids skipped."""
    import torch
    nq, m = ids.shape
    ids = ids.contiguous()
    scores = scores.contiguous()
    is_int = scores.dtype == torch.int32
    out_ids = torch.empty((nq, k_out), dtype=torch.int32, device=ids.device)
    out_sc = torch.empty((nq, k_out), dtype=scores.dtype, device=ids.device)
    out_cnt = torch.empty((nq,), dtype=torch.int32, device=ids.device)
    st = torch.cuda.current_stream(ids.device).cuda_stream
    N.check(N.lib().crs_select_topk(C.c_void_p(st), C.c_void_p(ids.data_ptr()), C.c_void_p(scores.data_ptr()),
                                    int(is_int), nq, m, int(k_out), C.c_void_p(out_ids.data_ptr()),
                                    C.c_void_p(out_sc.data_ptr()), C.c_void_p(out_cnt.data_ptr())))
    return out_ids, out_sc, out_cnt


def merge_gathered(gathered, nq: int, k_in: int, k_out: int, is_int: bool):
    """K7 straight out of an allgathered buffer: `gathered` int32 CUDA [G, 2, nq, k_in] where [:, 0] are
    the ranks' id blocks and [:, 1] their raw-score bits -> (ids, scores, counts); no copies."""
    import torch
    g = gathered.shape[0]
    dev = gathered.device
    out_ids = torch.empty((nq, k_out), dtype=torch.int32, device=dev)
    out_sc = torch.empty((nq, k_out), dtype=torch.int32 if is_int else torch.float32, device=dev)
    out_cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    base = gathered.data_ptr()
    N.check(N.lib().crs_merge_topk_strided(C.c_void_p(st), C.c_void_p(base), C.c_void_p(base + nq * k_in * 4), int(is_int),
                                           g, nq, k_in, int(k_out), 2 * nq * k_in, C.c_void_p(out_ids.data_ptr()),
                                           C.c_void_p(out_sc.data_ptr()), C.c_void_p(out_cnt.data_ptr())))
    return out_ids, out_sc, out_cnt


def merge_topk(ids, scores, k_out: int):
    """K7 on torch CUDA tensors: ids int32 [G, nq, k_in] (uint32 bit patterns), scores
    f32|i32 [G, nq, k_in] -> (ids [nq,k_out], scores [nq,k_out], counts [nq])."""
    import torch
    g, nq, k_in = ids.shape
    ids = ids.contiguous()
    scores = scores.contiguous()
    is_int = scores.dtype == torch.int32
    out_ids = torch.empty((nq, k_out), dtype=torch.int32, device=ids.device)
    out_sc = torch.empty((nq, k_out), dtype=scores.dtype, device=ids.device)
    out_cnt = torch.empty((nq,), dtype=torch.int32, device=ids.device)
    st = torch.cuda.current_stream(ids.device).cuda_stream
    N.check(N.lib().crs_merge_topk(C.c_void_p(st), C.c_void_p(ids.data_ptr()), C.c_void_p(scores.data_ptr()),
                                   int(is_int), g, nq, k_in, int(k_out), C.c_void_p(out_ids.data_ptr()),
                                   C.c_void_p(out_sc.data_ptr()), C.c_void_p(out_cnt.data_ptr())))
    return out_ids, out_sc, out_cnt
