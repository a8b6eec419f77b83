"""MultiDeviceIndex — one collection dealt out over several GPUs of one box, driven by ONE host thread.

This is what sits behind ``VectorStore({"devices": [0, 1, ...]})``: the reference builds one ``VectorStore``
in one process (rag/pipeline.py:66-74), so the multi-GPU search has to be reachable from there and not only
from ``torchrun`` (one process per GPU, ``sharded.ShardedSearcher``).

* every device holds one ``ShardIndex``; each ``add`` deals its rows out in G contiguous pieces, and
  ``crs_index_map_ids`` makes every shard report the INSERTION index as the global id, so "ties -> first
  inserted row" holds across devices exactly as on one device;
* a search is, per device, ``crs_index_search_push`` (local exact top-k + NVLink peer stores of the
  ``k x (id, score)`` candidates into every peer's receive buffer — no waiting), then ONE
  ``crs_exchange_merge`` on the first device (waits for the pushes of this step and merges).  Everything is
  enqueued asynchronously; the only host synchronisation is the final copy of ``[nq, k]`` results;
* ``fetch_rows`` addresses rows by global id and is routed to the owning shard (the MMR step reads the
  stored vectors of <= 2k survivors; no second collective).

The same device may appear several times (``devices=[0, 0]``): the shards then live side by side on one GPU,
which is how the multi-device path is tested on a one-GPU box.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np

from .index import ShardIndex
from .sharded import PeerExchange


class MultiDeviceIndex:
    def __init__(self, dim: int, dtype: str = "f16", metric: str = "cosine", devices: Optional[List[int]] = None,
                 max_nq: int = 256, max_k: int = 128):
        devices = [0] if not devices else [int(d) for d in devices]
        self.devices = devices
        self.shards = [ShardIndex(dim, dtype=dtype, metric=metric, device=d) for d in devices]
        s0 = self.shards[0]
        self.dim, self.dim_padded, self.row_bytes = s0.dim, s0.dim_padded, s0.row_bytes
        self.dtype, self.metric, self.is_int, self.similarity_scale = s0.dtype, s0.metric, s0.is_int, s0.similarity_scale
        self.device = devices[0]
        self._count = 0
        # segments: (first_global_id, n, shard, first_local_row), in insertion order
        self._seg_start: List[int] = []
        self._seg: List[tuple] = []
        self._cap = (int(max_nq), int(max_k))
        self._ex: Optional[List[PeerExchange]] = None
        self._next = 0                                   # shard that receives the next single row

    # ------------------------------------------------------------------ lifecycle
    def close(self) -> None:
        if self._ex:
            for e in self._ex:
                e.close()
            self._ex = None
        for s in self.shards:
            s.close()

    def __len__(self) -> int:
        return self._count

    def set_option(self, name: str, value: int) -> None:
        for s in self.shards:
            s.set_option(name, value)

    def similarity(self, raw):
        return self.shards[0].similarity(raw)

    def last_stats(self) -> dict:
        return self.shards[0].last_stats()

    # ------------------------------------------------------------------ ingest
    def add(self, rows) -> None:
        """rows: float32 [n, dim] numpy array or torch CUDA tensor (any device of the box)."""
        n = int(rows.shape[0])
        if n == 0:
            return
        g = len(self.shards)
        if n < g:                                        # a handful of rows: deal them out one by one
            pieces = [(i, i + 1, (self._next + i) % g) for i in range(n)]
            self._next = (self._next + n) % g
        else:
            per = (n + g - 1) // g
            pieces = [(lo, min(lo + per, n), j) for j, lo in enumerate(range(0, n, per))]
        for lo, hi, j in pieces:
            sh = self.shards[j]
            piece = rows[lo:hi]
            if hasattr(piece, "is_cuda") and piece.is_cuda and piece.device.index != sh.device:
                piece = piece.to(f"cuda:{sh.device}")    # peer copy over NVLink
            first_local = len(sh)
            sh.add(piece)
            sh.map_ids(first_local, hi - lo, self._count + lo)
            self._seg_start.append(self._count + lo)
            self._seg.append((self._count + lo, hi - lo, j, first_local))
        self._count += n

    # ------------------------------------------------------------------ id routing
    def _locate(self, gids: np.ndarray):
        """global ids -> (shard index, local row) arrays (pad ids -> shard -1)."""
        gids = np.asarray(gids, dtype=np.int64).reshape(-1)
        starts = np.asarray(self._seg_start, dtype=np.int64)
        seg = np.searchsorted(starts, gids, side="right") - 1
        shard = np.full(gids.shape, -1, dtype=np.int64)
        local = np.zeros(gids.shape, dtype=np.int64)
        ok = (gids >= 0) & (gids < self._count) & (seg >= 0)
        if ok.any():
            tab = np.asarray(self._seg, dtype=np.int64)          # [n_seg, 4]
            s = tab[seg[ok]]
            shard[ok] = s[:, 2]
            local[ok] = s[:, 3] + (gids[ok] - s[:, 0])
        return shard, local

    def _local_allow(self, allow) -> List[Optional[np.ndarray]]:
        """bool mask over the GLOBAL rows -> one bool mask per shard over its local rows."""
        a = np.asarray(allow, dtype=bool).reshape(-1)
        if a.shape[0] != self._count:
            raise ValueError(f"allow mask has {a.shape[0]} entries, index has {self._count} rows")
        out = [np.zeros(len(s), dtype=bool) for s in self.shards]
        for first, n, j, first_local in self._seg:
            out[j][first_local:first_local + n] = a[first:first + n]
        return out

    # ------------------------------------------------------------------ search
    def _exchanges(self, nq: int, k: int) -> List[PeerExchange]:
        if self._ex is None or nq > self._ex[0].max_nq or k > self._ex[0].max_k:
            if self._ex:
                for e in self._ex:
                    e.close()
            cap_nq, cap_k = max(nq, self._cap[0]), min(128, max(k, self._cap[1]))
            g = len(self.shards)
            self._ex = [PeerExchange(self.devices[r], r, g, cap_nq, cap_k) for r in range(g)]
            PeerExchange.wire_local(self._ex)
        return self._ex

    def search(self, queries, k: int, min_similarity: float = -math.inf, allow=None, out=None):
        """Same contract as ShardIndex.search: numpy in -> numpy out (synchronous); torch CUDA in -> torch
        CUDA out on the first device.  ids are insertion indices (global), padded with 0xFFFFFFFF."""
        import torch
        is_t = hasattr(queries, "is_cuda") and queries.is_cuda
        if is_t:
            q0 = queries if queries.dim() == 2 else queries[None, :]
            if q0.dtype != torch.float32 or q0.shape[1] != self.dim:
                raise ValueError(f"queries must be float32 [nq, {self.dim}]")
        else:
            q0 = np.ascontiguousarray(queries, dtype=np.float32)
            if q0.ndim == 1:
                q0 = q0[None, :]
            if q0.ndim != 2 or q0.shape[1] != self.dim:
                raise ValueError(f"queries must be float32 [nq, {self.dim}], got {q0.shape}")
        if k <= 0:
            raise ValueError("crs: nq must be >= 0 and k > 0")
        nq = int(q0.shape[0])
        exs = self._exchanges(nq, k)
        allows = self._local_allow(allow) if allow is not None else [None] * len(self.shards)
        # one copy of the queries per device (pinned host -> device, or a peer copy), all asynchronous
        if not is_t:
            host = torch.from_numpy(q0)
        # (all copies first: a device-to-device copy is ordered behind the work already enqueued on its source device,
        # so copying after device 0's search has been enqueued would serialise the devices)
        qds = []
        for sh in self.shards:
            dev = torch.device("cuda", sh.device)
            with torch.cuda.device(dev):
                qds.append(q0.to(dev, non_blocking=True) if is_t else host.to(dev, non_blocking=True))
        for r, sh in enumerate(self.shards):
            with torch.cuda.device(torch.device("cuda", sh.device)):
                sh.search_push(exs[r], qds[r], k, min_similarity, allow=allows[r])
        with torch.cuda.device(torch.device("cuda", self.devices[0])):
            ids, sc, cnt = exs[0].merge(nq, k, self.is_int)
            if is_t:
                return ids, sc, cnt
            torch.cuda.current_stream().synchronize()
            res = (ids.cpu().numpy().view(np.uint32), sc.cpu().numpy(), cnt.cpu().numpy())
        if out is not None:
            for dst, src in zip(out, res):
                dst[...] = src.view(dst.dtype) if dst.dtype != src.dtype else src
            return out
        return res

    def exchange_status(self):
        return [e.status() for e in (self._ex or [])]

    # ------------------------------------------------------------------ candidate vectors / MMR / rescoring
    def fetch_rows(self, ids, out: Optional[np.ndarray] = None) -> np.ndarray:
        gids = np.ascontiguousarray(ids, dtype=np.uint32).reshape(-1)
        if out is None:
            out = np.zeros((gids.shape[0], self.row_bytes), dtype=np.uint8)
        shard, local = self._locate(np.where(gids == 0xFFFFFFFF, -1, gids.astype(np.int64)))
        for j, sh in enumerate(self.shards):
            sel = np.nonzero(shard == j)[0]
            if sel.size:
                out[sel] = sh.fetch_rows((local[sel] + sh.row_base).astype(np.uint32))
        return out

    def mmr(self, vecs, relevance, lam: float, k_out: Optional[int] = None):
        return self.shards[0].mmr(vecs, relevance, lam, k_out)

    def pack_allow(self, allow):
        raise TypeError("MultiDeviceIndex takes a bool mask over the global rows (allow=...), not a packed bitmap")

    # ------------------------------------------------------------------ persistence (one blob per shard)
    def layout(self) -> dict:
        return {"devices": len(self.shards), "segments": [list(map(int, s)) for s in self._seg], "count": self._count}

    def save(self, path: str) -> None:
        for j, sh in enumerate(self.shards):
            sh.save(f"{path}.d{j}")

    def append_to(self, path: str) -> None:
        import os
        for j, sh in enumerate(self.shards):
            p = f"{path}.d{j}"
            sh.append_to(p) if os.path.exists(p) else sh.save(p)

    @classmethod
    def load(cls, path: str, layout: dict, devices: List[int]) -> "MultiDeviceIndex":
        g = int(layout["devices"])
        if len(devices) != g:
            raise ValueError(f"the stored collection is dealt out over {g} shards, {len(devices)} devices were given")
        shards = [ShardIndex.load(f"{path}.d{j}", device=devices[j]) for j in range(g)]
        self = cls.__new__(cls)
        self.devices = [int(d) for d in devices]
        self.shards = shards
        s0 = shards[0]
        self.dim, self.dim_padded, self.row_bytes = s0.dim, s0.dim_padded, s0.row_bytes
        self.dtype, self.metric, self.is_int, self.similarity_scale = s0.dtype, s0.metric, s0.is_int, s0.similarity_scale
        self.device = self.devices[0]
        self._seg, self._seg_start, self._count = [], [], 0
        self._cap, self._ex, self._next = (256, 128), None, 0
        for first, n, j, first_local in layout["segments"]:
            if first_local + n > len(shards[j]):
                break                                     # a shard blob holds fewer rows than the sidecar says: torn append
            shards[j].map_ids(first_local, n, first)
            self._seg.append((first, n, j, first_local))
            self._seg_start.append(first)
            self._count = first + n
        for j, sh in enumerate(shards):                  # rows no complete segment covers (torn append) are dropped
            keep = max([fl + n for (_f, n, jj, fl) in self._seg if jj == j], default=0)
            if keep < len(sh):
                sh.truncate(keep)
        return self
