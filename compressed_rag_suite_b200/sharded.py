"""Row-sharded multi-GPU search: one process per GPU, one collective per search.

GPU g holds the contiguous row block ``[g*ceil(N/G), (g+1)*ceil(N/G))``, so
global id = local row + row_base and "ties -> lowest id" survives sharding.  A search is:
local exact top-k on every rank (libcrs) -> ONE allgather of the packed ``[nq, k, 2]``
int32 candidates (id, raw-score bits) over NCCL/NVLink -> merge kernel K7 on every rank.
Raw scores are the canonical ones (identical on every rank), so the merged result is
bit-identical to a single-GPU search over the whole corpus.

The reference has no counterpart (it is single-node, rag/indexing.py); this is the
north star's multi-GPU extension of ``collection.query``.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np


def shard_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of rows owned by `rank`: [lo, hi)."""
    per = (n_total + world - 1) // world
    lo = min(rank * per, n_total)
    return lo, min(lo + per, n_total)


def pack_candidates(ids, scores):
    """(ids int32 [nq,k], scores f32|i32 [nq,k]) -> int32 [nq,k,2] (score bits kept verbatim)."""
    import torch
    return torch.stack((ids.view(torch.int32), scores.view(torch.int32)), dim=-1).contiguous()


def unpack_candidates(packed, is_int: bool):
    """int32 [G,nq,k,2] -> (ids int32 [G,nq,k], scores f32|i32 [G,nq,k])."""
    import torch
    ids = packed[..., 0].contiguous()
    sc = packed[..., 1].contiguous()
    return ids, (sc if is_int else sc.view(torch.float32))


def exchange_candidates(packed, group=None):
    """The one collective: allgather of each rank's packed local top-k.
    -> int32 [G, nq, k, 2].  Works on NCCL (CUDA tensors) and gloo (CPU tensors)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(packed.shape), dtype=packed.dtype, device=packed.device)
    if packed.is_cuda:
        dist.all_gather_into_tensor(out, packed, group=group)
    else:
        parts = [out[i] for i in range(world)]
        dist.all_gather(parts, packed, group=group)
    return out


class PeerExchangeUnavailable(RuntimeError):
    """Peer memory between the ranks' GPUs cannot be set up (raised on every rank together)."""


class PeerExchange:
    """One rank's ``crs_exchange``: a receive buffer every peer GPU stores its local top-k into (NVLink
    peer memory), so that the exchange + merge of a sharded search is one kernel of libcrs instead of an
    NCCL allgather launch followed by a merge launch.  Wiring the peers is the only host-side step:

    * ranks in different processes (torchrun): ``PeerExchange.from_process_group`` allgathers the 64-byte
      CUDA IPC handles of the receive buffers through torch.distributed and maps them;
    * ranks in one process (one host thread driving several GPUs): ``PeerExchange.wire_local``.
    """

    def __init__(self, device: int, rank: int, world: int, max_nq: int, max_k: int):
        import ctypes as C
        from . import _native as N
        self._lib = N.lib()
        self._h = C.c_void_p()
        N.check(self._lib.crs_exchange_create(C.byref(self._h), int(device), int(rank), int(world), int(max_nq), int(max_k)))
        self.device, self.rank, self.world, self.max_nq, self.max_k = int(device), int(rank), int(world), int(max_nq), int(max_k)
        self.has_shards = False

    def close(self) -> None:
        if self._h:
            self._lib.crs_exchange_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def ipc_handle(self) -> bytes:
        import ctypes as C
        from . import _native as N
        buf = (C.c_uint8 * 64)()
        N.check(self._lib.crs_exchange_ipc_handle(self._h, buf))
        return bytes(buf)

    def open_peers(self, handles: bytes) -> None:
        import ctypes as C
        from . import _native as N
        if len(handles) != 64 * self.world:
            raise ValueError("handles must hold world x 64 bytes")
        N.check(self._lib.crs_exchange_open_peers(self._h, C.c_char_p(handles)))

    def buffer(self) -> int:
        import ctypes as C
        from . import _native as N
        p = C.c_void_p()
        N.check(self._lib.crs_exchange_buffer(self._h, C.byref(p)))
        return p.value

    def status(self):
        """-> (timed_out: bool, step: int) after the current stream has drained."""
        import ctypes as C
        import torch
        from . import _native as N
        t, s = C.c_int(), C.c_uint32()
        st = torch.cuda.current_stream(self.device).cuda_stream
        N.check(self._lib.crs_exchange_status(self._h, C.c_void_p(st), C.byref(t), C.byref(s)))
        return bool(t.value), int(s.value)

    def merge(self, nq: int, k: int, is_int: bool):
        """Second half of a split sharded search (see ShardIndex.search_push): wait for the peers' pushes of
        this step and merge -> (ids, scores, counts) CUDA tensors on this exchange's device."""
        import ctypes as C
        import torch
        from . import _native as N
        dev = torch.device("cuda", self.device)
        ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
        sc = torch.empty((nq, k), dtype=torch.int32 if is_int else torch.float32, device=dev)
        cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        N.check(self._lib.crs_exchange_merge(self._h, C.c_void_p(st), int(nq), int(k), int(is_int), C.c_void_p(ids.data_ptr()),
                                             C.c_void_p(sc.data_ptr()), C.c_void_p(cnt.data_ptr())))
        return ids, sc, cnt

    # ---- the shards' stored rows, peer-mapped: candidate vectors without a second collective
    def register_shards_distributed(self, index, group=None) -> None:
        """Collective: every rank publishes the CUDA IPC handle, id base and row count of its shard of `index`
        and maps the others'.  Call after the corpus is built (growing an index re-allocates its rows)."""
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _native as N
        h = (C.c_uint8 * 64)()
        base, cnt = C.c_uint32(), C.c_int64()
        n_local = len(index)
        N.check(self._lib.crs_index_codes_handle(index._h, h if n_local else None, None, C.byref(base), C.byref(cnt)))
        dev = torch.device("cuda", self.device)
        mine = torch.zeros(64 + 16, dtype=torch.uint8, device=dev)
        mine[:64] = torch.frombuffer(bytearray(bytes(h)), dtype=torch.uint8).to(dev)
        meta = torch.tensor([base.value, cnt.value], dtype=torch.int64).view(torch.uint8).to(dev)
        mine[64:] = meta
        allm = torch.empty((self.world, 80), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allm, mine, group=group)
        allm = allm.cpu()
        handles = bytes(allm[:, :64].contiguous().numpy().tobytes())
        metas = allm[:, 64:].contiguous().view(torch.int64).numpy()
        bases = (C.c_uint32 * self.world)(*[int(v) for v in metas[:, 0]])
        counts = (C.c_int64 * self.world)(*[int(v) for v in metas[:, 1]])
        N.check(self._lib.crs_exchange_open_shards(self._h, index._h, C.c_char_p(handles), bases, counts))
        dist.barrier(group=group)
        self.has_shards = True

    @staticmethod
    def register_shards_local(exchanges, indexes) -> None:
        """Single process: exchanges[r] / indexes[r] belong to rank r."""
        import ctypes as C
        from . import _native as N
        world = len(exchanges)
        ptrs, bases, counts = (C.c_void_p * world)(), (C.c_uint32 * world)(), (C.c_int64 * world)()
        for r, ix in enumerate(indexes):
            p, b, c = C.c_void_p(), C.c_uint32(), C.c_int64()
            N.check(ix._lib.crs_index_codes_handle(ix._h, None, C.byref(p), C.byref(b), C.byref(c)))
            ptrs[r], bases[r], counts[r] = p.value or 0, b.value, c.value
        for r, e in enumerate(exchanges):
            N.check(e._lib.crs_exchange_set_shards(e._h, indexes[r]._h, ptrs, bases, counts))
            e.has_shards = True

    def fetch_rows(self, ids, row_bytes: int):
        """ids: int32 CUDA tensor (uint32 bit patterns, any shape) -> uint8 CUDA tensor ids.shape + (row_bytes,):
        every candidate's stored row read from the GPU that owns it (pad ids -> zero rows)."""
        import ctypes as C
        import torch
        from . import _native as N
        ids = ids.contiguous()
        out = torch.empty(tuple(ids.shape) + (int(row_bytes),), dtype=torch.uint8, device=ids.device)
        st = torch.cuda.current_stream(ids.device).cuda_stream
        N.check(self._lib.crs_exchange_fetch_rows(self._h, C.c_void_p(st), C.c_void_p(ids.data_ptr()), ids.numel(),
                                                  C.c_void_p(out.data_ptr())))
        return out

    def score_rows(self, index, queries, ids):
        """K8 on global ids wherever the rows live (crs_exchange_score_rows): -> [nq, m] canonical scores of
        `index`'s store dtype; pad ids get the absent score."""
        import ctypes as C
        import torch
        from . import _native as N
        q, i = queries.contiguous(), ids.contiguous()
        nq, m = i.shape
        out = torch.empty((nq, m), dtype=torch.int32 if index.is_int else torch.float32, device=q.device)
        index._use_torch_stream()
        N.check(self._lib.crs_exchange_score_rows(index._h, self._h, C.c_void_p(q.data_ptr()), nq, C.c_void_p(i.data_ptr()), m,
                                                  C.c_void_p(out.data_ptr())))
        absent = torch.iinfo(torch.int32).min if index.is_int else -math.inf
        return torch.where(i != -1, out, torch.full_like(out, absent))

    @classmethod
    def from_process_group(cls, device: int, max_nq: int, max_k: int, group=None) -> "PeerExchange":
        """Collective: every rank of the group creates its exchange and maps every peer's receive buffer.
        The ranks agree (MIN all-reduce of a success flag) after each local step, so a rank on which peer
        memory cannot be set up — no P2P between the devices, CUDA IPC not permitted in this container —
        makes EVERY rank raise ``PeerExchangeUnavailable`` together instead of leaving the others in a collective."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        dev = torch.device("cuda", device)

        def agree(ok: bool, what: str, err) -> None:
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 0:
                raise PeerExchangeUnavailable(f"{what} failed on at least one rank" + (f" (here: {err})" if err else ""))

        import os
        ex, err, handle = None, None, b"\0" * 64
        try:
            if os.environ.get("CRS_DISABLE_PEER_EXCHANGE"):      # knob (and test hook for the agreed fallback)
                raise RuntimeError("disabled by CRS_DISABLE_PEER_EXCHANGE")
            ex = cls(device, rank, world, max_nq, max_k)
            handle = ex.ipc_handle()
        except Exception as e:  # noqa: BLE001
            err = e
        try:
            agree(err is None, "creating the receive buffer / its IPC handle", err)
            mine = torch.frombuffer(bytearray(handle), dtype=torch.uint8).to(dev)
            allh = torch.empty((world, 64), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allh, mine, group=group)
            try:
                ex.open_peers(bytes(allh.cpu().numpy().tobytes()))
            except Exception as e:  # noqa: BLE001
                err = e
            agree(err is None, "mapping the peers' receive buffers", err)
        except PeerExchangeUnavailable:
            if ex is not None:
                ex.close()
            raise
        dist.barrier(group=group)
        return ex

    @staticmethod
    def wire_local(exchanges) -> None:
        """Single process: exchanges[r] is rank r's exchange (one per device); gives every rank the others'
        receive buffers (peer access between the devices is enabled by the library)."""
        import ctypes as C
        from . import _native as N
        world = len(exchanges)
        bufs = (C.c_void_p * world)(*[C.c_void_p(e.buffer()) for e in exchanges])
        for e in exchanges:
            if e.world != world:
                raise ValueError("every exchange must be created for the same world size")
            N.check(e._lib.crs_exchange_set_peer_buffers(e._h, bufs))


class ShardedSearcher:
    """Wraps this rank's ShardIndex; ``search`` returns the GLOBAL top-k on every rank.

    exchange: "peer" (default on CUDA with world > 1) — the library's one-kernel exchange + merge over NVLink
    peer memory; "nccl" — ``all_gather_into_tensor`` of the packed candidates followed by the merge kernel K7
    (the only form that also runs on gloo / CPU tensors).  Both return identical results."""

    def __init__(self, index, group=None, local_only: bool = False, exchange: str = "peer",
                 max_nq: int = 1024, max_k: int = 128):
        """local_only: the index holds the whole corpus; never exchange (even inside a process group)."""
        import torch.distributed as dist
        self.index = index
        self.group = group
        self.local_only = local_only
        self.world = 1 if local_only else (dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1)
        self.merge_launches = 0
        if exchange not in ("peer", "nccl"):
            raise ValueError("exchange must be 'peer' or 'nccl'")
        self.exchange = exchange
        self._peer = None
        self._peer_cap = (int(max_nq), int(max_k))

    def _peer_exchange(self, nq: int, k: int):
        """(Re)create the peer exchange when a search exceeds its capacity — a collective step: every rank
        sees the same nq and k, so every rank takes it together."""
        if self._peer is None or nq > self._peer.max_nq or k > self._peer.max_k:
            if self._peer is not None:
                self._peer.close()
            cap_nq, cap_k = max(nq, self._peer_cap[0]), min(128, max(k, self._peer_cap[1]))
            self._peer = None
            try:
                self._peer = PeerExchange.from_process_group(self.index.device, cap_nq, cap_k, self.group)
            except PeerExchangeUnavailable as e:
                # every rank lands here together: use the NCCL exchange from now on (same results, one more launch)
                import logging
                logging.getLogger(__name__).warning(f"peer-memory exchange unavailable ({e}); using the NCCL allgather + merge kernel")
                self.exchange = "nccl"
        return self._peer

    def search(self, queries, k: int, min_similarity: float = -math.inf):
        """queries: float32 CUDA tensor [nq, dim], the same on every rank.
        -> (ids int32 bit patterns [nq,k], raw scores [nq,k], counts [nq]) CUDA tensors."""
        import torch
        from .index import merge_gathered
        if self.world == 1:
            self.merge_launches = 0
            return self.index.search(queries, k, min_similarity)
        if self.exchange == "peer" and queries.is_cuda:
            nq = queries.shape[0] if queries.dim() > 1 else 1
            ex = self._peer_exchange(nq, k)              # may switch self.exchange to "nccl" (on every rank together)
            if ex is not None:
                self.merge_launches = 0                  # the exchange kernel is counted by the library
                return self.index.search_sharded(ex, queries, k, min_similarity)
        # the local result is written straight into the send buffer ([0] ids, [1] raw-score bits) and the
        # merge reads the gathered buffer in place: search -> allgather -> merge, no pack / unpack copies
        nq = queries.shape[0] if queries.dim() > 1 else 1
        send = torch.empty((2, nq, k), dtype=torch.int32, device=queries.device)
        cnt = torch.empty((nq,), dtype=torch.int32, device=queries.device)
        self.index.search_into(queries, k, min_similarity, send[0], send[1], cnt)
        gathered = exchange_candidates(send, self.group)             # [G, 2, nq, k]
        self.merge_launches = 1
        return merge_gathered(gathered, nq, k, k, self.index.is_int)

    def capture(self, nq: int, k: int, min_similarity: float = -math.inf):
        """The whole sharded step — encode, local search, exchange, merge — as ONE CUDA graph (peer exchange
        only: nothing in it is a library collective).  Returns a GraphedSearch: copy the queries into
        ``.queries``, ``.replay()``, read ``.ids / .scores / .counts``.  Every rank must capture and replay in
        step; the index must not grow while the graph is in use."""
        import torch
        from .index import GraphedSearch
        if self.world > 1 and self.exchange != "peer":
            raise ValueError("capture needs the peer exchange")
        ix = self.index
        dev = torch.device("cuda", ix.device)
        gen = torch.Generator(device=dev)                  # random warm-up queries, the same on every rank (see capture_search)
        gen.manual_seed(7)
        q = torch.randn((nq, ix.dim), dtype=torch.float32, device=dev, generator=gen)
        ex = self._peer_exchange(nq, k) if self.world > 1 else None
        if self.world > 1 and ex is None:
            raise ValueError("capture needs the peer exchange")
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                      # warm-up sizes every scratch buffer (and takes two steps on every rank)
            for _ in range(2):
                self.search(q, k, min_similarity)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        ix.set_option("profiling", 0)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            ids, sc, cnt = self.search(q, k, min_similarity)
        ix._use_torch_stream()
        return GraphedSearch(graph, q, ids, sc, cnt, len(ix))


# ------------------------------------------------------------------------------------------
# Candidate-set steps that follow the sharded top-k (BASELINE configs 4 and 5)

def assemble_over_shards(t, group=None):
    """Every rank contributed its own rows of `t` and the neutral element elsewhere (0 bytes for
    stored codes, -inf / INT32_MIN for scores): an elementwise MAX over the ranks assembles the
    full tensor on every rank.  In place; no-op without a process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t


def reference_relevance(sims):
    """The reference's `score` of a hit from its cosine similarity, in float64 exactly as Python
    evaluates it: Chroma distance d = float32(1 - cos) (rag/indexing.py:171-176), then
    rag/retrieval.py:75-77: d <- clamp(d, 0, 2); score = clamp(1 - d*d/2, 0, 1)."""
    import torch
    d = (1.0 - sims.to(torch.float32)).to(torch.float64)
    d = torch.clamp(d, 0.0, 2.0)
    return torch.clamp(1.0 - (d * d / 2.0), 0.0, 1.0)


def similarity_of(index, raw):
    """raw scores (torch) -> float32 cosine-domain similarity, same arithmetic as ShardIndex.similarity."""
    import torch
    if index.dtype == "i8":
        return raw.to(torch.float32) * float(np.float32(index.similarity_scale))
    if index.dtype == "b1":
        return (raw.to(torch.float64) / float(index.dim)).to(torch.float32)
    return raw.to(torch.float32)


class ShardedMMRSearcher(ShardedSearcher):
    """BASELINE config 4: global top-`fetch_k` over the row shards, then the reference's greedy MMR
    (rag/retrieval.py:219-277) over the STORED vectors of those candidates, first `k` of the
    greedy order.  With the peer exchange every rank reads the candidates' rows straight out of the
    owning GPU's HBM (NVLink peer loads, no second collective); with exchange="nccl" the owners gather
    their rows and one MAX all-reduce of [nq, fetch_k, row_bytes] bytes assembles them."""

    def _shard_exchange(self, nq: int, k: int):
        ex = self._peer_exchange(nq, k)
        if ex is not None and not ex.has_shards:
            ex.register_shards_distributed(self.index, self.group)
        return ex

    def search_mmr(self, queries, k: int, fetch_k: int, diversity_penalty: float,
                   min_similarity: float = -math.inf):
        """-> (ids int32 [nq,k] (pad -1), similarity f32 [nq,k], relevance f64 [nq,k], counts [nq])."""
        import torch
        ids, raw, cnt = self.search(queries, fetch_k, min_similarity)
        if self.world > 1 and self.exchange == "peer":
            nq = queries.shape[0] if queries.dim() > 1 else 1
            ex = self._shard_exchange(nq, fetch_k)
            if ex is not None:
                vecs = ex.fetch_rows(ids, self.index.row_bytes)
                return self.index.mmr_select(vecs, ids, raw, cnt, 1.0 - diversity_penalty, k)
        vecs = self.index.fetch_rows_device(ids)
        if self.world > 1:
            vecs = assemble_over_shards(vecs, self.group)
        # one launch: similarity + the reference's score transform + greedy MMR + gather of the picks
        return self.index.mmr_select(vecs, ids, raw, cnt, 1.0 - diversity_penalty, k)


class TwoStageSearcher:
    """BASELINE config 5: a coarse index (1-bit codes, Hamming) picks the global top-`fetch_k`,
    a fine index over the SAME rows (fp16) rescores those candidates with its canonical score
    (K8, crs_index_score_rows) and the best `k` by (fine score desc, id asc) are returned.
    Both indexes are row-sharded identically; per search: one allgather of the coarse
    candidates + one MAX all-reduce of [nq, fetch_k] fine scores."""

    def __init__(self, coarse, fine, group=None, local_only: bool = False, row_source=None, exchange: str = "peer"):
        """row_source: optional callable ids (int32 CUDA [nq, m], uint32 bit patterns, pad -1) ->
        float32 CUDA [nq, m, dim], the candidates' ORIGINAL vectors, for corpora whose fine rows
        are not resident in HBM (1 B x 1024-d fp16 = 2 TB: re-materialised from the generator, or
        read from host memory); `fine` then only carries dim / dtype / metric and stays empty.
        Every rank must be able to produce every candidate (no second collective)."""
        if row_source is None and (len(coarse) != len(fine) or coarse.row_base != fine.row_base):
            raise ValueError("coarse and fine index must hold the same rows")
        self.coarse = ShardedSearcher(coarse, group, local_only, exchange=exchange)
        self.fine = fine
        self.group = group
        self.row_source = row_source
        self._fine_ex = None                     # peer access to the fine shards (exchange="peer")

    def search(self, queries, k: int, fetch_k: int, min_similarity: float = -math.inf):
        """-> (ids int32 [nq,k] (uint32 bit patterns, pad -1), fine scores f32 [nq,k], counts [nq])."""
        import torch
        from .index import select_topk
        ids, _, _ = self.coarse.search(queries, fetch_k)
        if self.row_source is not None:
            fine = self.fine.score_vectors(queries, self.row_source(ids))
            fine = torch.where(ids != -1, fine, torch.full_like(fine, -math.inf))
        elif self.coarse.world > 1 and self.coarse.exchange == "peer":
            # the candidates' fine rows are read from the GPUs that own them (NVLink peer loads) and scored here:
            # no second collective
            if self._fine_ex is None:
                try:
                    self._fine_ex = PeerExchange.from_process_group(self.fine.device, 1, 1, self.group)
                    self._fine_ex.register_shards_distributed(self.fine, self.group)
                except PeerExchangeUnavailable:
                    self._fine_ex = False                 # every rank together: fall back to the all-reduce assembly
            if self._fine_ex:
                fine = self._fine_ex.score_rows(self.fine, queries, ids)
            else:
                fine = assemble_over_shards(self.fine.score_rows(queries, ids), self.group)
        else:
            fine = self.fine.score_rows(queries, ids)
            if self.coarse.world > 1:
                fine = assemble_over_shards(fine, self.group)
        if min_similarity > -math.inf:
            ids = torch.where(fine >= min_similarity, ids, torch.full_like(ids, -1))
        return select_topk(ids, fine, k)
