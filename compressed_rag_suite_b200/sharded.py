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


class ShardedSearcher:
    """Wraps this rank's ShardIndex; ``search`` returns the GLOBAL top-k on every rank."""

    def __init__(self, index, group=None):
        import torch.distributed as dist
        self.index = index
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.merge_launches = 0

    def search(self, queries, k: int, min_similarity: float = -math.inf):
        """queries: float32 CUDA tensor [nq, dim], the same on every rank.
        -> (ids int32 bit patterns [nq,k], raw scores [nq,k], counts [nq]) CUDA tensors."""
        from .index import merge_topk
        ids, sc, cnt = self.index.search(queries, k, min_similarity)
        if self.world == 1:
            self.merge_launches = 0
            return ids, sc, cnt
        gathered = exchange_candidates(pack_candidates(ids, sc), self.group)
        g_ids, g_sc = unpack_candidates(gathered, self.index.is_int)
        self.merge_launches = 1
        return merge_topk(g_ids, g_sc, k)
