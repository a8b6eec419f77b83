"""Oracle: the two candidate-set pipelines BASELINE.json configs 4 and 5 name, composed from
the canonical pieces in ``oracle/search.py`` / ``oracle/postprocess.py``.

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.

Neither pipeline exists in the reference as such (its only retrieval is k <= 2k hits from
Chroma, rag/retrieval.py:117-121); they are the north star's extensions and are defined here
first so the CUDA path has something to equal:

* ``search_then_mmr``: top-``fetch_k`` (score desc, id asc) -> the reference's ``score``
  transform (rag/retrieval.py:75-77 on the Chroma distance ``1 - sim``) -> the reference's
  greedy MMR (rag/retrieval.py:246-275, restated in ``postprocess.mmr_order``) over the STORED
  candidate vectors -> first ``k`` of the greedy order (Appendix A.5 of SURVEY.md: MMR never
  drops, so "top-100 -> MMR -> 10" is the prefix of the greedy order).
* ``two_stage``: coarse top-``fetch_k`` on one store (1-bit codes, Hamming), candidates
  rescored with the canonical score of a finer store (fp16) over the same rows, best ``k`` by
  (fine score desc, id asc).
"""
from __future__ import annotations

import numpy as np

from . import postprocess, search
from .encode import decode_rows, encode_rows


def search_then_mmr(x, q, store, k, fetch_k, diversity_penalty, min_similarity=-np.inf):
    """-> list over queries of (ids list, sims list, relevance list) in final order."""
    dim = x.shape[1]
    codes = encode_rows(x, store)
    qc = search.encode_queries(q, store)
    ids, raw, cnt = search.search(codes, qc, store, dim, fetch_k, min_similarity)
    out = []
    for i in range(len(qc)):
        c = int(cnt[i])
        if c == 0:
            out.append(([], [], []))
            continue
        cid = ids[i, :c].astype(np.int64)
        sims = search.similarity_from_raw(raw[i, :c], store, dim)
        rel = [postprocess.distance_to_similarity(float(np.float32(1.0) - np.float32(s))) for s in sims]
        dec = decode_rows(codes[cid], store)
        if store == "b1":
            dec = dec[:, :dim]
        order = postprocess.mmr_order(rel, postprocess.pairwise_sims_f32(dec), 1.0 - diversity_penalty,
                                      k_out=min(k, c))
        out.append(([int(cid[p]) for p in order], [float(sims[p]) for p in order], [rel[p] for p in order]))
    return out


def two_stage(x, q, k, fetch_k, coarse="b1", fine="f16", min_similarity=-np.inf):
    """-> (ids [nq,k] u32 padded 0xFFFFFFFF, fine raw scores [nq,k], counts [nq])."""
    dim = x.shape[1]
    c_codes = encode_rows(x, coarse)
    f_codes = encode_rows(x, fine)
    cand, _, cnt = search.search(c_codes, search.encode_queries(q, coarse), coarse, dim, fetch_k)
    fq = search.encode_queries(q, fine)
    rdt = np.float32 if fine in ("f16", "bf16") else np.int32
    out_ids = np.full((len(fq), k), 0xFFFFFFFF, dtype=np.uint32)
    out_raw = np.full((len(fq), k), search.pad_raw(rdt), dtype=rdt)
    counts = np.zeros(len(fq), dtype=np.int32)
    for i in range(len(fq)):
        cid = cand[i, :cnt[i]].astype(np.int64)
        raw = search.raw_scores(f_codes[cid], fq[i], fine, dim)
        sims = search.similarity_from_raw(raw, fine, dim)
        keep = sims >= np.float32(min_similarity) if np.isfinite(min_similarity) else np.ones(len(cid), bool)
        cid, raw = cid[keep], raw[keep]
        order = np.lexsort((cid, -raw.astype(np.float64)))[:k]
        counts[i] = len(order)
        out_ids[i, :len(order)] = cid[order]
        out_raw[i, :len(order)] = raw[order]
    return out_ids, out_raw, counts
