"""Oracle: exact brute-force vector search, MMR re-selection and shard merge, numpy on the CPU.
TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

What it follows:

* H5  Milvus-lite FLAT index, metric L2 (squared distance, ascending) as configured by
  langchain-milvus 0.1.3 when driven from ``server/RAGHelper.py:388-394`` and searched through the
  retriever built at ``:497-499``; H5b pgvector cosine distance for the
  ``postgres`` store (``:399-404``); plain inner product for the north-star's IP variant.
* H6  ``maximal_marginal_relevance`` (langchain-core, used by ``search_type="mmr"``; Appendix A.4),
  float64 like numpy on python-float embeddings, strict ``>`` so the lowest index wins ties.
* the all-gather + merge of per-shard top-k lists restated as a concatenate + sort.

Tie-break (the reference's stores do not document one; pinned here): better score first, then the
lower row / id.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def metric_values(q: np.ndarray, x: np.ndarray, metric: str, dtype=np.float32) -> np.ndarray:
    """[nq, D] x [N, D] -> [nq, N] metric values (ip / cosine similarity / squared L2 distance)."""
    q = np.asarray(q, dtype=dtype)
    x = np.asarray(x, dtype=dtype)
    if metric == "ip":
        return q @ x.T
    if metric == "cosine":
        qn = np.linalg.norm(q, axis=1, keepdims=True)
        xn = np.linalg.norm(x, axis=1, keepdims=True)
        den = qn * xn.T
        with np.errstate(divide="ignore", invalid="ignore"):
            c = (q @ x.T) / den
        return np.where(den > 0, c, 0).astype(dtype)
    if metric == "l2":
        # direct sum of squared differences (what a FLAT L2 scan computes), blocked to bound memory
        out = np.empty((q.shape[0], x.shape[0]), dtype=dtype)
        for i in range(q.shape[0]):
            d = x - q[i]
            out[i] = np.einsum("nd,nd->n", d, d)
        return out
    raise ValueError(metric)


def flat_search(q: np.ndarray, x: np.ndarray, k: int, metric: str, id_offset: int = 0,
                dtype=np.float32) -> Tuple[np.ndarray, np.ndarray]:
    """Exact top-k: returns (scores [nq,k], ids [nq,k] int64); missing -> id -1, score -/+inf."""
    q = np.atleast_2d(np.asarray(q))
    nq, n = q.shape[0], x.shape[0]
    missing = np.inf if metric == "l2" else -np.inf
    scores = np.full((nq, k), missing, dtype=np.float32)
    ids = np.full((nq, k), -1, dtype=np.int64)
    if n == 0:
        return scores, ids
    vals = metric_values(q, x, metric, dtype)
    rank = -vals if metric != "l2" else vals              # ascending sort key
    kk = min(k, n)
    for i in range(nq):
        order = np.lexsort((np.arange(n), rank[i]))[:kk]  # rank asc, then row asc
        scores[i, :kk] = vals[i, order]
        ids[i, :kk] = order + id_offset
    return scores, ids


def flat_search_blocked(q: np.ndarray, x: np.ndarray, k: int, metric: str, id_offset: int = 0,
                        block: int = 131072) -> Tuple[np.ndarray, np.ndarray]:
    """The same result as :func:`flat_search` (same fp32 values, same tie rule) for corpora too large to sort
    whole: the metric is evaluated in row blocks, every block keeps its k best plus everything tied with its
    k-th value, and the survivors are ordered by (rank, row) exactly as ``flat_search`` orders all rows.
    Cosine norms are taken per block (row norms do not depend on the block), so the values are bit-identical
    to ``metric_values`` on the whole matrix up to BLAS blocking of the inner products."""
    q = np.atleast_2d(np.asarray(q, dtype=np.float32))
    nq, n = q.shape[0], x.shape[0]
    missing = np.inf if metric == "l2" else -np.inf
    scores = np.full((nq, k), missing, dtype=np.float32)
    ids = np.full((nq, k), -1, dtype=np.int64)
    if n == 0:
        return scores, ids
    keep_v = [[] for _ in range(nq)]
    keep_r = [[] for _ in range(nq)]
    for b0 in range(0, n, block):
        xb = x[b0:b0 + block]
        vals = metric_values(q, xb, metric)
        rank = -vals if metric != "l2" else vals
        nb = xb.shape[0]
        kk = min(k, nb)
        for i in range(nq):
            if nb > kk:
                kth = np.partition(rank[i], kk - 1)[kk - 1]
                sel = np.nonzero(rank[i] <= kth)[0]          # the k best and every tie with the k-th
            else:
                sel = np.arange(nb)
            keep_v[i].append(vals[i, sel])
            keep_r[i].append(sel + b0)
    kk = min(k, n)
    for i in range(nq):
        v = np.concatenate(keep_v[i])
        r = np.concatenate(keep_r[i])
        rank = -v if metric != "l2" else v
        order = np.lexsort((r, rank))[:kk]                    # rank asc, then row asc
        scores[i, :kk] = v[order]
        ids[i, :kk] = r[order] + id_offset
    return scores, ids


def shard_merge(scores: np.ndarray, ids: np.ndarray, metric: str) -> Tuple[np.ndarray, np.ndarray]:
    """[R, nq, k] -> [nq, k]: best score first, then lower id; id -1 entries are padding."""
    R, nq, k = scores.shape
    out_s = np.full((nq, k), np.inf if metric == "l2" else -np.inf, dtype=np.float32)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    for i in range(nq):
        s = scores[:, i, :].reshape(-1)
        d = ids[:, i, :].reshape(-1)
        valid = d >= 0
        s, d = s[valid], d[valid]
        rank = s if metric == "l2" else -s
        order = np.lexsort((d, rank))[:k]
        out_s[i, :len(order)] = s[order]
        out_i[i, :len(order)] = d[order]
    return out_s, out_i


def cosine_similarity(X, Y) -> np.ndarray:
    """langchain_core.vectorstores.utils.cosine_similarity: float64, NaN/Inf -> 0."""
    X = np.array(X, dtype=np.float64)
    Y = np.array(Y, dtype=np.float64)
    if X.size == 0 or Y.size == 0:
        return np.array([])
    Xn = np.linalg.norm(X, axis=1)
    Yn = np.linalg.norm(Y, axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        sim = np.dot(X, Y.T) / np.outer(Xn, Yn)
    sim[np.isnan(sim) | np.isinf(sim)] = 0.0
    return sim


def mmr(query_embedding, embedding_list, lambda_mult: float = 0.5, k: int = 4) -> List[int]:
    """maximal_marginal_relevance (Appendix A.4)."""
    if min(k, len(embedding_list)) <= 0:
        return []
    q = np.array(query_embedding, dtype=np.float64)
    if q.ndim == 1:
        q = q[None]
    E = np.array(embedding_list, dtype=np.float64)
    sim_q = cosine_similarity(q, E)[0]
    most = int(np.argmax(sim_q))
    idxs = [most]
    selected = np.array([E[most]])
    while len(idxs) < min(k, len(E)):
        best, add = -np.inf, -1
        sim_s = cosine_similarity(E, selected)
        for i, qs in enumerate(sim_q):
            if i in idxs:
                continue
            sc = lambda_mult * qs - (1 - lambda_mult) * max(sim_s[i])
            if sc > best:
                best, add = sc, i
        idxs.append(add)
        selected = np.append(selected, [E[add]], axis=0)
    return idxs


def mmr_search(q: np.ndarray, x: np.ndarray, k: int, metric: str, fetch_k: int = 20,
               lambda_mult: float = 0.5) -> List[List[int]]:
    """The dense retriever as the reference wires it (``search_type="mmr"``): top-fetch_k by the
    store metric -> fetch those vectors -> greedy MMR -> k row ids in MMR order."""
    out = []
    _, ids = flat_search(q, x, fetch_k, metric)
    for i in range(np.atleast_2d(q).shape[0]):
        cand = [int(r) for r in ids[i] if r >= 0]
        sel = mmr(np.atleast_2d(q)[i], x[cand], lambda_mult=lambda_mult, k=k)
        out.append([cand[j] for j in sel])
    return out
