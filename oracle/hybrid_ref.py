"""Oracle for the sparse leg and the hybrid fusion that feed the reranker.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY STATUS: unpinned by the reference
repository (it holds no tests); the algorithms live in third-party wheels that are absent here and are
restated from their published sources:

* ``rank_bm25`` (``BM25Okapi``; pulled in by ``langchain_community.retrievers.BM25Retriever``):
  ``_initialize`` / ``_calc_idf`` / ``get_scores`` / ``get_top_n``.
* ``langchain_community.retrievers.bm25.BM25Retriever`` 0.2.10 (``server/requirements.txt:2``):
  ``from_texts`` with ``default_preprocessing_func = str.split`` and ``k = 4``.  The reference builds it at
  ``server/RAGHelper.py:436-443`` with no ``k`` (so 4 documents) and rebuilds it on every add (``:531-533``).
* ``langchain.retrievers.EnsembleRetriever.weighted_reciprocal_rank`` 0.2.11, wired at
  ``server/RAGHelper.py:501-503`` with ``retrievers=[sparse, dense]``, ``weights=[0.5, 0.5]``, ``c = 60``.
* ``langchain.retrievers.ContextualCompressionRetriever`` 0.2.11 (``server/RAGHelper.py:488-490``):
  ``docs = base_retriever.invoke(query); return list(base_compressor.compress_documents(docs, query))``.

Everything is float64 numpy / plain Python, exactly as the originals compute it.
"""
from __future__ import annotations

import math
from collections import defaultdict
from typing import Any, Callable, Dict, Hashable, List, Optional, Sequence

import numpy as np


def default_preprocessing_func(text: str) -> List[str]:
    return text.split()


class BM25Okapi:
    """rank_bm25.BM25Okapi restated (k1 = 1.5, b = 0.75, epsilon = 0.25)."""

    def __init__(self, corpus: Sequence[Sequence[str]], k1: float = 1.5, b: float = 0.75, epsilon: float = 0.25):
        self.k1, self.b, self.epsilon = k1, b, epsilon
        self.corpus_size = 0
        self.avgdl = 0.0
        self.doc_freqs: List[Dict[str, int]] = []
        self.idf: Dict[str, float] = {}
        self.doc_len: List[int] = []
        nd: Dict[str, int] = {}
        num_doc = 0
        for document in corpus:
            self.doc_len.append(len(document))
            num_doc += len(document)
            frequencies: Dict[str, int] = {}
            for word in document:
                frequencies[word] = frequencies.get(word, 0) + 1
            self.doc_freqs.append(frequencies)
            for word in frequencies:
                nd[word] = nd.get(word, 0) + 1
            self.corpus_size += 1
        self.avgdl = num_doc / self.corpus_size
        # _calc_idf: negative idfs are floored at epsilon * average idf
        idf_sum = 0.0
        negative = []
        for word, freq in nd.items():
            idf = math.log(self.corpus_size - freq + 0.5) - math.log(freq + 0.5)
            self.idf[word] = idf
            idf_sum += idf
            if idf < 0:
                negative.append(word)
        self.average_idf = idf_sum / len(self.idf)
        eps = self.epsilon * self.average_idf
        for word in negative:
            self.idf[word] = eps

    def get_scores(self, query: Sequence[str]) -> np.ndarray:
        score = np.zeros(self.corpus_size)
        doc_len = np.array(self.doc_len)
        for q in query:
            q_freq = np.array([(doc.get(q) or 0) for doc in self.doc_freqs])
            score += (self.idf.get(q) or 0) * (q_freq * (self.k1 + 1) /
                                               (q_freq + self.k1 * (1 - self.b + self.b * doc_len / self.avgdl)))
        return score

    def get_top_n(self, query: Sequence[str], documents: Sequence[Any], n: int = 5) -> List[Any]:
        scores = self.get_scores(query)
        # rank_bm25 uses np.argsort(scores)[::-1][:n]; numpy's default sort is not stable, so the order of EQUAL
        # scores is not defined by the original.  The stable sort pins it: equal scores -> larger index first.
        top_n = np.argsort(scores, kind="stable")[::-1][:n]
        return [documents[i] for i in top_n]


def bm25_topk(bm25: BM25Okapi, query_tokens: Sequence[str], n: int):
    """(scores float64 [n'], doc indices int64 [n']) in get_top_n order."""
    scores = bm25.get_scores(query_tokens)
    top = np.argsort(scores, kind="stable")[::-1][:n]
    return scores[top], top.astype(np.int64)


class BM25Retriever:
    """langchain_community BM25Retriever restated over plain (page_content, metadata) documents."""

    def __init__(self, vectorizer: BM25Okapi, docs: List[Any], k: int = 4,
                 preprocess_func: Callable[[str], List[str]] = default_preprocessing_func):
        self.vectorizer, self.docs, self.k, self.preprocess_func = vectorizer, docs, k, preprocess_func

    @classmethod
    def from_texts(cls, texts: Sequence[str], metadatas: Optional[Sequence[dict]] = None, make_doc=None,
                   bm25_params: Optional[dict] = None,
                   preprocess_func: Callable[[str], List[str]] = default_preprocessing_func, **kwargs: Any):
        texts_processed = [preprocess_func(t) for t in texts]
        vectorizer = BM25Okapi(texts_processed, **(bm25_params or {}))
        metadatas = metadatas or ({} for _ in texts)
        docs = [make_doc(t, m) if make_doc else (t, m) for t, m in zip(texts, metadatas)]
        return cls(vectorizer=vectorizer, docs=docs, preprocess_func=preprocess_func, **kwargs)

    def invoke(self, query: str) -> List[Any]:
        return self.vectorizer.get_top_n(self.preprocess_func(query), self.docs, n=self.k)


def weighted_reciprocal_rank(doc_lists: Sequence[Sequence[Any]], weights: Sequence[float], c: int = 60,
                             key: Callable[[Any], Hashable] = lambda d: d.page_content) -> List[Any]:
    """EnsembleRetriever.weighted_reciprocal_rank: score[key] += w / (rank + c), rank from 1; documents unique
    by key in first-seen order across the lists (chained), then a stable sort by score descending."""
    if len(doc_lists) != len(weights):
        raise ValueError("Number of rank lists must be equal to the number of weights.")
    rrf: Dict[Hashable, float] = defaultdict(float)
    for docs, w in zip(doc_lists, weights):
        for rank, d in enumerate(docs, start=1):
            rrf[key(d)] += w / (rank + c)
    seen = set()
    uniq = []
    for docs in doc_lists:
        for d in docs:
            kk = key(d)
            if kk not in seen:
                seen.add(kk)
                uniq.append(d)
    return sorted(uniq, key=lambda d: rrf[key(d)], reverse=True)


def ensemble_invoke(retrievers: Sequence[Any], weights: Sequence[float], query: str, c: int = 60,
                    key: Callable[[Any], Hashable] = lambda d: d.page_content) -> List[Any]:
    return weighted_reciprocal_rank([r.invoke(query) for r in retrievers], weights, c, key)


def contextual_compression_invoke(base_compressor: Any, base_retriever: Any, query: str) -> List[Any]:
    docs = base_retriever.invoke(query)
    if docs:
        return list(base_compressor.compress_documents(docs, query))
    return []
