"""The retriever composition of the reference's hybrid search.

``server/RAGHelper.py`` wires (``:436-443``, ``:488-503``)::

    sparse   = BM25Retriever.from_texts(texts, metadatas=...)                 # k = 4 (class default)
    dense    = db.as_retriever(search_type="mmr", search_kwargs={"k": vector_store_k})
    ensemble = EnsembleRetriever(retrievers=[sparse, dense], weights=[0.5, 0.5])
    rerank   = ContextualCompressionRetriever(base_compressor=compressor, base_retriever=ensemble)

Same class names, constructor keywords, ``invoke`` / ``|`` behaviour as langchain-community 0.2.10's
``BM25Retriever`` and langchain 0.2.11's ``EnsembleRetriever`` (weighted reciprocal rank fusion, c = 60,
duplicates merged by ``page_content``) and ``ContextualCompressionRetriever``.  BM25 scoring runs on the
GPU (``bm25.BM25Index`` -> ``csrc/rmu_bm25.cu``); the fusion itself is a few dozen dictionary updates per
query and stays on the host, as in the reference.
"""
from __future__ import annotations

from collections import defaultdict
from itertools import chain
from typing import Any, Callable, Dict, Hashable, Iterable, List, Optional, Sequence

from .bm25 import BM25Index
from .documents import Document, Runnable


def default_preprocessing_func(text: str) -> List[str]:
    return text.split()


class _Retriever(Runnable):
    def _get_relevant_documents(self, query: str) -> List[Document]:
        raise NotImplementedError

    def invoke(self, input: str, config: Any = None, **kwargs: Any) -> List[Document]:  # noqa: A002
        return self._get_relevant_documents(input)

    def get_relevant_documents(self, query: str, **_: Any) -> List[Document]:
        return self._get_relevant_documents(query)


class BM25Retriever(_Retriever):
    """langchain_community.retrievers.BM25Retriever over a GPU BM25Index."""

    def __init__(self, vectorizer: Any = None, docs: Optional[List[Document]] = None, k: int = 4,
                 preprocess_func: Callable[[str], List[str]] = default_preprocessing_func, **_: Any):
        self.vectorizer = vectorizer
        self.docs = list(docs or [])
        self.k = int(k)
        self.preprocess_func = preprocess_func

    @classmethod
    def from_texts(cls, texts: Iterable[str], metadatas: Optional[Iterable[dict]] = None,
                   bm25_params: Optional[Dict[str, Any]] = None,
                   preprocess_func: Callable[[str], List[str]] = default_preprocessing_func,
                   **kwargs: Any) -> "BM25Retriever":
        texts = list(texts)
        if preprocess_func is default_preprocessing_func:
            vectorizer = BM25Index.from_texts(texts, **(bm25_params or {}))     # str.split() + counting in C++
        else:
            vectorizer = BM25Index([preprocess_func(t) for t in texts], **(bm25_params or {}))
        metadatas = metadatas or ({} for _ in texts)
        docs = [Document(page_content=t, metadata=m) for t, m in zip(texts, metadatas)]
        return cls(vectorizer=vectorizer, docs=docs, preprocess_func=preprocess_func, **kwargs)

    @classmethod
    def from_documents(cls, documents: Iterable[Document], *, bm25_params: Optional[Dict[str, Any]] = None,
                       preprocess_func: Callable[[str], List[str]] = default_preprocessing_func,
                       **kwargs: Any) -> "BM25Retriever":
        documents = list(documents)
        return cls.from_texts([d.page_content for d in documents], [d.metadata for d in documents],
                              bm25_params=bm25_params, preprocess_func=preprocess_func, **kwargs)

    def _get_relevant_documents(self, query: str) -> List[Document]:
        return self.vectorizer.get_top_n(self.preprocess_func(query), self.docs, n=self.k)

    # additional batched path: Q queries in one device call
    def batch(self, queries: Sequence[str], config: Any = None, **kwargs: Any) -> List[List[Document]]:
        _, rows = self.vectorizer.search([self.preprocess_func(q) for q in queries], self.k)
        return [[self.docs[int(i)] for i in r] for r in rows]


def unique_by_key(iterable: Iterable[Any], key: Callable[[Any], Hashable]) -> Iterable[Any]:
    seen = set()
    for e in iterable:
        k = key(e)
        if k not in seen:
            seen.add(k)
            yield e


class EnsembleRetriever(_Retriever):
    """Weighted Reciprocal Rank Fusion over several retrievers (langchain 0.2.11 semantics)."""

    def __init__(self, retrievers: Sequence[Any], weights: Optional[Sequence[float]] = None, c: int = 60,
                 id_key: Optional[str] = None, **_: Any):
        self.retrievers = list(retrievers)
        self.weights = list(weights) if weights else [1 / len(self.retrievers)] * len(self.retrievers)
        self.c = int(c)
        self.id_key = id_key

    def _key(self, doc: Document) -> Hashable:
        return doc.page_content if self.id_key is None else doc.metadata[self.id_key]

    def weighted_reciprocal_rank(self, doc_lists: List[List[Document]]) -> List[Document]:
        if len(doc_lists) != len(self.weights):
            raise ValueError("Number of rank lists must be equal to the number of weights.")
        rrf_score: Dict[Hashable, float] = defaultdict(float)
        for doc_list, weight in zip(doc_lists, self.weights):
            for rank, doc in enumerate(doc_list, start=1):
                rrf_score[self._key(doc)] += weight / (rank + self.c)
        all_docs = chain.from_iterable(doc_lists)
        return sorted(unique_by_key(all_docs, self._key), reverse=True, key=lambda doc: rrf_score[self._key(doc)])

    def rank_fusion(self, query: str) -> List[Document]:
        doc_lists = [r.invoke(query) for r in self.retrievers]
        doc_lists = [[Document(page_content=d) if isinstance(d, str) else d for d in dl] for dl in doc_lists]
        return self.weighted_reciprocal_rank(doc_lists)

    def _get_relevant_documents(self, query: str) -> List[Document]:
        return self.rank_fusion(query)


class ContextualCompressionRetriever(_Retriever):
    """base_retriever -> base_compressor.compress_documents (langchain 0.2.11 semantics)."""

    def __init__(self, base_compressor: Any, base_retriever: Any, **_: Any):
        self.base_compressor = base_compressor
        self.base_retriever = base_retriever

    def _get_relevant_documents(self, query: str) -> List[Document]:
        docs = self.base_retriever.invoke(query)
        if docs:
            return list(self.base_compressor.compress_documents(docs, query))
        return []
