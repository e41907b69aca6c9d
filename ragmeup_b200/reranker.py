"""ScoredCrossEncoderReranker drop-in.

Behaviour of the reference's only in-repo hot-path class,
``server/ScoredCrossEncoderReranker.py:12-45``: score every ``(query, doc.page_content)`` pair with
``model.score``, order by score descending with Python's stable sort (ties keep input order), keep
``top_n`` and return copies whose metadata gains ``relevance_score``.  Constructed by the reference
at ``server/RAGHelper.py:483-486`` and called through ``ContextualCompressionRetriever``
(``:488-490``) and ``compute_rerank_provenance`` (``server/provenance.py:100-108``).
"""
from __future__ import annotations

from typing import Any, List, Optional, Sequence

from .documents import copy_document


class ScoredCrossEncoderReranker:
    """Document compressor that uses a cross-encoder for reranking."""

    def __init__(self, model: Any = None, top_n: int = 3, **extra: Any):
        if extra:
            # the reference's pydantic Config is extra="forbid" (ScoredCrossEncoderReranker.py:21-23)
            raise TypeError(f"unexpected fields: {sorted(extra)}")
        if model is None or not hasattr(model, "score"):
            raise TypeError("model must provide score(text_pairs)")
        self.model = model
        self.top_n = int(top_n)

    def compress_documents(self, documents: Sequence[Any], query: str, callbacks: Optional[Any] = None) -> List[Any]:
        pairs = [(query, d.page_content) for d in documents]
        scores = self.model.score(pairs)
        ranked = sorted(zip(documents, scores), key=lambda ds: ds[1], reverse=True)
        return [copy_document(d, {**d.metadata, "relevance_score": s}) for d, s in ranked[: self.top_n]]

    async def acompress_documents(self, documents: Sequence[Any], query: str, callbacks: Optional[Any] = None) -> List[Any]:
        return self.compress_documents(documents, query, callbacks)
