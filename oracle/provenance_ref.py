"""Oracle for the provenance helpers.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows ``server/provenance.py:100-108`` (``compute_rerank_provenance``) and ``:164-202``
(``DocumentSimilarityAttribution.compute_similarity``): ``SentenceTransformer.encode`` (restated in
``bert_ref.st_encode``), sklearn ``cosine_similarity`` (rows L2-normalised, zero rows left at zero, then a dot
product, all in the input dtype float32), the average with the query similarity, and the division by the sum.
PARITY STATUS: unpinned by the reference repository (no tests there); sentence-transformers is restated, the sklearn
cosine restatement is checked against the real scikit-learn in tests/test_hybrid_cpu.py.
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import numpy as np


def sk_cosine_similarity(X: np.ndarray, Y: np.ndarray) -> np.ndarray:
    def normalize(A):
        A = np.asarray(A, dtype=np.float32)
        n = np.sqrt((A * A).sum(axis=1))
        n[n == 0.0] = 1.0
        return A / n[:, None]
    return normalize(X) @ normalize(Y).T


def compute_similarity(encode: Callable[[Sequence[str]], np.ndarray], query: str, context: Sequence[str], answer: str,
                       include_query: bool = True) -> List[float]:
    answer_embedding = encode([answer])[0]
    context_embeddings = encode(list(context))
    if include_query:
        query_embedding = encode([query])[0]
    similarity_scores = []
    for doc_embedding in context_embeddings:
        doc_answer = sk_cosine_similarity([doc_embedding], [answer_embedding])[0][0]
        if include_query:
            doc_query = sk_cosine_similarity([doc_embedding], [query_embedding])[0][0]
            similarity_scores.append((doc_answer + doc_query) / 2)
        else:
            similarity_scores.append(doc_answer)
    total = sum(similarity_scores)
    return [s / total for s in similarity_scores] if total > 0 else similarity_scores


def compute_rerank_provenance(reranker, query: str, documents, answer: str, include_query: bool):
    full_text = query + "\n" + answer if include_query else answer
    return reranker.compress_documents(documents, full_text)
