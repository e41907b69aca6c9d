"""Provenance attribution helpers that reuse the encoder / cross-encoder / cosine primitives.

Drop-ins for the two attribution methods of ``server/provenance.py`` that sit on this path:

* ``compute_rerank_provenance`` (``server/provenance.py:100-108``): re-score the retrieved documents against
  ``answer`` (or ``query + "\\n" + answer`` when ``attribute_include_query == "True"``) with the reranker.
* ``DocumentSimilarityAttribution.compute_similarity`` (``server/provenance.py:164-202``): sentence-embed answer,
  query and documents, cosine of every document with the answer (and the query), average, normalise by the sum.

The reference's ``SentenceTransformer.encode`` + sklearn ``cosine_similarity`` become ``HuggingFaceEmbeddings``
(BERT on the GPU) + one brute-force cosine search over the documents (``FlatIndex``); nothing runs on the CPU but
the final handful of float32 additions.
"""
from __future__ import annotations

import os
from typing import Any, List, Optional, Sequence

import numpy as np

from . import _lib
from .index import FlatIndex


def compute_rerank_provenance(reranker: Any, query: str, documents: Sequence[Any], answer: str) -> Sequence[Any]:
    if os.getenv("attribute_include_query") == "True":
        full_text = query + "\n" + answer
    else:
        full_text = answer
    # same document list back, each with metadata["relevance_score"] (ScoredCrossEncoderReranker)
    return reranker.compress_documents(documents, full_text)


class DocumentSimilarityAttribution:
    def __init__(self, model_name: Optional[str] = None, embeddings: Any = None):
        if os.getenv("force_cpu") == "True":
            raise _lib.RmuError("force_cpu=True: ragmeup_b200 runs on CUDA (sm_100a) only")
        if embeddings is None:
            from .embeddings import HuggingFaceEmbeddings
            embeddings = HuggingFaceEmbeddings(model_name=model_name or os.getenv("provenance_similarity_llm"),
                                               model_kwargs={"device": "cuda"})
        self.model = embeddings

    def compute_similarity(self, query: str, context: Sequence[Any], answer: str) -> List[float]:
        include_query = os.getenv("attribute_include_query") != "False"
        texts = [c.page_content if hasattr(c, "page_content") else str(c) for c in context]
        n = len(texts)
        if n == 0:
            return []
        # SentenceTransformer.encode only strips; encode_tensor also maps "\n" to " " (langchain's embed_documents),
        # which the BERT normaliser (clean_text) does to every whitespace character anyway: same tokens
        enc = self.model.encode_tensor
        ctx = enc(texts)
        probes = enc([answer, query] if include_query else [answer])
        idx = FlatIndex(int(ctx.shape[1]), "cosine", device=ctx.device.index)
        idx.add(ctx)
        scores, ids = idx.search(probes, n)                   # every document, sorted by similarity
        scores, ids = scores.cpu().numpy(), ids.cpu().numpy()
        sim = np.zeros((probes.shape[0], n), dtype=np.float32)
        for r in range(probes.shape[0]):
            sim[r, ids[r]] = scores[r]
        if include_query:
            similarity = [(sim[0, i] + sim[1, i]) / 2 for i in range(n)]
        else:
            similarity = [sim[0, i] for i in range(n)]
        total = sum(similarity)
        out = [s / total for s in similarity] if total > 0 else similarity
        return [float(s) for s in out]
