"""SemanticChunker — the text splitter RAGMeUp selects with ``splitter=SemanticChunker``.

The reference builds ``langchain_experimental.text_splitter.SemanticChunker(self.embeddings,
breakpoint_threshold_type=…, breakpoint_threshold_amount=…, number_of_chunks=…)`` at
``server/RAGHelper.py:329-341`` and runs ``self.text_splitter.split_documents(docs)`` at ``:368``.  langchain_experimental
is a third-party dependency that is not part of the reference repository (``server/requirements.txt``); this module
restates its published algorithm [3P-recall], with the two array computations on the GPU:

* the sentence groups are embedded by ``HuggingFaceEmbeddings.encode_tensor`` (vectors stay in HBM),
* the distance between consecutive groups, ``1 - cosine_similarity`` in float64, is one kernel
  (``rmu_adjacent_cosine_distance``).

What stays on the host is what is host work in the reference too: the regex sentence split, the sliding window that joins a
sentence with its neighbours, the percentile / standard-deviation / interquartile / gradient threshold over the (few
hundred) distances, and the assembly of chunks.  There is no CPU path for the arithmetic: embeddings without
``encode_tensor`` are rejected.
"""
from __future__ import annotations

import copy
import re
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .documents import Document

BREAKPOINT_DEFAULTS: Dict[str, float] = {"percentile": 95, "standard_deviation": 3, "interquartile": 1.5, "gradient": 95}


def combine_sentences(sentences: List[Dict[str, Any]], buffer_size: int = 1) -> List[Dict[str, Any]]:
    """every sentence joined with ``buffer_size`` neighbours on either side (what gets embedded)"""
    n = len(sentences)
    for i in range(n):
        lo, hi = max(0, i - buffer_size), min(n, i + buffer_size + 1)
        sentences[i]["combined_sentence"] = " ".join(sentences[j]["sentence"] for j in range(lo, hi))
    return sentences


def adjacent_cosine_distance(vectors) -> np.ndarray:
    """CUDA fp32 [n, dim] -> float64 [n - 1] distances between consecutive rows (device kernel, one D2H copy)"""
    torch = _lib.require_cuda()
    if not (hasattr(vectors, "is_cuda") and vectors.is_cuda):
        raise TypeError("adjacent_cosine_distance needs a CUDA tensor (HuggingFaceEmbeddings.encode_tensor)")
    v = vectors.contiguous().float()
    n, dim = int(v.shape[0]), int(v.shape[1])
    if n < 2:
        return np.zeros((0,), np.float64)
    out = torch.empty((n - 1,), dtype=torch.float64, device=v.device)
    _lib.check(_lib.lib().rmu_adjacent_cosine_distance(v.data_ptr(), n, dim, out.data_ptr(), _lib.stream_ptr()),
               "rmu_adjacent_cosine_distance")
    return out.cpu().numpy()


class SemanticChunker:
    """Split text where consecutive sentence groups are semantically far apart (same constructor as
    langchain_experimental's)."""

    def __init__(self, embeddings: Any, buffer_size: int = 1, add_start_index: bool = False,
                 breakpoint_threshold_type: Optional[str] = "percentile",
                 breakpoint_threshold_amount: Optional[float] = None, number_of_chunks: Optional[int] = None,
                 sentence_split_regex: str = r"(?<=[.?!])\s+", min_chunk_size: Optional[int] = None):
        kind = breakpoint_threshold_type or "percentile"
        if kind not in BREAKPOINT_DEFAULTS:
            raise ValueError(f"Got unexpected `breakpoint_threshold_type`: {kind}")
        if not hasattr(embeddings, "encode_tensor"):
            raise TypeError("SemanticChunker needs ragmeup_b200.embeddings.HuggingFaceEmbeddings (device encode); "
                            "this path has no CPU fallback")
        self.embeddings = embeddings
        self.buffer_size = buffer_size
        self._add_start_index = add_start_index
        self.breakpoint_threshold_type = kind
        self.number_of_chunks = number_of_chunks
        self.sentence_split_regex = sentence_split_regex
        self.breakpoint_threshold_amount = (BREAKPOINT_DEFAULTS[kind] if breakpoint_threshold_amount is None
                                            else breakpoint_threshold_amount)
        self.min_chunk_size = min_chunk_size

    # ------------------------------------------------------------------ thresholds (host, a few hundred numbers)
    def _calculate_breakpoint_threshold(self, distances: Sequence[float]) -> Tuple[float, Sequence[float]]:
        kind, amount = self.breakpoint_threshold_type, self.breakpoint_threshold_amount
        if kind == "percentile":
            return float(np.percentile(distances, amount)), distances
        if kind == "standard_deviation":
            return float(np.mean(distances) + amount * np.std(distances)), distances
        if kind == "interquartile":
            q1, q3 = np.percentile(distances, [25, 75])
            return float(np.mean(distances) + amount * (q3 - q1)), distances
        grad = np.gradient(distances, range(0, len(distances)))
        return float(np.percentile(grad, amount)), grad

    def _threshold_from_clusters(self, distances: Sequence[float]) -> float:
        """``number_of_chunks`` -> a percentile by linear interpolation between (len, 0 %) and (1, 100 %)"""
        if self.number_of_chunks is None:
            raise ValueError("This should never be called if `number_of_chunks` is None.")
        x1, y1 = len(distances), 0.0
        x2, y2 = 1.0, 100.0
        x = max(min(self.number_of_chunks, x1), x2)
        y = y2 if x2 == x1 else y1 + ((y2 - y1) / (x2 - x1)) * (x - x1)
        y = min(max(y, 0), 100)
        return float(np.percentile(distances, y))

    # ------------------------------------------------------------------ distances (device)
    def _calculate_sentence_distances(self, single_sentences_list: List[str]) -> Tuple[List[float], List[Dict[str, Any]]]:
        sentences = combine_sentences([{"sentence": x, "index": i} for i, x in enumerate(single_sentences_list)],
                                      self.buffer_size)
        vectors = self.embeddings.encode_tensor([x["combined_sentence"] for x in sentences])
        return adjacent_cosine_distance(vectors).tolist(), sentences

    # ------------------------------------------------------------------ splitting
    def chunks_from_distances(self, sentences: List[Dict[str, Any]], distances: Sequence[float]) -> List[str]:
        """threshold, breakpoints and chunk assembly (pure host logic, unit-tested without a GPU)"""
        if self.number_of_chunks is not None:
            threshold, breakpoint_array = self._threshold_from_clusters(distances), distances
        else:
            threshold, breakpoint_array = self._calculate_breakpoint_threshold(distances)
        above = [i for i, x in enumerate(breakpoint_array) if x > threshold]
        chunks: List[str] = []
        start = 0
        for index in above:
            text = " ".join(d["sentence"] for d in sentences[start:index + 1])
            if self.min_chunk_size is not None and len(text) < self.min_chunk_size:
                continue
            chunks.append(text)
            start = index + 1
        if start < len(sentences):
            chunks.append(" ".join(d["sentence"] for d in sentences[start:]))
        return chunks

    def split_text(self, text: str) -> List[str]:
        single = re.split(self.sentence_split_regex, text)
        if len(single) == 1:
            return single
        if self.breakpoint_threshold_type == "gradient" and len(single) == 2:
            return single
        distances, sentences = self._calculate_sentence_distances(single)
        return self.chunks_from_distances(sentences, distances)

    def create_documents(self, texts: List[str], metadatas: Optional[List[dict]] = None) -> List[Document]:
        metas = metadatas or [{}] * len(texts)
        out: List[Document] = []
        for i, text in enumerate(texts):
            start_index = 0
            for chunk in self.split_text(text):
                meta = copy.deepcopy(metas[i])
                if self._add_start_index:
                    meta["start_index"] = start_index
                out.append(Document(page_content=chunk, metadata=meta))
                start_index += len(chunk)
        return out

    def split_documents(self, documents: Iterable[Document]) -> List[Document]:
        docs = list(documents)
        return self.create_documents([d.page_content for d in docs], [d.metadata for d in docs])

    def transform_documents(self, documents: Sequence[Document], **_: Any) -> Sequence[Document]:
        return self.split_documents(list(documents))
