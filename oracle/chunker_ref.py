"""TEST INFRASTRUCTURE ONLY — CPU restatement of langchain_experimental.text_splitter.SemanticChunker.

The reference constructs it at /root/reference/server/RAGHelper.py:329-341 (arguments: embeddings,
breakpoint_threshold_type, breakpoint_threshold_amount, number_of_chunks) and calls ``split_documents`` at :368.
langchain_experimental is a third-party dependency (reference ``server/requirements.txt``) that is neither vendored in the
reference nor installable here: this file restates its published algorithm from memory [3P-recall] — parity unpinned by
the reference's own repository, like the rest of the third-party legs — and is only ever imported by tests/.
Structure follows the upstream module: ``combine_sentences`` -> ``embed_documents`` -> ``calculate_cosine_distances``
(numpy, float64 on the python-float lists) -> threshold (percentile | standard_deviation | interquartile | gradient, or
``number_of_chunks``) -> chunks.
"""
import re

import numpy as np

DEFAULTS = {"percentile": 95, "standard_deviation": 3, "interquartile": 1.5, "gradient": 95}


def cosine_similarity(X, Y):
    """langchain_community.utils.math.cosine_similarity: float64 numpy, rows of X against rows of Y"""
    X = np.array(X, dtype=np.float64)
    Y = np.array(Y, dtype=np.float64)
    Xn = np.linalg.norm(X, axis=1)
    Yn = np.linalg.norm(Y, axis=1)
    sim = np.dot(X, Y.T) / np.outer(Xn, Yn)
    sim[np.isnan(sim) | np.isinf(sim)] = 0.0
    return sim


def sentence_groups(sentences, buffer_size=1):
    out = []
    for i in range(len(sentences)):
        combined = ""
        for j in range(i - buffer_size, i):
            if j >= 0:
                combined += sentences[j] + " "
        combined += sentences[i]
        for j in range(i + 1, i + 1 + buffer_size):
            if j < len(sentences):
                combined += " " + sentences[j]
        out.append(combined)
    return out


def distances_from_embeddings(embeddings):
    """embeddings: list of lists of python floats (what ``embed_documents`` returns)"""
    d = []
    for i in range(len(embeddings) - 1):
        d.append(1 - cosine_similarity([embeddings[i]], [embeddings[i + 1]])[0][0])
    return d


def split_text(text, embed_documents, kind="percentile", amount=None, number_of_chunks=None, buffer_size=1,
               regex=r"(?<=[.?!])\s+", min_chunk_size=None):
    amount = DEFAULTS[kind] if amount is None else amount
    single = re.split(regex, text)
    if len(single) == 1:
        return single
    if kind == "gradient" and len(single) == 2:
        return single
    groups = sentence_groups(single, buffer_size)
    distances = distances_from_embeddings(embed_documents(groups))
    if number_of_chunks is not None:
        x1, y1, x2, y2 = len(distances), 0.0, 1.0, 100.0
        x = max(min(number_of_chunks, x1), x2)
        y = y2 if x2 == x1 else y1 + ((y2 - y1) / (x2 - x1)) * (x - x1)
        threshold, arr = np.percentile(distances, min(max(y, 0), 100)), distances
    elif kind == "percentile":
        threshold, arr = np.percentile(distances, amount), distances
    elif kind == "standard_deviation":
        threshold, arr = np.mean(distances) + amount * np.std(distances), distances
    elif kind == "interquartile":
        q1, q3 = np.percentile(distances, [25, 75])
        threshold, arr = np.mean(distances) + amount * (q3 - q1), distances
    elif kind == "gradient":
        arr = np.gradient(distances, range(0, len(distances)))
        threshold = np.percentile(arr, amount)
    else:
        raise ValueError(kind)
    chunks, start = [], 0
    for index in [i for i, x in enumerate(arr) if x > threshold]:
        text_ = " ".join(single[start:index + 1])
        if min_chunk_size is not None and len(text_) < min_chunk_size:
            continue
        chunks.append(text_)
        start = index + 1
    if start < len(single):
        chunks.append(" ".join(single[start:]))
    return chunks, distances
