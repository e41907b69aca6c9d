"""Vector-store drop-ins: ``Milvus`` / ``PGVector`` look-alikes over the B200 flat index
(SURVEY.md §8 rows a6-a8, a12).

Surface kept from the reference's use of langchain-milvus 0.1.3 / langchain-postgres 0.0.12:

* ``Milvus.from_documents([], embeddings, drop_old=..., connection_args={"uri": ...},
  collection_name=...)``                                   — ``server/RAGHelper.py:388-394``
* ``PGVector(embeddings=, collection_name=, connection=, use_jsonb=True)``   — ``:399-404``
* ``db.add_documents(documents, ids=ids)``                                   — ``:431``, ``:525``
* ``db.as_retriever(search_type="mmr", search_kwargs={"k": k})``             — ``:497-499``, ``:533-535``
  → retriever usable as ``retriever.invoke(q)`` and ``retriever | fn``.

Semantics kept: Milvus-lite falls back to a FLAT index with metric L2,
returned score = squared L2 distance ascending; PGVector's default is cosine distance (1 - cos);
``search_type="mmr"`` = top-``fetch_k`` (20) by the store metric, fetch those vectors, greedy MMR
(lambda 0.5) down to k.  Result metadata carries the stored metadata plus ``pk`` (Milvus) like the
reference's stores do (``server/server.py:281-284`` reads ``source`` / ``pk``).

The corpus lives in HBM as fp32 ``[N, D]`` inside the C index; texts, ids and metadata stay on the
host.  All distance arithmetic runs in ``csrc/rmu_index.cu``.
"""
from __future__ import annotations

import threading
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .documents import Document, Runnable
from .index import FlatIndex, mmr_select


class B200Retriever(Runnable):
    """What ``VectorStore.as_retriever`` returns: ``invoke`` / ``get_relevant_documents``."""

    def __init__(self, vectorstore: "B200VectorStore", search_type: str = "similarity",
                 search_kwargs: Optional[Dict[str, Any]] = None):
        if search_type not in ("similarity", "mmr", "similarity_score_threshold"):
            raise ValueError(f"search_type of {search_type} not allowed.")
        self.vectorstore = vectorstore
        self.search_type = search_type
        self.search_kwargs = dict(search_kwargs or {})
        self.tags = [type(vectorstore).__name__]

    def _get_relevant_documents(self, query: str) -> List[Document]:
        if self.search_type == "mmr":
            return self.vectorstore.max_marginal_relevance_search(query, **self.search_kwargs)
        if self.search_type == "similarity_score_threshold":
            kw = dict(self.search_kwargs)
            thr = kw.pop("score_threshold")
            return [d for d, s in self.vectorstore.similarity_search_with_relevance_scores(query, **kw) if s >= thr]
        return self.vectorstore.similarity_search(query, **self.search_kwargs)

    def invoke(self, input: str, config: Any = None, **kwargs: Any) -> List[Document]:  # noqa: A002
        return self._get_relevant_documents(input)

    def get_relevant_documents(self, query: str, **_: Any) -> List[Document]:
        return self._get_relevant_documents(query)

    async def ainvoke(self, input: str, config: Any = None, **kwargs: Any) -> List[Document]:  # noqa: A002
        return self._get_relevant_documents(input)


class B200VectorStore:
    """Exact brute-force store on one B200.  ``metric``: 'l2' | 'cosine' | 'ip'."""

    id_field = "pk"

    def __init__(self, embedding_function: Any, metric: str = "l2", collection_name: str = "LangChainCollection",
                 device: Optional[int] = None):
        self.embedding_func = embedding_function
        self.embeddings = embedding_function
        self.metric = metric
        self.collection_name = collection_name
        self._device = device
        self.index: Optional[FlatIndex] = None
        self._pks: List[str] = []
        self._texts: List[str] = []
        self._metas: List[Dict[str, Any]] = []
        self._lock = threading.Lock()

    # ------------------------------------------------------------------ construction (reference forms)
    @classmethod
    def from_documents(cls, documents: Sequence[Document], embedding: Any, **kwargs: Any) -> "B200VectorStore":
        ids = kwargs.pop("ids", None)
        store = cls(embedding, **kwargs)
        if documents:
            store.add_documents(list(documents), ids=ids)
        return store

    @classmethod
    def from_texts(cls, texts: Sequence[str], embedding: Any, metadatas: Optional[List[dict]] = None,
                   **kwargs: Any) -> "B200VectorStore":
        ids = kwargs.pop("ids", None)
        store = cls(embedding, **kwargs)
        if texts:
            store.add_texts(list(texts), metadatas=metadatas, ids=ids)
        return store

    def __len__(self) -> int:
        return len(self._pks)

    # ------------------------------------------------------------------ insert
    def _ensure_index(self, dim: int) -> FlatIndex:
        if self.index is None:
            self.index = FlatIndex(dim, self.metric, device=self._device)
        elif self.index.dim != dim:
            raise ValueError(f"embedding dimension {dim} does not match the collection ({self.index.dim})")
        return self.index

    def add_embeddings(self, vectors, texts: Sequence[str], metadatas: Optional[Sequence[dict]] = None,
                       ids: Optional[Sequence[str]] = None) -> List[str]:
        """Append pre-computed vectors (CUDA/CPU tensor or numpy [n, D])."""
        n = len(texts)
        if ids is None:
            ids = [str(len(self._pks) + i) for i in range(n)]
        if len(ids) != n or (metadatas is not None and len(metadatas) != n):
            raise ValueError("texts, metadatas and ids must have the same length")
        if n == 0:
            return []
        dim = int(vectors.shape[1])
        with self._lock:
            self._ensure_index(dim).add(vectors)
            self._pks.extend(str(i) for i in ids)
            self._texts.extend(texts)
            self._metas.extend(dict(m) for m in (metadatas or [{} for _ in range(n)]))
        return [str(i) for i in ids]

    def add_texts(self, texts: Iterable[str], metadatas: Optional[List[dict]] = None, ids: Optional[List[str]] = None,
                  batch_size: int = 1000, **_: Any) -> List[str]:
        texts = list(texts)
        out: List[str] = []
        for s in range(0, len(texts), batch_size):          # langchain-milvus inserts in chunks of 1000
            chunk = texts[s:s + batch_size]
            if hasattr(self.embedding_func, "encode_tensor"):
                vecs = self.embedding_func.encode_tensor(chunk)          # stays in HBM
            else:
                vecs = np.asarray(self.embedding_func.embed_documents(chunk), dtype=np.float32)
            out += self.add_embeddings(vecs, chunk, None if metadatas is None else metadatas[s:s + batch_size],
                                       None if ids is None else ids[s:s + batch_size])
        return out

    def add_documents(self, documents: List[Document], ids: Optional[List[str]] = None, **kwargs: Any) -> List[str]:
        texts = [d.page_content for d in documents]
        metas = [d.metadata for d in documents]
        return self.add_texts(texts, metas, ids=ids, **kwargs)

    # ------------------------------------------------------------------ search (tensor level)
    def _score_out(self, scores):
        """index metric value -> the score the reference's store reports"""
        return (1.0 - scores) if self.metric == "cosine" else scores

    def search_tensor(self, queries, k: int):
        """CUDA fp32 [Q, D] -> (scores [Q,k], rows int64 [Q,k]) on the device; rows index insertion order."""
        if self.index is None:
            raise _lib.RmuError("the collection is empty")
        s, i = self.index.search(queries, k)
        return self._score_out(s), i

    def mmr_tensor(self, queries, k: int = 4, fetch_k: int = 20, lambda_mult: float = 0.5):
        """CUDA fp32 [Q, D] -> rows int64 [Q, k] in MMR order (-1 padded), all on the device."""
        if self.index is None:
            raise _lib.RmuError("the collection is empty")
        torch = self.index.torch
        q = queries.to(device=self.index.device, dtype=torch.float32).contiguous()
        _, rows = self.index.search(q, fetch_k)
        n_cand = (rows >= 0).sum(1).to(torch.int32)
        cand = self.index.gather(rows.clamp_min(0).view(-1)).view(q.shape[0], fetch_k, self.index.dim)
        sel = mmr_select(q, cand, n_cand, k, lambda_mult).to(torch.int64)
        picked = torch.gather(rows, 1, sel.clamp_min(0))
        return torch.where(sel >= 0, picked, torch.full_like(picked, -1))

    # ------------------------------------------------------------------ search (reference surface)
    def _doc(self, row: int) -> Document:
        meta = dict(self._metas[row])
        meta[self.id_field] = self._pks[row]
        return Document(page_content=self._texts[row], metadata=meta)

    def _embed_query_tensor(self, query: str):
        torch = _lib.require_cuda()
        if hasattr(self.embedding_func, "encode_tensor"):
            return self.embedding_func.encode_tensor([query])
        return torch.tensor([self.embedding_func.embed_query(query)], dtype=torch.float32, device="cuda")

    def similarity_search_with_score_by_vector(self, embedding: List[float], k: int = 4, **_: Any) -> List[Tuple[Document, float]]:
        if self.index is None or len(self) == 0:
            return []
        torch = self.index.torch
        q = torch.as_tensor(np.asarray(embedding, dtype=np.float32)[None], device=self.index.device)
        s, i = self.search_tensor(q, k)
        s, i = s[0].tolist(), i[0].tolist()
        return [(self._doc(r), float(sc)) for sc, r in zip(s, i) if r >= 0]

    def similarity_search_with_score(self, query: str, k: int = 4, **kw: Any) -> List[Tuple[Document, float]]:
        if self.index is None or len(self) == 0:
            return []
        s, i = self.search_tensor(self._embed_query_tensor(query), k)
        s, i = s[0].tolist(), i[0].tolist()
        return [(self._doc(r), float(sc)) for sc, r in zip(s, i) if r >= 0]

    def similarity_search(self, query: str, k: int = 4, **kw: Any) -> List[Document]:
        return [d for d, _ in self.similarity_search_with_score(query, k, **kw)]

    def similarity_search_by_vector(self, embedding: List[float], k: int = 4, **kw: Any) -> List[Document]:
        return [d for d, _ in self.similarity_search_with_score_by_vector(embedding, k, **kw)]

    def similarity_search_with_relevance_scores(self, query: str, k: int = 4, **kw: Any) -> List[Tuple[Document, float]]:
        out = []
        for d, s in self.similarity_search_with_score(query, k, **kw):
            if self.metric == "l2":
                rel = 1.0 - s / 2.0 ** 0.5          # LangChain's euclidean relevance for unit vectors
            elif self.metric == "cosine":
                rel = 1.0 - s
            else:
                rel = s
            out.append((d, rel))
        return out

    def max_marginal_relevance_search_by_vector(self, embedding: List[float], k: int = 4, fetch_k: int = 20,
                                                lambda_mult: float = 0.5, **_: Any) -> List[Document]:
        if self.index is None or len(self) == 0:
            return []
        torch = self.index.torch
        q = torch.as_tensor(np.asarray(embedding, dtype=np.float32)[None], device=self.index.device)
        rows = self.mmr_tensor(q, k, fetch_k, lambda_mult)[0].tolist()
        return [self._doc(r) for r in rows if r >= 0]

    def max_marginal_relevance_search(self, query: str, k: int = 4, fetch_k: int = 20, lambda_mult: float = 0.5,
                                      **_: Any) -> List[Document]:
        if self.index is None or len(self) == 0:
            return []
        rows = self.mmr_tensor(self._embed_query_tensor(query), k, fetch_k, lambda_mult)[0].tolist()
        return [self._doc(r) for r in rows if r >= 0]

    def as_retriever(self, **kwargs: Any) -> B200Retriever:
        return B200Retriever(self, kwargs.get("search_type", "similarity"), kwargs.get("search_kwargs"))

    # ------------------------------------------------------------------ persistence (vector_store_uri reuse)
    def save(self, path: str) -> None:
        import json
        vec = self.index.data().cpu().numpy() if self.index is not None else np.zeros((0, 0), np.float32)
        np.savez(path, vectors=vec, metric=self.metric,
                 table=np.frombuffer(json.dumps({"pks": self._pks, "texts": self._texts, "metas": self._metas}).encode(),
                                     dtype=np.uint8))

    @classmethod
    def load(cls, path: str, embedding: Any, **kwargs: Any) -> "B200VectorStore":
        import json
        z = np.load(path if path.endswith(".npz") else path + ".npz", allow_pickle=False)
        store = cls(embedding, **kwargs)
        table = json.loads(bytes(z["table"]).decode())
        if z["vectors"].size:
            store.add_embeddings(np.ascontiguousarray(z["vectors"]), table["texts"], table["metas"], table["pks"])
        return store


class Milvus(B200VectorStore):
    """Answers to ``langchain_milvus.vectorstores.Milvus`` as constructed at
    ``server/RAGHelper.py:388-394`` (FLAT / L2; ``drop_old`` clears a reused in-process collection)."""

    _collections: Dict[Tuple[str, str], "Milvus"] = {}

    def __init__(self, embedding_function: Any, collection_name: str = "LangChainCollection",
                 connection_args: Optional[Dict[str, Any]] = None, drop_old: bool = False, auto_id: bool = False,
                 device: Optional[int] = None, **_: Any):
        super().__init__(embedding_function, metric="l2", collection_name=collection_name, device=device)
        self.connection_args = dict(connection_args or {})
        self.drop_old = drop_old
        self.auto_id = auto_id


class PGVector(B200VectorStore):
    """Answers to ``langchain_postgres.vectorstores.PGVector`` as constructed at
    ``server/RAGHelper.py:399-404`` (default distance strategy COSINE, score = 1 - cos)."""

    id_field = "id"

    def __init__(self, embeddings: Any = None, collection_name: str = "langchain", connection: Any = None,
                 use_jsonb: bool = True, device: Optional[int] = None, **_: Any):
        super().__init__(embeddings, metric="cosine", collection_name=collection_name, device=device)
        self.connection = connection
        self.use_jsonb = use_jsonb
