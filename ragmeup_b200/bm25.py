"""BM25Index — Okapi BM25 over an in-memory corpus, scored on the GPU.

Stands where ``rank_bm25.BM25Okapi`` stands under ``langchain_community.retrievers.BM25Retriever``
(built by the reference at ``server/RAGHelper.py:436-443``).  The host part builds the inverted index
and the float64 statistics with the same operations, in the same order, as rank_bm25 (``_initialize``,
``_calc_idf``: document frequencies in first-seen order, ``idf = log(N - df + 0.5) - log(df + 0.5)``,
negative idfs floored at ``epsilon * average_idf``); scoring and top-n run in ``csrc/rmu_bm25.cu``
(float64, bit-identical sums).  There is no CPU scoring path.
"""
from __future__ import annotations

import ctypes as C
import math
from collections import Counter
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib

MAX_K = 256


class InvertedIndex:
    """Host-side build of the inverted index + float64 statistics (rank_bm25 ``_initialize`` / ``_calc_idf``)."""

    def __init__(self, corpus: Sequence[Sequence[str]], k1: float = 1.5, b: float = 0.75, epsilon: float = 0.25):
        self.k1, self.b, self.epsilon = float(k1), float(b), float(epsilon)
        self.vocab: Dict[str, int] = {}
        vocab = self.vocab
        n = len(corpus)
        doc_len = np.empty(n, dtype=np.int64)
        ent_term: List[int] = []
        ent_tf: List[int] = []
        ent_cnt = np.empty(n, dtype=np.int64)          # distinct terms per document
        num_doc = 0
        for i, document in enumerate(corpus):
            doc_len[i] = len(document)
            num_doc += len(document)
            freq = Counter(document)                   # first-seen order, like rank_bm25's per-document dict
            for w, f in freq.items():
                tid = vocab.get(w)
                if tid is None:
                    tid = len(vocab)
                    vocab[w] = tid
                ent_term.append(tid)
                ent_tf.append(f)
            ent_cnt[i] = len(freq)
        terms = np.asarray(ent_term, dtype=np.int64)
        tfs = np.asarray(ent_tf, dtype=np.int32)
        docs = np.repeat(np.arange(n, dtype=np.int32), ent_cnt)
        n_terms = len(vocab)
        df = np.bincount(terms, minlength=n_terms).astype(np.int64)
        order = np.argsort(terms, kind="stable")       # documents stay ascending inside a term
        self.post_ptr = np.zeros(n_terms + 1, dtype=np.int64)
        np.cumsum(df, out=self.post_ptr[1:])
        self.post_doc = np.ascontiguousarray(docs[order])
        self.post_tf = np.ascontiguousarray(tfs[order])
        self.doc_len = doc_len
        self._set_statistics()

    def _set_statistics(self) -> None:
        """avgdl, idf (with the epsilon floor) and the document-length denominators from doc_len / post_ptr, with the
        operations and the summation order of rank_bm25 (``_initialize`` / ``_calc_idf`` / ``get_scores``)."""
        n = self.corpus_size = len(self.doc_len)
        self.avgdl = int(self.doc_len.sum()) / self.corpus_size      # ZeroDivisionError on an empty corpus, as rank_bm25
        df = np.diff(self.post_ptr)
        # _calc_idf: sequential float64 sum in vocabulary (first-seen) order
        idf = [math.log(n - int(f) + 0.5) - math.log(int(f) + 0.5) for f in df]
        idf_sum = 0.0
        for v in idf:
            idf_sum += v
        self.average_idf = idf_sum / len(idf) if idf else 0.0
        eps = self.epsilon * self.average_idf
        self.idf = np.asarray([eps if v < 0 else v for v in idf], dtype=np.float64)
        # document-length part of the denominator, numpy float64 exactly as get_scores evaluates it
        self.den = np.ascontiguousarray(self.k1 * (1 - self.b + self.b * self.doc_len / self.avgdl), dtype=np.float64)

    @classmethod
    def from_texts(cls, texts: Sequence[str], k1: float = 1.5, b: float = 0.75, epsilon: float = 0.25, **kw):
        """Index of ``[t.split() for t in texts]`` built by the C++ host routine ``rmu_bm25_csr_build`` (Python
        ``str.split()`` semantics, first-seen vocabulary order): same arrays as ``cls([t.split() for t in texts])``."""
        try:
            enc = [t.encode("utf-8") for t in texts]
        except UnicodeEncodeError:                      # lone surrogates: only the Python path can represent them
            return cls([t.split() for t in texts], k1, b, epsilon, **kw)
        self = cls.__new__(cls)
        self._prepare(**kw)
        self.k1, self.b, self.epsilon = float(k1), float(b), float(epsilon)
        n = len(enc)
        lib = _lib.lib()
        ptrs = (C.c_char_p * max(n, 1))(*enc)
        lens = np.asarray([len(e) for e in enc], dtype=np.int64)
        h = C.c_void_p()
        _lib.check(lib.rmu_bm25_csr_build(C.cast(ptrs, C.c_void_p), lens.ctypes.data, n, C.byref(h)), "rmu_bm25_csr_build")
        try:
            nt, nnz, vb = C.c_int64(), C.c_int64(), C.c_int64()
            _lib.check(lib.rmu_bm25_csr_sizes(h, C.byref(nt), C.byref(nnz), C.byref(vb)), "rmu_bm25_csr_sizes")
            self.doc_len = np.empty(n, dtype=np.int64)
            self.post_ptr = np.empty(nt.value + 1, dtype=np.int64)
            self.post_doc = np.empty(nnz.value, dtype=np.int32)
            self.post_tf = np.empty(nnz.value, dtype=np.int32)
            voff = np.empty(nt.value + 1, dtype=np.int64)
            vbytes = C.create_string_buffer(max(vb.value, 1))
            _lib.check(lib.rmu_bm25_csr_export(h, self.doc_len.ctypes.data, self.post_ptr.ctypes.data,
                                               self.post_doc.ctypes.data, self.post_tf.ctypes.data, voff.ctypes.data,
                                               C.cast(vbytes, C.c_void_p)), "rmu_bm25_csr_export")
        finally:
            lib.rmu_bm25_csr_free(h)
        raw = vbytes.raw
        self.vocab = {raw[voff[i]:voff[i + 1]].decode("utf-8"): i for i in range(nt.value)}
        self._set_statistics()
        self._finish()
        return self

    def _prepare(self, **kw) -> None:      # hooks for the GPU subclass
        pass

    def _finish(self) -> None:
        pass

    def term_ids(self, query: Sequence[str]) -> List[int]:
        """query tokens -> ids of the terms that contribute (``(self.idf.get(q) or 0)``: unknown terms and terms whose
        idf is exactly 0 add nothing); order and repeats are kept."""
        out = []
        for q in query:
            tid = self.vocab.get(q)
            if tid is not None and self.idf[tid] != 0.0:
                out.append(tid)
        return out


class BM25Index(InvertedIndex):
    def __init__(self, corpus: Sequence[Sequence[str]], k1: float = 1.5, b: float = 0.75, epsilon: float = 0.25,
                 device: Optional[int] = None):
        self._prepare(device=device)
        super().__init__(corpus, k1, b, epsilon)
        self._finish()

    def _prepare(self, device: Optional[int] = None, **_) -> None:
        torch = _lib.require_cuda()
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)

    def _finish(self) -> None:
        self._upload()

    @classmethod
    def from_arrays(cls, post_ptr: np.ndarray, post_doc: np.ndarray, post_tf: np.ndarray, doc_len: np.ndarray,
                    k1: float = 1.5, b: float = 0.75, epsilon: float = 0.25, device: Optional[int] = None) -> "BM25Index":
        """Index over PRE-TOKENISED data: postings CSR by integer term id (documents ascending inside a term) and
        document lengths; statistics are derived exactly as in ``InvertedIndex`` (vocabulary = ``str(term id)``)."""
        self = cls.__new__(cls)
        self._prepare(device=device)
        self.k1, self.b, self.epsilon = float(k1), float(b), float(epsilon)
        self.post_ptr = np.ascontiguousarray(post_ptr, dtype=np.int64)
        self.post_doc = np.ascontiguousarray(post_doc, dtype=np.int32)
        self.post_tf = np.ascontiguousarray(post_tf, dtype=np.int32)
        self.doc_len = np.ascontiguousarray(doc_len, dtype=np.int64)
        self.vocab = {str(t): t for t in range(len(self.post_ptr) - 1)}
        self._set_statistics()
        self._finish()
        return self

    def _upload(self) -> None:
        torch = self.torch
        n, n_terms = self.corpus_size, len(self.post_ptr) - 1
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().rmu_bm25_create(n, n_terms, self.post_ptr.ctypes.data, self.post_doc.ctypes.data,
                                                  self.post_tf.ctypes.data, self.den.ctypes.data, self.idf.ctypes.data,
                                                  self.k1 + 1, C.byref(self._h)), "rmu_bm25_create")

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _lib.lib().rmu_bm25_destroy(h)
            except Exception:
                pass
            self._h = C.c_void_p()

    def __len__(self) -> int:
        return self.corpus_size

    def search(self, queries: Sequence[Sequence[str]], k: int) -> Tuple[np.ndarray, np.ndarray]:
        """tokenised queries -> (scores float64 [Q, k'], document numbers int64 [Q, k']), k' = min(k, corpus size),
        each row ordered as ``np.argsort(scores, kind='stable')[::-1][:k]``."""
        if k < 1:
            raise ValueError("k must be >= 1")
        if k > MAX_K:
            raise _lib.RmuError(f"BM25Index.search: k={k} exceeds {MAX_K}")
        nq = len(queries)
        kk = min(k, self.corpus_size)
        if nq == 0:
            return np.zeros((0, kk)), np.zeros((0, kk), dtype=np.int64)
        ids = [self.term_ids(q) for q in queries]
        q_ptr = np.zeros(nq + 1, dtype=np.int32)
        np.cumsum([len(t) for t in ids], out=q_ptr[1:])
        q_terms = np.asarray([t for row in ids for t in row], dtype=np.int32)
        if q_terms.size == 0:
            q_terms = np.zeros(1, dtype=np.int32)
        scores = np.empty((nq, k), dtype=np.float64)
        docs = np.empty((nq, k), dtype=np.int64)
        with self.torch.cuda.device(self.device):
            _lib.check(_lib.lib().rmu_bm25_search_host(self._h, q_ptr.ctypes.data, q_terms.ctypes.data, nq, k,
                                                       scores.ctypes.data, docs.ctypes.data, _lib.stream_ptr()),
                       "rmu_bm25_search_host")
        return scores[:, :kk], docs[:, :kk]

    # rank_bm25 surface used by BM25Retriever
    def get_top_n(self, query: Sequence[str], documents: Sequence, n: int = 5) -> List:
        _, docs = self.search([query], n)
        return [documents[int(i)] for i in docs[0]]
