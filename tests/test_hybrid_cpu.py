"""CPU: the sparse-leg oracle (BM25Okapi restatement), the host index build, and the fusion / compression retrievers
(reference wiring server/RAGHelper.py:436-443, 488-503)."""
import math
import sys

import numpy as np
import pytest

from oracle import hybrid_ref as H
from ragmeup_b200.bm25 import InvertedIndex
from ragmeup_b200.documents import Document
from ragmeup_b200.retrievers import ContextualCompressionRetriever, EnsembleRetriever
import cases


def test_bm25_oracle_known_answer():
    # worked by hand from the Okapi formula: N = 3, avgdl = 2, k1 = 1.5, b = 0.75, epsilon = 0.25
    corpus = [["a", "b"], ["a", "c", "c"], ["d"]]
    bm = H.BM25Okapi(corpus)
    pos = math.log(2.5) - math.log(1.5)            # df = 1
    neg = math.log(1.5) - math.log(2.5)            # df = 2 ("a"): negative -> epsilon * average idf
    eps = 0.25 * ((neg + 3 * pos) / 4)
    assert bm.avgdl == 2.0
    assert bm.idf == {"a": eps, "b": pos, "c": pos, "d": pos}
    s = bm.get_scores(["a", "c"])
    d0 = eps * (1 * 2.5 / (1 + 1.5 * (1 - 0.75 + 0.75 * 2 / 2)))
    d1 = eps * (1 * 2.5 / (1 + 1.5 * (1 - 0.75 + 0.75 * 3 / 2))) + pos * (2 * 2.5 / (2 + 1.5 * (1 - 0.75 + 0.75 * 3 / 2)))
    assert s.tolist() == [d0, d1, 0.0]
    assert bm.get_top_n(["a", "c"], ["D0", "D1", "D2"], n=2) == ["D1", "D0"]
    # no match at all: every score is 0 and the (stable) reversed argsort returns the LAST documents first
    assert bm.get_top_n(["zzz"], ["D0", "D1", "D2"], n=2) == ["D2", "D1"]


@pytest.mark.parametrize("case", cases.BM25_CASES, ids=lambda c: c["name"])
def test_inverted_index_reproduces_rank_bm25_statistics(case):
    corpus, queries = cases.bm25_inputs(case)
    if len(corpus) > 5000:
        pytest.skip("large case: GPU suite")
    bm = H.BM25Okapi(corpus)
    ii = InvertedIndex(corpus)
    assert ii.avgdl == bm.avgdl and ii.average_idf == bm.average_idf and ii.corpus_size == bm.corpus_size
    assert set(ii.vocab) == set(bm.idf)
    assert all(ii.idf[ii.vocab[w]] == v for w, v in bm.idf.items())
    # postings hold exactly the per-document frequencies, documents ascending inside a term
    for w, t in list(ii.vocab.items())[:50]:
        lo, hi = ii.post_ptr[t], ii.post_ptr[t + 1]
        docs = ii.post_doc[lo:hi]
        assert np.all(np.diff(docs) > 0)
        assert [bm.doc_freqs[d][w] for d in docs] == ii.post_tf[lo:hi].tolist()
        assert hi - lo == sum(1 for f in bm.doc_freqs if w in f)
    # the arithmetic the kernel performs (postings only, den precomputed, terms in query order) gives the SAME float64
    # sums as rank_bm25's dense pass
    for q in queries[:8]:
        want = bm.get_scores(q)
        got = np.zeros(len(corpus))
        for t in ii.term_ids(q):
            lo, hi = ii.post_ptr[t], ii.post_ptr[t + 1]
            d = ii.post_doc[lo:hi]
            tf = ii.post_tf[lo:hi].astype(np.float64)
            got[d] = got[d] + ii.idf[t] * ((tf * (ii.k1 + 1)) / (tf + ii.den[d]))
        assert np.array_equal(got, want)


class _Fixed:
    def __init__(self, docs):
        self.docs = docs

    def invoke(self, query, config=None, **kw):
        return list(self.docs)


def _docs(*texts):
    return [Document(page_content=t, metadata={"source": t}) for t in texts]


def test_ensemble_weighted_rrf_matches_oracle():
    sparse = _Fixed(_docs("a", "b", "c", "d"))
    dense = _Fixed(_docs("c", "a", "e"))
    ens = EnsembleRetriever(retrievers=[sparse, dense], weights=[0.5, 0.5])
    got = [d.page_content for d in ens.invoke("q")]
    want = [d.page_content for d in H.ensemble_invoke([sparse, dense], [0.5, 0.5], "q")]
    assert got == want
    # a: .5/61 + .5/62, c: .5/63 + .5/61 -> a > c > b (.5/62) > e (.5/63 from dense rank 3) == ... check by hand
    sc = {"a": .5 / 61 + .5 / 62, "b": .5 / 62, "c": .5 / 63 + .5 / 61, "d": .5 / 64, "e": .5 / 63}
    assert got == sorted(sc, key=lambda k: -sc[k])
    # equal scores keep first-seen order (sparse list first): stable sort
    ens2 = EnsembleRetriever(retrievers=[_Fixed(_docs("x")), _Fixed(_docs("y"))], weights=[0.5, 0.5])
    assert [d.page_content for d in ens2.invoke("q")] == ["x", "y"]
    # default weights are uniform; strings are wrapped into Documents; id_key switches the merge key
    ens3 = EnsembleRetriever(retrievers=[_Fixed(["s1", "s2"]), _Fixed(["s2"])])
    assert [d.page_content for d in ens3.invoke("q")] == ["s2", "s1"]
    a1, a2 = Document("same", {"id": 1}), Document("same", {"id": 2})
    ens4 = EnsembleRetriever(retrievers=[_Fixed([a1]), _Fixed([a2])], weights=[0.5, 0.5], id_key="id")
    assert len(ens4.invoke("q")) == 2
    with pytest.raises(ValueError):
        EnsembleRetriever(retrievers=[sparse, dense], weights=[1.0]).invoke("q")
    # retriever | fn composition used by the reference's chains (RAGHelper_local.py:157-159)
    assert (ens | (lambda docs: len(docs))).invoke("q") == 5


def test_contextual_compression_retriever():
    class Comp:
        def compress_documents(self, documents, query, callbacks=None):
            return [d for d in documents if d.page_content != "b"][:2]

    base = _Fixed(_docs("a", "b", "c", "d"))
    r = ContextualCompressionRetriever(base_compressor=Comp(), base_retriever=base)
    assert [d.page_content for d in r.invoke("q")] == ["a", "c"]
    assert [d.page_content for d in H.contextual_compression_invoke(Comp(), base, "q")] == ["a", "c"]
    assert ContextualCompressionRetriever(base_compressor=Comp(), base_retriever=_Fixed([])).invoke("q") == []


def test_install_registers_retriever_modules():
    from ragmeup_b200 import install
    c = install.install()
    from langchain_community.retrievers import BM25Retriever as B
    from langchain.retrievers import ContextualCompressionRetriever as CC, EnsembleRetriever as E
    assert B is c["BM25Retriever"] and CC is c["ContextualCompressionRetriever"] and E is c["EnsembleRetriever"]
    assert "langchain.retrievers" in sys.modules


def test_bm25_retriever_needs_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from ragmeup_b200 import _lib
    from ragmeup_b200.retrievers import BM25Retriever
    with pytest.raises(_lib.RmuError):
        BM25Retriever.from_texts(["a b", "c d"])


def test_bm25_golden_fixture_matches_oracle():
    g = np.load(cases.golden_path("bm25.npz"))
    for case in cases.BM25_CASES:
        corpus, queries = cases.bm25_inputs(case)
        if len(corpus) > 5000:
            continue
        bm = H.BM25Okapi(corpus)
        for qi, q in enumerate(queries):
            s, d = H.bm25_topk(bm, q, case["k"])
            assert np.array_equal(g[f"{case['name']}_ids"][qi][:len(d)], d)
            assert np.array_equal(g[f"{case['name']}_scores"][qi][:len(s)], s)


def test_rerank_provenance_text_selection(monkeypatch):
    """server/provenance.py:100-108: answer alone, or query + newline + answer when attribute_include_query == "True"."""
    from oracle import provenance_ref as P
    from ragmeup_b200.provenance import compute_rerank_provenance

    class Rec:
        def compress_documents(self, documents, query, callbacks=None):
            return [(d, query) for d in documents]

    monkeypatch.delenv("attribute_include_query", raising=False)
    assert compute_rerank_provenance(Rec(), "Q", ["d1", "d2"], "A") == [("d1", "A"), ("d2", "A")]
    assert compute_rerank_provenance(Rec(), "Q", ["d1"], "A") == P.compute_rerank_provenance(Rec(), "Q", ["d1"], "A", False)
    monkeypatch.setenv("attribute_include_query", "True")
    assert compute_rerank_provenance(Rec(), "Q", ["d1"], "A") == [("d1", "Q\nA")]
    assert compute_rerank_provenance(Rec(), "Q", ["d1"], "A") == P.compute_rerank_provenance(Rec(), "Q", ["d1"], "A", True)


def test_similarity_provenance_oracle_known_answer():
    from oracle import provenance_ref as P
    table = {"ans": [1.0, 0.0], "qry": [0.0, 1.0], "d0": [1.0, 0.0], "d1": [1.0, 1.0], "d2": [0.0, 2.0]}
    enc = lambda texts: np.asarray([table[t] for t in texts], dtype=np.float32)  # noqa: E731
    got = P.compute_similarity(enc, "qry", ["d0", "d1", "d2"], "ans", include_query=True)
    r = np.float32(1.0) / np.sqrt(np.float32(2.0))
    raw = [np.float32(0.5), (r + r) / 2, np.float32(0.5)]
    assert np.allclose(got, [x / sum(raw) for x in raw], atol=1e-7) and abs(sum(got) - 1) < 1e-6
    only_answer = P.compute_similarity(enc, "qry", ["d0", "d2"], "ans", include_query=False)
    assert np.allclose(only_answer, [1.0, 0.0])


_WS = [" ", "\t", "\n", "\x0b", "\x0c", "\r", "\x1c", "\x1d", "\x1e", "\x1f", "\x85", "\xa0", " ", " ", " ",
       " ", " ", " ", " ", " ", "　", "  \t "]
_WORDS = ["a", "bb", "héllo", "日本語", "x​y", "naïve", "\U0001F600", "tab", "q" * 40, "́acc",
          "zero\x00nul", "\x1bESC", "A", "a", "¡", "⁠wj", "﻿bom"]


def _check_same_index(texts):
    a = InvertedIndex([t.split() for t in texts])
    b = InvertedIndex.from_texts(texts)
    assert a.vocab == b.vocab
    for name in ("post_ptr", "post_doc", "post_tf", "doc_len", "idf", "den"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    assert a.avgdl == b.avgdl and a.average_idf == b.average_idf


def test_cpp_index_builder_equals_python_split_and_count():
    """rmu_bm25_csr_build (host C++): Python str.split() semantics on UTF-8 incl. every str.isspace() code point,
    NUL bytes, zero-width characters that are NOT whitespace, empty / all-blank documents; first-seen vocabulary."""
    rng = np.random.default_rng(0)
    texts = []
    for _ in range(1500):
        t = _WS[int(rng.integers(len(_WS)))] if rng.random() < 0.3 else ""
        for _ in range(int(rng.integers(0, 30))):
            t += _WORDS[int(rng.integers(len(_WORDS)))] + _WS[int(rng.integers(len(_WS)))]
        texts.append(t)
    texts += ["", "   ", "　　", "single", "trailing ", " leading"]
    _check_same_index(texts)
    # every whitespace code point Python knows, against every neighbouring non-space code point class
    spaces = [chr(c) for c in range(0x3001) if chr(c).isspace()]
    assert len(spaces) == 29
    _check_same_index(["x" + s + "y" + s + s + "x" for s in spaces] + ["p​q ᠎ r s"])
    # lone surrogates cannot be UTF-8 encoded: the Python path takes over
    odd = InvertedIndex.from_texts(["ok \ud800 fine", "fine"])
    assert set(odd.vocab) == {"ok", "\ud800", "fine"}


def test_cpp_index_builder_property():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.text(alphabet=st.characters(blacklist_categories=("Cs",)), max_size=40), min_size=1, max_size=12))
    def prop(texts):
        if not any(t.split() for t in texts):
            texts = texts + ["w"]
        _check_same_index(texts)

    prop()


def test_cosine_restatement_pinned_against_sklearn():
    """oracle/provenance_ref.sk_cosine_similarity vs the real sklearn.metrics.pairwise.cosine_similarity (installed in
    this image; the reference imports it at server/provenance.py:6) — float32 inputs, zero rows included."""
    sk = pytest.importorskip("sklearn.metrics.pairwise")
    from oracle import provenance_ref as P
    rng = np.random.default_rng(3)
    X = rng.standard_normal((17, 384)).astype(np.float32)
    Y = rng.standard_normal((5, 384)).astype(np.float32)
    X[4] = 0.0
    Y[2] = 0.0
    want = sk.cosine_similarity(X, Y)
    got = P.sk_cosine_similarity(X, Y)
    assert got.dtype == want.dtype == np.float32
    assert np.abs(got - want).max() < 2e-7 and np.all(got[4] == 0) and np.all(got[:, 2] == 0)
