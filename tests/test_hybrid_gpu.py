"""GPU: the BM25 sparse leg (csrc/rmu_bm25.cu through the C ABI) against the rank_bm25 restatement, bit for bit,
and the hybrid retriever stack as the reference wires it (server/RAGHelper.py:436-443, 488-503)."""
from dataclasses import asdict

import numpy as np
import pytest

import cases
from oracle import bert_ref, flat_ref
from oracle import hybrid_ref as H
from ragmeup_b200.documents import Document
from ragmeup_b200.tokenizer import synthetic_sentences, synthetic_vocab

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", cases.BM25_CASES, ids=lambda c: c["name"])
def test_bm25_scores_and_order_are_bit_identical(cuda, case):
    from ragmeup_b200.bm25 import BM25Index
    corpus, queries = cases.bm25_inputs(case)
    idx = BM25Index(corpus)
    bm = H.BM25Okapi(corpus)
    k = case["k"]
    scores, docs = idx.search(queries, k)
    assert scores.shape == (len(queries), min(k, len(corpus))) and docs.dtype == np.int64
    for qi, q in enumerate(queries):
        ws, wd = H.bm25_topk(bm, q, k)
        assert np.array_equal(docs[qi], wd), (case["name"], qi)
        assert np.array_equal(scores[qi], ws), (case["name"], qi)      # float64, bit-identical
    if case["n"] <= 10000:
        g = np.load(cases.golden_path("bm25.npz"))
        assert np.array_equal(docs, g[f"{case['name']}_ids"][:, :docs.shape[1]])
        assert np.array_equal(scores, g[f"{case['name']}_scores"][:, :docs.shape[1]])
    # one query at a time gives the same rows as the batch (get_top_n surface)
    assert idx.get_top_n(queries[0], list(range(len(corpus))), n=min(k, 7)) == docs[0][:min(k, 7)].tolist()


def test_bm25_edge_cases(cuda):
    from ragmeup_b200 import _lib
    from ragmeup_b200.bm25 import BM25Index
    # fewer documents than k; a single document; empty documents; all-identical documents (every score ties)
    corpus = [["a", "b"], ["a", "c", "c"], ["d"]]
    idx = BM25Index(corpus)
    bm = H.BM25Okapi(corpus)
    s, d = idx.search([["a", "c"], ["zzz"], [], ["d", "d", "a"]], 10)
    assert s.shape == (4, 3)
    for qi, q in enumerate([["a", "c"], ["zzz"], [], ["d", "d", "a"]]):
        ws, wd = H.bm25_topk(bm, q, 10)
        assert np.array_equal(d[qi], wd) and np.array_equal(s[qi], ws)
    one = BM25Index([["x", "y"]])
    s, d = one.search([["x"]], 4)
    ws, wd = H.bm25_topk(H.BM25Okapi([["x", "y"]]), ["x"], 4)
    assert np.array_equal(d[0], wd) and np.array_equal(s[0], ws)
    same = [["p", "q", "q"]] * 5000 + [[]] * 3
    idx = BM25Index(same)
    s, d = idx.search([["q"], ["p", "nope"]], 6)
    for qi, q in enumerate([["q"], ["p", "nope"]]):
        ws, wd = H.bm25_topk(H.BM25Okapi(same), q, 6)
        assert np.array_equal(d[qi], wd) and np.array_equal(s[qi], ws)
    with pytest.raises(_lib.RmuError):
        idx.search([["q"]], 257)
    with pytest.raises(ZeroDivisionError):
        BM25Index([])


def test_bm25_non_default_parameters(cuda):
    from ragmeup_b200.bm25 import BM25Index
    corpus, queries = cases.bm25_inputs(cases.BM25_CASES[0])
    for params in (dict(k1=1.2, b=0.5, epsilon=0.1), dict(k1=2.0, b=0.0, epsilon=0.25), dict(k1=0.9, b=1.0, epsilon=0.5)):
        idx = BM25Index(corpus, **params)
        bm = H.BM25Okapi(corpus, **params)
        s, d = idx.search(queries, 5)
        for qi, q in enumerate(queries):
            ws, wd = H.bm25_topk(bm, q, 5)
            assert np.array_equal(d[qi], wd) and np.array_equal(s[qi], ws)


def test_hybrid_stack_matches_reference_wiring(cuda):
    """sparse (BM25, k = 4) + dense (MMR, k = 10) -> weighted RRF -> cross-encoder rerank -> top_n, as
    RAGHelper._setup_retrievers / _initialize_reranker build it, against the oracle restatement of every stage."""
    from ragmeup_b200.cross_encoder import HuggingFaceCrossEncoder
    from ragmeup_b200.embeddings import HuggingFaceEmbeddings
    from ragmeup_b200.reranker import ScoredCrossEncoderReranker
    from ragmeup_b200.retrievers import BM25Retriever, ContextualCompressionRetriever, EnsembleRetriever
    from ragmeup_b200.vectorstore import Milvus
    from ragmeup_b200.weights import resolve_model
    emb = HuggingFaceEmbeddings(model_name="synthetic:all-MiniLM-L6-v2:0", model_kwargs={"device": "cuda"})
    ce = HuggingFaceCrossEncoder(model_name="synthetic:ms-marco-MiniLM-L-6-v2:1:4.0")
    vocab = synthetic_vocab(30522)
    texts = synthetic_sentences(vocab, 600, 30, 60, seed=17)
    queries = [" ".join(t.split()[3:9]) for t in texts[5:9]]          # queries share words with some documents
    documents = [Document(t, {"source": f"f{i % 5}.txt", "id": f"id{i}"}) for i, t in enumerate(texts)]
    db = Milvus.from_documents([], emb, drop_old=True, connection_args={"uri": "data.db"}, collection_name="c")
    db.add_documents(documents, ids=[d.metadata["id"] for d in documents])
    sparse = BM25Retriever.from_texts([d.page_content for d in documents], metadatas=[d.metadata for d in documents])
    dense = db.as_retriever(search_type="mmr", search_kwargs={"k": 10})
    ensemble = EnsembleRetriever(retrievers=[sparse, dense], weights=[0.5, 0.5])
    rerank = ContextualCompressionRetriever(base_compressor=ScoredCrossEncoderReranker(model=ce, top_n=3),
                                            base_retriever=ensemble)
    # oracle side
    ecfg, ew, *_ = resolve_model(emb.model_name, with_head=False)
    ccfg, cw, *_ = resolve_model(ce.model_name, with_head=True)
    ecfg, ccfg = bert_ref.BertCfg(**asdict(ecfg)), bert_ref.BertCfg(**asdict(ccfg))
    X = bert_ref.st_encode(ew, ecfg, emb.tokenizer, texts, "mean", True, 256)
    o_sparse = H.BM25Retriever.from_texts(texts, make_doc=lambda t, m: t)
    assert sparse.k == 4 and o_sparse.k == 4
    for q in queries:
        sp = o_sparse.invoke(q)
        assert [d.page_content for d in sparse.invoke(q)] == sp
        qv = np.asarray(bert_ref.hf_embed_query(ew, ecfg, emb.tokenizer, q, max_seq_length=256), dtype=np.float32)
        de = [texts[r] for r in flat_ref.mmr_search(qv[None], X, 10, "l2", fetch_k=20, lambda_mult=0.5)[0]]
        assert [d.page_content for d in dense.invoke(q)] == de
        fused = H.weighted_reciprocal_rank([sp, de], [0.5, 0.5], key=lambda t: t)
        got = ensemble.invoke(q)
        assert [d.page_content for d in got] == fused
        assert all("source" in d.metadata for d in got)
        logits = bert_ref.cross_encoder_predict(cw, ccfg, ce.tokenizer, [(q, t) for t in fused], max_length=512,
                                                activation=ce.activation)
        order = sorted(range(len(fused)), key=lambda i: logits[i], reverse=True)[:3]
        out = rerank.invoke(q)
        gaps = np.diff(np.sort(np.asarray(logits))[::-1])
        if np.all(np.abs(gaps) > 2e-3):          # only when the oracle's own order is unambiguous at the 1e-3 tolerance
            assert [d.page_content for d in out] == [fused[i] for i in order]
        assert np.abs(np.array([d.metadata["relevance_score"] for d in out]) -
                      np.sort(np.asarray(logits))[::-1][:3]).max() < 1e-3
    # batched sparse path == one query at a time
    rows = sparse.batch(queries)
    assert [[d.page_content for d in r] for r in rows] == [[d.page_content for d in sparse.invoke(q)] for q in queries]


def test_similarity_provenance_matches_oracle(cuda, monkeypatch):
    """DocumentSimilarityAttribution.compute_similarity (server/provenance.py:164-202) on the GPU encoder + cosine
    search vs the restatement on the oracle encoder."""
    from oracle import provenance_ref as P
    from ragmeup_b200.embeddings import HuggingFaceEmbeddings
    from ragmeup_b200.provenance import DocumentSimilarityAttribution
    from ragmeup_b200.weights import resolve_model
    emb = HuggingFaceEmbeddings(model_name="synthetic:all-MiniLM-L6-v2:0", model_kwargs={"device": "cuda"})
    ecfg, ew, *_ = resolve_model(emb.model_name, with_head=False)
    ecfg = bert_ref.BertCfg(**asdict(ecfg))
    vocab = synthetic_vocab(30522)
    context = synthetic_sentences(vocab, 7, 30, 60, seed=31)
    answer = " ".join(context[2].split()[:20])
    query = " ".join(context[4].split()[5:12])
    enc = lambda texts: bert_ref.st_encode(ew, ecfg, emb.tokenizer, list(texts), "mean", True, 256)  # noqa: E731
    attr = DocumentSimilarityAttribution(embeddings=emb)
    for flag in (None, "False"):
        if flag is None:
            monkeypatch.delenv("attribute_include_query", raising=False)
        else:
            monkeypatch.setenv("attribute_include_query", flag)
        got = attr.compute_similarity(query, [Document(t, {}) for t in context], answer)
        want = P.compute_similarity(enc, query, context, answer, include_query=flag is None)
        assert len(got) == 7 and np.abs(np.asarray(got) - np.asarray(want, dtype=np.float64)).max() < 1e-3
        assert int(np.argmax(got)) == int(np.argmax(want))
    assert attr.compute_similarity(query, [], answer) == []
