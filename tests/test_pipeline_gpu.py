"""GPU: the drop-in classes end to end against the oracle's restatement of the reference stack
(HuggingFaceEmbeddings -> Milvus/PGVector + MMR retriever -> ScoredCrossEncoderReranker)."""
from dataclasses import asdict

import numpy as np
import pytest
import torch

from oracle import bert_ref, flat_ref
from ragmeup_b200.documents import Document
from ragmeup_b200.tokenizer import synthetic_sentences, synthetic_vocab

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def stack(cuda):
    from ragmeup_b200.cross_encoder import HuggingFaceCrossEncoder
    from ragmeup_b200.embeddings import HuggingFaceEmbeddings
    emb = HuggingFaceEmbeddings(model_name="synthetic:all-MiniLM-L6-v2:0", model_kwargs={"device": "cuda"})
    ce = HuggingFaceCrossEncoder(model_name="synthetic:ms-marco-MiniLM-L-6-v2:1:4.0")
    vocab = synthetic_vocab(30522)
    docs = synthetic_sentences(vocab, 1000, 60, 90, seed=7)      # C1: 1k-doc corpus
    queries = synthetic_sentences(vocab, 16, 5, 20, seed=8)
    return emb, ce, docs, queries


def _oracle_models(emb, ce):
    from ragmeup_b200.weights import resolve_model
    ecfg, ew, *_ = resolve_model(emb.model_name, with_head=False)
    ccfg, cw, *_ = resolve_model(ce.model_name, with_head=True)
    return (ew, bert_ref.BertCfg(**asdict(ecfg))), (cw, bert_ref.BertCfg(**asdict(ccfg)))


def test_embed_documents_and_query_match_oracle(stack):
    emb, ce, docs, queries = stack
    (ew, ecfg), _ = _oracle_models(emb, ce)
    texts = docs[:40] + ["line one\nline two  ", ""]
    got = np.asarray(emb.embed_documents(texts), dtype=np.float32)
    ref = np.asarray(bert_ref.hf_embed_documents(ew, ecfg, emb.tokenizer, texts, pooling="mean", normalize=True,
                                                 max_seq_length=emb.max_seq_length), dtype=np.float32)
    assert got.shape == (42, 384) and np.abs(got - ref).max() < 1e-3
    one = np.asarray(emb.embed_query(queries[0]))
    assert np.abs(one - np.asarray(bert_ref.hf_embed_query(ew, ecfg, emb.tokenizer, queries[0], max_seq_length=256))).max() < 1e-3
    assert isinstance(emb.embed_query("x"), list) and isinstance(emb.embed_query("x")[0], float)
    assert emb.embed_documents([]) == []


def test_c1_dense_retrieval_matches_reference_wiring(stack):
    """C1 (1k docs, top-10, no rerank): Milvus (L2) store + MMR retriever as RAGHelper wires it
    (server/RAGHelper.py:388-394, 497-499) returns the same documents, in the same order, as the
    oracle's restatement of embed_query -> FLAT/L2 top-20 -> MMR(0.5) -> 10."""
    from ragmeup_b200.vectorstore import Milvus, PGVector
    emb, ce, docs, queries = stack
    (ew, ecfg), _ = _oracle_models(emb, ce)
    documents = [Document(t, {"source": f"f{i % 7}.txt", "id": f"id{i}"}) for i, t in enumerate(docs)]
    db = Milvus.from_documents([], emb, drop_old=True, connection_args={"uri": "data.db"}, collection_name="c")
    for a in range(0, 1000, 400):
        db.add_documents(documents[a:a + 400], ids=[d.metadata["id"] for d in documents[a:a + 400]])
    assert len(db) == 1000
    X = bert_ref.st_encode(ew, ecfg, emb.tokenizer, [t.replace("\n", " ") for t in docs], "mean", True, 256)
    retriever = db.as_retriever(search_type="mmr", search_kwargs={"k": 10})
    for qtext in queries[:8]:
        qv = np.asarray(bert_ref.hf_embed_query(ew, ecfg, emb.tokenizer, qtext, max_seq_length=256), dtype=np.float32)
        want = flat_ref.mmr_search(qv[None], X, 10, "l2", fetch_k=20, lambda_mult=0.5)[0]
        got = retriever.invoke(qtext)
        assert [d.metadata["pk"] for d in got] == [f"id{r}" for r in want]
        assert got[0].metadata["source"] == documents[want[0]].metadata["source"]
        # plain top-k with scores = squared L2 distance, ascending
        pairs = db.similarity_search_with_score(qtext, k=10)
        rs, ri = flat_ref.flat_search(qv[None], X, 10, "l2")
        assert [d.metadata["pk"] for d, _ in pairs] == [f"id{r}" for r in ri[0]]
        assert np.abs(np.array([s for _, s in pairs]) - rs[0]).max() < 1e-3
    chained = (retriever | (lambda ds: len(ds))).invoke(queries[0])
    assert chained == 10
    pg = PGVector(embeddings=emb, collection_name="c", connection="postgresql://x", use_jsonb=True)
    pg.add_documents(documents[:300], ids=[d.metadata["id"] for d in documents[:300]])
    qv = np.asarray(bert_ref.hf_embed_query(ew, ecfg, emb.tokenizer, queries[1], max_seq_length=256), dtype=np.float32)
    rs, ri = flat_ref.flat_search(qv[None], X[:300], 5, "cosine")
    got = pg.similarity_search_with_score(queries[1], k=5)
    assert [d.metadata["id"] for d, _ in got] == [f"id{r}" for r in ri[0]]
    assert np.abs(np.array([s for _, s in got]) - (1.0 - rs[0])).max() < 1e-3     # cosine DISTANCE


def test_rerank_matches_reference_semantics(stack, tmp_path):
    """C4-shaped: cross-encoder scores of (query, doc) pairs within 1e-3 of the oracle's
    CrossEncoder.predict restatement; ScoredCrossEncoderReranker keeps the same documents."""
    from ragmeup_b200.reranker import ScoredCrossEncoderReranker
    emb, ce, docs, queries = stack
    _, (cw, ccfg) = _oracle_models(emb, ce)
    pairs = [(queries[0], d) for d in docs[:70]] + [("  padded query ", docs[3] + " " + docs[4] * 9)]   # last one truncates at 512
    got = ce.score(pairs)
    ref = bert_ref.cross_encoder_predict(cw, ccfg, ce.tokenizer, pairs, max_length=512)
    assert got.dtype == np.float32 and got.shape == (71,)
    assert np.abs(got - ref).max() < 1e-3
    documents = [Document(t, {"source": "s"}) for t in docs[:70]]
    out = ScoredCrossEncoderReranker(model=ce, top_n=10).compress_documents(documents, queries[0])
    order = sorted(range(70), key=lambda j: ref[j], reverse=True)[:10]
    assert [d.page_content for d in out] == [docs[j] for j in order]
    assert abs(float(out[0].metadata["relevance_score"]) - float(ref[order[0]])) < 1e-3
    with pytest.raises(IndexError):
        ce.score([])
    # persistence round trip (vector_store_uri reuse)
    from ragmeup_b200.vectorstore import Milvus
    db = Milvus(emb, collection_name="p")
    db.add_documents(documents[:50], ids=[f"k{i}" for i in range(50)])
    db.save(str(tmp_path / "store"))
    db2 = Milvus.load(str(tmp_path / "store"), emb)
    a = [d.metadata["pk"] for d in db.similarity_search(queries[2], k=5)]
    b = [d.metadata["pk"] for d in db2.similarity_search(queries[2], k=5)]
    assert a == b and len(db2) == 50


def test_concurrent_callers_share_handles(stack):
    """Flask serves /chat on several threads with ONE shared RAGHelper (server/server.py:141-146,394):
    concurrent embed_query / search / score on the same handles must return the single-threaded answers."""
    import threading
    from ragmeup_b200.vectorstore import Milvus
    emb, ce, docs, queries = stack
    db = Milvus(emb, collection_name="t")
    db.add_documents([Document(t, {"source": "s"}) for t in docs[:500]], ids=[f"k{i}" for i in range(500)])
    want_docs = {q: [d.metadata["pk"] for d in db.similarity_search(q, k=5)] for q in queries[:6]}
    pairs = [(queries[0], d) for d in docs[:12]]
    want_scores = ce.score(pairs)
    want_emb = np.asarray(emb.embed_query(queries[1]))
    errors = []

    def worker(tid):
        try:
            for it in range(6):
                q = queries[(tid + it) % 6]
                got = [d.metadata["pk"] for d in db.similarity_search(q, k=5)]
                assert got == want_docs[q]
                assert np.abs(ce.score(pairs) - want_scores).max() < 1e-6
                assert np.abs(np.asarray(emb.embed_query(queries[1])) - want_emb).max() < 1e-6
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_growth_under_concurrent_search_and_scoring(stack):
    """Locking holes of round 1 (rmu_index_gather read the corpus pointer without the handle mutex; the *_host encoder
    entry points staged into buffers another thread's workspace growth could free): one thread keeps ADDING documents
    (the index reallocates and frees its old corpus) and scoring ever larger batches (the activation workspace is
    re-allocated) while other threads run MMR searches (search + gather) and score through the host entry points."""
    import threading
    from ragmeup_b200.vectorstore import Milvus
    emb, ce, docs, queries = stack
    db = Milvus(emb, collection_name="grow")
    db.add_documents([Document(t, {"source": "s"}) for t in docs[:64]], ids=[f"k{i}" for i in range(64)])
    probe = docs[3]                                         # an exact copy of a stored text: distance 0, always first
    pairs = [(queries[0], d) for d in docs[:8]]
    want_scores = ce.score(pairs)
    errors, stop = [], threading.Event()

    def writer():
        try:
            n = 64
            for step in range(10):
                grow = 64 * (step + 1)
                batch = [Document(docs[(n + j) % len(docs)] + f" #{n + j}", {"source": "s"}) for j in range(grow)]
                db.add_documents(batch, ids=[f"k{n + j}" for j in range(grow)])
                n += grow
                ce.score([(queries[1], d) for d in docs[: 40 * (step + 1)]])      # a larger token batch every time
        except Exception as e:  # noqa: BLE001
            errors.append("writer: " + repr(e))
        finally:
            stop.set()

    def reader():
        try:
            while not stop.is_set():
                got = db.max_marginal_relevance_search(probe, k=4, fetch_k=12)
                assert got and got[0].page_content == probe
                assert np.abs(ce.score(pairs) - want_scores).max() < 1e-6
        except Exception as e:  # noqa: BLE001
            errors.append("reader: " + repr(e))
            stop.set()

    threads = [threading.Thread(target=writer)] + [threading.Thread(target=reader) for _ in range(3)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert len(db) == 64 + sum(64 * (s + 1) for s in range(10))
