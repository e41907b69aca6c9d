"""Bulk-ingest throughput: documents -> WordPiece (host) -> BERT encode (GPU) -> FLAT store,
in the reference's 1000-document add_documents batches (server/RAGHelper.py:55,423-431), plus the BM25 index build
that _initialize_bm25retriever does over the same chunks (:436-443).  PROF_DOCS documents of ~PROF_WORDS words."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ragmeup_b200.documents import Document  # noqa: E402
from ragmeup_b200.embeddings import HuggingFaceEmbeddings  # noqa: E402
from ragmeup_b200.retrievers import BM25Retriever  # noqa: E402
from ragmeup_b200.tokenizer import synthetic_sentences, synthetic_vocab  # noqa: E402
from ragmeup_b200.vectorstore import Milvus  # noqa: E402


def main():
    n = int(os.environ.get("PROF_DOCS", 20000)); words = int(os.environ.get("PROF_WORDS", 70))
    emb = HuggingFaceEmbeddings(model_name="synthetic:all-MiniLM-L6-v2:0", model_kwargs={"device": "cuda"})
    texts = synthetic_sentences(synthetic_vocab(30522), n, words - 10, words + 10, seed=3)
    docs = [Document(t, {"source": f"f{i % 9}.txt", "id": f"id{i}"}) for i, t in enumerate(texts)]
    t0 = time.time()
    ids_, _, cu = emb._ragged(texts[:2000])
    tok_s = (time.time() - t0) / 2000
    print(f"tokenise only: {1 / tok_s:.0f} docs/s ({cu[-1] / 2000:.1f} tokens/doc)", flush=True)
    db = Milvus.from_documents([], emb, drop_old=True, connection_args={"uri": "x.db"}, collection_name="c")
    db.add_documents(docs[:1000], ids=[d.metadata["id"] for d in docs[:1000]])      # warm-up
    db = Milvus.from_documents([], emb, drop_old=True, connection_args={"uri": "x.db"}, collection_name="c")
    torch.cuda.synchronize()
    prof = None
    if os.environ.get("PROF_CPROFILE") == "1":
        import cProfile
        prof = cProfile.Profile()
        prof.enable()
    t0 = time.time()
    for a in range(0, n, 1000):
        db.add_documents(docs[a:a + 1000], ids=[d.metadata["id"] for d in docs[a:a + 1000]])
    torch.cuda.synchronize()
    dt = time.time() - t0
    if prof is not None:
        import pstats
        prof.disable()
        pstats.Stats(prof).sort_stats("cumulative").print_stats(18)
    print(f"add_documents: {n} docs in {dt:.2f}s = {n / dt:.0f} docs/s ({n / dt * cu[-1] / 2000:.0f} tokens/s), "
          f"store rows {len(db)}", flush=True)
    t0 = time.time()
    sp = BM25Retriever.from_texts([d.page_content for d in docs], metadatas=[d.metadata for d in docs])
    print(f"BM25 index build: {n} docs in {time.time() - t0:.2f}s ({len(sp.vectorizer.vocab)} terms, "
          f"{len(sp.vectorizer.post_doc)} postings)", flush=True)


if __name__ == "__main__":
    main()
