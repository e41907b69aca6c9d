"""The reference's retrieval stack end to end on ragmeup_b200, without LangChain (needs a B200):

    RAGHelper._setup_retrievers / _initialize_reranker (server/RAGHelper.py:436-503) wire
        sparse  = BM25Retriever.from_texts(chunks)                                   # GPU BM25, k = 4
        dense   = db.as_retriever(search_type="mmr", search_kwargs={"k": vector_store_k})
        fused   = EnsembleRetriever([sparse, dense], weights=[0.5, 0.5])             # weighted RRF, c = 60
        rerank  = ContextualCompressionRetriever(ScoredCrossEncoderReranker(HuggingFaceCrossEncoder(...), top_n), fused)

Model names: a local HuggingFace snapshot directory, or `synthetic:<preset>[:seed[:scale]]` for seeded random weights
(this script's default, since the build image has no network).  The same objects are what `ragmeup_b200.install()`
puts behind the reference's own import lines.
"""
import sys

from ragmeup_b200.cross_encoder import HuggingFaceCrossEncoder
from ragmeup_b200.documents import Document
from ragmeup_b200.embeddings import HuggingFaceEmbeddings
from ragmeup_b200.provenance import DocumentSimilarityAttribution
from ragmeup_b200.reranker import ScoredCrossEncoderReranker
from ragmeup_b200.retrievers import BM25Retriever, ContextualCompressionRetriever, EnsembleRetriever
from ragmeup_b200.tokenizer import synthetic_sentences, synthetic_vocab
from ragmeup_b200.vectorstore import Milvus


def main(n_docs: int = 5000) -> None:
    embeddings = HuggingFaceEmbeddings(model_name="synthetic:all-MiniLM-L6-v2:0", model_kwargs={"device": "cuda"})
    cross_encoder = HuggingFaceCrossEncoder(model_name="synthetic:ms-marco-MiniLM-L-6-v2:1:4.0")
    vocab = synthetic_vocab(30522)
    chunks = [Document(t, {"source": f"file{i % 13}.txt", "id": f"{i:08x}"})
              for i, t in enumerate(synthetic_sentences(vocab, n_docs, 60, 110, seed=1))]

    db = Milvus.from_documents([], embeddings, drop_old=True, connection_args={"uri": "data.db"}, collection_name="c")
    for a in range(0, n_docs, 1000):                                   # RAGHelper inserts in batches of 1000
        db.add_documents(chunks[a:a + 1000], ids=[d.metadata["id"] for d in chunks[a:a + 1000]])

    sparse = BM25Retriever.from_texts([d.page_content for d in chunks], metadatas=[d.metadata for d in chunks])
    dense = db.as_retriever(search_type="mmr", search_kwargs={"k": 10})
    fused = EnsembleRetriever(retrievers=[sparse, dense], weights=[0.5, 0.5])
    rerank = ContextualCompressionRetriever(base_compressor=ScoredCrossEncoderReranker(model=cross_encoder, top_n=3),
                                            base_retriever=fused)

    query = " ".join(chunks[42].page_content.split()[5:14])
    docs = rerank.invoke(query)
    for d in docs:
        print(f"{d.metadata['relevance_score']:+.4f}  {d.metadata['source']:<12s} {d.page_content[:70]}...")
    shares = DocumentSimilarityAttribution(embeddings=embeddings).compute_similarity(query, docs, docs[0].page_content[:200])
    print("similarity provenance:", [round(s, 3) for s in shares])


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 5000)
