"""Text-level end to end through the drop-in classes, tokenisation included: query strings -> embed_documents ->
top-R over a document store -> (query, document) pair strings -> HuggingFaceCrossEncoder.score -> top-10.
PROF_DOCS documents (default 50k, ~110 WordPiece tokens each), PROF_Q queries per step, PROF_R candidates."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ragmeup_b200.cross_encoder import HuggingFaceCrossEncoder  # noqa: E402
from ragmeup_b200.documents import Document  # noqa: E402
from ragmeup_b200.embeddings import HuggingFaceEmbeddings  # noqa: E402
from ragmeup_b200.tokenizer import synthetic_sentences, synthetic_vocab  # noqa: E402
from ragmeup_b200.vectorstore import PGVector  # noqa: E402


def main():
    n = int(os.environ.get("PROF_DOCS", 50000)); Q = int(os.environ.get("PROF_Q", 64)); R = int(os.environ.get("PROF_R", 100))
    emb = HuggingFaceEmbeddings(model_name="synthetic:all-MiniLM-L6-v2:0", model_kwargs={"device": "cuda"})
    ce = HuggingFaceCrossEncoder(model_name="synthetic:ms-marco-MiniLM-L-6-v2:1:4.0")
    vocab = synthetic_vocab(30522)
    texts = synthetic_sentences(vocab, n, 100, 125, seed=3)
    queries = synthetic_sentences(vocab, Q, 10, 16, seed=4)
    db = PGVector(embeddings=emb, collection_name="c", connection="postgresql://x", use_jsonb=True)
    for a in range(0, n, 1000):
        db.add_documents([Document(t, {"source": "s", "id": str(a + i)}) for i, t in enumerate(texts[a:a + 1000])],
                         ids=[str(a + i) for i in range(len(texts[a:a + 1000]))])
    torch.cuda.synchronize()

    def step():
        qv = emb.encode_tensor(queries)
        _, rows = db.search_tensor(qv, R)
        rows = rows.cpu().numpy()
        pairs = [(queries[i], texts[int(r)]) for i in range(Q) for r in rows[i]]
        scores = ce.score(pairs).reshape(Q, R)
        return np.argsort(-scores, axis=1, kind="stable")[:, :10]

    step()
    for label in ("pipelined",):
        t0 = time.time()
        it = 3
        for _ in range(it):
            step()
        dt = (time.time() - t0) / it
        print(f"text e2e ({label}): Q={Q} R={R}: {dt * 1e3:.1f} ms/step = {Q / dt:.0f} queries/s (tokenisation included)", flush=True)
    # the tokeniser alone on the same pairs
    rows = db.search_tensor(emb.encode_tensor(queries), R)[1].cpu().numpy()
    pairs = [(queries[i], texts[int(r)]) for i in range(Q) for r in rows[i]]
    t0 = time.time()
    ce._ragged([p[0] for p in pairs], [p[1] for p in pairs])
    print(f"WordPiece alone on {len(pairs)} pairs: {(time.time() - t0) * 1e3:.1f} ms", flush=True)


if __name__ == "__main__":
    main()
