"""GPU: SemanticChunker drop-in (device embeddings + adjacent cosine-distance kernel) against the oracle's restatement of
langchain_experimental's algorithm (server/RAGHelper.py:329-341,368)."""
from dataclasses import asdict

import numpy as np
import pytest

from oracle import bert_ref, chunker_ref
from ragmeup_b200.documents import Document
from ragmeup_b200.tokenizer import synthetic_sentences, synthetic_vocab

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def emb(cuda):
    from ragmeup_b200.embeddings import HuggingFaceEmbeddings
    return HuggingFaceEmbeddings(model_name="synthetic:all-MiniLM-L6-v2:0", model_kwargs={"device": "cuda"})


def _text(n, seed):
    vocab = synthetic_vocab(30522)
    words = [w for w in vocab if w.isalpha()]
    # topics: runs of sentences drawn from the same small word pool, so that consecutive groups differ in distance
    sents = []
    for t, k in enumerate(np.random.default_rng(seed).integers(2, 7, size=n)):
        pool = {w: vocab[w] for w in words[400 * (t % 9):400 * (t % 9) + 400]}
        sents += [s + "." for s in synthetic_sentences(pool, int(k), 6, 14, seed=seed * 100 + t)]
    return " ".join(sents)


def test_adjacent_distance_kernel_matches_numpy(emb):
    from ragmeup_b200.chunker import adjacent_cosine_distance
    import torch
    g = torch.Generator(device="cuda").manual_seed(1)
    for n, d in ((2, 384), (37, 384), (300, 768), (5, 33)):
        x = torch.randn(n, d, device="cuda", generator=g) * torch.rand(n, 1, device="cuda", generator=g) * 3
        got = adjacent_cosine_distance(x)
        ref = np.asarray(chunker_ref.distances_from_embeddings(x.cpu().numpy().tolist()))
        assert got.dtype == np.float64 and got.shape == (n - 1,)
        assert np.abs(got - ref).max() < 1e-12
    assert adjacent_cosine_distance(torch.zeros(1, 8, device="cuda")).shape == (0,)
    with pytest.raises(TypeError):
        adjacent_cosine_distance(np.zeros((3, 4), np.float32))


@pytest.mark.parametrize("kind,amount,nchunks", [("percentile", None, None), ("standard_deviation", 1.0, None),
                                                 ("interquartile", None, None), ("gradient", 80, None),
                                                 ("percentile", None, 6)])
def test_split_text_matches_oracle(emb, kind, amount, nchunks):
    from ragmeup_b200.chunker import SemanticChunker
    text = _text(12, 5)
    sc = SemanticChunker(emb, breakpoint_threshold_type=kind, breakpoint_threshold_amount=amount, number_of_chunks=nchunks)
    got = sc.split_text(text)
    ref, ref_dist = chunker_ref.split_text(text, emb.embed_documents, kind=kind, amount=amount, number_of_chunks=nchunks)
    assert got == ref and " ".join(got) == text and len(got) >= 2
    # the distances themselves against the full CPU oracle (its own BERT forward on the same synthetic weights)
    from ragmeup_b200.weights import resolve_model
    cfg, w, *_ = resolve_model(emb.model_name, with_head=False)
    import re
    groups = chunker_ref.sentence_groups(re.split(r"(?<=[.?!])\s+", text))
    full = bert_ref.hf_embed_documents(w, bert_ref.BertCfg(**asdict(cfg)), emb.tokenizer, groups, pooling="mean",
                                       normalize=True, max_seq_length=emb.max_seq_length)
    full_dist = np.asarray(chunker_ref.distances_from_embeddings(np.asarray(full, np.float32).tolist()))
    assert np.abs(np.asarray(ref_dist) - full_dist).max() < 1e-3


def test_split_documents_through_the_reference_import_path(emb):
    from ragmeup_b200 import install
    install.install()
    from langchain_experimental.text_splitter import SemanticChunker  # the line server/RAGHelper.py:27 runs
    sc = SemanticChunker(emb, breakpoint_threshold_type="percentile", breakpoint_threshold_amount=None, number_of_chunks=None)
    docs = [Document(_text(6, 2), {"source": "a.txt"}), Document("One sentence only", {"source": "b.txt"})]
    out = sc.split_documents(docs)
    assert [d.metadata["source"] for d in out].count("b.txt") == 1 and out[-1].page_content == "One sentence only"
    assert " ".join(d.page_content for d in out if d.metadata["source"] == "a.txt") == docs[0].page_content
