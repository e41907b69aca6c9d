"""CPU: host-side logic of the drop-in classes (no GPU arithmetic): reranker semantics, retriever
plumbing, tokeniser packing, weight resolution, shim installation."""
import sys

import numpy as np
import pytest

from ragmeup_b200.documents import Document, Runnable
from ragmeup_b200.reranker import ScoredCrossEncoderReranker
from ragmeup_b200.tokenizer import (CLS, SEP, build_wordpiece, encode_ragged, synthetic_sentences, synthetic_vocab)
from ragmeup_b200.weights import PRESETS, BertConfig, resolve_model, synthetic_bert_weights, weight_names


class FakeCE:
    def __init__(self, scores):
        self.scores = scores
        self.seen = None

    def score(self, pairs):
        self.seen = list(pairs)
        return np.asarray(self.scores[: len(pairs)], dtype=np.float32)


def test_reranker_semantics_match_reference():
    """server/ScoredCrossEncoderReranker.py:42-45: pairs=(query, page_content), stable sort desc,
    top_n, copies with relevance_score; ties keep input order."""
    docs = [Document(f"d{i}", {"source": f"s{i}"}) for i in range(5)]
    ce = FakeCE([0.5, 2.0, 0.5, -1.0, 2.0])
    out = ScoredCrossEncoderReranker(model=ce, top_n=4).compress_documents(docs, "q?")
    assert ce.seen == [("q?", f"d{i}") for i in range(5)]
    assert [d.page_content for d in out] == ["d1", "d4", "d0", "d2"]
    assert [float(d.metadata["relevance_score"]) for d in out] == [2.0, 2.0, 0.5, 0.5]
    assert out[0].metadata["source"] == "s1"
    assert "relevance_score" not in docs[1].metadata            # originals untouched
    assert ScoredCrossEncoderReranker(model=ce).top_n == 3      # reference default
    with pytest.raises(TypeError):
        ScoredCrossEncoderReranker(model=ce, top_n=2, bogus=1)  # extra="forbid"
    assert len(ScoredCrossEncoderReranker(model=ce, top_n=10).compress_documents(docs[:2], "q")) == 2


def test_runnable_pipe():
    class R(Runnable):
        def invoke(self, x, config=None, **kw):
            return [Document("a"), Document("b")]
    chain = R() | (lambda docs: "\n".join(d.page_content for d in docs))
    assert chain.invoke("q") == "a\nb"


def test_tokenizer_ragged_matches_padded():
    vocab = synthetic_vocab(2000)
    tok = build_wordpiece(vocab)
    texts = synthetic_sentences(vocab, 9, 1, 40, seed=1)
    ids, typ, cu = encode_ragged(tok, texts, None, 32)
    assert cu[0] == 0 and len(cu) == 10 and cu[-1] == len(ids) == len(typ)
    assert np.diff(cu).max() <= 32
    for b in range(9):
        seq = ids[cu[b]:cu[b + 1]]
        assert seq[0] == CLS and seq[-1] == SEP
    # pairs: token types 0 then 1, longest_first truncation keeps both sides
    a = synthetic_sentences(vocab, 3, 5, 10, seed=2)
    b = synthetic_sentences(vocab, 3, 60, 80, seed=3)
    ids, typ, cu = encode_ragged(tok, a, b, 48)
    for i in range(3):
        t = typ[cu[i]:cu[i + 1]]
        s = ids[cu[i]:cu[i + 1]]
        assert len(s) == 48 and t[0] == 0 and t[-1] == 1 and (np.diff(t) >= 0).all()
        assert (s == SEP).sum() == 2
    # same tokens as the padded call the oracle uses
    from oracle import bert_ref
    pi, pm, pt = bert_ref._tok_batch(tok, a, b, 48)
    for i in range(3):
        n = int(pm[i].sum())
        assert (pi[i, :n].numpy() == ids[cu[i]:cu[i + 1]]).all()
        assert (pt[i, :n].numpy() == typ[cu[i]:cu[i + 1]]).all()


def test_fast_tokenisation_and_packing_equal_the_plain_path():
    """encode_batch_fast (no offsets) and the zero-filled type ids of single texts give exactly what encode_batch +
    per-token type ids give: truncation, ragged lengths, pairs, empty strings, unknown words"""
    from tokenizers import Tokenizer
    from ragmeup_b200 import tokenizer as T
    vocab = synthetic_vocab(3000)
    tok = build_wordpiece(vocab)
    texts = synthetic_sentences(vocab, 60, 1, 90, seed=11) + ["", "zzzzqqqq unknownword", "a\nb  c"]
    other = synthetic_sentences(vocab, 63, 1, 90, seed=12)
    for max_len in (16, 64):
        ref_tok = Tokenizer.from_str(tok.to_str())
        ref_tok.enable_truncation(max_length=max_len, strategy="longest_first")
        ref_tok.no_padding()
        rt = T.RaggedTokenizer(tok, max_len)
        for a, b in ((texts, None), (texts, other)):
            enc = ref_tok.encode_batch(list(a) if b is None else list(zip(a, b)))
            ids, typ, cu = rt(a, b)
            assert cu[-1] == len(ids) == len(typ) and len(cu) == len(a) + 1
            for i, e in enumerate(enc):
                assert ids[cu[i]:cu[i + 1]].tolist() == e.ids and typ[cu[i]:cu[i + 1]].tolist() == e.type_ids
    if hasattr(tok, "encode_batch_fast"):                    # the wheel in this image has it: make sure it is what runs
        assert T._encode_batch.__doc__ and T._encode_batch(tok, ["a b"])[0].ids == tok.encode_batch(["a b"])[0].ids


def test_weight_resolution():
    cfg, w, pooling, normalize, max_len, act, src = resolve_model("synthetic:all-MiniLM-L6-v2:3", with_head=False)
    assert (cfg.hidden, cfg.layers, cfg.heads, cfg.ffn) == (384, 6, 12, 1536) and pooling == "mean" and max_len == 256
    assert set(w) == {n for n, _ in weight_names(cfg, False)}
    cfg2, w2, *_ = resolve_model("synthetic:ms-marco-MiniLM-L-6-v2", with_head=True)
    assert "classifier.weight" in w2 and w2["classifier.weight"].shape == (1, 384)
    a = synthetic_bert_weights(cfg, seed=3)
    assert all(np.array_equal(a[k], w[k]) for k in a)           # seeded, reproducible
    with pytest.raises(FileNotFoundError):
        resolve_model("sentence-transformers/all-MiniLM-L6-v2", with_head=False)
    assert PRESETS["bge-base-en-v1.5"][0].hidden == 768


def test_hf_snapshot_loader(tmp_path):
    import json
    from safetensors.numpy import save_file
    cfg = BertConfig(vocab_size=400, hidden=128, layers=1, heads=4, ffn=256, max_pos=64)
    w = synthetic_bert_weights(cfg, seed=1, with_head=True)
    d = tmp_path / "m"
    (d / "1_Pooling").mkdir(parents=True)
    json.dump({"vocab_size": 400, "hidden_size": 128, "num_hidden_layers": 1, "num_attention_heads": 4,
               "intermediate_size": 256, "max_position_embeddings": 64, "hidden_act": "gelu",
               "id2label": {"0": "LABEL_0"}, "sbert_ce_default_activation_function": "torch.nn.modules.linear.Identity"},
              open(d / "config.json", "w"))
    save_file({"bert." + k: v for k, v in w.items()}, str(d / "model.safetensors"))
    json.dump([{"type": "sentence_transformers.models.Transformer", "path": ""},
               {"type": "sentence_transformers.models.Pooling", "path": "1_Pooling"},
               {"type": "sentence_transformers.models.Normalize", "path": "2_Normalize"}], open(d / "modules.json", "w"))
    json.dump({"pooling_mode_cls_token": True, "pooling_mode_mean_tokens": False}, open(d / "1_Pooling" / "config.json", "w"))
    json.dump({"max_seq_length": 48}, open(d / "sentence_bert_config.json", "w"))
    c, ww, pooling, normalize, max_len, act, src = resolve_model(str(d), with_head=True)
    assert (c.hidden, c.num_labels, pooling, normalize, max_len, act) == (128, 1, "cls", True, 48, "identity")
    assert np.array_equal(ww["pooler.dense.weight"], w["pooler.dense.weight"])


def test_install_shims():
    from ragmeup_b200 import install
    c = install.install()
    from langchain_huggingface.embeddings import HuggingFaceEmbeddings as E
    from langchain_milvus.vectorstores import Milvus as M
    from langchain_postgres.vectorstores import PGVector as P
    from langchain_community.cross_encoders import HuggingFaceCrossEncoder as X
    from ScoredCrossEncoderReranker import ScoredCrossEncoderReranker as S
    assert E is c["HuggingFaceEmbeddings"] and M is c["Milvus"] and P is c["PGVector"]
    assert X is c["HuggingFaceCrossEncoder"] and S is c["ScoredCrossEncoderReranker"]
    assert "langchain_milvus.vectorstores" in sys.modules


def test_pipelined_overlaps_and_keeps_order():
    import threading
    import time
    from ragmeup_b200.tokenizer import pipelined
    spans = {}

    def tok(c):
        t0 = time.perf_counter()
        time.sleep(0.05)
        spans[("tok", c)] = (t0, time.perf_counter(), threading.current_thread() is threading.main_thread())
        return c * 2

    def run(t):
        t0 = time.perf_counter()
        time.sleep(0.05)
        spans[("run", t)] = (t0, time.perf_counter())
        return t + 1

    assert pipelined([1, 2, 3, 4], tok, run) == [3, 5, 7, 9]
    # tokenisation of chunk i+1 starts before the device call of chunk i ends, and runs off the calling thread
    for c in (1, 2, 3):
        assert spans[("tok", c + 1)][0] < spans[("run", 2 * c)][1]
    assert not any(v[2] for k, v in spans.items() if k[0] == "tok")
    assert [k[1] for k in sorted((k for k in spans if k[0] == "run"), key=lambda k: spans[k][0])] == [2, 4, 6, 8]
    assert pipelined([], tok, run) == [] and pipelined([5], tok, run) == [11]

    def bad(c):
        raise ValueError("boom")
    with pytest.raises(ValueError):
        pipelined([1, 2], bad, run)


def test_pair_assembler_equals_pair_tokenisation():
    """[CLS] a [SEP] b [SEP] built from cached per-text ids == HF pair tokenisation with truncation='longest_first',
    across truncation regimes (both sides cut, one side cut, none), empty texts and cache eviction."""
    from ragmeup_b200.tokenizer import PairAssembler, RaggedTokenizer, load_tokenizer, synthetic_sentences, synthetic_vocab
    tok = load_tokenizer(None, 30522)
    vocab = synthetic_vocab(30522)
    rng = np.random.default_rng(0)
    a = synthetic_sentences(vocab, 200, 0, 40, seed=1) + ["", "x", "  ", "Ünïcode çase Test", "a\tb\nc"]
    b = synthetic_sentences(vocab, 200, 0, 60, seed=2) + ["y", "", "z z z", "MiXed CASE words", " nbsp"]
    for max_length in (12, 20, 64, 512):
        rt, pa = RaggedTokenizer(tok, max_length), PairAssembler(tok, max_length, max_entries=64)
        for _ in range(3):                                   # repeated calls: cache hits and evictions
            ia = [a[int(i)].strip() for i in rng.integers(0, len(a), 300)]
            ib = [b[int(i)].strip() for i in rng.integers(0, len(b), 300)]
            want, got = rt(ia, ib), pa(ia, ib)
            assert all(np.array_equal(x, y) for x, y in zip(want, got)), max_length
        assert len(pa._cache) <= 64
    # every (len_a, len_b) combination around the budget
    words = [w for w in vocab if w.isalpha() and len(w) > 2][:40]
    rt, pa = RaggedTokenizer(tok, 16), PairAssembler(tok, 16)
    ia = [" ".join(words[:n]) for n in range(0, 20) for _ in range(20)]
    ib = [" ".join(words[20:20 + m]) for _ in range(20) for m in range(0, 20)]
    assert all(np.array_equal(x, y) for x, y in zip(rt(ia, ib), pa(ia, ib)))
    e = pa([], [])
    assert e[0].size == 0 and e[2].tolist() == [0]
