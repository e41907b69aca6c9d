"""CPU: the oracle restatements against the golden vectors minted from the independent
implementations in the image (tests/golden/make_golden.py) and hand-computed known answers."""
from dataclasses import asdict

import numpy as np
import pytest
import torch

from cases import FLAT_CASES, flat_case
from oracle import bert_ref, flat_ref
from ragmeup_b200.weights import PRESETS, BertConfig, synthetic_bert_weights

ENC_CASES = [("tiny", 0, 1.0), ("all-MiniLM-L6-v2", 0, 1.0), ("all-MiniLM-L6-v2", 1, 4.0), ("bge-base-en-v1.5", 0, 1.0)]
CE_CASES = [("tiny", 3, 4.0), ("ms-marco-MiniLM-L-6-v2", 0, 1.0), ("ms-marco-MiniLM-L-6-v2", 2, 6.0)]


def _cfg(preset):
    cfg = BertConfig(**asdict(PRESETS[preset][0]))
    return cfg, bert_ref.BertCfg(**asdict(cfg))


@pytest.mark.parametrize("preset,seed,scale", ENC_CASES)
def test_encoder_oracle_matches_hf(golden_dir, preset, seed, scale):
    g = np.load(f"{golden_dir}/encoder.npz")
    key = f"enc_{preset}_{seed}_{scale}"
    cfg, ocfg = _cfg(preset)
    w = synthetic_bert_weights(cfg, seed=seed, with_head=False, scale=scale)
    ids = torch.from_numpy(g[key + "_ids"].astype(np.int64))
    mask = torch.from_numpy(g[key + "_mask"].astype(np.int64))
    typ = torch.from_numpy(g[key + "_typ"].astype(np.int64))
    with torch.no_grad():
        h = bert_ref.bert_encoder_forward(w, ocfg, ids, mask, typ)
        mean = bert_ref.l2_normalize(bert_ref.pool(h, mask, "mean")).numpy()
        cls = bert_ref.l2_normalize(bert_ref.pool(h, mask, "cls")).numpy()
    lens = mask.sum(1).tolist()
    h_last = np.stack([h[b, n - 1].numpy() for b, n in enumerate(lens)])
    tol = 2e-5 if scale == 1.0 else 2e-4
    assert np.abs(mean - g[key + "_mean"]).max() < tol
    assert np.abs(cls - g[key + "_cls"]).max() < tol
    assert np.abs(h[:, 0].numpy() - g[key + "_h_first"]).max() < 50 * tol
    assert np.abs(h_last - g[key + "_h_last"]).max() < 50 * tol


@pytest.mark.parametrize("preset,seed,scale", CE_CASES)
def test_cross_encoder_oracle_matches_hf(golden_dir, preset, seed, scale):
    g = np.load(f"{golden_dir}/cross_encoder.npz")
    key = f"ce_{preset}_{seed}_{scale}"
    cfg, ocfg = _cfg(preset)
    w = synthetic_bert_weights(cfg, seed=seed, with_head=True, scale=scale)
    ids = torch.from_numpy(g[key + "_ids"].astype(np.int64))
    mask = torch.from_numpy(g[key + "_mask"].astype(np.int64))
    typ = torch.from_numpy(g[key + "_typ"].astype(np.int64))
    with torch.no_grad():
        logits = bert_ref.classifier_head(w, bert_ref.bert_encoder_forward(w, ocfg, ids, mask, typ)).numpy()
    assert np.abs(logits - g[key + "_logits"]).max() < (2e-5 if scale == 1.0 else 5e-4)


@pytest.mark.parametrize("name", sorted(FLAT_CASES))
@pytest.mark.parametrize("metric", ["ip", "cosine", "l2"])
def test_flat_oracle_matches_fp64_golden(golden_dir, name, metric):
    g = np.load(f"{golden_dir}/flat.npz")
    x, q, k = flat_case(name)
    s, i = flat_ref.flat_search(q, x, k, metric, dtype=np.float64)
    gi, gs = g[f"flat_{name}_{metric}_ids"], g[f"flat_{name}_{metric}_scores"]
    # identical sets; identical order wherever the golden score gaps are not fp32 ties
    for r in range(q.shape[0]):
        assert set(i[r].tolist()) == set(gi[r].tolist())
    valid = gi >= 0
    assert np.abs(s[valid] - gs[valid]).max() < 1e-5


def test_flat_tie_break_lower_row_first():
    x = np.zeros((6, 8), dtype=np.float32)
    x[:, 0] = 1.0                       # six identical rows
    q = x[:1].copy()
    for metric in ("ip", "cosine", "l2"):
        _, i = flat_ref.flat_search(q, x, 4, metric)
        assert i[0].tolist() == [0, 1, 2, 3]
    _, i = flat_ref.flat_search(q, x, 8, "ip", id_offset=100)
    assert i[0].tolist() == [100, 101, 102, 103, 104, 105, -1, -1]


def test_shard_merge_equals_unsharded():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((1003, 32)).astype(np.float32)
    x[700] = x[2]
    q = rng.standard_normal((4, 32)).astype(np.float32)
    q[0] = x[2]
    for metric in ("ip", "cosine", "l2"):
        full_s, full_i = flat_ref.flat_search(q, x, 10, metric)
        for R in (2, 4, 8):
            bounds = [len(x) * r // R for r in range(R + 1)]
            parts = [flat_ref.flat_search(q, x[bounds[r]:bounds[r + 1]], 10, metric, id_offset=bounds[r]) for r in range(R)]
            s = np.stack([p[0] for p in parts])
            i = np.stack([p[1] for p in parts])
            ms, mi = flat_ref.shard_merge(s, i, metric)
            assert (mi == full_i).all()
            assert np.allclose(ms, full_s, atol=1e-6)


def test_mmr_known_answer():
    # q along e0.  c0 and c1 are near-duplicates close to q, c2 is less similar but orthogonal to them.
    q = np.array([1.0, 0.0, 0.0])
    E = np.array([[1.0, 0.05, 0.0], [1.0, 0.06, 0.0], [0.6, 0.0, 0.8], [0.0, 1.0, 0.0]])
    # lambda=0.5: first = argmax cos = c0; then c1 scores .5*cos(q,c1) - .5*cos(c1,c0) ~ 0, c2 scores .5*.6 - .5*.6*~1 ~ 0.0005
    assert flat_ref.mmr(q, E, lambda_mult=0.5, k=2)[0] == 0
    sel = flat_ref.mmr(q, E, lambda_mult=0.5, k=3)
    assert sel[0] == 0 and set(sel) == {0, 1, 2} or sel[1] in (1, 2)
    # lambda=1 ignores redundancy -> pure similarity order
    assert flat_ref.mmr(q, E, lambda_mult=1.0, k=4) == [0, 1, 2, 3]
    # lambda=0 only avoids redundancy: after c0 pick the candidate least similar to c0 -> c3 (cos .05), then c2
    assert flat_ref.mmr(q, E, lambda_mult=0.0, k=3) == [0, 3, 2]
    # ties: identical candidates -> lowest index first (strict >)
    assert flat_ref.mmr(q, np.array([[0.5, 0.5, 0], [0.5, 0.5, 0], [0.5, 0.5, 0]]), k=2) == [0, 1]
    assert flat_ref.mmr(q, E, k=0) == []
    assert flat_ref.mmr(q, E[:1], k=5) == [0]


def test_st_encode_batch_composition_invariance():
    """SentenceTransformer.encode sorts by length and pads per 32-batch; the result per text must not
    depend on which batch it lands in (padding is masked) -- the property the ragged GPU path relies on."""
    from ragmeup_b200.tokenizer import build_wordpiece, synthetic_sentences, synthetic_vocab
    cfg, ocfg = _cfg("tiny")
    vocab = synthetic_vocab(cfg.vocab_size)
    tok = build_wordpiece(vocab)
    w = synthetic_bert_weights(cfg, seed=0)
    texts = synthetic_sentences(vocab, 40, 1, 30, seed=3)
    a = bert_ref.st_encode(w, ocfg, tok, texts, "mean", True, 64, batch_size=32)
    b = bert_ref.st_encode(w, ocfg, tok, texts, "mean", True, 64, batch_size=1)
    assert np.abs(a - b).max() < 1e-5
    one = np.asarray(bert_ref.hf_embed_query(w, ocfg, tok, texts[7] + "\n", pooling="mean", max_seq_length=64))
    assert np.abs(one - a[7]).max() < 1e-5
