"""Mint the golden vectors under tests/golden/ (run once in the BUILD container; commit the output).

The reference repo holds no tests or fixtures and its third-party stack is not installable here
(SURVEY.md §4, §8c), so the goldens come from the independent implementations that ARE in this
image:

* transformer math: ``transformers`` ``BertModel`` / ``BertForSequenceClassification``
  (attn_implementation="eager", fp32) fed the same seeded synthetic weights the tests rebuild
  from ``ragmeup_b200.weights.synthetic_bert_weights`` — only token ids and outputs are stored;
* flat search: torch float64 matmul + stable sort (independent of ``oracle/flat_ref.py``);
* hand-computed MMR / reranker cases live in the tests themselves.

Usage:  python tests/golden/make_golden.py
"""
import os
import sys
from dataclasses import asdict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ragmeup_b200.weights import PRESETS, BertConfig, synthetic_bert_weights  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def hf_models(cfg: BertConfig, w, with_head: bool):
    from transformers import BertConfig as HC, BertForSequenceClassification, BertModel
    hc = HC(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden, num_hidden_layers=cfg.layers,
            num_attention_heads=cfg.heads, intermediate_size=cfg.ffn, max_position_embeddings=cfg.max_pos,
            type_vocab_size=cfg.type_vocab, layer_norm_eps=cfg.ln_eps, num_labels=cfg.num_labels,
            hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, attn_implementation="eager")
    sd = {k: torch.from_numpy(v) for k, v in w.items()}
    if with_head:
        m = BertForSequenceClassification(hc).eval()
        sd = {("bert." + k if not k.startswith("classifier") else k): v for k, v in sd.items()}
        missing = m.load_state_dict(sd, strict=False)
    else:
        m = BertModel(hc, add_pooling_layer=False).eval()
        missing = m.load_state_dict(sd, strict=False)
    bad = [k for k in missing.missing_keys if "position_ids" not in k]
    assert not bad and not missing.unexpected_keys, (bad, missing.unexpected_keys)
    return m


def ragged_batch(rng, cfg, lens, pair_split=None):
    B, S = len(lens), max(lens)
    ids = np.zeros((B, S), dtype=np.int64)
    mask = np.zeros((B, S), dtype=np.int64)
    typ = np.zeros((B, S), dtype=np.int64)
    for b, n in enumerate(lens):
        ids[b, :n] = rng.integers(104, cfg.vocab_size, n)
        ids[b, 0] = 101
        ids[b, n - 1] = 102
        mask[b, :n] = 1
        if pair_split is not None:
            cut = max(2, int(n * pair_split[b]))
            ids[b, cut - 1] = 102
            typ[b, cut:n] = 1
    return ids, mask, typ


def encoder_goldens():
    rng = np.random.default_rng(1234)
    out = {}
    cases = [("tiny", 0, 1.0, [2, 5, 17, 64, 33]),
             ("all-MiniLM-L6-v2", 0, 1.0, [2, 9, 31, 64, 50]),
             ("all-MiniLM-L6-v2", 1, 4.0, [7, 40, 23]),
             ("bge-base-en-v1.5", 0, 1.0, [3, 24, 12])]
    for preset, seed, scale, lens in cases:
        cfg = BertConfig(**asdict(PRESETS[preset][0]))
        w = synthetic_bert_weights(cfg, seed=seed, with_head=False, scale=scale)
        m = hf_models(cfg, w, False)
        ids, mask, typ = ragged_batch(rng, cfg, lens)
        with torch.no_grad():
            h = m(input_ids=torch.from_numpy(ids), attention_mask=torch.from_numpy(mask),
                  token_type_ids=torch.from_numpy(typ)).last_hidden_state
        mk = torch.from_numpy(mask)[..., None].float()
        mean = (h * mk).sum(1) / mk.sum(1).clamp(min=1e-9)
        mean = torch.nn.functional.normalize(mean, p=2, dim=1)
        cls = torch.nn.functional.normalize(h[:, 0], p=2, dim=1)
        key = f"enc_{preset}_{seed}_{scale}"
        out[key + "_ids"] = ids.astype(np.int32)
        out[key + "_mask"] = mask.astype(np.int8)
        out[key + "_typ"] = typ.astype(np.int8)
        out[key + "_mean"] = mean.numpy()
        out[key + "_cls"] = cls.numpy()
        out[key + "_h_first"] = h[:, 0].numpy()          # hidden state of [CLS]
        out[key + "_h_last"] = np.stack([h[b, n - 1].numpy() for b, n in enumerate(lens)])
    return out


def cross_encoder_goldens():
    rng = np.random.default_rng(4321)
    out = {}
    for preset, seed, scale, lens in [("tiny", 3, 4.0, [9, 30, 64, 12]),
                                      ("ms-marco-MiniLM-L-6-v2", 0, 1.0, [12, 64, 35]),
                                      ("ms-marco-MiniLM-L-6-v2", 2, 6.0, [20, 48, 96, 33])]:
        cfg = BertConfig(**asdict(PRESETS[preset][0]))
        w = synthetic_bert_weights(cfg, seed=seed, with_head=True, scale=scale)
        m = hf_models(cfg, w, True)
        ids, mask, typ = ragged_batch(rng, cfg, lens, pair_split=rng.uniform(0.2, 0.6, len(lens)))
        with torch.no_grad():
            logits = m(input_ids=torch.from_numpy(ids), attention_mask=torch.from_numpy(mask),
                       token_type_ids=torch.from_numpy(typ)).logits
        key = f"ce_{preset}_{seed}_{scale}"
        out[key + "_ids"] = ids.astype(np.int32)
        out[key + "_mask"] = mask.astype(np.int8)
        out[key + "_typ"] = typ.astype(np.int8)
        out[key + "_logits"] = logits.numpy()
    return out


def flat_goldens():
    """exact top-k by torch float64 with planted duplicates / ties; tie rule: lower row first."""
    from cases import FLAT_CASES, flat_case
    out = {}
    for name in FLAT_CASES:
        x, q, k = flat_case(name)
        n, nq = x.shape[0], q.shape[0]
        xd, qd = torch.from_numpy(x).double(), torch.from_numpy(q).double()
        for metric in ("ip", "cosine", "l2"):
            if metric == "ip":
                v = qd @ xd.T
                rank = -v
            elif metric == "cosine":
                v = (qd @ xd.T) / (qd.norm(dim=1, keepdim=True) * xd.norm(dim=1)[None])
                rank = -v
            else:
                v = ((qd[:, None, :] - xd[None, :, :]) ** 2).sum(-1)
                rank = v
            # fp32-rounded values define the ties the fp32 implementations see
            rank32 = rank.float()
            order = torch.stack([torch.tensor(np.lexsort((np.arange(n), rank32[i].numpy()))) for i in range(nq)])
            kk = min(k, n)
            ids = torch.full((nq, k), -1, dtype=torch.long)
            ids[:, :kk] = order[:, :kk]
            sc = torch.full((nq, k), float("inf") if metric == "l2" else float("-inf"), dtype=torch.float64)
            sc[:, :kk] = torch.gather(v, 1, order[:, :kk])
            out[f"flat_{name}_{metric}_ids"] = ids.numpy()
            out[f"flat_{name}_{metric}_scores"] = sc.numpy().astype(np.float32)
    return out


def bm25_goldens():
    """BM25Okapi top-n (scores float64, document numbers) from the restatement in oracle/hybrid_ref.py — rank_bm25
    itself is not installed in this image, so these pin the restatement against regressions only; the hand-worked
    known-answer test lives in tests/test_hybrid_cpu.py."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from cases import BM25_CASES, bm25_inputs
    from oracle import hybrid_ref as H
    out = {}
    for case in BM25_CASES:
        if case["n"] > 10000:
            continue
        corpus, queries = bm25_inputs(case)
        bm = H.BM25Okapi(corpus)
        k = case["k"]
        ids = np.full((len(queries), k), -1, dtype=np.int64)
        sc = np.zeros((len(queries), k), dtype=np.float64)
        for qi, q in enumerate(queries):
            s, d = H.bm25_topk(bm, q, k)
            ids[qi, :len(d)] = d
            sc[qi, :len(s)] = s
        out[f"{case['name']}_ids"] = ids
        out[f"{case['name']}_scores"] = sc
    return out


if __name__ == "__main__":
    if "--bm25" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "bm25.npz"), **bm25_goldens())
        print("bm25.npz", os.path.getsize(os.path.join(HERE, "bm25.npz")) // 1024, "KiB")
        sys.exit(0)
    torch.manual_seed(0)
    torch.set_num_threads(8)
    np.savez_compressed(os.path.join(HERE, "encoder.npz"), **encoder_goldens())
    np.savez_compressed(os.path.join(HERE, "cross_encoder.npz"), **cross_encoder_goldens())
    np.savez_compressed(os.path.join(HERE, "flat.npz"), **flat_goldens())
    np.savez_compressed(os.path.join(HERE, "bm25.npz"), **bm25_goldens())
    for f in ("encoder.npz", "cross_encoder.npz", "flat.npz", "bm25.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
