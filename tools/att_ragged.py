"""Hang / parity probe for the attention kernels on ragged batches: python tools/att_ragged.py "<lens expr>" [mode]
compares RMU_ATTN_MODE=0 (tcgen05 where the shape allows) with the oracle on one cross-encoder forward."""
import os, sys
from dataclasses import asdict
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import bert_ref
from ragmeup_b200.encoder import BertEncoder
from ragmeup_b200.weights import PRESETS, BertConfig, synthetic_bert_weights
preset = os.environ.get("PRESET", "ms-marco-MiniLM-L-6-v2")
cfg = BertConfig(**asdict(PRESETS[preset][0]))
w = synthetic_bert_weights(cfg, seed=1, with_head=True, scale=4.0)
enc = BertEncoder(cfg, w, with_head=True)
rng = np.random.default_rng(0)
lens = eval(sys.argv[1])
ids = np.concatenate([rng.integers(104, cfg.vocab_size, n) for n in lens]).astype(np.int32)
typ = np.zeros_like(ids)
cu = np.r_[0, np.cumsum(lens)].astype(np.int32)
print("lens", lens[:12], "... n =", len(lens), flush=True)
lg = enc.classify_tokens(ids, typ, cu, int(max(lens))).cpu().numpy()
torch.cuda.synchronize()
print("ran", flush=True)
if len(lens) <= 64:
    S = max(lens)
    I = np.zeros((len(lens), S), np.int64); M = np.zeros((len(lens), S), np.int64)
    for b, n in enumerate(lens):
        I[b, :n] = ids[cu[b]:cu[b + 1]]; M[b, :n] = 1
    with torch.no_grad():
        h = bert_ref.bert_encoder_forward({k: torch.from_numpy(v) for k, v in w.items()}, bert_ref.BertCfg(**asdict(cfg)), torch.from_numpy(I), torch.from_numpy(M))
        ref = bert_ref.classifier_head({k: torch.from_numpy(v) for k, v in w.items()}, h).numpy()
    print("max |logit err|", float(np.abs(lg - ref).max()), flush=True)
