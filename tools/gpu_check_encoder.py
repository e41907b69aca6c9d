"""GPU diagnostic for the BERT encoder: hidden states / embeddings / logits vs the CPU oracle."""
import os, sys, time
from dataclasses import asdict
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ragmeup_b200.encoder import BertEncoder
from ragmeup_b200.weights import PRESETS, BertConfig, synthetic_bert_weights
from oracle import bert_ref

def log(*a): print(*a, flush=True)
rng = np.random.default_rng(0)

def batch(cfg, lens, pair=False):
    ids = []; typ = []; cu = [0]
    for n in lens:
        t = rng.integers(104, cfg.vocab_size, n); t[0] = 101; t[-1] = 102
        ty = np.zeros(n, dtype=np.int64)
        if pair and n > 4:
            cut = max(2, n // 3); t[cut - 1] = 102; ty[cut:] = 1
        ids.append(t); typ.append(ty); cu.append(cu[-1] + n)
    return np.concatenate(ids).astype(np.int32), np.concatenate(typ).astype(np.int32), np.array(cu, dtype=np.int32)

def padded(ids, typ, cu):
    B = len(cu) - 1; S = int(np.max(np.diff(cu)))
    I = np.zeros((B, S), dtype=np.int64); M = np.zeros((B, S), dtype=np.int64); T = np.zeros((B, S), dtype=np.int64)
    for b in range(B):
        n = cu[b + 1] - cu[b]; I[b, :n] = ids[cu[b]:cu[b + 1]]; T[b, :n] = typ[cu[b]:cu[b + 1]]; M[b, :n] = 1
    return torch.from_numpy(I), torch.from_numpy(M), torch.from_numpy(T)

for preset, seed, scale, lens, head in [("tiny", 0, 1.0, [2, 5, 17, 64, 33], False),
                                        ("all-MiniLM-L6-v2", 0, 1.0, [2, 9, 31, 64, 50, 130, 200], False),
                                        ("ms-marco-MiniLM-L-6-v2", 2, 6.0, [20, 48, 96, 33, 147, 147], True),
                                        ("bge-base-en-v1.5", 0, 1.0, [3, 24, 12, 140], False)]:
    cfg = BertConfig(**asdict(PRESETS[preset][0])); ocfg = bert_ref.BertCfg(**asdict(cfg))
    w = synthetic_bert_weights(cfg, seed=seed, with_head=head, scale=scale)
    t0 = time.time()
    enc = BertEncoder(cfg, w, with_head=head)
    ids, typ, cu = batch(cfg, lens, pair=head)
    h = enc.hidden_tokens(ids, typ, cu, int(np.max(np.diff(cu)))).cpu()
    torch.cuda.synchronize()
    I, M, T = padded(ids, typ, cu)
    with torch.no_grad():
        ho = bert_ref.bert_encoder_forward(w, ocfg, I, M, T)
    ho_r = torch.cat([ho[b, :cu[b + 1] - cu[b]] for b in range(len(lens))])
    log(f"{preset} scale={scale}: hidden max|d|={float((h - ho_r).abs().max()):.3e} (|h|max {float(ho_r.abs().max()):.2f})")
    for pooling in ("mean", "cls"):
        e = enc.embed_tokens(ids, typ, cu, int(np.max(np.diff(cu))), pooling, True).cpu()
        eo = bert_ref.l2_normalize(bert_ref.pool(ho, M, pooling))
        log(f"   emb[{pooling}] max|d|={float((e - eo).abs().max()):.3e}")
    eh = enc.embed_host(ids, typ, cu, "mean", True)
    log(f"   embed_host vs device: {float(np.abs(eh - enc.embed_tokens(ids, typ, cu, int(np.max(np.diff(cu))), 'mean', True).cpu().numpy()).max()):.3e}")
    if head:
        lg = enc.classify_tokens(ids, typ, cu, int(np.max(np.diff(cu)))).cpu()
        lo = bert_ref.classifier_head(w, ho)
        log(f"   logits max|d|={float((lg - lo).abs().max()):.3e}  range [{float(lo.min()):.3f},{float(lo.max()):.3f}]")
        lh = enc.classify_host(ids, typ, cu)
        log(f"   classify_host vs device: {float(np.abs(lh - lg.numpy()).max()):.3e}")
    del enc

# throughput: rerank-shaped batch (100 pairs x 147 tokens) and query-shaped batch (64 x 16)
cfg = BertConfig(**asdict(PRESETS["ms-marco-MiniLM-L-6-v2"][0]))
w = synthetic_bert_weights(cfg, seed=0, with_head=True)
enc = BertEncoder(cfg, w, with_head=True)
for B, S in ((100, 147), (200, 147), (64, 16), (256, 16)):
    ids, typ, cu = batch(cfg, [S] * B, pair=True)
    dids = torch.from_numpy(ids).cuda(); dtyp = torch.from_numpy(typ).cuda(); dcu = torch.from_numpy(cu).cuda()
    for _ in range(3): enc.classify_tokens(dids, dtyp, dcu, S)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): enc.classify_tokens(dids, dtyp, dcu, S)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    flops = B * S * (21.23e6 + 9216 * S)
    log(f"classify B={B} S={S}: {ms:.3f} ms  -> {flops / ms / 1e9:.1f} TFLOP/s algorithmic, {B * S / ms * 1e3:.0f} tok/s")
