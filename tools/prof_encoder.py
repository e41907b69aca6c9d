"""Profiling target: cross-encoder forward on a rerank-shaped ragged batch (for ncu)."""
import os, sys
from dataclasses import asdict
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ragmeup_b200.encoder import BertEncoder
from ragmeup_b200.weights import PRESETS, BertConfig, synthetic_bert_weights
B = int(os.environ.get("PROF_B", 400)); S = int(os.environ.get("PROF_S", 147)); IT = int(os.environ.get("PROF_ITERS", 3))
cfg = BertConfig(**asdict(PRESETS["ms-marco-MiniLM-L-6-v2"][0]))
enc = BertEncoder(cfg, synthetic_bert_weights(cfg, seed=1, with_head=True, scale=4.0), with_head=True)
rng = np.random.default_rng(0)
ids = torch.from_numpy(rng.integers(104, cfg.vocab_size, B * S).astype(np.int32)).cuda()
typ = torch.zeros(B * S, dtype=torch.int32, device="cuda")
cu = (torch.arange(B + 1, dtype=torch.int32) * S).cuda()
for _ in range(2): enc.classify_tokens(ids, typ, cu, S)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(IT): enc.classify_tokens(ids, typ, cu, S)
e1.record(); torch.cuda.synchronize()
print(f"classify B={B} S={S}: {e0.elapsed_time(e1)/IT:.3f} ms/iter", flush=True)
if os.environ.get("PROF_CLASSES") == "1":
    from ragmeup_b200 import _lib
    _lib.profile_enable(True); _lib.profile_reset()
    for _ in range(IT): enc.classify_tokens(ids, typ, cu, S)
    torch.cuda.synchronize()
    print({k: (round(v[0] / IT, 4), v[1] // IT) for k, v in _lib.profile_read().items() if v[1]}, flush=True)
    _lib.profile_enable(False)
