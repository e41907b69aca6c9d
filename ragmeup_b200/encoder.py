"""BertEncoder — torch/numpy view of the C-ABI BERT encoder (``rmu_encoder_*``).

Weights go to the library in the canonical order of ``weights.weight_names`` (HuggingFace
``BertModel`` tensor order: embeddings, then per layer q/k/v/attention-output/LayerNorm/
intermediate/output/LayerNorm, then pooler + classifier when a head is present).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import _lib
from .weights import BertConfig, weight_names

POOL = {"mean": 0, "cls": 1}
MAX_TOKENS_PER_CALL = 131072    # bounds the activation workspace (~19 KB/token for MiniLM shapes => ~2.5 GB)


class BertEncoder:
    def __init__(self, cfg: BertConfig, weights: Dict[str, np.ndarray], with_head: bool, device: Optional[int] = None):
        torch = _lib.require_cuda()
        self.torch = torch
        self.cfg = cfg
        self.with_head = bool(with_head)
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        names = weight_names(cfg, with_head)
        arrs = []
        for name, shape in names:
            a = np.ascontiguousarray(weights[name], dtype=np.float32)
            if tuple(a.shape) != tuple(shape):
                raise ValueError(f"{name}: expected shape {shape}, got {a.shape}")
            arrs.append(a)
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        cc = _lib.BertConfigC(cfg.vocab_size, cfg.hidden, cfg.layers, cfg.heads, cfg.ffn, cfg.max_pos, cfg.type_vocab,
                              cfg.num_labels, cfg.ln_eps)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().rmu_encoder_create(C.byref(cc), ptrs, len(arrs), int(self.with_head), C.byref(self._h)),
                       "rmu_encoder_create")

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _lib.lib().rmu_encoder_destroy(h)
            except Exception:
                pass
            self._h = C.c_void_p()

    # ------------------------------------------------------------------ device-tensor entry points
    def _dev(self, a, dtype):
        torch = self.torch
        if isinstance(a, np.ndarray):
            a = torch.from_numpy(np.ascontiguousarray(a))
        return a.to(device=self.device, dtype=dtype, non_blocking=True).contiguous()

    def embed_tokens(self, ids, type_ids, cu_seqlens, max_seqlen: int, pooling: str = "mean", normalize: bool = True):
        """ragged token batch -> CUDA fp32 [B, hidden] sentence embeddings."""
        torch = self.torch
        ids = self._dev(ids, torch.int32)
        typ = None if type_ids is None else self._dev(type_ids, torch.int32)
        cu = self._dev(cu_seqlens, torch.int32)
        B = cu.numel() - 1
        out = torch.empty((B, self.cfg.hidden), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().rmu_encoder_embed(self._h, ids.data_ptr(), None if typ is None else typ.data_ptr(),
                                                    cu.data_ptr(), B, ids.numel(), int(max_seqlen), POOL[pooling],
                                                    int(bool(normalize)), out.data_ptr(), _lib.stream_ptr()),
                       "rmu_encoder_embed")
        return out

    def classify_tokens(self, ids, type_ids, cu_seqlens, max_seqlen: int):
        """ragged (pair) token batch -> CUDA fp32 [B, num_labels] raw logits."""
        torch = self.torch
        ids = self._dev(ids, torch.int32)
        typ = None if type_ids is None else self._dev(type_ids, torch.int32)
        cu = self._dev(cu_seqlens, torch.int32)
        B = cu.numel() - 1
        out = torch.empty((B, self.cfg.num_labels), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().rmu_encoder_classify(self._h, ids.data_ptr(), None if typ is None else typ.data_ptr(),
                                                       cu.data_ptr(), B, ids.numel(), int(max_seqlen), out.data_ptr(),
                                                       _lib.stream_ptr()), "rmu_encoder_classify")
        return out

    def hidden_tokens(self, ids, type_ids, cu_seqlens, max_seqlen: int):
        """last hidden state, CUDA fp32 [total_tokens, hidden]."""
        torch = self.torch
        ids = self._dev(ids, torch.int32)
        typ = None if type_ids is None else self._dev(type_ids, torch.int32)
        cu = self._dev(cu_seqlens, torch.int32)
        B = cu.numel() - 1
        out = torch.empty((ids.numel(), self.cfg.hidden), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().rmu_encoder_hidden(self._h, ids.data_ptr(), None if typ is None else typ.data_ptr(),
                                                     cu.data_ptr(), B, ids.numel(), int(max_seqlen), out.data_ptr(),
                                                     _lib.stream_ptr()), "rmu_encoder_hidden")
        return out

    # ------------------------------------------------------------------ host-buffer entry points
    @staticmethod
    def _chunks(cu: np.ndarray, max_tokens: int):
        """split sequences [0, B) into runs whose token count stays under max_tokens"""
        B = len(cu) - 1
        s = 0
        while s < B:
            e = s + 1
            while e < B and cu[e + 1] - cu[s] <= max_tokens:
                e += 1
            yield s, e
            s = e

    def embed_host(self, ids: np.ndarray, type_ids: Optional[np.ndarray], cu: np.ndarray, pooling: str = "mean",
                   normalize: bool = True) -> np.ndarray:
        """host int32 token batch -> host fp32 [B, hidden] (H2D/D2H inside the C call)."""
        B = len(cu) - 1
        out = np.empty((B, self.cfg.hidden), dtype=np.float32)
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        typ = None if type_ids is None else np.ascontiguousarray(type_ids, dtype=np.int32)
        cu = np.ascontiguousarray(cu, dtype=np.int32)
        with self.torch.cuda.device(self.device):
            for s, e in self._chunks(cu, MAX_TOKENS_PER_CALL):
                t0, t1 = int(cu[s]), int(cu[e])
                sub_cu = np.ascontiguousarray(cu[s:e + 1] - cu[s], dtype=np.int32)
                sub_ids = ids[t0:t1]
                sub_typ = None if typ is None else typ[t0:t1]
                o = out[s:e]
                maxlen = int(np.max(np.diff(sub_cu)))
                _lib.check(_lib.lib().rmu_encoder_embed_host(
                    self._h, sub_ids.ctypes.data, None if sub_typ is None else sub_typ.ctypes.data, sub_cu.ctypes.data,
                    e - s, t1 - t0, maxlen, POOL[pooling], int(bool(normalize)), o.ctypes.data, _lib.stream_ptr()),
                    "rmu_encoder_embed_host")
        return out

    def classify_host(self, ids: np.ndarray, type_ids: Optional[np.ndarray], cu: np.ndarray) -> np.ndarray:
        """host int32 pair-token batch -> host fp32 [B, num_labels] raw logits."""
        B = len(cu) - 1
        out = np.empty((B, self.cfg.num_labels), dtype=np.float32)
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        typ = None if type_ids is None else np.ascontiguousarray(type_ids, dtype=np.int32)
        cu = np.ascontiguousarray(cu, dtype=np.int32)
        with self.torch.cuda.device(self.device):
            for s, e in self._chunks(cu, MAX_TOKENS_PER_CALL):
                t0, t1 = int(cu[s]), int(cu[e])
                sub_cu = np.ascontiguousarray(cu[s:e + 1] - cu[s], dtype=np.int32)
                sub_ids = ids[t0:t1]
                sub_typ = None if typ is None else typ[t0:t1]
                o = out[s:e]
                maxlen = int(np.max(np.diff(sub_cu)))
                _lib.check(_lib.lib().rmu_encoder_classify_host(
                    self._h, sub_ids.ctypes.data, None if sub_typ is None else sub_typ.ctypes.data, sub_cu.ctypes.data,
                    e - s, t1 - t0, maxlen, o.ctypes.data, _lib.stream_ptr()), "rmu_encoder_classify_host")
        return out
