"""Host-side WordPiece tokenisation.

The reference tokenises with HuggingFace ``tokenizers`` (Rust) through
``BertTokenizerFast`` inside sentence-transformers (called from
``server/RAGHelper_local.py:114-117`` / ``server/RAGHelper.py:484``); it is not
GPU arithmetic and stays on the host here too, using the same library.  This
module only builds the ``tokenizers.Tokenizer`` (from a snapshot's
``tokenizer.json`` / ``vocab.txt`` or a seeded synthetic vocab) and packs
batches into the ragged ``cu_seqlens`` layout the CUDA encoder consumes
(no padding tokens are ever sent to the device).
"""
from __future__ import annotations

import os
import threading
from concurrent.futures import ThreadPoolExecutor
from itertools import chain
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
from tokenizers import Tokenizer, models, normalizers, pre_tokenizers, processors, decoders

PAD, UNK, CLS, SEP, MASK = 0, 100, 101, 102, 103

_SYL = ["ba", "ce", "di", "fo", "gu", "ha", "je", "ki", "lo", "mu", "na", "pe", "qi", "ro", "su",
        "ta", "ve", "wi", "xo", "yu", "za", "bri", "cle", "dro", "fla", "gri", "ple", "sto", "tru", "vla"]


def synthetic_vocab(size: int = 30522, seed: int = 7) -> Dict[str, int]:
    """bert-base-uncased-like id layout ([PAD]=0, [UNK]=100, [CLS]=101, [SEP]=102,
    [MASK]=103) filled with seeded pseudo-words and ##suffixes."""
    assert size >= 400
    vocab: Dict[str, int] = {}
    specials = {PAD: "[PAD]", UNK: "[UNK]", CLS: "[CLS]", SEP: "[SEP]", MASK: "[MASK]"}
    i = 0
    u = 0
    while i < 104:
        if i in specials:
            vocab[specials[i]] = i
        else:
            vocab[f"[unused{u}]"] = i
            u += 1
        i += 1
    for ch in "abcdefghijklmnopqrstuvwxyz0123456789.,;:!?'\"()-":
        vocab[ch] = i
        i += 1
    for ch in "abcdefghijklmnopqrstuvwxyz0123456789":
        vocab["##" + ch] = i
        i += 1
    rng = np.random.default_rng(seed)
    while i < size:
        n = int(rng.integers(1, 4))
        word = "".join(_SYL[int(j)] for j in rng.integers(0, len(_SYL), n))
        if rng.random() < 0.25:
            word = "##" + word
        if word not in vocab:
            vocab[word] = i
            i += 1
    return vocab


def build_wordpiece(vocab: Dict[str, int], lowercase: bool = True) -> Tokenizer:
    tok = Tokenizer(models.WordPiece(vocab=vocab, unk_token="[UNK]", max_input_chars_per_word=100))
    tok.normalizer = normalizers.BertNormalizer(clean_text=True, handle_chinese_chars=True,
                                                strip_accents=None, lowercase=lowercase)
    tok.pre_tokenizer = pre_tokenizers.BertPreTokenizer()
    tok.post_processor = processors.TemplateProcessing(
        single="[CLS] $A [SEP]", pair="[CLS] $A [SEP] $B:1 [SEP]:1",
        special_tokens=[("[CLS]", vocab["[CLS]"]), ("[SEP]", vocab["[SEP]"])])
    tok.decoder = decoders.WordPiece(prefix="##")
    return tok


def load_tokenizer(source: Optional[str], vocab_size: int = 30522) -> Tokenizer:
    """``source`` = snapshot dir (tokenizer.json or vocab.txt) or None -> synthetic vocab."""
    if source is None:
        return build_wordpiece(synthetic_vocab(vocab_size))
    tj = os.path.join(source, "tokenizer.json")
    if os.path.exists(tj):
        return Tokenizer.from_file(tj)
    vt = os.path.join(source, "vocab.txt")
    if os.path.exists(vt):
        with open(vt, encoding="utf-8") as f:
            vocab = {line.rstrip("\n"): i for i, line in enumerate(f)}
        return build_wordpiece(vocab)
    raise FileNotFoundError(f"{source}: neither tokenizer.json nor vocab.txt")


def synthetic_sentences(vocab: Dict[str, int], n: int, min_words: int, max_words: int, seed: int) -> List[str]:
    """Seeded pseudo-sentences made of whole vocabulary words (for tests / benches)."""
    words = [w for w in vocab if not w.startswith("[") and not w.startswith("##") and len(w) > 1]
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        k = int(rng.integers(min_words, max_words + 1))
        out.append(" ".join(words[int(j)] for j in rng.integers(0, len(words), k)))
    return out


class RaggedTokenizer:
    """A PRIVATE copy of a tokenizer, configured once (truncation='longest_first' to max_length, no
    padding) and never mutated afterwards, so any number of request threads can tokenise through it
    (Flask serves /chat on several threads over one shared RAGHelper: server/server.py:141-146,394;
    mutating a shared `tokenizers.Tokenizer` while another thread encodes raises "Already borrowed")."""

    def __init__(self, tok: Tokenizer, max_length: int):
        self.max_length = int(max_length)
        self._tok = Tokenizer.from_str(tok.to_str())
        self._tok.enable_truncation(max_length=self.max_length, strategy="longest_first")
        self._tok.no_padding()
        # single texts carry type id 0 throughout with BERT's post-processor; checked once instead of assumed
        probe = self._tok.encode("a b")
        self._single_zero = not any(probe.type_ids)

    def __call__(self, texts_a: Sequence[str], texts_b: Optional[Sequence[str]] = None):
        if texts_b is None:
            return _pack(_encode_batch(self._tok, list(texts_a)), single_segment=self._single_zero)
        return _pack(_encode_batch(self._tok, list(zip(texts_a, texts_b))))


def encode_ragged(tok: Tokenizer, texts_a: Sequence[str], texts_b: Optional[Sequence[str]],
                  max_length: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Tokenise with truncation='longest_first', NO padding, and pack (single-threaded helper: it
    re-configures `tok`; the drop-in classes use RaggedTokenizer instead).

    Returns (ids int32 [T], type_ids int32 [T], cu_seqlens int32 [B+1])."""
    tok.enable_truncation(max_length=max_length, strategy="longest_first")
    tok.no_padding()
    if texts_b is None:
        enc = _encode_batch(tok, list(texts_a))
    else:
        enc = _encode_batch(tok, list(zip(texts_a, texts_b)))
    return _pack(enc)


def _encode_batch(tok: Tokenizer, inputs, **kw):
    """``encode_batch_fast`` (tokenizers >= 0.20) skips the character-offset bookkeeping this path never reads: the same
    ids / type ids / truncation, about a third of the time on 60-word chunks — and WordPiece is what bounds bulk ingest
    and the text-level rerank.  Older wheels fall back to ``encode_batch``."""
    fast = getattr(tok, "encode_batch_fast", None)
    return fast(inputs, **kw) if fast is not None else tok.encode_batch(inputs, **kw)


def _pack(enc, single_segment: bool = False) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    id_lists = [e.ids for e in enc]                      # each .ids builds a Python list: fetch it once
    lens = np.fromiter(map(len, id_lists), dtype=np.int64, count=len(id_lists))
    cu = np.zeros(len(enc) + 1, dtype=np.int32)
    np.cumsum(lens, out=cu[1:])
    total = int(cu[-1])
    ids = np.fromiter(chain.from_iterable(id_lists), dtype=np.int32, count=total)
    if single_segment:                                   # [CLS] a [SEP]: BERT's template gives every token type 0
        typ = np.zeros(total, dtype=np.int32)
    else:
        typ = np.fromiter(chain.from_iterable(e.type_ids for e in enc), dtype=np.int32, count=total)
    return ids, typ, cu


def pipelined(chunks: Sequence, tokenize, run) -> List:
    """``[run(tokenize(c)) for c in chunks]`` with the tokenisation of chunk i+1 (host, HF `tokenizers` releases the
    GIL) overlapped with the device call of chunk i (ctypes releases the GIL while it waits on the stream).  On a B200
    the encoder consumes tokens faster than WordPiece produces them, so for bulk calls the host side is what is left
    to hide."""
    chunks = list(chunks)
    if len(chunks) <= 1:
        return [run(tokenize(c)) for c in chunks]
    out = []
    with ThreadPoolExecutor(max_workers=1) as ex:
        fut = ex.submit(tokenize, chunks[0])
        for i in range(len(chunks)):
            tok = fut.result()
            if i + 1 < len(chunks):
                fut = ex.submit(tokenize, chunks[i + 1])
            out.append(run(tok))
    return out


class PairAssembler:
    """(text_a, text_b) pairs -> the ragged ``[CLS] a [SEP] b [SEP]`` batch of ``RaggedTokenizer(a, b)``, built from
    PER-TEXT WordPiece ids that are cached.  A reranker scores the same documents against many queries (and the same
    query against many documents): with the cache only unseen texts reach the tokeniser, and on a B200 the tokeniser,
    not the encoder, bounds a text-level rerank (6400 pairs: 214 ms of WordPiece vs 90 ms of GPU).

    Equality with pair tokenisation rests on two facts checked in tests/test_host_logic.py: BERT normalisation and
    pre-tokenisation act on each segment separately, and `truncation="longest_first"` in HF `tokenizers` is the closed
    form below (shorter side n1, longer n2: n2 = max(n1, B - n1); if still too long n1 = B // 2, n2 = n1 + B % 2), with
    B = max_length minus the three special tokens."""

    def __init__(self, tok: Tokenizer, max_length: int, max_entries: int = 1 << 18):
        self.max_length = int(max_length)
        self._tok = Tokenizer.from_str(tok.to_str())
        self._tok.no_truncation()
        self._tok.no_padding()
        self._cls = tok.token_to_id("[CLS]")
        self._sep = tok.token_to_id("[SEP]")
        if self._cls is None or self._sep is None:
            raise ValueError("PairAssembler needs a BERT vocabulary with [CLS] and [SEP]")
        self._cache: Dict[str, np.ndarray] = {}
        self._max_entries = int(max_entries)
        self._lock = threading.Lock()

    def _ids(self, texts: Sequence[str]) -> List[np.ndarray]:
        got: Dict[str, np.ndarray] = {}
        with self._lock:
            for t in dict.fromkeys(texts):
                v = self._cache.get(t)
                if v is not None:
                    got[t] = v
        missing = [t for t in dict.fromkeys(texts) if t not in got]
        if missing:
            enc = _encode_batch(self._tok, missing, add_special_tokens=False)    # outside the lock: releases the GIL
            fresh = {t: np.asarray(e.ids, dtype=np.int32) for t, e in zip(missing, enc)}
            got.update(fresh)
            with self._lock:
                self._cache.update(fresh)
                while len(self._cache) > self._max_entries:                      # oldest first
                    self._cache.pop(next(iter(self._cache)))
        return [got[t] for t in texts]

    def __call__(self, texts_a: Sequence[str], texts_b: Sequence[str]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        ia, ib = self._ids(list(texts_a)), self._ids(list(texts_b))
        la = np.fromiter(map(len, ia), dtype=np.int64, count=len(ia))
        lb = np.fromiter(map(len, ib), dtype=np.int64, count=len(ib))
        budget = self.max_length - 3
        swap = la > lb
        n1, n2 = np.where(swap, lb, la), np.where(swap, la, lb)
        over = la + lb > budget
        n2t = np.maximum(n1, budget - n1)
        both = n1 + n2t > budget
        n1t = np.where(both, budget // 2, n1)
        n2t = np.where(both, budget // 2 + budget % 2, n2t)
        na = np.where(over, np.where(swap, n2t, n1t), la)
        nb = np.where(over, np.where(swap, n1t, n2t), lb)
        cu = np.zeros(len(ia) + 1, dtype=np.int32)
        np.cumsum(na + nb + 3, out=cu[1:])
        cls = np.asarray([self._cls], dtype=np.int32)
        sep = np.asarray([self._sep], dtype=np.int32)
        pieces = []
        for a, b, x, y in zip(ia, ib, na, nb):
            pieces += (cls, a[:x], sep, b[:y], sep)
        ids = np.concatenate(pieces) if pieces else np.zeros(0, dtype=np.int32)
        typ = np.zeros(int(cu[-1]), dtype=np.int32)
        starts = cu[:-1] + na.astype(np.int32) + 2                 # first token of segment b
        for s0, e0 in zip(starts, cu[1:]):
            typ[s0:e0] = 1
        return ids, typ, cu
