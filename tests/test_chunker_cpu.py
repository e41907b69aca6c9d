"""CPU: host logic of the SemanticChunker drop-in against the oracle's restatement of langchain_experimental's algorithm
(thresholds, breakpoints, chunk assembly, sentence windows) — the embedding + distance arithmetic is the GPU test's."""
import re

import numpy as np
import pytest

from oracle import chunker_ref
from ragmeup_b200 import chunker
from ragmeup_b200.documents import Document


class _DeviceEmb:
    """stands for the B200 embeddings class: has encode_tensor (never called in these tests)"""

    def encode_tensor(self, texts):  # pragma: no cover
        raise AssertionError("the CPU tests must not reach the device path")


def _sentences(n, seed):
    rng = np.random.default_rng(seed)
    return [" ".join(f"w{rng.integers(0, 50)}" for _ in range(rng.integers(2, 9))) + "." for _ in range(n)]


def test_sentence_windows_match_the_reference_algorithm():
    for n in (1, 2, 3, 7):
        for buf in (0, 1, 2):
            sents = _sentences(n, n + buf)
            got = chunker.combine_sentences([{"sentence": s, "index": i} for i, s in enumerate(sents)], buf)
            assert [g["combined_sentence"] for g in got] == chunker_ref.sentence_groups(sents, buf)


@pytest.mark.parametrize("kind", ["percentile", "standard_deviation", "interquartile", "gradient"])
@pytest.mark.parametrize("amount", [None, 50])
def test_breakpoints_and_chunks_match_oracle(kind, amount):
    for seed in range(6):
        rng = np.random.default_rng(seed)
        n = int(rng.integers(3, 60))
        sents = _sentences(n, seed)
        text = " ".join(sents)
        dist = rng.random(n - 1).tolist()
        fake = {g: i for i, g in enumerate(chunker_ref.sentence_groups(re.split(r"(?<=[.?!])\s+", text)))}

        def embed(groups, dist=dist):
            # unit vectors on a circle whose consecutive angles give exactly the wanted cosine distances
            ang = np.concatenate([[0.0], np.cumsum(np.arccos(1.0 - np.asarray(dist)))])
            return [[float(np.cos(a)), float(np.sin(a))] for a in ang[:len(groups)]]

        ref_chunks, ref_dist = chunker_ref.split_text(text, embed, kind=kind, amount=amount)
        sc = chunker.SemanticChunker(_DeviceEmb(), breakpoint_threshold_type=kind, breakpoint_threshold_amount=amount)
        single = re.split(sc.sentence_split_regex, text)
        sentences = chunker.combine_sentences([{"sentence": s, "index": i} for i, s in enumerate(single)], 1)
        assert len(sentences) == len(fake)
        got = sc.chunks_from_distances(sentences, ref_dist)
        assert got == ref_chunks and " ".join(got) == text


def test_number_of_chunks_and_min_chunk_size():
    rng = np.random.default_rng(3)
    sents = _sentences(40, 3)
    text = " ".join(sents)
    dist = rng.random(39).tolist()

    def embed(groups):
        ang = np.concatenate([[0.0], np.cumsum(np.arccos(1.0 - np.asarray(dist)))])
        return [[float(np.cos(a)), float(np.sin(a))] for a in ang[:len(groups)]]

    for k in (1, 5, 39, 100):
        ref_chunks, ref_dist = chunker_ref.split_text(text, embed, number_of_chunks=k)
        sc = chunker.SemanticChunker(_DeviceEmb(), number_of_chunks=k)
        sentences = chunker.combine_sentences([{"sentence": s, "index": i} for i, s in enumerate(sents)], 1)
        got = sc.chunks_from_distances(sentences, ref_dist)
        assert got == ref_chunks
        assert len(got) <= max(k, 1) + 1
    ref_chunks, ref_dist = chunker_ref.split_text(text, embed, kind="percentile", amount=30, min_chunk_size=80)
    sc = chunker.SemanticChunker(_DeviceEmb(), breakpoint_threshold_amount=30, min_chunk_size=80)
    sentences = chunker.combine_sentences([{"sentence": s, "index": i} for i, s in enumerate(sents)], 1)
    got = sc.chunks_from_distances(sentences, ref_dist)
    assert got == ref_chunks and all(len(c) >= 80 for c in got[:-1])


def test_short_texts_constructor_errors_and_documents(monkeypatch):
    sc = chunker.SemanticChunker(_DeviceEmb(), breakpoint_threshold_type=None)        # reference passes None from the .env
    assert sc.breakpoint_threshold_type == "percentile" and sc.breakpoint_threshold_amount == 95
    assert sc.split_text("only one sentence") == ["only one sentence"]
    g = chunker.SemanticChunker(_DeviceEmb(), breakpoint_threshold_type="gradient")
    assert g.split_text("One. Two.") == ["One.", "Two."]
    with pytest.raises(ValueError):
        chunker.SemanticChunker(_DeviceEmb(), breakpoint_threshold_type="nope")
    with pytest.raises(TypeError):
        chunker.SemanticChunker(object())                                              # no device encode: no CPU path
    monkeypatch.setattr(sc, "split_text", lambda t: [t[:4], t[4:]])
    sc._add_start_index = True
    docs = sc.split_documents([Document("abcdefgh", {"source": "a"}), Document("ijklmnop", {"source": "b"})])
    assert [d.page_content for d in docs] == ["abcd", "efgh", "ijkl", "mnop"]
    assert [d.metadata for d in docs] == [{"source": "a", "start_index": 0}, {"source": "a", "start_index": 4},
                                          {"source": "b", "start_index": 0}, {"source": "b", "start_index": 4}]


def test_install_registers_the_text_splitter_shim():
    import importlib
    import sys
    from ragmeup_b200 import install
    install.install()
    mod = importlib.import_module("langchain_experimental.text_splitter")
    assert mod.SemanticChunker is chunker.SemanticChunker or "langchain_experimental" in sys.modules
