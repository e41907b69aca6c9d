"""Seeded inputs shared by make_golden.py and the tests (numpy Generator streams are stable across
platforms, so only the expected OUTPUTS need to be stored in the fixtures)."""
import numpy as np

FLAT_CASES = {  # name: (rows, dim, queries, k)
    "a": (1000, 384, 5, 10),
    "b": (7, 64, 3, 10),
    "c": (4099, 768, 4, 50),
    "d": (1, 32, 2, 3),
    "e": (20000, 384, 6, 100),
}


def flat_case(name):
    n, d, nq, k = FLAT_CASES[name]
    rng = np.random.default_rng({"a": 11, "b": 12, "c": 13, "d": 14, "e": 15}[name])
    x = rng.standard_normal((n, d)).astype(np.float32)
    q = rng.standard_normal((nq, d)).astype(np.float32)
    if n >= 1000:
        x[10] = x[3]
        x[500] = x[3]                      # exact duplicates -> tied scores, pins the tie-break
        q[0] = x[3] + 0.01 * rng.standard_normal(d).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return np.ascontiguousarray(x, dtype=np.float32), np.ascontiguousarray(q, dtype=np.float32), k


# ---------------------------------------------------------------------------------------------- BM25 (sparse leg)
import os as _os

BM25_CASES = [
    # n docs, vocabulary, doc length range, distinct texts (0 = all distinct), queries, k
    dict(name="small", n=300, vocab=80, lens=(1, 40), distinct=0, nq=6, k=4, seed=21),
    dict(name="dups", n=2000, vocab=300, lens=(3, 25), distinct=50, nq=5, k=10, seed=22),
    dict(name="blocks", n=10000, vocab=2000, lens=(5, 60), distinct=0, nq=4, k=100, seed=23),
    dict(name="rounds", n=70000, vocab=5000, lens=(5, 30), distinct=0, nq=3, k=256, seed=24),
]


def golden_path(name):
    return _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), name)


def bm25_inputs(case):
    """(tokenised corpus, tokenised queries): Zipf-distributed pseudo-words, so frequent words get NEGATIVE idf
    (df > N/2) and hit rank_bm25's epsilon floor; queries mix frequent, rare, repeated and unknown words."""
    rng = np.random.default_rng(case["seed"])
    v = case["vocab"]
    words = [f"w{i}" for i in range(v)]
    p = 1.0 / (np.arange(v) + 1.0)
    p /= p.sum()
    lo, hi = case["lens"]

    def make_doc():
        return [words[j] for j in rng.choice(v, size=int(rng.integers(lo, hi + 1)), p=p)]

    if case["distinct"]:
        pool = [make_doc() for _ in range(case["distinct"])]
        corpus = [pool[int(j)] for j in rng.integers(0, len(pool), case["n"])]
    else:
        corpus = [make_doc() for _ in range(case["n"])]
    queries = []
    for qi in range(case["nq"]):
        q = [words[j] for j in rng.choice(v, size=int(rng.integers(1, 9)), p=p)]
        if qi % 3 == 0:
            q += [q[0], "never-seen-word"]            # repeated term + unknown term
        if qi % 3 == 1:
            q += [words[int(rng.integers(v // 2, v))]]  # a rare word
        queries.append(q)
    queries[-1] = ["never-seen-word"]                  # nothing matches: all scores 0
    return corpus, queries
