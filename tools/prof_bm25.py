"""Time the BM25 sparse leg (csrc/rmu_bm25.cu) on a synthetic pre-tokenised corpus and check it against a numpy
float64 evaluation of the same postings.  PROF_N documents (default 1M), PROF_LEN tokens per document, PROF_V
vocabulary, PROF_Q queries of PROF_T terms, top PROF_K."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ragmeup_b200.bm25 import BM25Index  # noqa: E402


def main():
    N = int(os.environ.get("PROF_N", 1_000_000)); L = int(os.environ.get("PROF_LEN", 40))
    V = int(os.environ.get("PROF_V", 50_000)); Q = int(os.environ.get("PROF_Q", 64))
    T = int(os.environ.get("PROF_T", 8)); K = int(os.environ.get("PROF_K", 4))
    rng = np.random.default_rng(5)
    p = 1.0 / (np.arange(V) + 1.0); p /= p.sum()
    t0 = time.time()
    lens = rng.integers(L // 2, L + L // 2 + 1, N)
    doc = np.repeat(np.arange(N, dtype=np.int64), lens)
    term = rng.choice(V, size=int(lens.sum()), p=p).astype(np.int64)
    key, tf = np.unique(term * N + doc, return_counts=True)       # sorted by (term, doc)
    pterm, pdoc = key // N, key % N
    post_ptr = np.zeros(V + 1, dtype=np.int64)
    np.cumsum(np.bincount(pterm, minlength=V), out=post_ptr[1:])
    print(f"corpus: {N} docs, {len(key)} postings, built in {time.time() - t0:.1f}s", flush=True)
    idx = BM25Index.from_arrays(post_ptr, pdoc.astype(np.int32), tf.astype(np.int32), lens)
    queries = [[str(t) for t in rng.choice(V, size=T, p=p)] for _ in range(Q)]
    s, d = idx.search(queries, K)
    # numpy check of the first queries (same float64 operation order)
    for qi in range(min(Q, 4)):
        sc = np.zeros(N)
        for t in idx.term_ids(queries[qi]):
            lo, hi = post_ptr[t], post_ptr[t + 1]
            dd = pdoc[lo:hi]; f = tf[lo:hi].astype(np.float64)
            sc[dd] = sc[dd] + idx.idf[t] * ((f * (idx.k1 + 1)) / (f + idx.den[dd]))
        top = np.argsort(sc, kind="stable")[::-1][:K]
        assert np.array_equal(top, d[qi]) and np.array_equal(sc[top], s[qi]), qi
    for name, qs in (("batch", queries), ("single", queries[:1])):
        for _ in range(3):
            idx.search(qs, K)
        torch.cuda.synchronize()
        t0 = time.time()
        it = 20
        for _ in range(it):
            idx.search(qs, K)
        dt = (time.time() - t0) / it
        print(f"bm25 {name}: Q={len(qs)} k={K} terms={T}: {dt * 1e3:.3f} ms/call ({len(qs) / dt:.0f} q/s), host buffers", flush=True)


if __name__ == "__main__":
    main()
