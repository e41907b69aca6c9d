"""CPU: the certified coarse-to-exact search ALGORITHM of the tensor scan, emulated in numpy.

coarse keys = TF32-reduced inner products -> KEEP best candidates -> exact fp32 re-score -> certificate
`score(T) + eps < exact k-th` (T = KEEP-th best coarse key, eps = 2.2e-3 |q| max|x| + 1e-6 (1 + |k-th|)) -> accept, or
flag the query for the exact scan.  The property the product relies on: whenever the certificate accepts, the top-k
taken from the candidates IS the exact fp32 top-k over all rows — on random data and on adversarial near-duplicate
clusters where TF32 ordering and fp32 ordering disagree."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import flat_ref


def _tf32(a, mode):
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).copy()
    if mode == "rn":
        u = u + np.uint32(0x0FFF) + ((u >> np.uint32(13)) & np.uint32(1))
    return (u & np.uint32(0xFFFFE000)).view(np.float32)


def certified_topk(q, x, k, keep, mode):
    """-> (ids or None when flagged, accepted: bool) for one query, inner-product metric"""
    coarse = (_tf32(q, mode).astype(np.float64) @ _tf32(x, mode).astype(np.float64).T).astype(np.float32)
    order = np.lexsort((np.arange(len(x)), -coarse))                 # key desc, row asc
    cand = order[:keep]
    if len(x) <= keep:
        return None, False                                           # fewer rows than KEEP: the kernel keeps everything
    T = coarse[order[keep - 1]]
    exact = flat_ref.metric_values(q[None], x[cand], "ip")[0]        # fp32, the oracle's arithmetic
    o2 = np.lexsort((cand, -exact))
    kth = exact[o2[k - 1]]
    eps = 2.2e-3 * float(np.linalg.norm(q)) * float(np.linalg.norm(x, axis=1).max()) + 1e-6 * (1.0 + abs(float(kth)))
    if not (float(T) + eps < float(kth)):
        return None, False
    return cand[o2[:k]], True


@settings(max_examples=60, deadline=None)
@given(n=st.integers(400, 3000), d=st.sampled_from([32, 384]), k=st.sampled_from([1, 10]), keep=st.sampled_from([32, 64]),
       kind=st.sampled_from(["random", "clusters", "planted"]), mode=st.sampled_from(["trunc", "rn"]), seed=st.integers(0, 10_000))
def test_accepted_certificate_implies_exact_topk(n, d, k, keep, kind, mode, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    q = rng.standard_normal(d).astype(np.float32)
    if kind == "clusters":                                   # hundreds of near-duplicates of the best rows
        base = x[:8]
        x[8:8 + 40 * 8] = (np.repeat(base, 40, axis=0) * (1 + 2e-4 * rng.standard_normal((320, 1)))).astype(np.float32)
        q = (base[0] + 0.05 * rng.standard_normal(d)).astype(np.float32)
    elif kind == "planted":                                  # k clear winners, everything else far below
        q = q / np.linalg.norm(q)
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        x[:k] = (q[None] + 0.02 * rng.standard_normal((k, d))).astype(np.float32)
    ids, ok = certified_topk(q, x, k, keep, mode)
    es, ei = flat_ref.flat_search(q[None], x, k, "ip")
    if ok:
        assert ids.tolist() == ei[0].tolist()
    if kind == "planted":
        assert ok                                            # well-separated data must not be sent to the fallback


def test_certificate_rejects_when_the_cut_is_ambiguous():
    """3 * KEEP rows whose scores differ by less than the TF32 error: the certificate must flag (the exact scan decides)."""
    rng = np.random.default_rng(1)
    d, keep, k = 64, 32, 10
    q = rng.standard_normal(d).astype(np.float32)
    x = np.repeat(q[None] / np.linalg.norm(q), 96, axis=0).astype(np.float32)
    x *= (1 + 1e-5 * rng.standard_normal((96, 1))).astype(np.float32)
    x = np.concatenate([x, 0.01 * rng.standard_normal((200, d)).astype(np.float32)])
    ids, ok = certified_topk(q, x, k, keep, "trunc")
    assert not ok and ids is None
