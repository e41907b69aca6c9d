"""CPU property tests (hypothesis) on the oracle's search / merge / MMR restatements — the same
size-independent properties the GPU tests rely on at BASELINE sizes."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import flat_ref


@settings(max_examples=25, deadline=None)
@given(n=st.integers(1, 300), d=st.sampled_from([8, 32]), nq=st.integers(1, 4), k=st.integers(1, 20),
       R=st.integers(1, 6), metric=st.sampled_from(["ip", "cosine", "l2"]), seed=st.integers(0, 10_000))
def test_shard_merge_equals_full_search(n, d, nq, k, R, metric, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    if n > 3:
        x[n - 1] = x[0]                                   # duplicate across shards
    q = rng.standard_normal((nq, d)).astype(np.float32)
    fs, fi = flat_ref.flat_search(q, x, k, metric)
    bounds = [n * r // R for r in range(R + 1)]
    parts = [flat_ref.flat_search(q, x[bounds[r]:bounds[r + 1]], k, metric, id_offset=bounds[r]) for r in range(R)]
    ms, mi = flat_ref.shard_merge(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]), metric)
    assert (mi == fi).all()
    valid = fi >= 0
    assert np.allclose(ms[valid], fs[valid], atol=1e-6)


@settings(max_examples=25, deadline=None)
@given(n=st.integers(2, 200), k=st.integers(1, 10), seed=st.integers(0, 10_000))
def test_results_sorted_and_unit_vector_metrics_agree(n, k, seed):
    """on unit vectors L2^2 = 2 - 2 ip, so all three metrics return the same ids (the reason the
    reference's Milvus-L2 and PGVector-cosine stores rank alike)"""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, 16)).astype(np.float64)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    q = rng.standard_normal((2, 16)).astype(np.float64)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    si, ii = flat_ref.flat_search(q, x, k, "ip", dtype=np.float64)
    sc, ic = flat_ref.flat_search(q, x, k, "cosine", dtype=np.float64)
    sl, il = flat_ref.flat_search(q, x, k, "l2", dtype=np.float64)
    kk = min(k, n)
    assert (np.diff(si[:, :kk], axis=1) <= 1e-12).all() and (np.diff(sl[:, :kk], axis=1) >= -1e-12).all()
    gaps = np.abs(np.diff(si[:, :kk], axis=1))
    if gaps.size == 0 or gaps.min() > 1e-6:               # no near-ties: the orders must coincide
        assert (ii == ic).all() and (ii == il).all()
        assert np.allclose(sl[:, :kk], 2 - 2 * si[:, :kk], atol=1e-6)


@settings(max_examples=30, deadline=None)
@given(n=st.integers(1, 20), k=st.integers(0, 25), lam=st.floats(0.0, 1.0), seed=st.integers(0, 10_000))
def test_mmr_is_a_duplicate_free_prefix_selection(n, k, lam, seed):
    rng = np.random.default_rng(seed)
    E = rng.standard_normal((n, 12))
    q = rng.standard_normal(12)
    sel = flat_ref.mmr(q, E, lambda_mult=lam, k=k)
    assert len(sel) == min(max(k, 0), n) and len(set(sel)) == len(sel)
    if sel:
        assert sel[0] == int(np.argmax(flat_ref.cosine_similarity(q[None], E)[0]))
        assert flat_ref.mmr(q, E, lambda_mult=lam, k=len(sel) + 1)[: len(sel)] == sel      # greedy: prefixes nest
