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


def _tf32(a: np.ndarray, mode: str) -> np.ndarray:
    """fp32 -> TF32 (10 explicit mantissa bits) by truncation or round-to-nearest-even, as float32."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).copy()
    if mode == "rn":
        u = u + np.uint32(0x0FFF) + ((u >> np.uint32(13)) & np.uint32(1))
    return (u & np.uint32(0xFFFFE000)).view(np.float32)


@settings(max_examples=40, deadline=None)
@given(d=st.sampled_from([4, 96, 384, 768]), scale_q=st.sampled_from([1e-3, 1.0, 37.0]), scale_x=st.sampled_from([1e-2, 1.0, 250.0]),
       shape=st.sampled_from(["normal", "same_sign", "heavy"]), mode=st.sampled_from(["trunc", "rn"]), seed=st.integers(0, 10_000))
def test_tf32_coarse_key_error_is_inside_the_certificate_margin(d, scale_q, scale_x, shape, mode, seed):
    """The certificate of the tensor scan assumes |<tf32(q), tf32(x)> - <q, x>| <= 2.2e-3 * |q| * |x|
    whichever way the tensor core reduces fp32 operands to TF32 (truncation or rounding of both operands:
    (1 + 2^-10)^2 - 1 < 2^-9 per product, Cauchy-Schwarz over the sum).  Checked here on adversarial shapes,
    including all-same-sign vectors where the errors cannot cancel."""
    rng = np.random.default_rng(seed)
    def draw(n):
        if shape == "normal":
            v = rng.standard_normal((n, d))
        elif shape == "same_sign":
            v = np.abs(rng.standard_normal((n, d))) + 0.999          # mantissas just below a power of two lose the most
        else:
            v = rng.standard_cauchy((n, d))
        return v.astype(np.float32)
    q = draw(4) * np.float32(scale_q)
    x = draw(64) * np.float32(scale_x)
    exact = q.astype(np.float64) @ x.astype(np.float64).T
    coarse = _tf32(q, mode).astype(np.float64) @ _tf32(x, mode).astype(np.float64).T
    bound = 2.2e-3 * np.linalg.norm(q.astype(np.float64), axis=1)[:, None] * np.linalg.norm(x.astype(np.float64), axis=1)[None, :]
    # fp32 accumulation of d products adds at most ~d * 2^-24 relative to sum |q_i x_i| <= |q||x|
    slack = d * 2.0 ** -24 * np.linalg.norm(q.astype(np.float64), axis=1)[:, None] * np.linalg.norm(x.astype(np.float64), axis=1)[None, :]
    assert np.all(np.abs(coarse - exact) <= bound - slack)


@settings(max_examples=30, deadline=None)
@given(k=st.sampled_from([64, 384, 1536]), scale=st.sampled_from([1e-2, 1.0, 30.0]), seed=st.integers(0, 10_000))
def test_fp16_split_three_term_product_is_fp32_grade(k, scale, seed):
    """The encoder GEMMs carry fp32 operands as fp16 planes v = hi + lo and sum hi*hi + lo*hi + hi*lo.
    hi + lo reproduces v to max(2^-21 |v|, 2^-25) (the lo plane of a value below ~0.06 is an fp16 subnormal, spacing
    2^-24: an ABSOLUTE 3e-8, harmless for a 1e-3 bar), and the dropped lo*lo term is 2^-22 of the product, so the 3-term
    sum is within a few 2^-21 |a||w| (+ that absolute floor) of the exact dot product — the fp32 noise floor — whereas a
    single fp16 term is ~2^-11."""
    rng = np.random.default_rng(seed)
    a = (rng.standard_normal((8, k)) * scale).astype(np.float32)
    w = (rng.standard_normal((16, k)) * 0.05).astype(np.float32)

    def split(v):
        hi = v.astype(np.float16)
        lo = (v - hi.astype(np.float32)).astype(np.float16)
        return hi.astype(np.float64), lo.astype(np.float64)

    ah, al = split(a)
    wh, wl = split(w)
    assert np.all(np.abs((ah + al) - a) <= np.maximum(2.0 ** -21 * np.abs(a), 2.0 ** -25))
    exact = a.astype(np.float64) @ w.astype(np.float64).T
    three = ah @ wh.T + al @ wh.T + ah @ wl.T
    one = ah @ wh.T
    norm = np.linalg.norm(a.astype(np.float64), axis=1)[:, None] * np.linalg.norm(w.astype(np.float64), axis=1)[None, :]
    floor = 2.0 ** -25 * (np.abs(w).sum(axis=1)[None, :] + np.abs(a).sum(axis=1)[:, None])
    assert np.all(np.abs(three - exact) <= 4 * 2.0 ** -21 * norm + floor)
    assert np.abs(one - exact).max() > 20 * np.abs(three - exact).max()
