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
    valid = fi >= 0
    assert (mi >= 0).tolist() == valid.tolist()
    assert np.allclose(ms[valid], fs[valid], atol=1e-6)
    # ids are identical wherever the scores decide; the planted duplicate has mathematically EQUAL scores in two shards,
    # and numpy's BLAS may round the same dot product differently for a 1-row and a 2-row shard (found by hypothesis:
    # n=4, R=3), so inside a tie (|gap| <= 1e-6) only the set of rows is compared
    for qi in range(nq):
        f_ids, m_ids, f_s = fi[qi][valid[qi]], mi[qi][valid[qi]], fs[qi][valid[qi]].astype(np.float64)
        lo = 0
        for j in range(1, len(f_ids)):
            if abs(f_s[j] - f_s[j - 1]) > 1e-6:          # a tie group that ends inside the list: same rows in it
                assert sorted(f_ids[lo:j].tolist()) == sorted(m_ids[lo:j].tolist())
                lo = j
        # (the last group may be cut by k inside a tie: its members are only known to carry the same scores)


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


@settings(max_examples=40, deadline=None)
@given(h=st.sampled_from([384, 768]), mean=st.sampled_from([0.0, 3.0, -250.0]), std=st.sampled_from([1e-3, 1.0, 40.0]),
       seed=st.integers(0, 10_000))
def test_chunked_chan_statistics_match_two_pass_layernorm(h, mean, std, seed):
    """The fused GEMM+LayerNorm epilogue never sees a whole row: each warp forms exact two-pass
    (mean, M2) per 32-column chunk in fp32, merges its chunks, and the 2*CL column slices of a row are merged again after
    the cluster exchange (Chan et al.).  Emulated in fp32: the merged statistics give the same normalised row as the
    reference's two-pass LayerNorm to fp32 rounding, also when |mean| >> std (where E[x^2] - E[x]^2 would lose everything)."""
    rng = np.random.default_rng(seed)
    f = np.float32
    x = (rng.standard_normal(h) * std + mean).astype(f)
    part_cols = 96
    parts = []
    for p0 in range(0, h, part_cols):
        m_loc, m2 = f(0), f(0)
        for cc, c0 in enumerate(range(p0, p0 + part_cols, 32)):
            y = x[c0:c0 + 32]
            mc = f(y.sum(dtype=f) * f(1 / 32))
            qc = f(((y - mc) ** 2).sum(dtype=f))
            na, nn = f(32 * cc), f(32 * cc + 32)
            delta = f(mc - m_loc)
            m_loc = f(m_loc + delta * f(32) / nn)
            m2 = f(m2 + qc + delta * delta * (na * f(32) / nn))
        parts.append((m_loc, m2))
    mean_all = f(sum(p[0] for p in parts) / f(len(parts)))
    M2 = f(0)
    for m_i, m2_i in parts:
        d = f(m_i - mean_all)
        M2 = f(M2 + m2_i + f(part_cols) * d * d)
    eps = f(1e-12)
    got = (x - mean_all) * f(1.0 / np.sqrt(f(M2 / f(h)) + eps))
    x64 = x.astype(np.float64)
    want = (x64 - x64.mean()) / np.sqrt(x64.var() + 1e-12)
    # the reference computes this in fp32 as well: compare at a few fp32 ulps of the normalised values (|.| <~ 5),
    # plus the cancellation (x - mean) inherits from fp32 inputs: |mean| * 2^-24 / std
    tol = 4e-6 + 4 * abs(mean) * 2.0 ** -24 / std
    assert np.abs(got - want).max() <= tol
