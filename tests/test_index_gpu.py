"""GPU parity: flat index (tcgen05 scan + exact path + merge + MMR) vs the oracle and the goldens.
ids: identical; scores: |d| <= 1e-3 (BASELINE.json), in practice ~1e-6."""
import numpy as np
import pytest
import torch

from cases import FLAT_CASES, flat_case
from oracle import flat_ref

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _idx(cuda, x, metric):
    from ragmeup_b200.index import FlatIndex
    ix = FlatIndex(x.shape[1], metric)
    ix.add(x if isinstance(x, torch.Tensor) else np.ascontiguousarray(x))
    return ix


def _check(s, i, rs, ri):
    i = i.cpu().numpy() if isinstance(i, torch.Tensor) else i
    s = s.cpu().numpy() if isinstance(s, torch.Tensor) else s
    assert (i == ri).all()
    valid = ri >= 0
    assert np.abs(s[valid] - rs[valid]).max() <= TOL
    assert np.isinf(s[~valid]).all()


def _check_topk_fp64(s, i, q, x, k, metric):
    """ids identical to an exact search up to fp32 near-ties: judged in float64 — returned scores
    within TOL, returned order sorted, and no excluded row better than the k-th returned one, with a
    slack of a few fp32 ulps of the score magnitude (unnormalised L2 distances are ~1e3, so two
    fp32 implementations may order rows whose distances differ by < 1e-4 either way)."""
    i = i.cpu().numpy()
    s = s.cpu().numpy().astype(np.float64)
    v = flat_ref.metric_values(q, x, metric, dtype=np.float64)
    rank = v if metric == "l2" else -v
    eps = 8 * np.finfo(np.float32).eps * max(1.0, float(np.abs(v).max()))
    for r in range(q.shape[0]):
        ids = i[r]
        assert (ids >= 0).all() and len(set(ids.tolist())) == len(ids)
        got = rank[r, ids]
        assert np.abs(s[r] - v[r, ids]).max() <= TOL * max(1.0, float(np.abs(v[r, ids]).max()) * 1e-2)
        assert (got[:-1] <= got[1:] + eps).all()
        rest = np.delete(rank[r], ids)
        assert rest.min() >= got[-1] - eps


@pytest.mark.parametrize("name", sorted(FLAT_CASES))
@pytest.mark.parametrize("metric", ["ip", "cosine", "l2"])
def test_golden_cases(cuda, golden_dir, name, metric):
    """planted duplicates / ties, k > N, N = 1, D = 768, N big enough for the tensor scan (case e)"""
    g = np.load(f"{golden_dir}/flat.npz")
    x, q, k = flat_case(name)
    ix = _idx(cuda, x, metric)
    s, i = ix.search(torch.from_numpy(q).cuda(), k)
    gi, gs = g[f"flat_{name}_{metric}_ids"], g[f"flat_{name}_{metric}_scores"]
    i = i.cpu().numpy()
    for r in range(q.shape[0]):                       # identical sets (golden order is fp64)
        assert set(i[r].tolist()) == set(gi[r].tolist())
    rs, ri = flat_ref.flat_search(q, x, k, metric)
    _check(s, i, rs, ri)                              # identical order vs the fp32 oracle incl. the tie rule


@pytest.mark.parametrize("metric", ["ip", "cosine", "l2"])
@pytest.mark.parametrize("n,nq,k", [(16384, 3, 10), (50_001, 130, 20), (40_000, 1, 100), (30_000, 64, 50)])
def test_tensor_scan_equals_exact_and_oracle(cuda, metric, n, nq, k):
    from ragmeup_b200.index import MODE_AUTO, MODE_EXACT
    g = torch.Generator(device="cuda").manual_seed(n + nq)
    x = torch.randn(n, 384, device="cuda", generator=g)
    if metric != "l2":
        x = torch.nn.functional.normalize(x, dim=1)
    x[n // 2] = x[7]                                  # duplicate rows -> tie decided by the lower row
    q = torch.randn(nq, 384, device="cuda", generator=g)
    q[0] = x[7]
    ix = _idx(cuda, x, metric)
    s0, i0 = ix.search(q, k, mode=MODE_EXACT)
    s1, i1 = ix.search(q, k, mode=MODE_AUTO, want_stats=True)
    assert ix.last_stats[1] >= 1                      # the tcgen05 scan ran
    assert (i0 == i1).all() and (s0 == s1).all()      # bit-identical to the exact path
    if metric == "l2":
        _check_topk_fp64(s1, i1, q.cpu().numpy(), x.cpu().numpy(), k, metric)
    else:
        rs, ri = flat_ref.flat_search(q.cpu().numpy(), x.cpu().numpy(), k, metric)
        _check(s1, i1, rs, ri)


@pytest.mark.parametrize("metric", ["ip", "cosine", "l2"])
@pytest.mark.parametrize("dim,n,nq,k", [(768, 20_000, 5, 10), (400, 17_000, 130, 20), (768, 33_000, 64, 50)])
def test_wide_vectors_on_the_tensor_scan(cuda, metric, dim, n, nq, k):
    """384 < dim <= 768 (bge-base 768-d, config 5): fewer queries per CTA (the resident query operand must fit shared
    memory), clusters with TMA multicast make up for it; results must still be bit-identical to the exact fp32 scan"""
    from ragmeup_b200.index import MODE_AUTO, MODE_EXACT
    g = torch.Generator(device="cuda").manual_seed(dim + n)
    x = torch.nn.functional.normalize(torch.randn(n, dim, device="cuda", generator=g), dim=1)
    if metric == "l2":
        x = x * (1.0 + 0.2 * torch.rand(n, 1, device="cuda", generator=g))     # non-unit rows: per-row bias path
    x[n - 3] = x[5]
    q = torch.nn.functional.normalize(torch.randn(nq, dim, device="cuda", generator=g), dim=1)
    q[0] = x[5]
    ix = _idx(cuda, x, metric)
    s0, i0 = ix.search(q, k, mode=MODE_EXACT)
    s1, i1 = ix.search(q, k, mode=MODE_AUTO, want_stats=True)
    assert ix.last_stats[1] >= 1                      # the tcgen05 scan ran
    assert (i0 == i1).all() and (s0 == s1).all()
    _check_topk_fp64(s1, i1, q.cpu().numpy(), x.cpu().numpy(), k, metric)


def test_certificate_failure_is_answered_by_the_second_pass(cuda):
    """more near-ties than the coarse pass re-scores: the certificate must fail; the second tensor pass (every row whose
    coarse key can reach the k-th exact score, re-scored exactly) must answer without the CUDA-core scan, bit-identical to
    MODE_EXACT and a valid top-k in float64 up to fp32 near-ties"""
    from ragmeup_b200.index import MODE_AUTO, MODE_EXACT, MODE_TENSOR_NOFALLBACK
    g = torch.Generator(device="cuda").manual_seed(3)
    base = torch.nn.functional.normalize(torch.randn(1, 384, device="cuda", generator=g), dim=1)
    x = torch.nn.functional.normalize(torch.randn(20000, 384, device="cuda", generator=g), dim=1)
    x[1000:1400] = torch.nn.functional.normalize(base + 1e-4 * torch.randn(400, 384, device="cuda", generator=g), dim=1)
    ix = _idx(cuda, x, "ip")
    s, i = ix.search(base, 10, mode=MODE_AUTO, want_stats=True)
    assert ix.last_stats[0] == 1                      # flagged by the certificate ...
    assert ix.last_stats[2] == 0                      # ... and answered by the second pass, not by the exact scan
    s0, i0 = ix.search(base, 10, mode=MODE_EXACT)
    assert (i == i0).all() and (s == s0).all()
    assert ((i >= 1000) & (i < 1400)).all()
    _check_topk_fp64(s, i, base.cpu().numpy(), x.cpu().numpy(), 10, "ip")
    ix.search(base, 10, mode=MODE_TENSOR_NOFALLBACK, want_stats=True)
    assert ix.last_stats[0] == 1


@pytest.mark.parametrize("metric", ["ip", "cosine", "l2"])
def test_second_pass_on_clustered_corpus_and_pool_overflow(cuda, metric):
    """a clustered corpus (every query sits in a cluster of ~600 rows within 1e-3 of each other): all queries flagged,
    all answered by the second pass, results bit-identical to the exact scan; a cluster larger than the pool (20 000
    copies of one row) overflows it and falls through to the exact scan, still exact"""
    from ragmeup_b200.index import MODE_AUTO, MODE_EXACT
    g = torch.Generator(device="cuda").manual_seed(17)
    n, d, nq = 60_000, 384, 33
    centers = torch.nn.functional.normalize(torch.randn(100, d, device="cuda", generator=g), dim=1)
    cid = torch.randint(0, 100, (n,), device="cuda", generator=g)
    x = torch.nn.functional.normalize(centers[cid] + 2e-4 * torch.randn(n, d, device="cuda", generator=g), dim=1)
    if metric == "l2":
        x = x * (1.0 + 0.1 * torch.rand(n, 1, device="cuda", generator=g))
    q = torch.nn.functional.normalize(centers[:nq] + 2e-4 * torch.randn(nq, d, device="cuda", generator=g), dim=1)
    ix = _idx(cuda, x, metric)
    s, i = ix.search(q, 20, mode=MODE_AUTO, want_stats=True)
    flagged, _, to_exact = ix.last_stats
    s0, i0 = ix.search(q, 20, mode=MODE_EXACT)
    assert (i == i0).all() and (s == s0).all()
    if metric != "l2":
        assert flagged == nq and to_exact == 0
    big = torch.nn.functional.normalize(torch.randn(1, d, device="cuda", generator=g), dim=1)
    xb = torch.cat([x, big.repeat(20_000, 1)])
    ixb = _idx(cuda, xb, "ip")
    s, i = ixb.search(big, 10, mode=MODE_AUTO, want_stats=True)
    assert ixb.last_stats[0] == 1 and ixb.last_stats[2] == 1          # pool overflow -> the exact scan
    s0, i0 = ixb.search(big, 10, mode=MODE_EXACT)
    assert (i == i0).all() and (s == s0).all()
    assert (i.cpu().numpy()[0] == np.arange(n, n + 10)).all()         # 20 000 exact ties: the lowest rows win


def test_empty_add_growth_offsets_and_host_api(cuda):
    from ragmeup_b200.index import FlatIndex
    ix = FlatIndex(64, "l2")
    q = torch.randn(2, 64, device="cuda")
    s, i = ix.search(q, 3)
    assert (i == -1).all() and torch.isinf(s).all()
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2500, 64)).astype(np.float32)
    for a in range(0, 2500, 700):                     # repeated adds force re-allocation
        ix.add(x[a:a + 700])
    assert len(ix) == 2500
    s, i = ix.search(q, 5, id_offset=1000)
    rs, ri = flat_ref.flat_search(q.cpu().numpy(), x, 5, "l2", id_offset=1000)
    _check(s, i, rs, ri)
    hs, hi = ix.search_host(q.cpu().numpy(), 5, id_offset=1000)
    _check(hs, hi, rs, ri)
    assert np.array_equal(ix.data().cpu().numpy(), x)
    got = ix.gather(torch.tensor([3, 2499, 0], device="cuda")).cpu().numpy()
    assert np.array_equal(got, x[[3, 2499, 0]])
    ix.clear()
    assert len(ix) == 0


def test_dim_not_multiple_of_four_uses_exact_path(cuda):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((20000, 50)).astype(np.float32)
    q = rng.standard_normal((3, 50)).astype(np.float32)
    ix = _idx(cuda, x, "cosine")
    s, i = ix.search(torch.from_numpy(q).cuda(), 7, want_stats=True)
    assert ix.last_stats[1] == 0
    _check(s, i, *flat_ref.flat_search(q, x, 7, "cosine"))


@pytest.mark.parametrize("R", [2, 4, 8])
@pytest.mark.parametrize("metric", ["ip", "l2"])
def test_shard_merge_kernel(cuda, R, metric):
    """per-shard search + rmu_topk_merge == unsharded search == oracle (KA6)"""
    from ragmeup_b200.index import topk_merge
    g = torch.Generator(device="cuda").manual_seed(R)
    n = 40_000
    x = torch.nn.functional.normalize(torch.randn(n, 384, device="cuda", generator=g), dim=1)
    x[n - 5] = x[11]                                  # duplicate in another shard
    q = torch.nn.functional.normalize(torch.randn(9, 384, device="cuda", generator=g), dim=1)
    q[0] = x[11]
    k = 10
    full = _idx(cuda, x, metric)
    fs, fi = full.search(q, k)
    bounds = [n * r // R for r in range(R + 1)]
    ss, ii = [], []
    for r in range(R):
        sh = _idx(cuda, x[bounds[r]:bounds[r + 1]], metric)
        s, i = sh.search(q, k, id_offset=bounds[r])
        ss.append(s)
        ii.append(i)
    ms, mi = topk_merge(torch.stack(ss), torch.stack(ii), metric)
    assert (mi == fi).all() and (ms == fs).all()
    os_, oi = flat_ref.shard_merge(torch.stack(ss).cpu().numpy(), torch.stack(ii).cpu().numpy(), metric)
    assert (mi.cpu().numpy() == oi).all()


def test_mmr_matches_oracle(cuda):
    from ragmeup_b200.index import mmr_select
    g = torch.Generator(device="cuda").manual_seed(5)
    q = torch.randn(6, 384, device="cuda", generator=g)
    cand = torch.randn(6, 20, 384, device="cuda", generator=g) + q[:, None, :] * 0.7
    cand[0, 3] = cand[0, 1]                           # duplicate candidate
    n_cand = torch.tensor([20, 20, 20, 7, 1, 20], dtype=torch.int32)
    sel = mmr_select(q, cand, n_cand, 10, 0.5).cpu().numpy()
    for b in range(6):
        n = int(n_cand[b])
        want = flat_ref.mmr(q[b].cpu().numpy(), cand[b, :n].cpu().numpy(), 0.5, 10)
        assert sel[b, :len(want)].tolist() == want
        assert (sel[b, len(want):] == -1).all()
    for lam in (0.0, 1.0, 0.3):
        sel = mmr_select(q, cand, None, 4, lam).cpu().numpy()
        for b in range(6):
            assert sel[b].tolist() == flat_ref.mmr(q[b].cpu().numpy(), cand[b].cpu().numpy(), lam, 4)


def test_large_corpus_properties(cuda):
    """BASELINE size class (1M x 384, Q=64): planted neighbours must come back first; results are
    sorted; searching twice is idempotent; shard-merge of two halves equals the full search."""
    from ragmeup_b200.index import FlatIndex, topk_merge
    n, d, nq, k = 1_000_000, 384, 64, 10
    g = torch.Generator(device="cuda").manual_seed(1234)
    ix = FlatIndex(d, "cosine")
    ix.reserve(n)
    halves = [FlatIndex(d, "cosine"), FlatIndex(d, "cosine")]
    planted = None
    for c in range(4):
        x = torch.nn.functional.normalize(torch.randn(n // 4, d, device="cuda", generator=g), dim=1)
        if c == 2:
            planted = x[1000:1000 + nq].clone()
        ix.add(x)
        halves[c // 2].add(x)
    q = torch.nn.functional.normalize(planted + 0.05 * torch.randn(nq, d, device="cuda", generator=g), dim=1)
    s, i = ix.search(q, k, want_stats=True)
    assert ix.last_stats[1] >= 1
    want0 = torch.arange(nq, device="cuda") + 2 * (n // 4) + 1000
    assert (i[:, 0] == want0).all()
    assert (s[:, :-1] >= s[:, 1:]).all()
    s2, i2 = ix.search(q, k)
    assert (i2 == i).all() and (s2 == s).all()
    a = halves[0].search(q, k)
    b = halves[1].search(q, k, id_offset=n // 2)
    ms, mi = topk_merge(torch.stack([a[0], b[0]]), torch.stack([a[1], b[1]]), "cosine")
    assert (mi == i).all() and (ms == s).all()


@pytest.mark.parametrize("nq,k,metric", [(64, 10, "cosine"), (256, 10, "ip"), (64, 100, "cosine"), (64, 50, "l2")])
def test_million_rows_against_the_oracle(cuda, nq, k, metric):
    """BASELINE size class against the CPU oracle itself (not only through properties): 1M x 384, Q = 64 / 256,
    k = 10 / 50 / 100, unit-norm rows as sentence-transformers' Normalize emits them — ids identical (order too wherever
    the oracle's own scores differ by more than 2e-6), scores within 1e-3 (SURVEY.md §8c tolerance)."""
    from ragmeup_b200.index import FlatIndex
    n, d = 1_000_000, 384
    g = torch.Generator(device="cuda").manual_seed(77)
    ix = FlatIndex(d, metric)
    ix.reserve(n)
    for _ in range(4):
        ix.add(torch.nn.functional.normalize(torch.randn(n // 4, d, device="cuda", generator=g), dim=1))
    q = torch.nn.functional.normalize(torch.randn(nq, d, device="cuda", generator=g), dim=1)
    q[0] = ix.data()[123_456]                            # an exact hit
    s, i = ix.search(q, k, want_stats=True)
    assert ix.last_stats[1] >= 1
    rs, ri = flat_ref.flat_search_blocked(q.cpu().numpy(), ix.data().cpu().numpy(), k, metric)
    s, i = s.cpu().numpy(), i.cpu().numpy()
    assert i[0, 0] == 123_456
    scale = 1.0
    assert np.abs(s - rs).max() <= TOL
    for r in range(nq):
        if set(i[r].tolist()) != set(ri[r].tolist()):
            assert _near_tie_sets(i[r], ri[r], rs[r])
            continue
        for j in np.nonzero(i[r] != ri[r])[0]:           # order may differ only inside fp32 near-ties
            pos = np.nonzero(i[r] == ri[r, j])[0]
            assert len(pos) == 1 and abs(float(rs[r, j]) - float(rs[r, pos[0]])) <= 2e-6 * scale


def _near_tie_sets(got, want, want_scores):
    """two fp32 implementations may disagree on the k-th row when its score ties the (k+1)-th to the last ulp"""
    extra = set(got.tolist()) ^ set(want.tolist())
    return len(extra) <= 2 and abs(float(want_scores[-1]) - float(want_scores[-2])) < 1e-6
