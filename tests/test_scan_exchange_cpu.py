"""CPU: numpy emulation of the scan's candidate lists + threshold exchange (csrc/rmu_scan.cuh), the part of the search
that decides which rows are ever re-scored.  Per (CTA, query): a list that is sorted and cut to its best 64 when it
passes 95 entries at a tile boundary; the list's floor = the 64th key at a cut; every sort publishes the list's 16th
key into gmax[query][cta % groups] (max); a query's threshold is tau = max(min over groups of gmax, own floor); rows
scoring <= tau are rejected on the spot.
The properties the product's certificate rests on:
  (1) tau never exceeds the (16 * groups)-th best score of the whole corpus: nothing in the global top-KSEL is rejected
      by the exchange (only a list's own cut can lose such a row, and then its floor reports it);
  (2) every row that is in no list scores <= bound0 = max(tau_final, max floor); with bound = max(bound0, the
      (KSEL+1)-th best candidate), `k-th best candidate > bound` implies the candidates' top-k IS the global top-k."""
import numpy as np
import pytest

KEEP, TRIG, PUB = 64, 95, 16


def emulate(scores, nctas, groups, order_seed=0):
    """scores [N] of one query; rows are dealt to CTAs in interleaved 128-row tiles.  CTAs advance in a random
    interleaving (any interleaving must be valid).  Returns (candidates, tau_final, max floor, taus used)."""
    n = len(scores)
    ntiles = (n + 127) // 128
    tiles_of = [list(range(c, ntiles, nctas)) for c in range(nctas)]
    lists = [[] for _ in range(nctas)]
    floors = np.full(nctas, -np.inf)
    gmax = np.full(max(groups, 1), -np.inf)
    pos = [0] * nctas
    rng = np.random.default_rng(order_seed)
    taus = []

    def sort_cut(c):
        lst = sorted(lists[c], key=lambda e: (-e[0], e[1]))
        if len(lst) > KEEP:
            floors[c] = max(floors[c], lst[KEEP - 1][0])
            lst = lst[:KEEP]
        if len(lst) >= PUB and groups > 0:
            g = c % groups
            gmax[g] = max(gmax[g], lst[PUB - 1][0])
        lists[c] = lst

    live = [c for c in range(nctas) if tiles_of[c]]
    while live:
        c = live[rng.integers(len(live))]
        t = tiles_of[c][pos[c]]
        tau = max(gmax.min() if groups > 0 else -np.inf, floors[c])
        taus.append(tau)
        for r in range(t * 128, min(n, t * 128 + 128)):
            if scores[r] > tau:
                lists[c].append((scores[r], r))
        assert len(lists[c]) < 256                       # the list capacity of the kernel
        if len(lists[c]) > TRIG:
            sort_cut(c)
        pos[c] += 1
        if pos[c] == len(tiles_of[c]):
            live.remove(c)
    tau_fin = gmax.min() if groups > 0 else -np.inf
    cand = [e for c in range(nctas) for e in lists[c] if e[0] > tau_fin]
    return cand, tau_fin, floors.max(), taus


def corpus(kind, n, rng):
    if kind == "random":
        return rng.standard_normal(n).astype(np.float32)
    if kind == "clustered":                              # bursts of near-duplicates, adjacent in insertion order
        s = rng.standard_normal(n).astype(np.float32)
        for _ in range(20):
            a = rng.integers(0, n - 300)
            s[a:a + 300] = 3.0 + 0.001 * rng.standard_normal(300)
        return s
    if kind == "sorted":                                 # best rows last: thresholds keep rising until the end
        return np.sort(rng.standard_normal(n)).astype(np.float32)
    if kind == "ties":
        return np.round(rng.standard_normal(n), 1).astype(np.float32)
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["random", "clustered", "sorted", "ties"])
@pytest.mark.parametrize("ksel,k", [(64, 10), (256, 100)])
def test_exchange_never_rejects_the_global_top_and_the_bound_holds(kind, ksel, k):
    rng = np.random.default_rng(sum(map(ord, kind)) + ksel)
    n, nctas = 60_000, 37
    s = corpus(kind, n, rng)
    groups = ksel // PUB
    cand, tau_fin, floor_max, taus = emulate(s, nctas, groups, order_seed=ksel)
    order = np.lexsort((np.arange(n), -s))
    kth_global = s[order[ksel - 1]]
    assert max(taus) <= kth_global or floor_max >= max(taus)          # (1): the exchange alone never passes the KSEL-th key
    assert tau_fin <= kth_global
    in_list = {r for _, r in cand}
    bound0 = max(tau_fin, floor_max)
    outside = np.array([r for r in range(n) if r not in in_list])
    assert (s[outside] <= bound0).all()                               # (2) every row in no list is under bound0
    cs = sorted(cand, key=lambda e: (-e[0], e[1]))
    bound = max(bound0, cs[ksel][0]) if len(cs) > ksel else bound0
    top = cs[:ksel]
    if len(top) >= k and top[k - 1][0] > bound:                       # certificate accepts
        assert [r for _, r in top[:k]] == order[:k].tolist()
    if kind == "random":
        assert len(top) >= k and top[k - 1][0] > bound                # well-spread data must certify


def test_small_grids_run_without_exchange():
    rng = np.random.default_rng(3)
    s = rng.standard_normal(20_000).astype(np.float32)
    cand, tau_fin, floor_max, _ = emulate(s, 5, 0)
    assert tau_fin == -np.inf
    order = np.lexsort((np.arange(len(s)), -s))
    cs = sorted(cand, key=lambda e: (-e[0], e[1]))
    assert cs[9][0] > floor_max and [r for _, r in cs[:10]] == order[:10].tolist()
