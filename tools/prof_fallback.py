"""How often does the certificate of the tensor scan fail on clustered data, and what does a flagged query cost?
(VERDICT round 1, weak #3.)  Corpora of N x 384 unit vectors: iid Gaussian; a mixture of `C` tight Gaussian clusters
(sigma chosen so that a cluster's members are ~0.99 cosine to their centre: a real embedding corpus of near-topics);
and the iid corpus with bursts of near-duplicates (sigma 1e-3) planted around the queries' nearest rows.
Prints flagged queries / search latency per corpus and k; writes gpurun_out/fallback.json."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ragmeup_b200.index import FlatIndex, MODE_AUTO, MODE_TENSOR_NOFALLBACK  # noqa: E402

N = int(os.environ.get("PROF_N", 10_000_000)); D = 384; Q = 64
dev = torch.device("cuda")
out = {}


def unit(x):
    return torch.nn.functional.normalize(x, dim=1)


def build(kind):
    ix = FlatIndex(D, "cosine"); ix.reserve(N)
    g = torch.Generator(device=dev).manual_seed(11)
    centers = unit(torch.randn(2000, D, device=dev, generator=g))
    for b in range(0, N, 1_000_000):
        n = min(1_000_000, N - b)
        if kind == "clusters":
            cid = torch.randint(0, 2000, (n,), device=dev, generator=g)
            x = unit(centers[cid] + 0.007 * torch.randn(n, D, device=dev, generator=g))      # |noise| ~ 0.14 -> cos ~ 0.99
        else:
            x = unit(torch.randn(n, D, device=dev, generator=g))
        ix.add(x)
    return ix, centers


def timed(ix, q, k, mode):
    ts = []
    for _ in range(4):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); ix.search(q, k, mode=mode); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ix.search(q, k, mode=MODE_TENSOR_NOFALLBACK, want_stats=True)
    return min(ts[1:]), ix.last_stats[0]


for kind in ("iid", "clusters", "bursts"):
    ix, centers = build("clusters" if kind == "clusters" else "iid")
    g = torch.Generator(device=dev).manual_seed(5)
    if kind == "clusters":
        q = unit(centers[:Q] + 0.007 * torch.randn(Q, D, device=dev, generator=g))           # queries inside clusters of ~5000 rows
    else:
        q = unit(torch.randn(Q, D, device=dev, generator=g))
    if kind == "bursts":                                    # 400 near-copies (cos > 0.9999) of every query, adjacent rows
        for i in range(Q):
            rows = torch.arange(i * 50_000, i * 50_000 + 400, device=dev)
            ix.set_rows(rows, unit(q[i][None] + 1e-3 * torch.randn(400, D, device=dev, generator=g)))
    for k in (10, 100):
        t_auto, flagged = timed(ix, q, k, MODE_AUTO)
        t_tensor, _ = timed(ix, q, k, MODE_TENSOR_NOFALLBACK)
        print(f"{kind:9s} k={k:3d}: flagged {flagged:2d}/{Q}  search {t_auto:.3f} ms (certified part alone {t_tensor:.3f} ms)", flush=True)
        out[f"{kind}_k{k}"] = {"flagged": flagged, "queries": Q, "ms_auto": t_auto, "ms_tensor_only": t_tensor}
    del ix
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/fallback.json", "w"), indent=1)
