"""GPU check + timing of the flat-index scan (run under gpurun): accumulator dump vs fp64, tensor scan vs exact
scan on the launch geometries (NQ 16/32/64, clusters 1/2/4, dims 100..1536), then search timings per config.
Writes gpurun_out/scan_check.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ragmeup_b200 import _lib  # noqa: E402
from ragmeup_b200.index import FlatIndex, MODE_AUTO, MODE_EXACT  # noqa: E402

dev = torch.device("cuda")
out = {}


def log(*a):
    print(*a, flush=True)


def unit(n, d, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn(n, d, generator=g, device=dev), dim=1)


# ---- 1. raw accumulators of tile 0
x = unit(20000, 384, 1)
q = unit(64, 384, 2)
idx = FlatIndex(384, "ip")
idx.add(x)
dbg = torch.zeros(64, 256, device=dev)
_lib.check(_lib.lib().rmu_debug_scan_tile(idx._h, q.data_ptr(), 64, dbg.data_ptr(), _lib.stream_ptr()), "debug tile")
torch.cuda.synchronize()
ref = (q.double() @ x[:256].double().T).float()
err = (dbg - ref).abs().max().item()
log("tile dump max|err| vs fp64:", err)
out["tile_max_err"] = err
if err > 5e-3:
    log("dbg[0,:8]", dbg[0, :8].tolist(), "ref[0,:8]", ref[0, :8].tolist())
    log("dbg[1,:8]", dbg[1, :8].tolist(), "ref[1,:8]", ref[1, :8].tolist())
    json.dump(out, open("gpurun_out/scan_check.json", "w"))
    sys.exit(1)

# ---- 2. tensor scan == exact scan over the geometries
bad = 0
for (d, n, nq, k, metric) in [(384, 20000, 1, 10, "ip"), (384, 50001, 64, 100, "cosine"), (384, 40000, 130, 20, "l2"),
                              (384, 70000, 256, 10, "ip"), (384, 33000, 300, 10, "cosine"), (768, 33000, 64, 50, "ip"), (384, 16385, 128, 100, "cosine"),
                              (768, 20000, 5, 10, "l2"), (768, 30000, 128, 10, "cosine"), (1024, 20000, 40, 10, "ip"),
                              (1536, 20000, 9, 10, "ip"), (3072, 20000, 3, 10, "ip"), (100, 30000, 20, 10, "cosine"), (400, 17000, 130, 20, "l2")]:
    g = torch.Generator(device=dev).manual_seed(d + n + nq)
    xs = torch.randn(n, d, device=dev, generator=g)
    if metric != "l2":
        xs = torch.nn.functional.normalize(xs, dim=1)
    xs[n // 2] = xs[7]
    qs = torch.randn(nq, d, device=dev, generator=g)
    qs[0] = xs[7]
    ix = FlatIndex(d, metric)
    ix.add(xs)
    s0, i0 = ix.search(qs, k, mode=MODE_EXACT)
    s1, i1 = ix.search(qs, k, mode=MODE_AUTO, want_stats=True)
    ok = bool((i0 == i1).all() and (s0 == s1).all())
    log(f"d={d} n={n} nq={nq} k={k} {metric}: equal={ok} flagged={ix.last_stats[0]} launches={ix.last_stats[1]}")
    bad += 0 if ok else 1
    del ix, xs
out["geometry_mismatches"] = bad

# ---- 3. timings
def timed(ix, qs, k, iters=5):
    ts = []
    for _ in range(iters):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        ix.search(qs, k, mode=MODE_AUTO)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    _lib.profile_enable(True)
    _lib.profile_reset()
    for _ in range(5):
        ix.search(qs, k, mode=MODE_AUTO, want_stats=True)
    torch.cuda.synchronize()
    cls = {kk: round(v[0] / 5, 4) for kk, v in _lib.profile_read().items() if v[1]}
    _lib.profile_enable(False)
    return ts, cls, ix.last_stats


cfgs = [(10_000_000, 384, 64, 100, "cosine"), (10_000_000, 384, 64, 10, "cosine"), (10_000_000, 384, 128, 10, "ip"),
        (10_000_000, 384, 256, 10, "ip"), (1_000_000, 384, 64, 10, "cosine"), (1_250_000, 384, 256, 10, "cosine"),
        (1_250_000, 768, 64, 50, "cosine"), (5_000_000, 768, 64, 50, "cosine")]
if os.environ.get("SCAN_CHECK_SMALL") == "1":
    cfgs = cfgs[4:7]
cur = None
for (n, d, nq, k, metric) in cfgs:
    if cur is None or cur[0] != (n, d, metric):
        cur = None
        torch.cuda.empty_cache()
        ix = FlatIndex(d, metric)
        ix.reserve(n)
        for b in range(0, n, 1_000_000):
            ix.add(unit(min(1_000_000, n - b), d, 100 + b // 1_000_000))
        cur = ((n, d, metric), ix)
    ix = cur[1]
    qs = unit(nq, d, 5)
    ts, cls, st = timed(ix, qs, k)
    gbs = 4.0 * n * d / (min(ts) * 1e-3) / 1e9
    scan_gbs = 4.0 * n * d / (cls.get("scan", 1e9) * 1e-3) / 1e9
    log(f"N={n} D={d} Q={nq} k={k} {metric}: search ms {['%.3f' % t for t in ts]} -> {gbs:.0f} GB/s whole search; "
        f"classes {cls}; scan alone {scan_gbs:.0f} GB/s; flagged={st[0]}")
    out[f"time_{n}_{d}_{nq}_{k}"] = {"ms": ts, "classes": cls, "gbs_search": gbs, "gbs_scan": scan_gbs, "flagged": st[0]}
json.dump(out, open("gpurun_out/scan_check.json", "w"), indent=1)
log("done")
