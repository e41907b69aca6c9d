"""GPU diagnostic for the flat index (run under gpurun): accumulator dump vs numpy, exact path vs
oracle, tensor scan vs exact scan, and first timings.  Writes gpurun_out/check_index.json."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ragmeup_b200 import _lib  # noqa: E402
from ragmeup_b200.index import FlatIndex, MODE_AUTO, MODE_EXACT, MODE_TENSOR_NOFALLBACK  # noqa: E402
from oracle import flat_ref  # noqa: E402

out = {}
dev = torch.device("cuda")
torch.manual_seed(0)


def log(*a):
    print(*a, flush=True)


def unit(n, d, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(n, d, generator=g, device=dev, dtype=torch.float32)
    return torch.nn.functional.normalize(x, dim=1)


# ---- 1. raw accumulators of one tile
D = 384
x = unit(20000, D, 1)
q = unit(128, D, 2)
idx = FlatIndex(D, "ip")
idx.add(x)
q = q[:64].contiguous()
dbg = torch.zeros(64, 256, device=dev)                      # rmu_debug_scan_tile: out [nq <= 64, 256] = the first 256 corpus rows
_lib.check(_lib.lib().rmu_debug_scan_tile(idx._h, q.data_ptr(), 64, dbg.data_ptr(), _lib.stream_ptr()), "debug tile")
torch.cuda.synchronize()
ref = (q.double() @ x[:256].double().T).float()
err = (dbg - ref).abs().max().item()
log("tile dump max|err| vs fp64:", err, " ref absmax:", ref.abs().max().item())
out["tile_max_err"] = err
if err > 5e-3:
    log("dbg[0,:8]", dbg[0, :8].tolist())
    log("ref[0,:8]", ref[0, :8].tolist())

# ---- 2. exact path vs oracle (small)
for metric in ("ip", "cosine", "l2"):
    xs = torch.randn(3000, D, device=dev) * (1.0 if metric != "cosine" else 3.0)
    qs = torch.randn(7, D, device=dev)
    ix = FlatIndex(D, metric)
    ix.add(xs)
    s, i = ix.search(qs, 10, mode=MODE_EXACT)
    torch.cuda.synchronize()
    rs, ri = flat_ref.flat_search(qs.cpu().numpy(), xs.cpu().numpy(), 10, metric)
    ok_ids = bool((i.cpu().numpy() == ri).all())
    ds = float(np.abs(s.cpu().numpy() - rs).max())
    log(f"exact[{metric}] ids equal: {ok_ids}  max|dscore|: {ds:.2e}")
    out[f"exact_{metric}"] = [ok_ids, ds]

# ---- 3. tensor scan vs exact scan, 200k rows
for metric in ("ip", "cosine", "l2"):
    n = 200_000
    xs = unit(n, D, 11) if metric != "l2" else torch.randn(n, D, device=dev) * 0.3
    qs = unit(64, D, 12)
    ix = FlatIndex(D, metric)
    ix.add(xs)
    s0, i0 = ix.search(qs, 10, mode=MODE_EXACT)
    s1, i1 = ix.search(qs, 10, mode=MODE_TENSOR_NOFALLBACK, want_stats=True)
    st1 = ix.last_stats
    s2, i2 = ix.search(qs, 10, mode=MODE_AUTO, want_stats=True)
    st2 = ix.last_stats
    torch.cuda.synchronize()
    log(f"tensor[{metric}] nofallback ids equal exact: {bool((i0 == i1).all())} flagged={st1}; "
        f"auto ids equal: {bool((i0 == i2).all())} scores equal: {bool((s0 == s2).all())} flagged={st2}")
    out[f"tensor_{metric}"] = [bool((i0 == i1).all()), list(st1), bool((i0 == i2).all()), bool((s0 == s2).all())]
    if not bool((i0 == i1).all()):
        bad = (i0 != i1).any(1).nonzero().view(-1)[:3].tolist()
        for b in bad:
            log("  q", b, "exact", i0[b].tolist(), "tensor", i1[b].tolist())


# ---- 4. timings
def time_search(ix, qs, k, mode, iters=10):
    for _ in range(3):
        ix.search(qs, k, mode=mode)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for it in range(iters):
        ix.search(qs, k, mode=mode)
        ev[it + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
    return float(np.median(ts)), float(np.min(ts))


for n, nq in ((1_000_000, 64), (1_000_000, 1), (10_000_000, 128), (10_000_000, 256)):
    try:
        xs = None
        ix = FlatIndex(D, "ip")
        ix.reserve(n)
        for b in range(0, n, 1_000_000):
            ix.add(unit(min(1_000_000, n - b), D, 100 + b // 1_000_000))
        qs = unit(nq, D, 5)
        med, mn = time_search(ix, qs, 10, MODE_AUTO)
        gbs = 4.0 * n * D / (med * 1e-3) / 1e9
        ix.search(qs, 10, mode=MODE_AUTO, want_stats=True)
        log(f"N={n} Q={nq} k=10 auto: median {med:.3f} ms (min {mn:.3f})  -> {gbs:.0f} GB/s algorithmic "
            f"({ix.last_stats[1]} scan launches, {ix.last_stats[0]} flagged), {nq / med * 1e3:.0f} q/s")
        out[f"time_{n}_{nq}"] = [med, mn, gbs, list(ix.last_stats)]
        if n == 1_000_000 and nq == 64:
            s0, i0 = ix.search(qs, 10, mode=MODE_EXACT)
            s2, i2 = ix.search(qs, 10, mode=MODE_AUTO)
            torch.cuda.synchronize()
            log("   1M: auto == exact ids:", bool((i0 == i2).all()), "scores:", bool((s0 == s2).all()))
            t0 = time.time()
            ix.search(qs, 10, mode=MODE_EXACT)
            torch.cuda.synchronize()
            log(f"   1M exact path: {(time.time() - t0) * 1e3:.1f} ms")
        del ix
        torch.cuda.empty_cache()
    except Exception as e:  # keep going, report
        log("timing failed", n, nq, repr(e))
        out[f"time_{n}_{nq}"] = repr(e)

os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/check_index.json", "w") as f:
    json.dump(out, f, indent=1)
log("launches:", _lib.launch_count())
