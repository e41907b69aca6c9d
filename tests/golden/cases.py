"""Seeded inputs shared by make_golden.py and the tests (numpy Generator streams are stable across
platforms, so only the expected OUTPUTS need to be stored in the fixtures)."""
import numpy as np

FLAT_CASES = {  # name: (rows, dim, queries, k)
    "a": (1000, 384, 5, 10),
    "b": (7, 64, 3, 10),
    "c": (4099, 768, 4, 50),
    "d": (1, 32, 2, 3),
    "e": (20000, 384, 6, 100),
}


def flat_case(name):
    n, d, nq, k = FLAT_CASES[name]
    rng = np.random.default_rng({"a": 11, "b": 12, "c": 13, "d": 14, "e": 15}[name])
    x = rng.standard_normal((n, d)).astype(np.float32)
    q = rng.standard_normal((nq, d)).astype(np.float32)
    if n >= 1000:
        x[10] = x[3]
        x[500] = x[3]                      # exact duplicates -> tied scores, pins the tie-break
        q[0] = x[3] + 0.01 * rng.standard_normal(d).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return np.ascontiguousarray(x, dtype=np.float32), np.ascontiguousarray(q, dtype=np.float32), k
