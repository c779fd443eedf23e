"""Diagnostic: per-kernel time of the C5 skewed batches (Zipf x Zipf, monotone hot columns) on the C2 matrix."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dsa_b200 as D  # noqa: E402
from prof_c3 import prof  # noqa: E402  (runs the C3 rounds on import: harmless warm-up)

rng = np.random.default_rng(0xD5A00005)
m = n = 100_000
nnz = 10_000_000
I, J = rng.integers(1, m + 1, nnz), rng.integers(1, n + 1, nnz)
V = rng.random(nnz) + 1e-3
gm = D.dynamicsparse(I, J, V, m=m, n=n)
w = 1.0 / np.arange(1, m + 1)
cdf = np.cumsum(w) / w.sum()
nb = 1_000_000
for it in range(2):
    I2, J2 = np.searchsorted(cdf, rng.random(nb)) + 1, np.searchsorted(cdf, rng.random(nb)) + 1
    V2 = rng.random(nb) + 1e-3
    prof(f"zipf batch {it}", lambda: gm.set_batch(I2, J2, V2))
    print("   info", {k: gm.info(0)[k] for k in ("capacity", "nb_elements")})
hot = rng.choice(n, 100, replace=False) + 1
base = m + 1
for it in range(2):
    I3 = np.concatenate([np.arange(base, base + 10_000) for _ in hot])
    base += 10_000
    J3 = np.repeat(hot, 10_000)
    V3 = rng.random(len(I3)) + 1e-3
    prof(f"monotone batch {it}", lambda: gm.set_batch(I3, J3, V3))
    print("   info", {k: gm.info(0)[k] for k in ("capacity", "nb_elements")})
