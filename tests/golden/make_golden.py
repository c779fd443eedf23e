"""Generates tests/golden/*.npz: small seeded inputs and the layouts the CPU oracle produces for them
(bulk build = reference layout; batches = oracle batch policy).  Run from the repo root:
    python tests/golden/make_golden.py
The fixtures are committed so the GPU tests do not depend on a regenerated oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

out = os.path.dirname(os.path.abspath(__file__))


def vec_case(name, seed, n0, batches):
    rng = np.random.default_rng(seed)
    I = rng.integers(1, 5000, n0)
    V = rng.integers(1, 64, n0) / 4.0
    v = O.Vec(I, V)
    d = dict(kind="vec", I=I, V=V, nbatches=len(batches))
    for b, nb in enumerate(batches):
        bk = rng.integers(1, 6000, nb)
        bv = np.where(rng.random(nb) < 0.4, 0.0, rng.integers(1, 64, nb) / 4.0)
        v.set_batch_policy(bk, bv)
        d[f"bk{b}"], d[f"bv{b}"] = bk, bv
    tag, key, val = v.export()
    np.savez_compressed(os.path.join(out, name), tag=tag, key=key, val=val, **d)


def mat_case(name, seed, m, n, nnz, batches, ndel):
    rng = np.random.default_rng(seed)
    I, J = rng.integers(1, m + 1, nnz), rng.integers(1, n + 1, nnz)
    V = rng.integers(1, 64, nnz) / 4.0
    M = O.Matrix(I, J, V)
    d = dict(kind="mat", I=I, J=J, V=V, nbatches=len(batches))
    for b, nb in enumerate(batches):
        bi, bj = rng.integers(1, m + 20, nb), rng.integers(1, n + 20, nb)
        bv = np.where(rng.random(nb) < 0.35, 0.0, rng.integers(1, 64, nb) / 4.0)
        M.set_batch_policy(bi, bj, bv)
        d[f"bi{b}"], d[f"bj{b}"], d[f"bv{b}"] = bi, bj, bv
    e = M.export(0)
    live = e["col_keys"][e["col_live"] == 1]
    delcols = rng.choice(live[:-1], ndel, replace=False) if ndel else np.array([], np.int64)
    if ndel:
        M.delete_columns_policy(delcols)
    d["delcols"] = delcols
    for which in (0, 1):
        e = M.export(which)
        d[f"tag{which}"], d[f"key{which}"], d[f"val{which}"] = e["tag"], e["key"], e["val"]
        d[f"sem{which}"], d[f"live{which}"] = e["semaphores"] * e["col_live"], e["col_live"]
    mm, nn = M.size
    x = rng.integers(0, 8, nn) / 2.0
    d["x"], d["y"] = x, M.mul_dense(x, mm)
    np.savez_compressed(os.path.join(out, name), **d)


vec_case("vec_small.npz", 1, 300, [50, 400, 30])
vec_case("vec_grow.npz", 2, 10, [2000, 3000])
mat_case("mat_small.npz", 3, 40, 60, 500, [100, 300], 5)
mat_case("mat_wide.npz", 4, 20, 900, 4000, [1500], 40)
mat_case("mat_build_only.npz", 5, 200, 150, 6000, [], 0)
print("ok")
