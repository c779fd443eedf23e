"""Diagnostic: per-kernel time (CUDA events around every launch) of one C1 flush, one C3 round and one C5 batch."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dsa_b200 as D  # noqa: E402

L = D.lib()


def prof(label, f):
    f_warm = f
    L.dsa_prof_reset()
    L.dsa_prof_enable(C.c_int(1))
    t0 = time.perf_counter()
    f_warm()
    wall = time.perf_counter() - t0
    L.dsa_prof_enable(C.c_int(0))
    need = L.dsa_prof_dump(None, C.c_int64(0))
    buf = C.create_string_buffer(int(need) + 16)
    L.dsa_prof_dump(buf, C.c_int64(len(buf)))
    rows = [ln.split(",") for ln in buf.value.decode().strip().splitlines()]
    rows = sorted(((n, int(c), float(ms)) for n, c, ms in rows), key=lambda r: -r[2])
    print(f"== {label}: wall {1e3 * wall:.2f} ms, kernels {sum(r[2] for r in rows):.2f} ms in {sum(r[1] for r in rows)} launches")
    for n, c, ms in rows[:8]:
        print(f"   {n:24s} x{c:<4d} {ms:9.3f} ms")


rng = np.random.default_rng(1)
keys = np.unique(rng.integers(1, 10_000_000_000, 1_000_000))
gv = D.dynamicsparsevec(keys, rng.random(len(keys)) + 1)
for it in range(3):
    bk = np.concatenate([rng.integers(1, 10_000_000_000, 50_000), rng.choice(keys, 50_000, replace=False)])
    bv = np.concatenate([rng.random(50_000) + 1, np.zeros(50_000)])
    if it < 2:
        gv.set_batch(bk, bv)
    else:
        prof("C1 flush of 100k ops", lambda: gv.set_batch(bk, bv))

m = 100_000
def cols(first):
    J = np.repeat(np.arange(first, first + 10_000), 50)
    return rng.integers(1, m + 1, len(J)), J, rng.random(len(J)) + 0.01
I, J, V = cols(1)
gm = D.dynamicsparse(I, J, V, m=m)
nxt = 10_001
for it in range(3):
    I, J, V = cols(nxt)
    nxt += 10_000
    if it < 2:
        gm.set_batch(I, J, V)
    else:
        prof("C3 append 10k columns x 50", lambda: gm.set_batch(I, J, V))
dead = rng.choice(np.arange(1, nxt - 1), 1500, replace=False)
prof("C3 deletecolumn! x1500", lambda: D.deletecolumn(gm, dead))
