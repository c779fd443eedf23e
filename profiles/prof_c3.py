"""Diagnostic: the C3 column-generation rounds with per-kernel event timing; prints every call slower than 30 ms."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dsa_b200 as D  # noqa: E402

L = D.lib()
PROF = os.environ.get("PROF", "1") == "1"


def prof(label, f):
    L.dsa_prof_reset()
    L.dsa_prof_enable(C.c_int(1 if PROF else 0))
    t0 = time.perf_counter()
    f()
    wall = time.perf_counter() - t0
    L.dsa_prof_enable(C.c_int(0))
    need = L.dsa_prof_dump(None, C.c_int64(0))
    buf = C.create_string_buffer(int(need) + 16)
    L.dsa_prof_dump(buf, C.c_int64(len(buf)))
    rows = [ln.split(",") for ln in buf.value.decode().strip().splitlines() if ln]
    rows = sorted(((n, int(c), float(ms)) for n, c, ms in rows), key=lambda r: -r[2])
    print(f"== {label}: wall {1e3 * wall:.2f} ms, kernels {sum(r[2] for r in rows):.2f} ms in {sum(r[1] for r in rows)} launches", flush=True)
    if wall > 0.015:
        for n, c, ms in rows[:9]:
            print(f"   {n:24s} x{c:<4d} {ms:9.3f} ms")


rng = np.random.default_rng(0xD5A00003)
m = 100_000


def cols(first):
    J = np.repeat(np.arange(first, first + 10_000), 50)
    return rng.integers(1, m + 1, len(J)), J, rng.random(len(J)) + 0.01


I, J, V = cols(1)
gm = D.dynamicsparse(I, J, V, m=m)
live = list(range(1, 10_001))
nxt = 10_001
for r in range(1, 10):
    I, J, V = cols(nxt)
    prof(f"round {r} append", lambda: gm.set_batch(I, J, V))
    live += list(range(nxt, nxt + 10_000))
    nxt += 10_000
    dead = rng.choice(np.array(live[:-1]), len(live) // 20, replace=False)
    prof(f"round {r} delete {len(dead)}", lambda: D.deletecolumn(gm, dead))
    ds = set(dead.tolist())
    live = [c for c in live if c not in ds]
    print("   info", {k: gm.info(0)[k] for k in ("capacity", "nb_elements", "nb_partitions")}, {k: gm.info(1)[k] for k in ("capacity", "nb_elements")})
