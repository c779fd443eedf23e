#!/usr/bin/env python
"""Secondary measurements for BASELINE.md §5 (not the bench.py contract line): BASELINE.json configs C1, C3 and C5 on one B200,
each next to the CPU oracle (C++ restatement of the reference, 1 core) on the same inputs.  Prints one JSON object per config.
    python profiles/bench_configs.py            (run under gpurun)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dsa_b200 as D  # noqa: E402
from oracle import oracle as O  # noqa: E402


def timed(f, reps=1):
    t0 = time.perf_counter()
    for _ in range(reps):
        r = f()
    return (time.perf_counter() - t0) / reps, r


def c1():
    rng = np.random.default_rng(0xD5A00001)
    keys = np.unique(rng.integers(1, 10_000_000_000, 1_000_000))
    vals = rng.integers(10, 100001, len(keys)) / 10.0
    t_build_g, gv = timed(lambda: D.dynamicsparsevec(keys, vals))
    t_build_o, ov = timed(lambda: O.Vec(keys, vals))
    nb, rounds = 100_000, 10
    tg = to = 0.0
    live = keys
    for r in range(rounds):
        ins_k = rng.integers(1, 10_000_000_000, nb // 2)
        del_k = rng.choice(live, nb // 2, replace=False)
        bk = np.concatenate([ins_k, del_k])
        bv = np.concatenate([rng.integers(10, 100001, nb // 2) / 10.0, np.zeros(nb // 2)])
        p = rng.permutation(nb)
        bk, bv = bk[p], bv[p]
        dt, _ = timed(lambda: gv.set_batch(bk, bv))
        tg += dt
        dt, _ = timed(lambda: ov.set_many(bk, bv))
        to += dt
        live = np.setdiff1d(np.union1d(live, ins_k), del_k)
    q = rng.choice(live, 1_000_000)
    t_get_g, a = timed(lambda: gv.get_batch(q))
    t_get_o, b = timed(lambda: ov.get_many(q))
    assert np.array_equal(a, b)
    return {"config": "C1 dynamicsparsevec PMA, 1M keys, 100k mixed ops per flush (host buffers, synchronous call)",
            "gpu_Mupdates_s": nb * rounds / tg / 1e6, "cpu_Mupdates_s": nb * rounds / to / 1e6,
            "gpu_build_s": t_build_g, "cpu_build_s": t_build_o, "gpu_Mfinds_s": 1.0 / t_get_g, "cpu_Mfinds_s": 1.0 / t_get_o}


def c3():
    """GPU rounds run back to back (an idle GPU drops to a low-power state and the next call pays the wake-up), then the CPU
    oracle replays the same pre-generated rounds."""
    rng = np.random.default_rng(0xD5A00003)
    m, cols_per_round, nnz_per_col, rounds = 100_000, 10_000, 50, 10

    def new_columns(first_id):
        J = np.repeat(np.arange(first_id, first_id + cols_per_round), nnz_per_col)
        I = rng.integers(1, m + 1, len(J))
        return I, J, rng.random(len(I)) + 0.01

    plan, live, nxt = [], [], 1
    for r in range(rounds):
        I, J, V = new_columns(nxt)
        live += list(range(nxt, nxt + cols_per_round))
        nxt += cols_per_round
        dead = rng.choice(np.array(live[:-1]), len(live) // 20, replace=False) if r > 0 else np.array([], np.int64)
        ds = set(dead.tolist())
        live = [c for c in live if c not in ds]
        plan.append((I, J, V, dead, rng.random(nxt), rng.random(m)))
    upd = sum(len(p[0]) + len(p[3]) * nnz_per_col for p in plan)

    def run(build, setb, delete, mul):
        t_upd = t_mul = 0.0
        ys = []
        M = None
        for r, (I, J, V, dead, x, pi) in enumerate(plan):
            t0 = time.perf_counter()
            if r == 0:
                M = build(I, J, V)
            else:
                setb(M, I, J, V)
                delete(M, dead)
            t_upd += time.perf_counter() - t0
            t0 = time.perf_counter()
            ys.append(mul(M, x, pi))
            t_mul += time.perf_counter() - t0
        return M, t_upd, t_mul, ys

    def g_mul(M, x, pi):
        mm, nn = M.size
        return M.mul_dense(x[:nn]), M.mul_dense(pi[:mm], trans=True)

    def o_mul(M, x, pi):
        mm, nn = M.size
        return M.mul_dense(x[:nn], mm), M.mul_dense(pi[:mm], nn, trans=True)

    def o_del(M, dead):
        for c in dead:
            M.deletecolumn(int(c))

    D.dynamicsparse([1], [1], [1.0])   # context warm-up
    # first pass: warms the library's block pool (a cold process pays ~65 cudaMalloc calls of ~1.4 ms each: 0.7 s for the 10
    # rounds against 0.2 s warm); the reported pass is the second one
    cold = run(lambda I, J, V: D.dynamicsparse(I, J, V, m=m), lambda M, I, J, V: M.set_batch(I, J, V),
               lambda M, d: D.deletecolumn(M, d) if len(d) else None, g_mul)
    t_cold = cold[1] + cold[2]
    del cold
    gm, tg, tgm, yg = run(lambda I, J, V: D.dynamicsparse(I, J, V, m=m), lambda M, I, J, V: M.set_batch(I, J, V),
                          lambda M, d: D.deletecolumn(M, d) if len(d) else None, g_mul)
    om, to, tom, yo = run(lambda I, J, V: O.Matrix(I, J, V, m=m), lambda M, I, J, V: M.set_many(I, J, V), o_del, o_mul)
    for (a1, a2), (b1, b2) in zip(yg, yo):
        assert np.allclose(a1, b1, rtol=1e-12) and np.allclose(a2, b2, rtol=1e-12)
    return {"config": "C3 column generation: 10 rounds x (append 10k cols x 50 nnz, deletecolumn! 5%, A*x and A'*pi), host buffers",
            "gpu_Mupdates_s": upd / tg / 1e6, "cpu_Mupdates_s": upd / to / 1e6, "gpu_total_s": tg + tgm, "gpu_total_s_cold_process": t_cold, "cpu_total_s": to + tom,
            "gpu_spmv_pair_ms": 1e3 * tgm / rounds, "cpu_spmv_pair_ms": 1e3 * tom / rounds, "live_columns": len(live), "nnz": D.nnz(gm)}


def _prof(f):
    """per-kernel CUDA-event times of one call (name -> (launches, ms))"""
    import ctypes as C
    L = D.lib()
    L.dsa_prof_reset()
    L.dsa_prof_enable(C.c_int(1))
    f()
    L.dsa_prof_enable(C.c_int(0))
    need = L.dsa_prof_dump(None, C.c_int64(0))
    buf = C.create_string_buffer(int(need) + 16)
    L.dsa_prof_dump(buf, C.c_int64(len(buf)))
    out = {}
    for ln in buf.value.decode().strip().splitlines():
        name, cnt, ms = ln.split(",")
        out[name] = (int(cnt), round(float(ms), 3))
    return dict(sorted(out.items(), key=lambda kv: -kv[1][1]))


def c5():
    """Skewed inserts on the C2 matrix.  Every regime runs several batches back to back (the first one pays the growth of the
    workspace buffers and is reported separately); the last batch of each regime is also run under the per-kernel profile."""
    rng = np.random.default_rng(0xD5A00005)
    m = n = 100_000
    nnz = 10_000_000
    I, J = rng.integers(1, m + 1, nnz), rng.integers(1, n + 1, nnz)
    V = rng.random(nnz) + 1e-3
    gm = D.dynamicsparse(I, J, V, m=m, n=n)
    w = 1.0 / np.arange(1, m + 1)
    cdf = np.cumsum(w) / w.sum()
    nb = 1_000_000

    def zipf():
        return np.searchsorted(cdf, rng.random(nb)) + 1, np.searchsorted(cdf, rng.random(nb)) + 1, rng.random(nb) + 1e-3

    import ctypes as C
    import torch
    dev = torch.device("cuda", 0)
    L = D.lib()

    def dev_batches(bs):
        return [tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in b) for b in bs]

    def set_d(M, b):   # device-resident batch through the _d entry point, timed to completion
        D._lib.check(L.dsa_matrix_set_batch_d(M._h, C.c_void_p(b[0].data_ptr()), C.c_void_p(b[1].data_ptr()), C.c_void_p(b[2].data_ptr()),
                                              C.c_int64(b[0].numel())))
        torch.cuda.synchronize()

    zb = [zipf() for _ in range(5)]
    tz = [timed(lambda b=b: gm.set_batch(*b))[0] for b in zb[:4]]
    prof_z = _prof(lambda: gm.set_batch(*zb[4]))
    gz = D.dynamicsparse(I, J, V, m=m, n=n)      # the same regime with device-resident batches on a fresh matrix
    torch.cuda.synchronize()
    tzd = [timed(lambda b=b: set_d(gz, b))[0] for b in dev_batches(zb)]
    del gz
    hot = rng.choice(n, 100, replace=False) + 1
    mono, base = [], m + 1
    for _ in range(4):
        I3 = np.concatenate([np.arange(base, base + 10_000) for _ in hot])
        mono.append((I3, np.repeat(hot, 10_000), rng.random(len(I3)) + 1e-3))
        base += 10_000
    tm = [timed(lambda b=b: gm.set_batch(*b))[0] for b in mono[:3]]
    prof_m = _prof(lambda: gm.set_batch(*mono[3]))
    gmo = D.dynamicsparse(I, J, V, m=m, n=n)
    torch.cuda.synchronize()
    tmd = [timed(lambda b=b: set_d(gmo, b))[0] for b in dev_batches(mono)]
    cap_mono = gmo.info(0)["capacity"]
    del gmo
    print(f"C5 zipf batches: {[round(1e3 * t, 2) for t in tz]} ms; monotone: {[round(1e3 * t, 2) for t in tm]} ms", file=sys.stderr)
    sample = 200_000
    om = O.Matrix(I, J, V, m=m, n=n)
    to, _ = timed(lambda: om.set_many(zb[0][0][:sample], zb[0][1][:sample], zb[0][2][:sample]))
    to2, _ = timed(lambda: om.set_many(mono[0][0][:sample], mono[0][1][:sample], mono[0][2][:sample]))
    return {"config": "C5 skew on the C2 matrix: 1M Zipf(1.0) x Zipf(1.0) inserts per batch; 1M monotone inserts into 100 hot columns per batch "
                      "(host buffers, synchronous call)",
            "gpu_zipf_Mupdates_s_first": nb / tz[0] / 1e6, "gpu_zipf_Mupdates_s_warm": nb / min(tz[1:]) / 1e6,
            "gpu_zipf_ms": [1e3 * t for t in tz], "gpu_zipf_device_resident_ms": [1e3 * t for t in tzd],
            "gpu_zipf_device_resident_Mupdates_s_warm": nb / float(np.median(tzd[1:])) / 1e6, "cpu_zipf_Mupdates_s": sample / to / 1e6,
            "gpu_monotone_Mupdates_s_first": nb / tm[0] / 1e6, "gpu_monotone_Mupdates_s_warm": nb / min(tm[1:]) / 1e6,
            "gpu_monotone_ms": [1e3 * t for t in tm], "gpu_monotone_device_resident_ms": [1e3 * t for t in tmd],
            "gpu_monotone_device_resident_Mupdates_s_warm": nb / float(np.median(tmd[1:])) / 1e6, "capacity_after_monotone_fresh": cap_mono,
            "cpu_monotone_Mupdates_s": sample / to2 / 1e6, "cpu_sample": sample,
            "capacity_after": gm.info(0)["capacity"], "kernels_zipf": prof_z, "kernels_monotone": prof_m}


if __name__ == "__main__":
    D.require_gpu()
    only = sys.argv[1:] or ["c1", "c3", "c5"]
    for f in [g for g in (c1, c3, c5) if g.__name__ in only]:
        print(json.dumps(f()), flush=True)
