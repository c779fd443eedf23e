"""Config-2 style update steps with device-resident batches, for profiling (profiles/capture_ncu.sh) and for comparing kernel
switches (DSA_* environment variables are read once per process): prints the time per step and saves a digest of the final
layouts, so that two runs can be compared for bit-identical state.
Device memory comes from libcudart through ctypes (no torch import: keeps the run short).
usage: python profiles/run_c2_steps.py OUT.npz [m nnz batch steps]"""
import ctypes as C
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dsa_b200 as D  # noqa: E402
from dsa_b200._lib import check  # noqa: E402


def _cudart():
    for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so", "/usr/local/cuda/lib64/libcudart.so.12"):
        try:
            return C.CDLL(name)
        except OSError:
            continue
    raise RuntimeError("libcudart not found")


def main():
    out = sys.argv[1]
    m, nnz, batch, steps = (int(a) for a in sys.argv[2:6]) if len(sys.argv) >= 6 else (100_000, 10_000_000, 1_000_000, 10)
    D.require_gpu()
    rt = _cudart()
    L = D.lib()

    def to_dev(a):
        p = C.c_void_p()
        assert rt.cudaMalloc(C.byref(p), C.c_size_t(a.nbytes)) == 0
        assert rt.cudaMemcpy(p, a.ctypes.data_as(C.c_void_p), C.c_size_t(a.nbytes), C.c_int(1)) == 0
        return p

    rng = np.random.default_rng(77)
    A = D.dynamicsparse(rng.integers(1, m + 1, nnz), rng.integers(1, m + 1, nnz), rng.random(nnz) + 1e-3, m=m, n=m)
    warm = 3
    half = batch // 2
    pi, pj = rng.integers(1, m + 1, half), rng.integers(1, m + 1, half)
    dev = []
    for _ in range(warm + steps):   # 50 % inserts, 50 % deletes of the previous batch's inserts (stationary nnz), shuffled
        ii, jj, vv = rng.integers(1, m + 1, half), rng.integers(1, m + 1, half), rng.random(half) + 1e-3
        p = rng.permutation(2 * half)
        r, c, v = np.concatenate([ii, pi])[p], np.concatenate([jj, pj])[p], np.concatenate([vv, np.zeros(half)])[p]
        dev.append((to_dev(np.ascontiguousarray(r)), to_dev(np.ascontiguousarray(c)), to_dev(np.ascontiguousarray(v)), len(r)))
        pi, pj = ii, jj
    x = to_dev(rng.random(m))
    y = to_dev(np.zeros(m))

    def step(s, with_spmv):
        r, c, v, n = dev[s]
        check(L.dsa_matrix_set_batch_d(A._h, r, c, v, C.c_int64(n)))
        if with_spmv:
            check(L.dsa_matrix_spmv_dense_d(A._h, C.c_int(0), x, C.c_int64(m), y, C.c_int64(m)))

    for s in range(warm):
        step(s, True)
    rt.cudaDeviceSynchronize()
    profiling = os.environ.get("DSA_PROFILE_STEPS") == "1"   # under `ncu --profile-from-start off`: capture the timed steps only
    if profiling:
        rt.cudaProfilerStart()
    t0 = time.perf_counter()
    for s in range(warm, warm + steps):
        step(s, True)
    rt.cudaDeviceSynchronize()
    ms = (time.perf_counter() - t0) / steps * 1e3
    if profiling:
        rt.cudaProfilerStop()
        print(f"profiled {steps} steps")
        return
    digest = {}
    for which, name in ((0, "col"), (1, "row")):
        e = A.export(which)
        h = hashlib.sha256()
        for k in ("tag", "key", "val", "semaphores", "col_keys", "col_live"):
            h.update(np.ascontiguousarray(e[k]).tobytes())
        digest[name] = h.hexdigest()
    yh = np.zeros(m)
    assert rt.cudaMemcpy(yh.ctypes.data_as(C.c_void_p), y, C.c_size_t(yh.nbytes), C.c_int(2)) == 0
    np.savez(out, ms=ms, col=digest["col"], row=digest["row"], y=yh, nnz=D.nnz(A))
    sw = {k: v for k, v in sorted(os.environ.items()) if k.startswith("DSA_")}
    print(f"switches {sw}: col {digest['col'][:12]} row {digest['row'][:12]} "
          f"{ms:.3f} ms per step (wall clock, {batch} updates + SpMV) = {batch / ms / 1e3:.0f} Mupdates/s, nnz {D.nnz(A)}")


if __name__ == "__main__":
    main()
