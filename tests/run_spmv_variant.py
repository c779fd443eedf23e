"""Helper of tests/test_zz_experimental.py (not collected): computes A*x and transpose(A)*x at a given size with whatever
SpMV variant the environment selects (DSA_SPMV_BULK / DSA_SPMV_STEPS are read once per process), after a few update batches
so that the array holds gaps, tombstones and rows that straddle chunks, and saves the result bits.
usage: python tests/run_spmv_variant.py OUT.npz [m n nnz]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dsa_b200 as D  # noqa: E402


def main():
    out = sys.argv[1]
    m, n, nnz = (int(a) for a in sys.argv[2:5]) if len(sys.argv) >= 5 else (20_000, 30_000, 2_000_000)
    D.require_gpu()
    rng = np.random.default_rng(2024)
    J = rng.integers(1, n + 1, nnz)
    A = D.dynamicsparse(rng.integers(1, m + 1, nnz), J, rng.random(nnz) + 0.5, m=m, n=n)
    for _ in range(3):
        nb = max(nnz // 20, 1)
        A.set_batch(rng.integers(1, m + 1, nb), rng.integers(1, n + 1, nb), np.where(rng.random(nb) < 0.4, 0.0, rng.random(nb)))
    D.deletecolumn(A, np.unique(J[:40]).tolist())     # tombstones: the deleted partitions' slots stay in the column map
    x, xt = rng.random(n), rng.random(m)
    y, yt = A.mul_dense(x), A.mul_dense(xt, trans=True)
    # kernel time from the library's per-launch CUDA events (prof mode), after warm-up
    import ctypes as C
    L = D.lib()
    for _ in range(3):
        A.mul_dense(x)
    L.dsa_prof_reset()
    L.dsa_prof_enable(C.c_int(1))
    t0 = time.perf_counter()
    for _ in range(20):
        A.mul_dense(x)
    ms = (time.perf_counter() - t0) / 20 * 1e3
    L.dsa_prof_enable(C.c_int(0))
    need = L.dsa_prof_dump(None, C.c_int64(0))
    buf = C.create_string_buffer(int(need) + 16)
    L.dsa_prof_dump(buf, C.c_int64(len(buf)))
    kern = {ln.split(",")[0]: 1e3 * float(ln.split(",")[2]) / int(ln.split(",")[1]) for ln in buf.value.decode().strip().splitlines()}
    xs_k = np.sort(rng.choice(np.arange(1, n + 1), n // 3, replace=False))
    ys = A @ (xs_k, rng.random(len(xs_k)))            # sparse x (mask path)
    np.savez(out, y=y, yt=yt, ys_k=ys.nzind, ys_v=ys.nzval, ms=ms, cap=A.info(1)["capacity"])
    print(f"variant BULK={os.environ.get('DSA_SPMV_BULK', '0')} STEPS={os.environ.get('DSA_SPMV_STEPS', '4')}: capacity {A.info(1)['capacity']}, "
          f"{ms:.3f} ms per host-pointer mul_dense; kernel us per launch: "
          + ", ".join(f"{k}={v:.1f}" for k, v in sorted(kern.items()) if k.startswith("spmv")))


if __name__ == "__main__":
    main()
