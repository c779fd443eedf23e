"""Per-kernel event times of configs[1] as written (one batched insert of 1M new entries into the fresh config-2 matrix + SpMV).
usage (GPU box): python profiles/prof_insert_only.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
import dsa_b200 as D  # noqa: E402

L = D.lib()
dev = torch.device("cuda:0")
x, coo, batches, insert_only = B.make_workload(1)
A = D.dynamicsparse(coo[0], coo[1], coo[2], m=B.M_ROWS, n=B.N_COLS)
d_io = tuple(torch.from_numpy(a).to(dev) for a in insert_only)
d_x = torch.from_numpy(x).to(dev)
d_y = torch.zeros(B.M_ROWS, dtype=torch.float64, device=dev)
vp = lambda t: C.c_void_p(t.data_ptr())
for rep in range(4):
    h = C.c_void_p()
    D._lib.check(L.dsa_matrix_clone(A._h, C.byref(h)))
    torch.cuda.synchronize()
    if rep == 3:
        L.dsa_prof_reset()
        L.dsa_prof_enable(C.c_int(1))
    D._lib.check(L.dsa_matrix_set_batch_d(h, vp(d_io[0]), vp(d_io[1]), vp(d_io[2]), C.c_int64(B.BATCH)))
    D._lib.check(L.dsa_matrix_spmv_dense_d(h, C.c_int(0), vp(d_x), C.c_int64(B.N_COLS), vp(d_y), C.c_int64(B.M_ROWS)))
    torch.cuda.synchronize()
    if rep == 3:
        L.dsa_prof_enable(C.c_int(0))
        need = L.dsa_prof_dump(None, C.c_int64(0))
        buf = C.create_string_buffer(int(need) + 16)
        L.dsa_prof_dump(buf, C.c_int64(len(buf)))
        tot = 0.0
        for ln in sorted(buf.value.decode().strip().splitlines(), key=lambda l: -float(l.split(",")[2])):
            name, cnt, ms = ln.split(",")
            tot += float(ms)
            print(f"{name:24s} x{cnt:>3s}  {1e3 * float(ms):8.1f} us")
        print(f"sum {1e3 * tot:.1f} us")
    L.dsa_matrix_destroy(h)
