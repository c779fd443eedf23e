"""Diagnostic: how much of the SpMV time is the x gather?  Times dsa_matrix_spmv_dense_d on the config-2 matrix with the
real x, with nx = 0 (no gathers at all: pure stream of the gapped array) and with a tiny x (all gathers hit one line)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
import dsa_b200 as D  # noqa: E402

coo, _, x = B.make_workload(0)
A = D.dynamicsparse(coo[0], coo[1], coo[2], m=B.M_ROWS, n=B.N_COLS)
L = D.lib()
st = torch.cuda.current_stream()
L.dsa_matrix_set_stream(A._h, C.c_void_p(st.cuda_stream))
dx = torch.from_numpy(x).cuda()
dy = torch.zeros(B.M_ROWS, dtype=torch.float64, device="cuda")


def run(nx, reps=20):
    for _ in range(3):
        L.dsa_matrix_spmv_dense_d(A._h, C.c_int(0), C.c_void_p(dx.data_ptr()), C.c_int64(nx), C.c_void_p(dy.data_ptr()), C.c_int64(B.M_ROWS))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        L.dsa_matrix_spmv_dense_d(A._h, C.c_int(0), C.c_void_p(dx.data_ptr()), C.c_int64(nx), C.c_void_p(dy.data_ptr()), C.c_int64(B.M_ROWS))
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


for nx, label in ((B.N_COLS, "full x (1e5 doubles)"), (0, "no gather"), (16, "x of 16 doubles (gathers for keys <= 16 only)")):
    print(f"{label:50s} {run(nx):8.1f} us per spmv call (memset + flat + fixup + to_dense)")
a = torch.empty(1 << 25, dtype=torch.float64, device="cuda")
b = torch.empty_like(a)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    s = a.sum()
e1.record()
torch.cuda.synchronize()
print(f"torch sum of 268 MB: {1e3 * e0.elapsed_time(e1) / 10:.1f} us  -> {a.numel() * 8 / (e0.elapsed_time(e1) / 10 * 1e-3) / 1e9:.0f} GB/s read")
