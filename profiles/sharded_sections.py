"""Diagnostic (torchrun, >= 2 GPUs): wall time of each section of ShardedMatrix.set_batch / spmv, averaged over steps."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
import bench_dist as BD  # noqa: E402
import dsa_b200 as D  # noqa: E402
from dsa_b200 import sharded as S  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
D.lib().dsa_set_device(C.c_int(local))
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
per = B.M_ROWS
m = n = per * world
nnzb = B.NNZ0 // world
A = S.ShardedMatrix(m, n, S.LibdsaBackend(dev))
cI, cJ, cV = (np.concatenate(x) for x in zip(*[BD._block(B.SEED, a, rank, per, per, nnzb) for a in range(world)]))
A.local.build(0, cI, cJ, cV)
rI, rJ, rV = (np.concatenate(x) for x in zip(*[BD._block(B.SEED, rank, b, per, per, nnzb) for b in range(world)]))
A.local.build(1, rJ, rI, rV)
rng = np.random.default_rng([1, rank])
T = {}


def tick(name, t0):
    torch.cuda.synchronize()
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
    return time.perf_counter()


steps = 12
x = torch.from_numpy(np.random.default_rng(3).random(n)).to(dev)
for s in range(steps):
    rows = torch.from_numpy(rng.integers(1, m + 1, B.BATCH)).to(dev)
    cols = torch.from_numpy(rng.integers(1, n + 1, B.BATCH)).to(dev)
    vals = torch.from_numpy(rng.random(B.BATCH) + 0.01).to(dev)
    if s == 2:
        T.clear()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    W = world
    pk_c = torch.empty((B.BATCH, 3), dtype=torch.int64, device=dev)
    pk_r = torch.empty((B.BATCH, 3), dtype=torch.int64, device=dev)
    cc, cr = np.zeros(W, np.int64), np.zeros(W, np.int64)
    isc = np.asarray(A.col_split[1:-1], dtype=np.int64)
    isr = np.asarray(A.row_split[1:-1], dtype=np.int64)
    S.check(S.lib().dsa_route_batch2_d(C.c_void_p(rows.data_ptr()), C.c_void_p(cols.data_ptr()), C.c_void_p(vals.data_ptr()), C.c_int64(B.BATCH),
                                       C.c_void_p(isc.ctypes.data), C.c_void_p(isr.ctypes.data), C.c_int(W), C.c_void_p(pk_c.data_ptr()),
                                       C.c_void_p(pk_r.data_ptr()), C.c_void_p(cc.ctypes.data), C.c_void_p(cr.ctypes.data),
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    t0 = tick("1 route2", t0)
    sc = torch.tensor([v for p in zip(cc.tolist(), cr.tolist()) for v in p], dtype=torch.int64, device=dev)
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc)
    rcl = rc.view(W, 2).tolist()
    t0 = tick("2 counts a2a", t0)
    out = []
    for (packed, snd, rcv) in ((pk_c, cc.tolist(), [p[0] for p in rcl]), (pk_r, cr.tolist(), [p[1] for p in rcl])):
        recv = torch.empty((int(sum(rcv)), 3), dtype=torch.int64, device=dev)
        dist.all_to_all_single(recv, packed, output_split_sizes=rcv, input_split_sizes=snd)
        t0 = tick("3 data a2a", t0)
        c3 = recv.t().contiguous()
        out.append((c3[0], c3[1], c3[2].view(torch.float64)))
        t0 = tick("4 unpack", t0)
    A.local.set_batch_two(out[0], out[1])
    t0 = tick("5 set_batch_two", t0)
    y = A.spmv(x)
    t0 = tick("6 spmv + all_gather", t0)
if rank == 0:
    tot = sum(T.values())
    for k in sorted(T):
        print(f"{k:24s} {1e6 * T[k] / (steps - 2):8.1f} us/step")
    print(f"{'total':24s} {1e6 * tot / (steps - 2):8.1f} us/step (sections individually synchronised)")
dist.destroy_process_group()
