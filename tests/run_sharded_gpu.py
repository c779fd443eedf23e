"""Multi-GPU parity check of the sharded path (run under torchrun on >= 2 GPUs; not collected by pytest):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_sharded_gpu.py
Every rank replays the global batches on a replicated CPU oracle matrix and checks its own shard + the gathered SpMV.
DSA_DIST_PIPELINE=1 drives the batches through the background router (submit / apply_next, batch s+1 routed while batch s is
applied) instead of the synchronous set_batch — run it under `timeout`: that path has only passed the gloo test so far."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dsa_b200 as D  # noqa: E402
from dsa_b200.sharded import LibdsaBackend, ShardedMatrix  # noqa: E402
from oracle import oracle as O  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
D.lib().dsa_set_device(C.c_int(local))
import datetime  # noqa: E402
dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=60))
dev = torch.device("cuda", local)
PIPE = os.environ.get("DSA_DIST_PIPELINE", "0") == "1"
if PIPE:
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))   # the main work must not sit on the legacy default stream
m, n = 3000, 2600
A = ShardedMatrix(m, n, LibdsaBackend(dev))
G = O.Matrix(fill_mode=False)
rng = np.random.default_rng(5)
rounds = []
for rnd in range(5):
    nb = 40_000
    I, J = rng.integers(1, m + 1, nb), rng.integers(1, n + 1, nb)
    V = np.where(rng.random(nb) < 0.3, 0.0, rng.integers(1, 9, nb).astype(float))
    lin = I * (n + 1) + J
    _, first = np.unique(lin[::-1], return_index=True)     # keep the globally last write of every (i, j)
    keep = np.zeros(nb, bool)
    keep[nb - 1 - first] = True
    sel = np.nonzero(keep)[0]
    sel = sel[(sel >= rank * nb // world) & (sel < (rank + 1) * nb // world)]
    share = (torch.from_numpy(I[sel]).to(dev), torch.from_numpy(J[sel]).to(dev), torch.from_numpy(V[sel]).to(dev))
    rounds.append((I, J, V, share, rng.integers(0, 4, n).astype(float), rng.integers(0, 4, m).astype(float)))
if PIPE:
    A.submit(*rounds[0][3])
for rnd, (I, J, V, share, x, xt) in enumerate(rounds):
    G.set_batch_policy(I, J, V)
    if PIPE:
        if rnd + 1 < len(rounds):
            A.submit(*rounds[rnd + 1][3])
        A.apply_next()
    else:
        A.set_batch(*share)
    y = A.spmv(torch.from_numpy(x).to(dev)).cpu().numpy()
    assert np.array_equal(y, G.mul_dense(x, m)), f"rank {rank}: A*x differs in round {rnd}"
    yt = A.spmv(torch.from_numpy(xt).to(dev), trans=True).cpu().numpy()
    assert np.array_equal(yt, G.mul_dense(xt, n, trans=True)), f"rank {rank}: A'*x differs in round {rnd}"
infc, infr = A.local.info(0), A.local.info(1)
tot = torch.tensor([infc["nnz"], infr["nnz"]], dtype=torch.int64, device=dev)
dist.all_reduce(tot)
assert tot[0].item() == tot[1].item() == G.nnz(), (tot.tolist(), G.nnz())
if rank == 0:
    print(f"sharded parity ok on {world} GPUs ({'pipelined router' if PIPE else 'synchronous routing'}): nnz={G.nnz()}")
A.close()
dist.destroy_process_group()
