"""Multi-GPU parity worker of the sharded path (dsa_dmatrix_* through the C ABI), one process per GPU under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/run_sharded_gpu.py
tests/test_gpu_sharded.py launches it on min(device_count, 8) GPUs (N = 1 exercises the same routing / unpack / device-side
count path without NCCL).

Every rank holds the whole global input and checks ITS shards against CPU oracle matrices restricted to its key ranges:
  * routed bulk build: layout bit-exact (the reference's bulk layout) per shard and orientation
  * routed batches whose shares overlap in (i, j) ACROSS ranks — no de-duplication: the global op order is rank-major, arrival
    within a rank, and the last writer in that order must win — layout bit-exact against the oracle's batch policy, contents
    against the reference's sequential loop
  * A*x and transpose(A)*x against the replicated global oracle (integer-valued data: exact)
  * routed getindex, routed deletecolumn! / deleterow! (incl. the collective error path), nnz / partition totals
"""
import ctypes as C
import datetime
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dsa_b200 as D  # noqa: E402
from dsa_b200.sharded import DistContext, DistMatrix, even_splitters, owner_of  # noqa: E402
from oracle import oracle as O  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
D.lib().dsa_set_device(C.c_int(local))
dev = torch.device("cuda", local)
dist.init_process_group("nccl" if world > 1 else "gloo", device_id=dev if world > 1 else None, timeout=datetime.timedelta(seconds=120))
ctx = DistContext()


def shard_equal(g, o, which, what):
    for f in ("capacity", "segment_capacity", "nb_segments", "nb_elements", "height", "nb_partitions"):
        assert g[f] == o[f], (what, which, f, g[f], o[f])
    assert np.array_equal(g["tag"], o["tag"]), (what, which, "gap pattern differs")
    mk = o["tag"].astype(bool)
    assert np.array_equal(g["key"][mk], o["key"][mk]) and np.array_equal(g["val"][mk], o["val"][mk]), (what, which, "cells differ")
    lv = o["col_live"].astype(bool)
    assert np.array_equal(g["col_live"], o["col_live"]) and np.array_equal(g["col_keys"][lv], o["col_keys"][lv]), (what, which, "column map")
    assert np.array_equal(g["semaphores"][lv], o["semaphores"][lv]), (what, which, "semaphores")


def contents(e):
    mk = e["tag"].astype(bool)
    return e["key"][mk], e["val"][mk]


def run(m, n, nnz0, nb, nrounds, max_share, uneven, seed):
    rng = np.random.default_rng(seed)   # same stream on every rank: everybody knows the whole global input
    if uneven:   # unequal shards (what sampled splitters produce): exercises the padded SpMV gather
        def cuts(dim):
            c = np.sort(rng.choice(np.arange(2, dim), world - 1, replace=False)) if world > 1 else np.zeros(0, np.int64)
            return [1] + c.tolist() + [dim + 1]
        rs, cs = cuts(m), cuts(n)
    else:
        rs, cs = even_splitters(m, world), even_splitters(n, world)
    A = DistMatrix(ctx, m, n, max_share, row_split=rs, col_split=cs)
    I, J = rng.integers(1, m + 1, nnz0), rng.integers(1, n + 1, nnz0)
    V = rng.integers(1, 9, nnz0).astype(float) / 4.0
    # shares: contiguous slices of the global COO -> global fold order of duplicates = the plain input order
    lo, hi = rank * nnz0 // world, (rank + 1) * nnz0 // world
    A.build_coo(I[lo:hi], J[lo:hi], V[lo:hi])
    G = O.Matrix(I, J, V, m=m, n=n)                               # replicated global checker
    mc = owner_of(J, cs) == rank
    mr = owner_of(I, rs) == rank
    Gc = O.Matrix(I[mc], J[mc], V[mc], m=m, n=n)                  # this rank's columns  -> its column-major shard
    Gr = O.Matrix(I[mr], J[mr], V[mr], m=m, n=n)                  # this rank's rows     -> its row-major shard
    Sc, Sr = Gc.clone(), Gr.clone()                               # the same, driven by the reference's sequential loop
    L = A.local
    shard_equal(L.export(0), Gc.export(0), 0, "build")
    shard_equal(L.export(1), Gr.export(1), 1, "build")
    for rnd in range(nrounds):
        I2, J2 = rng.integers(1, m + 1, nb), rng.integers(1, n + 1, nb)
        # a narrow hot range so that the SAME (i, j) is written from several ranks in one batch
        hot = rng.random(nb) < 0.3
        I2[hot], J2[hot] = rng.integers(1, 12, hot.sum()), rng.integers(1, 9, hot.sum())
        V2 = np.where(rng.random(nb) < 0.3, 0.0, rng.integers(1, 9, nb).astype(float))
        if rnd == 2:   # keys beyond 32 bits: the shares that hold them travel as 24-byte triples, the others stay packed (16 bytes)
            J2[5], I2[nb // 2 + 9], V2[5], V2[nb // 2 + 9] = (1 << 40) + 7, (1 << 35) + 3, 2.0, 3.0
        lo, hi = rank * nb // world, (rank + 1) * nb // world
        if rnd % 3 == 0:   # device-resident share
            A.set_batch(torch.from_numpy(I2[lo:hi].copy()).to(dev), torch.from_numpy(J2[lo:hi].copy()).to(dev),
                        torch.from_numpy(V2[lo:hi].copy()).to(dev))
        elif rnd % 3 == 1:  # host share
            A.set_batch(I2[lo:hi], J2[lo:hi], V2[lo:hi])
        else:              # the same batch in two halves: routed + pushed on the side stream, applied later
            nnz_before = A.info()["nnz"]
            if rnd % 2 == 0:
                A.stage_batch(torch.from_numpy(I2[lo:hi].copy()).to(dev), torch.from_numpy(J2[lo:hi].copy()).to(dev),
                              torch.from_numpy(V2[lo:hi].copy()).to(dev))
            else:
                A.stage_batch(I2[lo:hi].copy(), J2[lo:hi].copy(), V2[lo:hi].copy())
            assert A.info()["nnz"] == nnz_before          # a staged batch is not visible yet
            A.apply_staged()
        G.set_batch_policy(I2, J2, V2)
        mc, mr = owner_of(J2, cs) == rank, owner_of(I2, rs) == rank
        Gc.set_batch_policy(I2[mc], J2[mc], V2[mc])
        Gr.set_batch_policy(I2[mr], J2[mr], V2[mr])
        Sc.set_many(I2[mc], J2[mc], V2[mc])
        Sr.set_many(I2[mr], J2[mr], V2[mr])
        shard_equal(L.export(0), Gc.export(0), 0, f"batch {rnd}")
        shard_equal(L.export(1), Gr.export(1), 1, f"batch {rnd}")
        for which, S in ((0, Sc), (1, Sr)):
            gk, gv = contents(L.export(which))
            sk, sv = contents(S.export(which))
            assert np.array_equal(gk, sk) and np.array_equal(gv, sv), ("contents vs sequential reference", which, rnd)
        x = rng.integers(0, 4, n).astype(float)
        xt = rng.integers(0, 4, m).astype(float)
        y = A.spmv(torch.from_numpy(x).to(dev)).cpu().numpy()
        assert np.array_equal(y, G.mul_dense(x, m)), f"rank {rank}: A*x differs in round {rnd}"
        yt = A.spmv(xt, trans=True)                                # host-pointer variant
        assert np.array_equal(yt, G.mul_dense(xt, n, trans=True)), f"rank {rank}: A'*x differs in round {rnd}"
        # routed reads: every rank asks for different pairs (different counts too)
        nq = 500 + 37 * rank
        qr, qc = np.random.default_rng([seed, rnd, rank]).integers(1, m + 1, nq), np.random.default_rng([seed, rnd, rank, 1]).integers(1, n + 1, nq)
        assert np.array_equal(A.get_batch(qr, qc), G.get_many(qr, qc))
        assert np.array_equal(A.get_batch(qr, qc, which=1), G.get_many(qr, qc, which=1))
    inf = A.info()
    assert inf["nnz"] == G.nnz(), (inf, G.nnz())
    # deletecolumn! / deleterow! with the same list on every rank
    live_c = G.export(0)
    live_c = live_c["col_keys"][live_c["col_live"].astype(bool)]
    dead_c = rng.choice(live_c, max(1, len(live_c) // 10), replace=False)
    A.deletecolumn(dead_c)
    G.delete_columns_policy(dead_c)
    Gc.delete_columns_policy(dead_c[owner_of(dead_c, cs) == rank])
    # row shard: replay on the restricted oracle through the same policy entry point (it deletes from its col-major twin and
    # routes the deletes to its row-major structure as one batch, like the owner ranks do collectively)
    present = Gr.export(0)
    present = set(present["col_keys"][present["col_live"].astype(bool)].tolist())
    Gr.delete_columns_policy(np.array([c for c in dead_c.tolist() if c in present], dtype=np.int64))
    shard_equal(L.export(0), Gc.export(0), 0, "deletecolumn")
    gk, gv = contents(L.export(1))
    ok_, ov_ = contents(Gr.export(1))
    assert np.array_equal(gk, ok_) and np.array_equal(gv, ov_), "row-major shard after deletecolumn!"
    live_r = G.export(1)
    live_r = live_r["col_keys"][live_r["col_live"].astype(bool)]
    dead_r = rng.choice(live_r, max(1, len(live_r) // 10), replace=False)
    A.deleterow(dead_r)
    G.delete_rows_policy(dead_r)
    x = rng.integers(0, 4, n).astype(float)
    assert np.array_equal(A.spmv(x), G.mul_dense(x, m))
    xt = rng.integers(0, 4, m).astype(float)
    assert np.array_equal(A.spmv(xt, trans=True), G.mul_dense(xt, n, trans=True))
    assert A.info()["nnz"] == G.nnz()
    qr, qc = rng.integers(1, m + 1, 2000), rng.integers(1, n + 1, 2000)
    assert np.array_equal(A.get_batch(qr, qc), G.get_many(qr, qc))
    # collective error path: a column that does not exist -> ArgumentError on EVERY rank, nothing changed
    nnz_before = A.info()["nnz"]
    try:
        A.deletecolumn([int(dead_c[0])])
        raise AssertionError("deleting a deleted column must fail")
    except D.ArgumentError:
        pass
    try:
        A.set_batch(np.array([0], np.int64) if rank == world - 1 else np.zeros(0, np.int64),
                    np.array([3], np.int64) if rank == world - 1 else np.zeros(0, np.int64),
                    np.array([1.0]) if rank == world - 1 else np.zeros(0))
        raise AssertionError("key 0 must be refused")
    except D.ArgumentError:
        pass
    try:   # a share larger than max_share on ONE rank: refused on every rank (the flag travels with the send counts)
        big = max_share + 1 if rank == world - 1 else 0
        A.set_batch(np.ones(big, np.int64), np.ones(big, np.int64), np.ones(big))
        raise AssertionError("an oversize share must be refused")
    except D.ArgumentError:
        pass
    assert A.info()["nnz"] == nnz_before
    x = rng.integers(0, 4, n).astype(float)
    assert np.array_equal(A.spmv(x), G.mul_dense(x, m))       # the group is still in step after three refused collectives
    A.close()


run(m=3000, n=2600, nnz0=120_000, nb=40_000, nrounds=6, max_share=60_000, uneven=False, seed=5)
run(m=700, n=900, nnz0=30_000, nb=9_000, nrounds=4, max_share=max(4_000, -(-9_000 // world)), uneven=True, seed=6)   # several build rounds (few ranks), unequal shards
if rank == 0:
    print(f"sharded parity ok on {world} GPU(s), transport {ctx.info()['transport']}, nccl {ctx.info()['nccl_version']}")
ctx.close()
dist.destroy_process_group()
