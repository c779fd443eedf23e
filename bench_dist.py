"""bench.py's multi-GPU arm (launched by torchrun, one rank per GPU): BASELINE.json configs[3] scaled to fit N GPUs with the
per-GPU work of configs[1] ("weak" scaling): a (1e5*N) x (1e5*N) PCSR with 1e7*N nnz sharded by column range; one step =
a global batch of 1M*N logical updates (each rank contributes 1M, routed to the owners of both orientations over NCCL) +
one SpMV A*x with the y slices all-gathered."""
import ctypes as C
import json
import os
import time

import numpy as np


def _block(seed, a, b, rows_per, cols_per, nnz):
    """entries of global block (row range a, col range b): same stream on every rank"""
    rng = np.random.default_rng([seed, a, b])
    I = rng.integers(1 + a * rows_per, 1 + (a + 1) * rows_per, nnz)
    J = rng.integers(1 + b * cols_per, 1 + (b + 1) * cols_per, nnz)
    V = rng.random(nnz) + 1e-3
    return I, J, V


def main_dist(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import bench as B
    import dsa_b200 as D
    from dsa_b200.sharded import LibdsaBackend, ShardedMatrix

    # stdout carries exactly ONE JSON line: everything else (NCCL banners, warnings) goes to stderr
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    D.lib().dsa_set_device(C.c_int(local_rank))
    import datetime
    # a short collective timeout: a hang must abort within minutes instead of holding N GPUs for the default 10
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=90))
    dev = torch.device("cuda", local_rank)
    # Default: every step routes, exchanges and applies its batch synchronously (validated on 2/4/8 GPUs).
    # DSA_DIST_PIPELINE=1 (experimental): the routing + NCCL exchange of batch s+1 runs on a background router (own stream and
    # communicator) and overlaps the application of batch s; the main work then runs on a non-default stream so that nothing
    # the router does can synchronise with it implicitly.
    PIPE = os.environ.get("DSA_DIST_PIPELINE", "0") == "1"
    if PIPE:
        torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    L = D.lib()
    K, W = args.steps, max(args.warmup, 3)
    per = B.M_ROWS                       # rows / cols per rank
    m = n = per * world
    nnz_block = B.NNZ0 // world          # every (a, b) block holds nnz/world entries -> each shard holds ~1e7 per orientation
    A = ShardedMatrix(m, n, LibdsaBackend(dev))
    # column-major shard = blocks (a, rank) for all a ; row-major shard = blocks (rank, b) for all b  (one consistent global matrix)
    cI, cJ, cV = (np.concatenate(x) for x in zip(*[_block(B.SEED, a, rank, per, per, nnz_block) for a in range(world)]))
    A.local.build(0, cI, cJ, cV)
    rI, rJ, rV = (np.concatenate(x) for x in zip(*[_block(B.SEED, rank, b, per, per, nnz_block) for b in range(world)]))
    A.local.build(1, rJ, rI, rV)
    del cI, cJ, cV, rI, rJ, rV
    # this rank's share of every global batch: 50% inserts anywhere in the global matrix, 50% deletes of its previous inserts
    rng = np.random.default_rng([B.SEED, 7, rank])
    half = B.BATCH // 2
    nsteps = 2 * (K + W)
    prev_i, prev_j = rng.integers(1, m + 1, half), rng.integers(1, n + 1, half)
    shares = []
    for _ in range(nsteps):
        ii, jj, vv = rng.integers(1, m + 1, half), rng.integers(1, n + 1, half), rng.random(half) + 1e-3
        p = rng.permutation(B.BATCH)
        shares.append((np.concatenate([ii, prev_i])[p], np.concatenate([jj, prev_j])[p], np.concatenate([vv, np.zeros(half)])[p]))
        prev_i, prev_j = ii, jj
    x_h = np.random.default_rng([B.SEED, 9]).random(n)
    d_x = torch.from_numpy(x_h).to(dev)
    d_sh = [tuple(torch.from_numpy(a).to(dev) for a in s) for s in shares[:K + W]]

    def step_dev(s, last):
        if not PIPE:
            A.set_batch(*d_sh[s])
            return A.spmv(d_x)
        if s + 1 < last:
            A.submit(*d_sh[s + 1])
        A.apply_next()
        return A.spmv(d_x)

    sampler = B.ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()   # samples every 100 ms from the warm-up to the end of the e2e loop (the timed regions are ~10 ms each)
    if PIPE:
        A.submit(*d_sh[0])
    for s in range(W):
        y = step_dev(s, W + K)
    torch.cuda.synchronize()
    dist.barrier()
    launches0 = L.dsa_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    dist.barrier()
    e0.record()
    for s in range(W, W + K):
        y = step_dev(s, W + K)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = L.dsa_launch_count() - launches0
    ms_step = float(ms.item()) / K
    value = B.BATCH * world / (ms_step * 1e-3) / 1e6
    checksum = float(y.sum().item())

    # e2e: pinned host shares, H2D + step + D2H of y inside the timed region
    h_sh = [tuple(torch.from_numpy(a).pin_memory() for a in s) for s in shares[K + W:]]
    h_x = torch.from_numpy(x_h).pin_memory()

    def step_host(s, last):
        xx = h_x.to(dev, non_blocking=True)
        if not PIPE:
            bi, bj, bv = (t.to(dev, non_blocking=True) for t in h_sh[s])
            A.set_batch(bi, bj, bv)
            return A.spmv(xx).cpu()
        if s + 1 < last:
            A.submit(*h_sh[s + 1])            # pinned host share: its H2D copy runs on the router's stream
        A.apply_next()
        return A.spmv(xx).cpu()

    if PIPE:
        A.submit(*h_sh[0])
    for s in range(W):
        yh = step_host(s, W + K)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for s in range(W, W + K):
        yh = step_host(s, W + K)
    torch.cuda.synchronize()
    dist.barrier()
    wall = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(wall, op=dist.ReduceOp.MAX)
    ms_e2e = 1e3 * float(wall.item()) / K
    clocks = sampler.stop() if sampler else None
    e2e = {"value": B.BATCH * world / (ms_e2e * 1e-3) / 1e6, "unit": "Mupdates/s", "h2d_bytes_per_step": (24 * B.BATCH + 8 * n) * world,
           "d2h_bytes_per_step": 8 * m * world, "ms_per_step": ms_e2e, "checksum": float(yh.sum().item())}

    # SpMV kernel roofline on rank 0 (CUDA events around every launch, outside the timed region)
    spmv = None
    inf = A.local.info(1)
    L.dsa_prof_reset()
    L.dsa_prof_enable(C.c_int(1))
    for _ in range(3):
        A.spmv(d_x)
    torch.cuda.synchronize()
    L.dsa_prof_enable(C.c_int(0))
    need = L.dsa_prof_dump(None, C.c_int64(0))
    buf = C.create_string_buffer(int(need) + 16)
    L.dsa_prof_dump(buf, C.c_int64(len(buf)))
    for ln in buf.value.decode().strip().splitlines():
        name, cnt, tms = ln.split(",")
        if name in ("spmv_flat", "spmv_bulk"):
            us = 1e3 * float(tms) / int(cnt)
            peak, src = B.measured_peak_gbs()
            alg = 16 * (inf["nnz"] + inf["nb_partitions"]) + 8 * (n + per)
            a = alg / (us * 1e-6) / 1e9
            spmv = {"kernel": name, "bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak,
                    "avg_us": us, "algorithmic_bytes": alg, "traffic": None, "peak_source": src, "scope": "rank 0 shard"}
    if rank == 0:
        cfg = B.workload_config(world)
        cfg["workload"] = (f"C4-style weak scaling: PCSR {m} x {n}, {nnz_block * world * world} nnz sharded by column range over {world} GPUs; "
                           f"step = {B.BATCH * world} updates routed to both orientations (NCCL all-to-all) + SpMV with all-gather")
        cfg.update(rows=m, cols=n, nnz=nnz_block * world * world, batch=B.BATCH * world,
                   routing="pipelined (background router)" if PIPE else "synchronous per step")
        os.dup2(saved_stdout, 1)
        print(json.dumps({
            "metric": "batched PCSR insert/delete Mupdates/s", "value": value, "unit": "Mupdates/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64 keys / f64 values", "data": "synthetic", "config": cfg, "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(launches), "roofline": spmv, "spmv": spmv, "cpu_baseline": None, "checksum": checksum,
            "shard_nnz_rank0": inf["nnz"]}), flush=True)
        os.dup2(2, 1)
    A.close()
    dist.destroy_process_group()
