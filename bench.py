#!/usr/bin/env python
"""bench.py — BASELINE.json configs[1]: PCSR 1e5 x 1e5 with 1e7 nnz resident in HBM; one step = one batched update of
1M logical entries (both orientations) followed by one SpMV (A * x, dense x).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (contract in the task statement).  `value` = whole-job Mupdates/s with inputs resident in HBM;
`e2e` = the same step through the host-pointer C-ABI calls (H2D of the batch + x, D2H of y inside the timed region);
`roofline` = the dominant kernel of the step against the measured HBM peak; `spmv` = the SpMV kernel's own roofline;
`cpu_baseline` = the CPU oracle (C++ restatement of the reference, 1 core) on one full step of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_ROWS = N_COLS = 100_000
NNZ0 = 10_000_000
BATCH = 1_000_000
SEED = 0xD5A00002
BYTES_PER_UPDATE = 80          # 2 orientations x (24 B triple read + 16 B cell write)
BYTES_PER_UPDATE_ONE = 40      # one orientation (what a single kernel launch of the update pipeline processes)


def make_workload(nsteps, m=M_ROWS, n=N_COLS, nnz0=NNZ0, nb=BATCH, seed=SEED):
    """Initial COO + `nsteps` batches.  Each batch = nb/2 inserts/overwrites of uniform (i, j, v) and nb/2 deletes (v = 0.0)
    of the previous batch's inserts (first batch: of initial entries), shuffled — the structure stays at ~1e7 nnz."""
    rng = np.random.default_rng(seed)
    I = rng.integers(1, m + 1, nnz0)
    J = rng.integers(1, n + 1, nnz0)
    V = rng.random(nnz0) + 1e-3
    half = nb // 2
    sel = rng.choice(nnz0, half, replace=False)
    prev_i, prev_j = I[sel], J[sel]
    batches = []
    for _ in range(nsteps):
        ins_i, ins_j = rng.integers(1, m + 1, half), rng.integers(1, n + 1, half)
        ins_v = rng.random(half) + 1e-3
        bi = np.concatenate([ins_i, prev_i])
        bj = np.concatenate([ins_j, prev_j])
        bv = np.concatenate([ins_v, np.zeros(half)])
        p = rng.permutation(nb)
        batches.append((np.ascontiguousarray(bi[p]), np.ascontiguousarray(bj[p]), np.ascontiguousarray(bv[p])))
        prev_i, prev_j = ins_i, ins_j
    x = rng.random(n)
    return (I, J, V), batches, x


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def ncu_traffic(kernel):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed ncu --set full capture."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["kernels"][kernel]["dram_bytes"]
    except Exception:
        return None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference_step(coo, batch, x, m, n):
    """The reference's CPU path on one full step: loop of setindex! over the batch (the reference has no batched update,
    matrix.jl:119-121) + mat * x.  Timed on one host core through the C++ oracle (Julia is not installed)."""
    from oracle import oracle as O
    t0 = time.perf_counter()
    M = O.Matrix(coo[0], coo[1], coo[2], m=m, n=n)
    t_build = time.perf_counter() - t0
    return M, t_build


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (C++ oracle port), rank 0 only."""
    if rank != 0:
        return
    from oracle import oracle as O
    coo, batches, x = make_workload(args.steps + args.warmup)
    M = O.Matrix(coo[0], coo[1], coo[2], m=M_ROWS, n=N_COLS)
    times = []
    for s, (bi, bj, bv) in enumerate(batches):
        t0 = time.perf_counter()
        M.set_many(bi, bj, bv)
        y = M.mul_dense(x, M_ROWS)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
    total = sum(times)
    val = BATCH * len(times) / total / 1e6
    line = {
        "impl": "reference", "metric": "batched PCSR insert/delete Mupdates/s", "value": val, "unit": "Mupdates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64 keys / f64 values", "data": "synthetic",
        "config": workload_config(1),
        "cpu_baseline": {"value": val, "unit": "Mupdates/s", "cores": 1, "kind": "port",
                         "sample": f"{len(times)} full steps (1M-update batch as a loop of setindex! + SpMV) on 1 host core; "
                                   "C++ restatement of the reference (Julia not installed)"},
        "e2e": {"value": val, "unit": "Mupdates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "checksum": float(np.sum(y)),
    }
    print(json.dumps(line))


def workload_config(n_gpus):
    return {"workload": "C2: dynamicsparse PCSR 1e5 x 1e5, 1e7 nnz resident; step = 1M-update batch (50% insert/overwrite, 50% delete, "
                        "both orientations) + SpMV A*x (dense x)", "rows": M_ROWS, "cols": N_COLS, "nnz": NNZ0, "batch": BATCH,
            "l2_policy": "inputs larger than L2 (2 x 268 MB gapped arrays per matrix vs 126 MB L2)", "seed": hex(SEED),
            "parallelism": f"column-range shards x{n_gpus}" if n_gpus > 1 else "single GPU",
            # experimental kernel switches in effect (empty = the validated defaults)
            "switches": {k: os.environ[k] for k in ("DSA_SPMV_BULK", "DSA_SPMV_STEPS", "DSA_TWO_STREAMS", "DSA_SCAN_ONEPASS", "DSA_ILP",
                                                    "DSA_DIST_PIPELINE") if k in os.environ}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        from bench_dist import main_dist   # sharded path (NCCL routing + all-gather)
        main_dist(args, rank, world, local_rank)
        return

    import torch

    import dsa_b200 as D
    D.require_gpu()
    L = D.lib()
    torch.cuda.set_device(local_rank)
    L.dsa_set_device(C.c_int(local_rank))
    K, W = args.steps, max(args.warmup, 3)
    nsteps_each = K + W
    prof_steps, sync_steps = 3, 5
    coo, batches, x = make_workload(2 * nsteps_each + prof_steps + sync_steps)   # one chain, consumed in order
    A = D.dynamicsparse(coo[0], coo[1], coo[2], m=M_ROWS, n=N_COLS)
    stream = torch.cuda.current_stream()
    D._lib.check(L.dsa_matrix_set_stream(A._h, C.c_void_p(stream.cuda_stream)))
    dev = torch.device("cuda", local_rank)

    def vp(t):
        return C.c_void_p(t.data_ptr())

    # ---- value: inputs resident in HBM --------------------------------------------------------------------
    d_batches = [(torch.from_numpy(bi).to(dev), torch.from_numpy(bj).to(dev), torch.from_numpy(bv).to(dev))
                 for bi, bj, bv in batches[:nsteps_each]]
    d_x = torch.from_numpy(x).to(dev)
    d_y = torch.zeros(M_ROWS, dtype=torch.float64, device=dev)

    def step_dev(s):
        bi, bj, bv = d_batches[s]
        D._lib.check(L.dsa_matrix_set_batch_d(A._h, vp(bi), vp(bj), vp(bv), C.c_int64(BATCH)))
        D._lib.check(L.dsa_matrix_spmv_dense_d(A._h, C.c_int(0), vp(d_x), C.c_int64(N_COLS), vp(d_y), C.c_int64(M_ROWS)))

    sampler = ClockSampler(local_rank)
    sampler.start()   # samples every 100 ms from the warm-up to the end of the e2e loop (each timed region is ~10 ms)
    for s in range(W):
        step_dev(s)
    torch.cuda.synchronize()
    launches0 = L.dsa_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for s in range(W, W + K):
        step_dev(s)
    e1.record(stream)
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    launches = L.dsa_launch_count() - launches0
    ms_step = ms_total / K
    value = BATCH / (ms_step * 1e-3) / 1e6
    checksum = float(d_y.sum().item())

    # ---- per-kernel durations (CUDA events around every launch, outside the timed region) -------------------
    L.dsa_prof_reset()
    L.dsa_prof_enable(C.c_int(1))
    extra = batches[nsteps_each:nsteps_each + prof_steps]
    for bi, bj, bv in extra:
        tb = (torch.from_numpy(bi).to(dev), torch.from_numpy(bj).to(dev), torch.from_numpy(bv).to(dev))
        D._lib.check(L.dsa_matrix_set_batch_d(A._h, vp(tb[0]), vp(tb[1]), vp(tb[2]), C.c_int64(BATCH)))
        D._lib.check(L.dsa_matrix_spmv_dense_d(A._h, C.c_int(0), vp(d_x), C.c_int64(N_COLS), vp(d_y), C.c_int64(M_ROWS)))
    torch.cuda.synchronize()
    L.dsa_prof_enable(C.c_int(0))
    need = L.dsa_prof_dump(None, C.c_int64(0))
    buf = C.create_string_buffer(int(need) + 16)
    L.dsa_prof_dump(buf, C.c_int64(len(buf)))
    kernels = {}
    for ln in buf.value.decode().strip().splitlines():
        name, cnt, ms = ln.split(",")
        kernels[name] = {"launches_per_step": int(cnt) / prof_steps, "ms_per_step": float(ms) / prof_steps,
                         "avg_us": 1e3 * float(ms) / int(cnt)}
    inf = A.info(1)
    peak, peak_src = measured_peak_gbs()
    spmv_alg_bytes = 16 * (inf["nnz"] + inf["nb_partitions"]) + 8 * (N_COLS + M_ROWS)
    spmv_phys_bytes = 16 * inf["capacity"] + 8 * (N_COLS + M_ROWS)
    spmv_name = "spmv_bulk" if "spmv_bulk" in kernels else "spmv_flat"   # DSA_SPMV_BULK selects the shared-memory-staged variant
    spmv_us = kernels.get(spmv_name, {}).get("avg_us")
    spmv = None
    if spmv_us:
        a = spmv_alg_bytes / (spmv_us * 1e-6) / 1e9
        spmv = {"kernel": spmv_name, "bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak,
                "frac_of_8TBs": a / 8000.0, "physical_gbs": spmv_phys_bytes / (spmv_us * 1e-6) / 1e9, "avg_us": spmv_us,
                "algorithmic_bytes": spmv_alg_bytes, "traffic": ncu_traffic(spmv_name), "peak_source": peak_src}
    dom = max(kernels.items(), key=lambda kv: kv[1]["ms_per_step"]) if kernels else (None, None)
    roofline = None
    if dom[0]:
        name, k = dom
        if name == spmv_name:
            alg = spmv_alg_bytes
        else:   # a kernel of the update pipeline processes one orientation's share of the batch per launch
            alg = BYTES_PER_UPDATE_ONE * BATCH
        a = alg / (k["avg_us"] * 1e-6) / 1e9
        roofline = {"kernel": name, "bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak, "traffic": ncu_traffic(name),
                    "avg_us": k["avg_us"], "algorithmic_bytes": alg, "share_of_step": k["ms_per_step"] / sum(v["ms_per_step"] for v in kernels.values()),
                    "peak_source": peak_src}
        if roofline["traffic"]:   # what the kernel physically moves (ncu DRAM bytes of the committed capture) over the live duration
            roofline["physical_gbs"] = roofline["traffic"] / (k["avg_us"] * 1e-6) / 1e9
            roofline["physical_frac"] = roofline["physical_gbs"] / peak

    # ---- e2e: host buffers through the host-pointer C-ABI calls ---------------------------------------------
    e2e = None
    if not args.no_e2e:
        hb = []
        for bi, bj, bv in batches[nsteps_each + prof_steps:2 * nsteps_each + prof_steps]:
            hb.append((torch.from_numpy(bi).pin_memory(), torch.from_numpy(bj).pin_memory(), torch.from_numpy(bv).pin_memory()))
        h_x = torch.from_numpy(x).pin_memory()
        h_y = torch.zeros(M_ROWS, dtype=torch.float64).pin_memory()
        # re-align the delete half of the first e2e batch with the last applied device batch: it is by construction
        # (batches form one chain), so the structure stays stationary

        # The flush is double-buffered (dsa_matrix_stage_batch / dsa_matrix_apply_staged): the PCIe transfer of batch s+1 overlaps
        # the kernels of batch s.  Every copy (batch, x in, y out) happens inside the timed region.
        def step_host(s, last):
            if s + 1 < last:
                bi, bj, bv = hb[s + 1]
                D._lib.check(L.dsa_matrix_stage_batch(A._h, vp(bi), vp(bj), vp(bv), C.c_int64(BATCH)))
            D._lib.check(L.dsa_matrix_apply_staged(A._h))
            D._lib.check(L.dsa_matrix_spmv_dense(A._h, C.c_int(0), vp(h_x), C.c_int64(N_COLS), vp(h_y), C.c_int64(M_ROWS)))

        bi, bj, bv = hb[0]
        D._lib.check(L.dsa_matrix_stage_batch(A._h, vp(bi), vp(bj), vp(bv), C.c_int64(BATCH)))
        for s in range(W):
            step_host(s, W + K)
        torch.cuda.synchronize()
        e0.record(stream)
        t0 = time.perf_counter()
        for s in range(W, W + K):
            step_host(s, W + K)
        e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms_e2e = max(e0.elapsed_time(e1), wall * 1e3) / K
        e2e = {"value": BATCH / (ms_e2e * 1e-3) / 1e6, "unit": "Mupdates/s", "h2d_bytes_per_step": 24 * BATCH + 8 * N_COLS,
               "d2h_bytes_per_step": 8 * M_ROWS, "ms_per_step": ms_e2e, "checksum": float(h_y.sum().item()),
               "api": "dsa_matrix_stage_batch + dsa_matrix_apply_staged + dsa_matrix_spmv_dense (host pinned buffers)"}
        # the same step through the plain synchronous call (no overlap of the copy), for reference
        hs = [(torch.from_numpy(bi).pin_memory(), torch.from_numpy(bj).pin_memory(), torch.from_numpy(bv).pin_memory())
              for bi, bj, bv in batches[2 * nsteps_each + prof_steps:]]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for bi, bj, bv in hs:
            D._lib.check(L.dsa_matrix_set_batch(A._h, vp(bi), vp(bj), vp(bv), C.c_int64(BATCH)))
            D._lib.check(L.dsa_matrix_spmv_dense(A._h, C.c_int(0), vp(h_x), C.c_int64(N_COLS), vp(h_y), C.c_int64(M_ROWS)))
        torch.cuda.synchronize()
        ms_sync = 1e3 * (time.perf_counter() - t0) / len(hs)
        e2e["synchronous_call"] = {"value": BATCH / (ms_sync * 1e-3) / 1e6, "ms_per_step": ms_sync, "steps": len(hs),
                                   "api": "dsa_matrix_set_batch + dsa_matrix_spmv_dense"}

    clocks = sampler.stop()

    # ---- CPU baseline: the oracle on one full step of the same workload ----------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import oracle as O
        OM = O.Matrix(coo[0], coo[1], coo[2], m=M_ROWS, n=N_COLS)
        bi, bj, bv = batches[0]
        t0 = time.perf_counter()
        OM.set_many(bi, bj, bv)
        t_upd = time.perf_counter() - t0
        t0 = time.perf_counter()
        yo = OM.mul_dense(x, M_ROWS)
        t_spmv = time.perf_counter() - t0
        cpu = {"value": BATCH / (t_upd + t_spmv) / 1e6, "unit": "Mupdates/s", "cores": 1, "kind": "port",
               "sample": "1 full step (1M-update batch as a loop of setindex! + SpMV) on 1 host core; C++ restatement of the "
                         "reference (Julia not installed)",
               "update_s": t_upd, "spmv_s": t_spmv, "spmv_effective_gbs": spmv_alg_bytes / t_spmv / 1e9,
               "host_cores_available": os.cpu_count()}

    line = {
        "metric": "batched PCSR insert/delete Mupdates/s", "value": value, "unit": "Mupdates/s", "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64 keys / f64 values", "data": "synthetic", "config": workload_config(1), "clocks": clocks, "e2e": e2e,
        "gpu_launches": int(launches), "roofline": roofline, "spmv": spmv, "cpu_baseline": cpu,
        "update_effective_gbs": BYTES_PER_UPDATE * BATCH / (ms_step * 1e-3) / 1e9, "kernels": kernels, "checksum": checksum,
        "nnz_after": inf["nnz"],
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
