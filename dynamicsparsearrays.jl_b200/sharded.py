"""Column-range sharding of a DynamicSparseMatrix across the GPUs of one box.

One process per GPU.  Rank r owns the column-major PCSR of the columns in ``col_split[r] .. col_split[r+1]-1`` and the
row-major PCSR of the rows in ``row_split[r] .. row_split[r+1]-1``.  A batch of logical updates ``A[i, j] = v`` is therefore
routed twice — to owner(j) for the column-major structure, to owner(i) for the row-major one — with one stable partition
by owner (``dsa_route_batch_d``, on the device) and one ``all_to_all_single`` per array over NCCL.  ``A * x`` is computed
from the row-major shards (each rank produces its own slice of y, written straight into the all-gather buffer);
``transpose(A) * x`` uses the column-major shards.  No other collective exists on the path.

The local structure is a ``LocalBackend`` (libdsa on the GPU).  The routing / exchange logic is backend-agnostic so that
``tests/test_sharded_gloo.py`` can drive it on CPU tensors over gloo with the oracle as the per-rank checker.
"""
import ctypes as C
import queue
import threading

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, lib


def even_splitters(n_keys, world):
    """split[r] = first key owned by rank r (keys are 1-based); split[world] = n_keys + 1."""
    per = -(-n_keys // world)
    return [1 + r * per for r in range(world)] + [max(n_keys, per * world) + 1]


def sampled_splitters(local_sample, n_keys, world, group=None):
    """Splitters from sampled quantiles (equal key ranges leave skewed data unbalanced).  Every rank passes a
    sample of the partition keys it will submit (any length, any rank may pass none); the samples are all-gathered, sorted,
    and cut at the world-quantiles.  Collective: every rank gets the same list.  split[0] = 1, split[world] = n_keys + 1."""
    mine = np.asarray(local_sample, dtype=np.int64).ravel()
    gathered = [None] * world
    dist.all_gather_object(gathered, mine, group=group)
    allk = np.sort(np.concatenate(gathered)) if sum(len(g) for g in gathered) else np.zeros(0, np.int64)
    if len(allk) == 0:
        return even_splitters(n_keys, world)
    cuts = [int(allk[min(len(allk) - 1, (r * len(allk)) // world)]) for r in range(1, world)]
    split = [1]
    for c in cuts:   # non-decreasing, inside [1, n_keys + 1]; equal splitters = an empty shard
        split.append(min(max(c, split[-1]), n_keys + 1))
    split.append(n_keys + 1)
    return split


def owner_of(keys, split):
    """owner(key) = number of interior splitters <= key."""
    inner = np.asarray(split[1:-1], dtype=np.int64)
    return np.searchsorted(inner, np.asarray(keys, dtype=np.int64), side="right")


class LibdsaBackend:
    """The per-rank structure on the GPU: one dsa_matrix handle whose two orientations are updated independently."""

    def __init__(self, device):
        self.device = device
        self.h = C.c_void_p()
        check(lib().dsa_matrix_create(C.byref(self.h)))
        check(lib().dsa_matrix_set_stream(self.h, C.c_void_p(torch.cuda.current_stream(device).cuda_stream)))

    def __del__(self):
        try:
            if self.h:
                lib().dsa_matrix_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def build(self, which, inkeys, partkeys, vals):
        inkeys, partkeys, vals = (np.ascontiguousarray(a) for a in (inkeys, partkeys, vals))
        check(lib().dsa_matrix_build_one(self.h, C.c_int(which), C.c_void_p(inkeys.ctypes.data), C.c_void_p(partkeys.ctypes.data),
                                         C.c_void_p(vals.ctypes.data), C.c_int64(len(vals)), C.c_int(0)))

    def set_batch(self, which, inkeys, partkeys, vals):
        n = inkeys.numel()
        if n:
            check(lib().dsa_matrix_set_batch_one_d(self.h, C.c_int(which), C.c_void_p(inkeys.data_ptr()), C.c_void_p(partkeys.data_ptr()),
                                                   C.c_void_p(vals.data_ptr()), C.c_int64(n)))

    def set_batch_two(self, col_triple, row_triple):
        """col_triple / row_triple = (rows, cols, vals) for the column-major / row-major structure; one call, shared syncs."""
        (rc, cc, vc), (rr, cr, vr) = col_triple, row_triple
        check(lib().dsa_matrix_set_batch_two_d(self.h, C.c_void_p(rc.data_ptr()), C.c_void_p(cc.data_ptr()), C.c_void_p(vc.data_ptr()),
                                               C.c_int64(rc.numel()), C.c_void_p(rr.data_ptr()), C.c_void_p(cr.data_ptr()),
                                               C.c_void_p(vr.data_ptr()), C.c_int64(rr.numel())))

    def spmv_range(self, trans, x, y_slice, key_lo, key_hi):
        check(lib().dsa_matrix_spmv_dense_range_d(self.h, C.c_int(1 if trans else 0), C.c_void_p(x.data_ptr()), C.c_int64(x.numel()),
                                                  C.c_void_p(y_slice.data_ptr()), C.c_int64(key_lo), C.c_int64(key_hi)))

    def info(self, which):
        out = np.zeros(10, np.int64)
        check(lib().dsa_matrix_info(self.h, C.c_int(which), C.c_void_p(out.ctypes.data)))
        return dict(capacity=int(out[0]), nb_elements=int(out[3]), nb_partitions=int(out[5]), nnz=int(out[9]))


def route(route_keys, rows, cols, vals, split, world):
    """Stable partition of a batch by owner(route_keys).  Returns (rows, cols, vals) in rank order + per-rank counts."""
    n = rows.numel()
    if rows.is_cuda:
        orows, ocols, ovals = torch.empty_like(rows), torch.empty_like(cols), torch.empty_like(vals)
        counts = np.zeros(world, np.int64)
        inner = np.asarray(split[1:-1], dtype=np.int64)
        check(lib().dsa_route_batch_d(C.c_void_p(route_keys.data_ptr()), C.c_void_p(rows.data_ptr()), C.c_void_p(cols.data_ptr()),
                                      C.c_void_p(vals.data_ptr()), C.c_int64(n), C.c_void_p(inner.ctypes.data), C.c_int(world),
                                      C.c_void_p(orows.data_ptr()), C.c_void_p(ocols.data_ptr()), C.c_void_p(ovals.data_ptr()),
                                      C.c_void_p(counts.ctypes.data), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return orows, ocols, ovals, counts.tolist()
    own = owner_of(route_keys.numpy(), split)
    order = np.argsort(own, kind="stable")
    counts = np.bincount(own, minlength=world).tolist()
    idx = torch.from_numpy(order)
    return rows[idx], cols[idx], vals[idx], counts


def exchange(arrays, send_counts, group=None):
    """all-to-all of the routed arrays: counts first (one small all_to_all), then one all_to_all_single per array."""
    world = dist.get_world_size(group)
    dev = arrays[0].device
    sc = torch.tensor(send_counts, dtype=torch.int64, device=dev)
    rc = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = rc.tolist()
    out = []
    for a in arrays:
        r = torch.empty(int(sum(recv_counts)), dtype=a.dtype, device=dev)
        dist.all_to_all_single(r, a.contiguous(), output_split_sizes=recv_counts, input_split_sizes=list(send_counts), group=group)
        out.append(r)
    return out, recv_counts


class ShardedMatrix:
    """set_batch(rows, cols, vals) routes + exchanges + applies one batch (validated on 2, 4 and 8 GPUs).

    EXPERIMENTAL: submit(...) hands the routing and the NCCL exchange of a batch to a background router (own CUDA stream,
    own communicator) so that they overlap with the application of the previous batch; apply_next() applies the oldest
    submitted batch.  The first version measured 2 043 Mupdates/s on 2 GPUs (86 % weak-scaling efficiency) but DEADLOCKED
    on 8 GPUs: the router's blocking host<->device copies synchronised with the legacy default stream, which was itself
    waiting on the main thread's all-gather, while other ranks waited for this router's all-to-all (two communicators,
    inconsistent order).  This version (pinned non-blocking count exchange, all-gather issued only after the router has
    enqueued the collectives of every submitted batch; the caller should also run its main work on a non-default stream)
    passes the gloo test but has not run on GPUs yet, so bench.py uses the synchronous path unless DSA_DIST_PIPELINE=1."""

    def __init__(self, m, n, backend, group=None, row_split=None, col_split=None):
        """row_split / col_split: optional splitter lists (world + 1 entries, e.g. from sampled_splitters) for skewed keys;
        the default is equal key ranges."""
        self.group = group
        self._router = None
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.m, self.n = int(m), int(n)
        self._even = row_split is None and col_split is None
        self.row_split = list(row_split) if row_split is not None else even_splitters(self.m, self.world)
        self.col_split = list(col_split) if col_split is not None else even_splitters(self.n, self.world)
        assert len(self.row_split) == self.world + 1 and len(self.col_split) == self.world + 1
        self.local = backend
        # slice length of the SpMV gather buffer: the widest shard (all shards are equally wide with even splitters)
        self.rows_per = max(b - a for a, b in zip(self.row_split[:-1], self.row_split[1:]))
        self.cols_per = max(b - a for a, b in zip(self.col_split[:-1], self.col_split[1:]))

    # rank-local key ranges
    def my_rows(self):
        return self.row_split[self.rank], min(self.row_split[self.rank + 1], self.m + 1)

    def my_cols(self):
        return self.col_split[self.rank], min(self.col_split[self.rank + 1], self.n + 1)

    def set_batch(self, rows, cols, vals):
        """rows/cols/vals: this rank's share of the global batch (tensors on the shard's device)."""
        if not rows.is_cuda:
            return self._set_batch_simple(rows, cols, vals)
        # the column-major structure lives with owner(col), the row-major one with owner(row): both stable partitions (packed
        # (n, 3) int64 rows, one library call, one host sync for the 2 x W send counts), one small all-to-all for the counts,
        # one all-to-all per orientation for the triples
        out = self._route_exchange(rows, cols, vals, self.group)
        self.local.set_batch_two(out[0], out[1])

    # ---- pipelined use -------------------------------------------------------------------------------------------
    def _route_exchange(self, rows, cols, vals, group):
        """route both orientations, exchange counts and packed triples; returns the two received (rows, cols, vals) triples"""
        W = self.world
        n = rows.numel()
        dev = rows.device
        pk_c = torch.empty((n, 3), dtype=torch.int64, device=dev)
        pk_r = torch.empty((n, 3), dtype=torch.int64, device=dev)
        cc, cr = np.zeros(W, np.int64), np.zeros(W, np.int64)
        isc = np.asarray(self.col_split[1:-1], dtype=np.int64)
        isr = np.asarray(self.row_split[1:-1], dtype=np.int64)
        check(lib().dsa_route_batch2_d(C.c_void_p(rows.data_ptr()), C.c_void_p(cols.data_ptr()), C.c_void_p(vals.data_ptr()), C.c_int64(n),
                                       C.c_void_p(isc.ctypes.data), C.c_void_p(isr.ctypes.data), C.c_int(W), C.c_void_p(pk_c.data_ptr()),
                                       C.c_void_p(pk_r.data_ptr()), C.c_void_p(cc.ctypes.data), C.c_void_p(cr.ctypes.data),
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        cnt_c, cnt_r = cc.tolist(), cr.tolist()
        sc = torch.tensor([x for pair in zip(cnt_c, cnt_r) for x in pair], dtype=torch.int64, device=dev)
        rc = torch.empty_like(sc)
        dist.all_to_all_single(rc, sc, group=group)
        rc = rc.view(W, 2).tolist()
        out = []
        for (packed, snd, rcv) in ((pk_c, cnt_c, [p[0] for p in rc]), (pk_r, cnt_r, [p[1] for p in rc])):
            recv = torch.empty((int(sum(rcv)), 3), dtype=torch.int64, device=dev)
            dist.all_to_all_single(recv, packed, output_split_sizes=rcv, input_split_sizes=list(snd), group=group)
            cols3 = recv.t().contiguous()
            out.append((cols3[0], cols3[1], cols3[2].view(torch.float64)))
        return out

    def _route_exchange_router(self, rows, cols, vals, group):
        """_route_exchange as issued by the router thread.  Differences that matter for deadlock freedom: no blocking
        pageable copy (those synchronise with the legacy default stream, which may be waiting on the main thread's
        all-gather): the counts travel through pinned buffers with non-blocking copies and a *stream* synchronisation."""
        W = self.world
        dev = rows.device
        if dev.type != "cuda":   # gloo / CPU tensors (tests): same protocol, host routing
            r1, c1, v1, cnt_c = route(cols, rows, cols, vals, self.col_split, W)
            r2, c2, v2, cnt_r = route(rows, rows, cols, vals, self.row_split, W)
            sc = torch.tensor([x for pair in zip(cnt_c, cnt_r) for x in pair], dtype=torch.int64)
            rc = torch.empty_like(sc)
            dist.all_to_all_single(rc, sc, group=group)
            rc = rc.view(W, 2).tolist()
            out = []
            for (r, c, v, snd, rcv) in ((r1, c1, v1, cnt_c, [p[0] for p in rc]), (r2, c2, v2, cnt_r, [p[1] for p in rc])):
                packed = torch.stack((r, c, v.view(torch.int64)), dim=1).contiguous()
                recv = torch.empty((int(sum(rcv)), 3), dtype=torch.int64)
                dist.all_to_all_single(recv, packed, output_split_sizes=rcv, input_split_sizes=list(snd), group=group)
                cols3 = recv.t().contiguous()
                out.append((cols3[0], cols3[1], cols3[2].view(torch.float64)))
            return out
        n = rows.numel()
        st = torch.cuda.current_stream()
        pk_c = torch.empty((n, 3), dtype=torch.int64, device=dev)
        pk_r = torch.empty((n, 3), dtype=torch.int64, device=dev)
        cc, cr = np.zeros(W, np.int64), np.zeros(W, np.int64)
        isc = np.asarray(self.col_split[1:-1], dtype=np.int64)
        isr = np.asarray(self.row_split[1:-1], dtype=np.int64)
        check(lib().dsa_route_batch2_d(C.c_void_p(rows.data_ptr()), C.c_void_p(cols.data_ptr()), C.c_void_p(vals.data_ptr()), C.c_int64(n),
                                       C.c_void_p(isc.ctypes.data), C.c_void_p(isr.ctypes.data), C.c_int(W), C.c_void_p(pk_c.data_ptr()),
                                       C.c_void_p(pk_r.data_ptr()), C.c_void_p(cc.ctypes.data), C.c_void_p(cr.ctypes.data),
                                       C.c_void_p(st.cuda_stream)))
        cnt_c, cnt_r = cc.tolist(), cr.tolist()
        h_sc = torch.tensor([x for pair in zip(cnt_c, cnt_r) for x in pair], dtype=torch.int64).pin_memory()
        sc = h_sc.to(dev, non_blocking=True)
        rc = torch.empty_like(sc)
        dist.all_to_all_single(rc, sc, group=group)
        h_rc = torch.empty(2 * W, dtype=torch.int64).pin_memory()
        h_rc.copy_(rc, non_blocking=True)
        st.synchronize()
        rcl = h_rc.view(W, 2).tolist()
        out = []
        for (packed, snd, rcv) in ((pk_c, cnt_c, [p[0] for p in rcl]), (pk_r, cnt_r, [p[1] for p in rcl])):
            recv = torch.empty((int(sum(rcv)), 3), dtype=torch.int64, device=dev)
            dist.all_to_all_single(recv, packed, output_split_sizes=rcv, input_split_sizes=list(snd), group=group)
            cols3 = recv.t().contiguous()
            out.append((cols3[0], cols3[1], cols3[2].view(torch.float64)))
        return out

    def _start_router(self, device):
        self._tasks, self._routed = queue.Queue(), queue.Queue()
        self._comm_group = dist.new_group(ranks=list(range(self.world)))   # own communicator: never shared with the main thread
        self._cuda = device is not None and device.type == "cuda"
        self._comm_stream = torch.cuda.Stream(device=device) if self._cuda else None
        # collectives of the two communicators must be enqueued in the same order on every rank: the main thread's
        # all-gather waits until the router has enqueued the collectives of every batch submitted so far
        self._order = threading.Condition()
        self._unissued = 0

        def loop():
            if self._cuda:
                torch.cuda.set_device(device)
                lib().dsa_set_device(C.c_int(device.index))
            while True:
                task = self._tasks.get()
                if task is None:
                    return
                try:
                    if self._cuda:
                        with torch.cuda.stream(self._comm_stream):
                            rows, cols, vals = (t.to(device, non_blocking=True) for t in task)   # pinned host shares are copied here
                            out = self._route_exchange_router(rows, cols, vals, self._comm_group)
                            ev = torch.cuda.Event()
                            ev.record(self._comm_stream)
                    else:
                        out, ev = self._route_exchange_router(*task, self._comm_group), None
                    self._routed.put((out, ev, None))
                except Exception as ex:   # surfaced by apply_next
                    self._routed.put((None, None, ex))
                finally:
                    with self._order:
                        self._unissued -= 1
                        self._order.notify_all()

        self._router = threading.Thread(target=loop, daemon=True)
        self._router.start()

    def submit(self, rows, cols, vals):
        """enqueue this rank's share of the next global batch (device tensors, or pinned host tensors)"""
        if self._router is None:
            self._start_router(getattr(self.local, "device", None))
        with self._order:
            self._unissued += 1
        self._tasks.put((rows, cols, vals))

    def _wait_router_issued(self):
        if self._router is not None:
            with self._order:
                while self._unissued > 0:
                    self._order.wait()

    def apply_next(self):
        """apply the oldest submitted batch to the local shards (blocks until its routing has been issued)"""
        out, ev, ex = self._routed.get()
        if ex is not None:
            raise ex
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
        if hasattr(self.local, "set_batch_two"):
            self.local.set_batch_two(out[0], out[1])
        else:   # backends with one call per orientation (the CPU backend of the gloo tests)
            (r1, c1, v1), (r2, c2, v2) = out
            self.local.set_batch(_lib.COLMAJOR, r1, c1, v1)
            self.local.set_batch(_lib.ROWMAJOR, c2, r2, v2)
        if ev is not None:
            for triple in out:   # the tensors were allocated on the router's stream
                for t in triple:
                    t.record_stream(torch.cuda.current_stream())

    def close(self):
        if self._router is not None:
            self._tasks.put(None)
            self._router.join(timeout=10)
            self._router = None

    def _set_batch_simple(self, rows, cols, vals):
        r1, c1, v1, cnt = route(cols, rows, cols, vals, self.col_split, self.world)
        (r1, c1, v1), _ = exchange([r1, c1, v1], cnt, self.group)
        self.local.set_batch(_lib.COLMAJOR, r1, c1, v1)          # in-array key = row, partition key = col
        r2, c2, v2, cnt = route(rows, rows, cols, vals, self.row_split, self.world)
        (r2, c2, v2), _ = exchange([r2, c2, v2], cnt, self.group)
        self.local.set_batch(_lib.ROWMAJOR, c2, r2, v2)          # in-array key = col, partition key = row

    def spmv(self, x, trans=False):
        """y = A * x (trans=False, x of length n) or transpose(A) * x; x replicated on every rank, y returned replicated."""
        per = self.cols_per if trans else self.rows_per
        if not self._even:
            return self._spmv_uneven(x, trans, per)
        lo = (self.col_split if trans else self.row_split)[self.rank]
        y = torch.zeros(per * self.world, dtype=torch.float64, device=x.device)
        y_slice = y[self.rank * per:(self.rank + 1) * per]
        self.local.spmv_range(trans, x, y_slice, lo, lo + per)   # epilogue writes this rank's slice of the gather buffer
        self._wait_router_issued()                               # keep the cross-communicator collective order identical on all ranks
        if y.is_cuda:
            dist.all_gather_into_tensor(y, y_slice, group=self.group)      # in place: the slice already sits at its offset
        else:                                                              # gloo (CPU tests)
            parts = [torch.empty(per, dtype=torch.float64) for _ in range(self.world)]
            dist.all_gather(parts, y_slice.clone(), group=self.group)
            y = torch.cat(parts)
        return y[: (self.n if trans else self.m)]

    def _spmv_uneven(self, x, trans, per):
        """sampled splitters: shards differ in width; every rank still contributes a slice of `per` entries (padded), and the
        result is assembled from the valid prefix of every slice"""
        split = self.col_split if trans else self.row_split
        lo, hi = split[self.rank], split[self.rank + 1]
        y = torch.zeros(per * self.world, dtype=torch.float64, device=x.device)
        y_slice = y[self.rank * per:(self.rank + 1) * per]
        if hi > lo:
            self.local.spmv_range(trans, x, y_slice[: hi - lo], lo, hi)
        self._wait_router_issued()
        if y.is_cuda:
            dist.all_gather_into_tensor(y, y_slice, group=self.group)
        else:
            parts = [torch.empty(per, dtype=torch.float64) for _ in range(self.world)]
            dist.all_gather(parts, y_slice.clone(), group=self.group)
            y = torch.cat(parts)
        out = torch.cat([y[r * per: r * per + (split[r + 1] - split[r])] for r in range(self.world)])
        return out[: (self.n if trans else self.m)]
