"""Column-range sharding of a DynamicSparseMatrix across the GPUs of one box, one process per GPU.

Rank r owns the column-major PCSR of the columns in ``col_split[r] .. col_split[r+1]-1`` and the row-major PCSR of the rows in
``row_split[r] .. row_split[r+1]-1``.  A batch of logical updates ``A[i, j] = v`` is routed twice — to owner(j) for the
column-major structure, to owner(i) for the row-major one.  The op order of a global batch is rank-major, then arrival within
a rank: last writer wins in that order.  ``A * x`` is computed from the row-major shards (each rank produces its own slice of
y) and all-gathered; ``transpose(A) * x`` uses the column-major shards.

``DistMatrix`` is the product path: a ctypes binding of libdsa's ``dsa_dmatrix_*`` entry points (include/dsa.h), where the
routed updates travel as NVLink peer-memory stores fused into the routing kernel and NCCL carries the send counts and the
SpMV all-gather.  torch.distributed is only the plumbing that hands the NCCL unique id to the ranks.

``ShardedMatrix`` is the same protocol in Python over torch.distributed collectives with a pluggable per-rank backend: the
host-logic mirror that ``tests/test_sharded_gloo.py`` drives on CPU tensors over gloo with the oracle as the per-rank checker.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, lib


def even_splitters(n_keys, world):
    """split[r] = first key owned by rank r (keys are 1-based); split[world] = n_keys + 1."""
    per = max(-(-n_keys // world), 1)
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


# ---------------------------------------------------------------------------------------------------------------------
# product path: libdsa's multi-GPU entry points
# ---------------------------------------------------------------------------------------------------------------------
def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _hp(a):
    return C.c_void_p(a.ctypes.data)


def _dp(t):
    return C.c_void_p(t.data_ptr())


class DistContext:
    """This rank's seat in the group (dsa_dist_t).  Collective constructor: rank 0 draws the NCCL unique id, torch.distributed
    hands it to the others (any host-side transport would do: the Julia glue uses MPI or a file)."""

    def __init__(self, group=None, device=None):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if device is not None:
            check(lib().dsa_set_device(C.c_int(device.index if hasattr(device, "index") else int(device))))
        uid = (C.c_uint8 * 128)()
        if self.world > 1:
            box = [None]
            if self.rank == 0:
                check(lib().dsa_dist_unique_id(uid))
                box[0] = bytes(uid)
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            uid = (C.c_uint8 * 128).from_buffer_copy(box[0])
        self._h = C.c_void_p()
        check(lib().dsa_dist_init(uid, C.c_int(self.rank), C.c_int(self.world), C.byref(self._h)))

    def info(self):
        out = np.zeros(4, np.int64)
        check(lib().dsa_dist_info(self._h, _hp(out)))
        return dict(rank=int(out[0]), world=int(out[1]), transport="nccl" if out[2] else "peer-memory", nccl_version=int(out[3]))

    def close(self):
        if self._h:
            lib().dsa_dist_destroy(self._h)
            self._h = None


class DistMatrix:
    """DynamicSparseMatrix sharded by key range over a DistContext (dsa_dmatrix_t).  Every method is a collective: all ranks
    call it in the same order.  Batches are this rank's SHARE of a global batch (numpy arrays on the host, or torch tensors on
    this rank's GPU)."""

    def __init__(self, ctx, m, n, max_share, row_split=None, col_split=None, stream=None):
        self.ctx, self.m, self.n = ctx, int(m), int(n)
        self.row_split = list(row_split) if row_split is not None else even_splitters(self.m, ctx.world)
        self.col_split = list(col_split) if col_split is not None else even_splitters(self.n, ctx.world)
        rs, cs = _i64(self.row_split), _i64(self.col_split)
        self._h = C.c_void_p()
        check(lib().dsa_dmatrix_create(ctx._h, C.c_int64(self.m), C.c_int64(self.n), _hp(rs), _hp(cs), C.c_int64(int(max_share)),
                                       C.byref(self._h)))
        # Work is enqueued on `stream` (a raw cudaStream_t).  Default: torch's current stream, so that torch tensors handed to
        # set_batch / spmv and the results they read back are ordered with the library's kernels without extra synchronisation.
        if stream is None and torch.cuda.is_available():
            stream = torch.cuda.current_stream().cuda_stream
        if stream is not None:
            check(lib().dsa_dmatrix_set_stream(self._h, C.c_void_p(stream)))

    def close(self):
        if self._h:
            lib().dsa_dmatrix_destroy(self._h)
            self._h = None

    # rank-local key ranges [lo, hi)
    def my_rows(self):
        return self.row_split[self.ctx.rank], self.row_split[self.ctx.rank + 1]

    def my_cols(self):
        return self.col_split[self.ctx.rank], self.col_split[self.ctx.rank + 1]

    @property
    def local(self):
        """this rank's shards as a (borrowed) DynamicSparseMatrix: info / export / views for parity checks"""
        from .api import DynamicSparseMatrix
        lib().dsa_dmatrix_local.restype = C.c_void_p
        M = DynamicSparseMatrix(C.c_void_p(lib().dsa_dmatrix_local(self._h)))
        M._owner = False
        return M

    def build_coo(self, rows, cols, vals, combine="+"):
        rows, cols, vals = _i64(rows), _i64(cols), _f64(vals)
        check(lib().dsa_dmatrix_build_coo(self._h, _hp(rows), _hp(cols), _hp(vals), C.c_int64(len(vals)), C.c_int(_lib.COMBINE[combine])))

    def build_local(self, which, inkeys, partkeys, vals, combine="+"):
        if isinstance(inkeys, torch.Tensor) and inkeys.is_cuda:
            check(lib().dsa_dmatrix_build_local_d(self._h, C.c_int(which), _dp(inkeys), _dp(partkeys), _dp(vals), C.c_int64(inkeys.numel()),
                                                  C.c_int(_lib.COMBINE[combine])))
            return
        inkeys, partkeys, vals = _i64(inkeys), _i64(partkeys), _f64(vals)
        check(lib().dsa_dmatrix_build_local(self._h, C.c_int(which), _hp(inkeys), _hp(partkeys), _hp(vals), C.c_int64(len(vals)),
                                            C.c_int(_lib.COMBINE[combine])))

    def set_batch(self, rows, cols, vals):
        if isinstance(rows, torch.Tensor) and rows.is_cuda:
            check(lib().dsa_dmatrix_set_batch_d(self._h, _dp(rows), _dp(cols), _dp(vals), C.c_int64(rows.numel())))
            return
        rows, cols, vals = _i64(rows), _i64(cols), _f64(vals)
        check(lib().dsa_dmatrix_set_batch(self._h, _hp(rows), _hp(cols), _hp(vals), C.c_int64(len(vals))))

    def stage_batch(self, rows, cols, vals):
        """Route this rank's share of the NEXT global batch and push it into the owners' receive regions on a side stream;
        returns at once.  apply_staged() applies the oldest staged batch: staging batch k+1 before applying batch k overlaps the
        exchange with the kernels of batch k.  The arrays must stay alive until the matching apply_staged() has returned."""
        self._staged = getattr(self, "_staged", [])
        if isinstance(rows, torch.Tensor) and rows.is_cuda:
            check(lib().dsa_dmatrix_stage_batch_d(self._h, _dp(rows), _dp(cols), _dp(vals), C.c_int64(rows.numel())))
        elif isinstance(rows, torch.Tensor):    # pinned host tensors
            check(lib().dsa_dmatrix_stage_batch(self._h, _dp(rows), _dp(cols), _dp(vals), C.c_int64(rows.numel())))
        else:
            rows, cols, vals = _i64(rows), _i64(cols), _f64(vals)
            check(lib().dsa_dmatrix_stage_batch(self._h, _hp(rows), _hp(cols), _hp(vals), C.c_int64(len(vals))))
        self._staged.append((rows, cols, vals))

    def apply_staged(self):
        check(lib().dsa_dmatrix_apply_staged(self._h))
        if getattr(self, "_staged", None):
            self._staged.pop(0)

    def spmv(self, x, trans=False, out=None):
        """y = A * x (x of length n) or transpose(A) * x; x replicated on every rank, y returned on every rank."""
        ny = self.n if trans else self.m
        if isinstance(x, torch.Tensor) and x.is_cuda:
            y = out if out is not None else torch.empty(ny, dtype=torch.float64, device=x.device)
            check(lib().dsa_dmatrix_spmv_dense_d(self._h, C.c_int(1 if trans else 0), _dp(x), C.c_int64(x.numel()), _dp(y), C.c_int64(ny)))
            return y
        x = _f64(x)
        y = np.zeros(max(ny, 1))
        check(lib().dsa_dmatrix_spmv_dense(self._h, C.c_int(1 if trans else 0), _hp(x), C.c_int64(len(x)), _hp(y), C.c_int64(ny)))
        return y[:ny]

    def get_batch(self, rows, cols, which=_lib.COLMAJOR):
        rows, cols = _i64(rows), _i64(cols)
        out = np.zeros(max(len(rows), 1))
        check(lib().dsa_dmatrix_get_batch(self._h, C.c_int(which), _hp(rows), _hp(cols), C.c_int64(len(rows)), _hp(out)))
        return out[:len(rows)]

    def deletecolumn(self, cols):
        cols = _i64(np.atleast_1d(cols))
        check(lib().dsa_dmatrix_delete_columns(self._h, _hp(cols), C.c_int64(len(cols))))

    def deleterow(self, rows):
        rows = _i64(np.atleast_1d(rows))
        check(lib().dsa_dmatrix_delete_rows(self._h, _hp(rows), C.c_int64(len(rows))))

    def info(self):
        out = np.zeros(8, np.int64)
        check(lib().dsa_dmatrix_info(self._h, _hp(out)))
        return dict(m=int(out[0]), n=int(out[1]), nnz=int(out[2]), col_partitions=int(out[3]), row_partitions=int(out[4]),
                    max_share=int(out[5]), nnz_colmajor=int(out[6]), transport="nccl" if out[7] else "peer-memory")


# ---------------------------------------------------------------------------------------------------------------------
# host-logic mirror over torch.distributed (CPU / gloo tests)
# ---------------------------------------------------------------------------------------------------------------------
def route(route_keys, rows, cols, vals, split, world):
    """Stable partition of a batch by owner(route_keys).  Returns (rows, cols, vals) in rank order + per-rank counts."""
    own = owner_of(route_keys.numpy(), split)
    order = np.argsort(own, kind="stable")
    counts = np.bincount(own, minlength=world).tolist()
    idx = torch.from_numpy(order)
    return rows[idx], cols[idx], vals[idx], counts


def exchange(arrays, send_counts, group=None):
    """all-to-all of the routed arrays: counts first, then one all_to_all_single per array.  What a rank receives is ordered
    by source rank, arrival order within a source: the global op order of the batch."""
    world = dist.get_world_size(group)
    sc = torch.tensor(send_counts, dtype=torch.int64)
    rc = torch.empty(world, dtype=torch.int64)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = rc.tolist()
    out = []
    for a in arrays:
        r = torch.empty(int(sum(recv_counts)), dtype=a.dtype)
        dist.all_to_all_single(r, a.contiguous(), output_split_sizes=recv_counts, input_split_sizes=list(send_counts), group=group)
        out.append(r)
    return out, recv_counts


class ShardedMatrix:
    """The sharding protocol in Python (CPU tensors, any torch.distributed backend) over a per-rank backend with
    ``set_batch(which, inkeys, partkeys, vals)`` and ``spmv_range(trans, x, y_slice, key_lo, key_hi)``."""

    def __init__(self, m, n, backend, group=None, row_split=None, col_split=None):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.m, self.n = int(m), int(n)
        self.row_split = list(row_split) if row_split is not None else even_splitters(self.m, self.world)
        self.col_split = list(col_split) if col_split is not None else even_splitters(self.n, self.world)
        assert len(self.row_split) == self.world + 1 and len(self.col_split) == self.world + 1
        self.local = backend
        # slice length of the SpMV gather buffer: the widest shard (all shards are equally wide with even splitters)
        self.rows_per = max(b - a for a, b in zip(self.row_split[:-1], self.row_split[1:]))
        self.cols_per = max(b - a for a, b in zip(self.col_split[:-1], self.col_split[1:]))

    def my_rows(self):
        return self.row_split[self.rank], min(self.row_split[self.rank + 1], self.m + 1)

    def my_cols(self):
        return self.col_split[self.rank], min(self.col_split[self.rank + 1], self.n + 1)

    def set_batch(self, rows, cols, vals):
        """rows/cols/vals: this rank's share of the global batch"""
        r1, c1, v1, cnt = route(cols, rows, cols, vals, self.col_split, self.world)
        (r1, c1, v1), _ = exchange([r1, c1, v1], cnt, self.group)
        self.local.set_batch(_lib.COLMAJOR, r1, c1, v1)          # in-array key = row, partition key = col
        r2, c2, v2, cnt = route(rows, rows, cols, vals, self.row_split, self.world)
        (r2, c2, v2), _ = exchange([r2, c2, v2], cnt, self.group)
        self.local.set_batch(_lib.ROWMAJOR, c2, r2, v2)          # in-array key = col, partition key = row

    def spmv(self, x, trans=False):
        """y = A * x (trans=False, x of length n) or transpose(A) * x; x replicated on every rank, y returned replicated."""
        per = self.cols_per if trans else self.rows_per
        split = self.col_split if trans else self.row_split
        lo, hi = split[self.rank], split[self.rank + 1]
        y_slice = torch.zeros(per, dtype=torch.float64)
        if hi > lo:
            self.local.spmv_range(trans, x, y_slice[: hi - lo], lo, hi)
        parts = [torch.empty(per, dtype=torch.float64) for _ in range(self.world)]
        dist.all_gather(parts, y_slice, group=self.group)
        out = torch.cat([parts[r][: split[r + 1] - split[r]] for r in range(self.world)])
        return out[: (self.n if trans else self.m)]

    def close(self):
        pass
