"""Host-side mirror of DynamicSparseArrays.jl's public API (src/DynamicSparseArrays.jl:5-16) over libdsa's C ABI.

Julia is not installed in the build image, so this Python layer is the executable statement of the glue a maintainer
writes in vector.jl / matrix.jl / buffer.jl (julia/DynamicSparseArraysB200.jl holds the same calls as `ccall`s):
same names, same argument meaning, same error behaviour; the structures themselves live in HBM behind opaque handles.

    reference (Julia)                         here (Python)
    dynamicsparsevec(I, V[, combine, n])      dynamicsparsevec(I, V, combine="+", n=None)
    dynamicsparse(I, J, V[, m, n])            dynamicsparse(I, J, V, m=None, n=None)
    dynamicsparse(Ti, Tj, Tv; fill_mode)      dynamicsparse(fill_mode=True)
    v[k] / v[k] = x / m[i, j] / m[i, j] = x   same (single writes are queued and flushed as one batch before any read)
    deletecolumn!(m, j) / deleterow!(m, i)    deletecolumn(m, j) / deleterow(m, i)   (lists accepted: one bulk call)
    addrow!(m, i, cols, vals)                 addrow(m, i, cols, vals)
    closefillmode!(m)                         closefillmode(m)
    view(m, :, j) / view(m, i, :)             m.col(j) / m.row(i)  -> (keys, vals) in ascending key order
    m * v, transpose(m) * v, v * m, ...       m @ v, m.T @ v, v @ m, v @ m.T -> SparseVector
    nnz, nbpartitions, shrink_size!           nnz(x), nbpartitions(m.colmajor), shrink_size(v)
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import ArgumentError, ErrorException, check, lib

__all__ = ["DynamicSparseVector", "DynamicSparseMatrix", "DynamicMatrixColView", "SparseVector", "dynamicsparsevec",
           "dynamicsparse", "nbpartitions", "deletepartition", "deletecolumn", "deleterow", "addrow", "closefillmode",
           "shrink_size", "nnz"]


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _combine(c):
    if callable(c):
        import operator
        c = {operator.add: "+", operator.mul: "*", max: "max", min: "min"}.get(c, c)
    if c not in _lib.COMBINE:
        raise ArgumentError(_lib.DSA_ERR_ARGUMENT, f"unsupported combine operator {c!r}; supported: {sorted(_lib.COMBINE)}")
    return _lib.COMBINE[c]


class SparseVector:
    """Result of a product (SparseArrays.SparseVector in the reference, operations.jl:11-12): sorted indices + values."""

    def __init__(self, n, nzind, nzval):
        self.n, self.nzind, self.nzval = int(n), nzind, nzval

    def __len__(self):
        return self.n

    def __getitem__(self, k):
        i = np.searchsorted(self.nzind, k)
        return float(self.nzval[i]) if i < len(self.nzind) and self.nzind[i] == k else 0.0

    def todense(self):
        y = np.zeros(self.n)
        m = (self.nzind >= 1) & (self.nzind <= self.n)
        y[self.nzind[m] - 1] = self.nzval[m]
        return y

    def __eq__(self, other):   # SparseArrays `==` ignores stored zeros
        return self.n == other.n and np.array_equal(self.todense(), other.todense())

    __hash__ = None


class _PendingWrites:
    """Single setindex! calls are queued on the host and applied as ONE batched call before any read.  The batch is
    last-writer-wins in arrival order, so the result equals the reference's loop of setindex! (matrix.jl:119-121)."""

    def __init__(self):
        self.a, self.b, self.v = [], [], []

    def __len__(self):
        return len(self.v)

    def clear(self):
        self.a, self.b, self.v = [], [], []


class DynamicSparseVector:
    """DynamicSparseVector{Int64,Float64} (vector.jl:1-4) backed by a device PMA."""

    def __init__(self, handle, owner=True):
        self._h = handle
        self._pending = _PendingWrites()
        self.flush_threshold = 1 << 20

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().dsa_vec_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- writes -------------------------------------------------------------------------------
    def __setitem__(self, key, value):   # vector.jl:76-81
        self._pending.a.append(int(key))
        self._pending.v.append(float(value))
        if len(self._pending) >= self.flush_threshold:
            self.flush()

    def set_batch(self, keys, vals):
        """Batched setindex!: last writer wins, 0.0 deletes (pma.jl:196-213)."""
        self.flush()
        keys, vals = _i64(keys), _f64(vals)
        if len(keys) != len(vals):
            raise ArgumentError(_lib.DSA_ERR_ARGUMENT, "keys & values must have same length.")
        check(lib().dsa_vec_set_batch(self._h, _p(keys), _p(vals), C.c_int64(len(keys))))

    def flush(self):
        if len(self._pending):
            k, v = _i64(self._pending.a), _f64(self._pending.v)
            self._pending.clear()
            check(lib().dsa_vec_set_batch(self._h, _p(k), _p(v), C.c_int64(len(k))))

    # -- reads --------------------------------------------------------------------------------
    def __getitem__(self, key):   # vector.jl:72-73
        if isinstance(key, slice) and key == slice(None):
            return self
        return float(self.get_batch([key])[0])

    def get_batch(self, keys):
        self.flush()
        keys = _i64(keys)
        out = np.zeros(len(keys), np.float64)
        check(lib().dsa_vec_get_batch(self._h, _p(keys), C.c_int64(len(keys)), _p(out)))
        return out

    def info(self):
        self.flush()
        out = np.zeros(6, np.int64)
        check(lib().dsa_vec_info(self._h, _p(out)))
        return dict(capacity=int(out[0]), segment_capacity=int(out[1]), nb_segments=int(out[2]), nnz=int(out[3]),
                    height=int(out[4]), n=int(out[5]))

    def __len__(self):       # vector.jl:68
        return self.info()["n"]

    @property
    def size(self):          # vector.jl:69
        return (len(self),)

    def nonzeros(self):
        """(nonzeroinds, nonzeros) = iteration order of the PMA (vector.jl:93-109, pma.jl:165-180)."""
        self.flush()
        n = self.info()["nnz"]
        k, v = np.zeros(max(n, 1), np.int64), np.zeros(max(n, 1), np.float64)
        cnt = C.c_int64()
        check(lib().dsa_vec_nonzeros(self._h, _p(k), _p(v), C.c_int64(n), C.byref(cnt)))
        return k[:cnt.value], v[:cnt.value]

    def __iter__(self):      # vector.jl:71
        k, v = self.nonzeros()
        return iter(zip(k.tolist(), v.tolist()))

    def export(self):
        """Raw layout (occupied, keys, vals) for parity checks."""
        cap = self.info()["capacity"]
        occ, k, v = np.zeros(cap, np.uint8), np.zeros(cap, np.int64), np.zeros(cap, np.float64)
        check(lib().dsa_vec_export(self._h, _p(occ), _p(k), _p(v)))
        return occ, k, v

    def __eq__(self, other):   # vector.jl:85-87 + pma.jl:262-266: same length, same stored sequence
        if not isinstance(other, DynamicSparseVector):
            return NotImplemented
        if len(self) != len(other):
            return False
        (ka, va), (kb, vb) = self.nonzeros(), other.nonzeros()
        return np.array_equal(ka, kb) and np.array_equal(va, vb)

    __hash__ = None

    def __copy__(self):        # vector.jl:90
        raise ErrorException(_lib.DSA_ERR_ERROR, "copy of a dynamic sparse vector not implemented.")

    def __deepcopy__(self, memo):
        self.flush()
        h = C.c_void_p()
        check(lib().dsa_vec_clone(self._h, C.byref(h)))
        return DynamicSparseVector(h)

    def filter(self, f):       # pma.jl:224-234
        k, v = self.nonzeros()
        keep = [i for i, e in enumerate(zip(k.tolist(), v.tolist())) if f(e)]
        return dynamicsparsevec(k[keep], v[keep])

    def __matmul__(self, mat):   # v * mat / v * transpose(mat)  (operations.jl:38-60)
        if isinstance(mat, _Transposed):
            return mat.array._mul(self, trans=False, n=mat.array.size[0])
        if isinstance(mat, DynamicSparseMatrix):
            return mat._mul(self, trans=True, n=mat.size[1])
        return NotImplemented


def dynamicsparsevec(I, V, combine="+", n=None):
    """dynamicsparsevec(I, V, [combine, n]) (vector.jl:44-62)."""
    I, V = _i64(I), _f64(V)
    if len(I) != len(V):
        raise ArgumentError(_lib.DSA_ERR_ARGUMENT, "keys & nonzeros vectors must have same length.")
    h = C.c_void_p()
    check(lib().dsa_vec_build(_p(I), _p(V), C.c_int64(len(I)), C.c_int(_combine(combine)), C.c_int64(0 if n is None else n),
                              C.c_int(0 if n is None else 1), C.byref(h)))
    return DynamicSparseVector(h)


def shrink_size(v):   # shrink_size! (vector.jl:64)
    v.flush()
    out = C.c_int64()
    check(lib().dsa_vec_shrink_size(v._h, C.byref(out)))
    return out.value


# ---------------------------------------------------------------------------------------------
class Buffer:
    """Fill-mode write buffer (buffer.jl:1-50): a host dict row -> (colids, vals); flushed as COO by closefillmode!."""

    def __init__(self):
        self.rowmajor_coo = {}
        self.length = 0

    def addrow(self, rowid, colids, vals):   # buffer.jl:10-18
        if rowid in self.rowmajor_coo:
            raise ErrorException(_lib.DSA_ERR_ERROR, f"Row with id {rowid} already written in dynamic sparse matrix buffer.")
        colids, vals = _i64(colids), _f64(vals)
        p = np.argsort(colids, kind="stable")
        self.rowmajor_coo[rowid] = (colids[p].tolist(), vals[p].tolist())
        self.length += len(vals)

    def addelem(self, rowid, colid, val):    # buffer.jl:20-31
        r = self.rowmajor_coo.setdefault(rowid, ([], []))
        r[0].append(int(colid))
        r[1].append(float(val))
        self.length += 1

    def get_rowids_colids_vals(self):        # buffer.jl:33-50
        rows = np.zeros(self.length, np.int64)
        cols = np.zeros(self.length, np.int64)
        vals = np.zeros(self.length, np.float64)
        cur = 0
        for rowid, (bc, bv) in self.rowmajor_coo.items():
            k = len(bv)
            rows[cur:cur + k] = rowid
            cols[cur:cur + k] = bc
            vals[cur:cur + k] = bv
            cur += k
        return rows, cols, vals

    def row(self, row):                      # view(buffer, row, :)  (views.jl:44-49): combined with +
        bc, bv = self.rowmajor_coo.get(row, ([], []))
        out = {}
        for c, v in sorted(zip(bc, bv), key=lambda e: e[0]):
            out[c] = out.get(c, 0.0) + v
        return _i64(list(out.keys())), _f64(list(out.values()))


class _Orientation:
    """Read-only facade of one MappedPackedCSC (matrix.colmajor / matrix.rowmajor, matrix.jl:6-7)."""

    def __init__(self, matrix, which):
        self._m, self.which = matrix, which

    def info(self):
        return self._m.info(self.which)

    def export(self):
        return self._m.export(self.which)

    def __getitem__(self, idx):
        a, b = idx
        r, c = (a, b) if self.which == _lib.COLMAJOR else (b, a)
        return float(self._m.get_batch([r], [c], which=self.which)[0])


class DynamicMatrixColView:
    """view(matrix, :, col) / view(matrix, row, :) (views.jl:3-35): iterates (key, value) over the stored cells."""

    def __init__(self, keys, vals):
        self.keys, self.vals = keys, vals

    def __iter__(self):
        return iter(zip(self.keys.tolist(), self.vals.tolist()))

    def __len__(self):
        return len(self.keys)


class _Transposed:   # operations.jl:1-9
    def __init__(self, array):
        self.array = array

    @property
    def size(self):
        return tuple(reversed(self.array.size))

    def __getitem__(self, idx):
        return self.array[idx[1], idx[0]]

    def __setitem__(self, idx, val):
        self.array[idx[1], idx[0]] = val

    def __matmul__(self, v):   # transpose(mat) * v  (operations.jl:26-36)
        return self.array._mul(v, trans=True, n=self.array.size[1])


class DynamicSparseMatrix:
    """DynamicSparseMatrix{Int64,Int64,Float64} (matrix.jl:1-8): both orientations live on the device."""

    def __init__(self, handle=None, fill_mode=False):
        self._h = handle
        self.fillmode = fill_mode
        self.buffer = Buffer() if fill_mode else None
        self._m = self._n = 0            # dims while in fill mode (matrix.jl:44-47)
        self._pending = _PendingWrites()
        self.flush_threshold = 1 << 20
        self.colmajor = _Orientation(self, _lib.COLMAJOR)
        self.rowmajor = _Orientation(self, _lib.ROWMAJOR)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().dsa_matrix_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- writes -------------------------------------------------------------------------------
    def __setitem__(self, idx, val):   # matrix.jl:43-62
        row, col = idx
        val = float(val)
        if self.fillmode:
            if val != 0.0:
                self._m, self._n = max(self._m, row), max(self._n, col)
            self.buffer.addelem(row, col, val)
            return
        self._pending.a.append(int(row))
        self._pending.b.append(int(col))
        self._pending.v.append(val)
        if len(self._pending) >= self.flush_threshold:
            self.flush()

    def set_batch(self, rows, cols, vals):
        """Batched setindex! on both orientations: last writer wins, 0.0 deletes, absent rows/columns are created."""
        self._not_fillmode("Cannot apply a batch in fill mode")
        self.flush()
        rows, cols, vals = _i64(rows), _i64(cols), _f64(vals)
        if not (len(rows) == len(cols) == len(vals)):
            raise ArgumentError(_lib.DSA_ERR_ARGUMENT, "rows, columns, and nonzeros do not have same length.")
        check(lib().dsa_matrix_set_batch(self._h, _p(rows), _p(cols), _p(vals), C.c_int64(len(rows))))

    def flush(self):
        if len(self._pending):
            r, c, v = _i64(self._pending.a), _i64(self._pending.b), _f64(self._pending.v)
            self._pending.clear()
            check(lib().dsa_matrix_set_batch(self._h, _p(r), _p(c), _p(v), C.c_int64(len(r))))

    def stage_batch(self, rows, cols, vals):
        """Start the host->device copy of a batch and return at once; apply_staged() applies it (double-buffered flush).
        The arrays must stay alive (and should be pinned) until the matching apply_staged() returns."""
        self._not_fillmode("Cannot apply a batch in fill mode")
        self.flush()
        rows, cols, vals = _i64(rows), _i64(cols), _f64(vals)
        if not (len(rows) == len(cols) == len(vals)):
            raise ArgumentError(_lib.DSA_ERR_ARGUMENT, "rows, columns, and nonzeros do not have same length.")
        check(lib().dsa_matrix_stage_batch(self._h, _p(rows), _p(cols), _p(vals), C.c_int64(len(rows))))
        self._staged = getattr(self, "_staged", [])
        self._staged.append((rows, cols, vals))   # keep the host arrays alive until the copy has been consumed

    def apply_staged(self):
        check(lib().dsa_matrix_apply_staged(self._h))
        if getattr(self, "_staged", None):
            self._staged.pop(0)

    def _not_fillmode(self, msg):
        if self.fillmode:
            raise ErrorException(_lib.DSA_ERR_ERROR, msg)

    # -- reads --------------------------------------------------------------------------------
    def __getitem__(self, idx):   # matrix.jl:64-68
        row, col = idx
        if self.fillmode:
            if isinstance(col, slice):       # buffer[row, :] (buffer.jl:52-55): PMA of the row, duplicates NOT combined
                bc, bv = self.buffer.rowmajor_coo[row]
                return dynamicsparsevec(bc, bv, combine="last")
            raise ErrorException(_lib.DSA_ERR_ERROR, "getindex(row, col) is not available in fill mode")
        if isinstance(col, slice):
            k, v = self.row(row)
            return dynamicsparsevec(k, v)
        if isinstance(row, slice):
            k, v = self.col(col)
            return dynamicsparsevec(k, v)
        return float(self.get_batch([row], [col])[0])

    def get_batch(self, rows, cols, which=_lib.COLMAJOR):
        self._not_fillmode("Matrix is in fill mode")
        self.flush()
        rows, cols = _i64(rows), _i64(cols)
        out = np.zeros(len(rows), np.float64)
        check(lib().dsa_matrix_get_batch(self._h, C.c_int(which), _p(rows), _p(cols), C.c_int64(len(rows)), _p(out)))
        return out

    def _span(self, fn, key):
        self.flush()
        cnt = C.c_int64()
        check(fn(self._h, C.c_int64(key), None, None, C.c_int64(0), C.byref(cnt)))
        n = cnt.value
        k, v = np.zeros(max(n, 1), np.int64), np.zeros(max(n, 1), np.float64)
        if n:
            check(fn(self._h, C.c_int64(key), _p(k), _p(v), C.c_int64(n), C.byref(cnt)))
        return k[:n], v[:n]

    def col(self, col):
        """view(matrix, :, col) (matrix.jl:83-88)."""
        self._not_fillmode("View of a column not available in fill mode.")
        return self._span(lib().dsa_matrix_column, col)

    def row(self, row):
        """view(matrix, row, :) (matrix.jl:70-81), served from the row-major twin."""
        self._not_fillmode("Matrix is in fill mode, cannot create a view. However, you can use the view method on the buffer.")
        return self._span(lib().dsa_matrix_row, row)

    def view(self, row, col):
        if isinstance(row, slice):
            return DynamicMatrixColView(*self.col(col))
        return DynamicMatrixColView(*self.row(row))

    def info(self, which=_lib.COLMAJOR):
        self.flush()
        out = np.zeros(10, np.int64)
        check(lib().dsa_matrix_info(self._h, C.c_int(which), _p(out)))
        return dict(capacity=int(out[0]), segment_capacity=int(out[1]), nb_segments=int(out[2]), nb_elements=int(out[3]),
                    height=int(out[4]), nb_partitions=int(out[5]), nb_semaphores=int(out[6]), m=int(out[7]), n=int(out[8]),
                    nnz=int(out[9]))

    def export(self, which=_lib.COLMAJOR):
        inf = self.info(which)
        cap, ns = inf["capacity"], inf["nb_semaphores"]
        occ, k, v = np.zeros(cap, np.uint8), np.zeros(cap, np.int64), np.zeros(cap, np.float64)
        sem, ck, cl = np.zeros(max(ns, 1), np.int64), np.zeros(max(ns, 1), np.int64), np.zeros(max(ns, 1), np.uint8)
        check(lib().dsa_matrix_export(self._h, C.c_int(which), _p(occ), _p(k), _p(v), _p(sem), _p(ck), _p(cl)))
        return dict(tag=occ, key=k, val=v, semaphores=sem[:ns], col_keys=ck[:ns], col_live=cl[:ns], **inf)

    @property
    def size(self):   # matrix.jl:92
        if self.fillmode:
            return (self._m, self._n)
        inf = self.info()
        return (inf["m"], inf["n"])

    @property
    def T(self):      # transpose (operations.jl:5)
        return _Transposed(self)

    def __deepcopy__(self, memo):
        self._not_fillmode("deepcopy in fill mode is not supported")
        self.flush()
        h = C.c_void_p()
        check(lib().dsa_matrix_clone(self._h, C.byref(h)))
        return DynamicSparseMatrix(h)

    # -- products -----------------------------------------------------------------------------
    def _mul(self, x, trans, n):
        self._not_fillmode("Matrix is in fill mode")
        self.flush()
        if isinstance(x, DynamicSparseVector):
            xk, xv = x.nonzeros()
        elif isinstance(x, SparseVector):
            xk, xv = x.nzind, x.nzval
        else:
            xk, xv = x
        xk, xv = _i64(xk), _f64(xv)
        cap = self.info(_lib.COLMAJOR if trans else _lib.ROWMAJOR)["nb_semaphores"]
        yk, yv = np.zeros(max(cap, 1), np.int64), np.zeros(max(cap, 1), np.float64)
        cnt = C.c_int64()
        check(lib().dsa_matrix_spmv(self._h, C.c_int(1 if trans else 0), _p(xk), _p(xv), C.c_int64(len(xk)), _p(yk), _p(yv),
                                    C.c_int64(cap), C.byref(cnt)))
        return SparseVector(n, yk[:cnt.value].copy(), yv[:cnt.value].copy())

    def __matmul__(self, x):   # mat * v  (operations.jl:14-24)
        return self._mul(x, trans=False, n=self.size[0])

    def mul_dense(self, x, trans=False):
        """Dense-x product: x[j-1] for every j in 1..len(x); returns dense y of length size[0] (or size[1] if trans)."""
        self._not_fillmode("Matrix is in fill mode")
        self.flush()
        x = _f64(x)
        ny = self.size[1] if trans else self.size[0]
        y = np.zeros(max(ny, 1), np.float64)
        check(lib().dsa_matrix_spmv_dense(self._h, C.c_int(1 if trans else 0), _p(x), C.c_int64(len(x)), _p(y), C.c_int64(ny)))
        return y[:ny]


def dynamicsparse(I=None, J=None, V=None, m=None, n=None, fill_mode=True, combine="+"):
    """dynamicsparse(I, J, V, [m, n]) (matrix.jl:15-19) or dynamicsparse(Ti, Tj, Tv; fill_mode) (matrix.jl:31-41)."""
    if I is None:
        if fill_mode:
            return DynamicSparseMatrix(None, fill_mode=True)
        h = C.c_void_p()
        check(lib().dsa_matrix_create(C.byref(h)))
        return DynamicSparseMatrix(h)
    I, J, V = _i64(I), _i64(J), _f64(V)
    if not (len(I) == len(J) == len(V)):
        raise ArgumentError(_lib.DSA_ERR_ARGUMENT, "rows, columns, and nonzeros do not have same length.")
    h = C.c_void_p()
    given = m is not None
    check(lib().dsa_matrix_build_coo(_p(I), _p(J), _p(V), C.c_int64(len(I)), C.c_int64(m or 0), C.c_int64(n or 0),
                                     C.c_int(1 if given else 0), C.c_int(_combine(combine)), C.byref(h)))
    return DynamicSparseMatrix(h)


def closefillmode(matrix):   # closefillmode! (matrix.jl:126-134)
    if not matrix.fillmode:
        raise ErrorException(_lib.DSA_ERR_ERROR, "Cannot close fill mode because matrix is not in fill mode.")
    I, J, V = matrix.buffer.get_rowids_colids_vals()
    h = C.c_void_p()
    check(lib().dsa_matrix_build_coo(_p(I), _p(J), _p(V), C.c_int64(len(I)), C.c_int64(matrix._m), C.c_int64(matrix._n),
                                     C.c_int(1), C.c_int(_lib.COMBINE["+"]), C.byref(h)))
    matrix._h = h
    matrix.fillmode = False
    matrix.buffer = None
    return True


def addrow(matrix, row, colids, vals):   # addrow! (matrix.jl:113-124)
    if matrix.fillmode:
        # like the reference, addrow! in fill mode does not touch the dimensions (matrix.jl:116-117)
        matrix.buffer.addrow(row, colids, vals)
    else:
        for c, v in zip(colids, vals):
            matrix[row, c] = v
    return True


def _as_list(x):
    return [int(x)] if np.isscalar(x) else [int(e) for e in x]


def deletecolumn(matrix, col):   # deletecolumn! (matrix.jl:95-102); a list deletes in one bulk call
    matrix._not_fillmode("Cannot delete a column in fill mode")
    matrix.flush()
    ids = _i64(_as_list(col))
    check(lib().dsa_matrix_delete_columns(matrix._h, _p(ids), C.c_int64(len(ids))))
    return True


def deleterow(matrix, row):      # deleterow! (matrix.jl:104-111)
    matrix._not_fillmode("Cannot delete a row in fill mode")
    matrix.flush()
    ids = _i64(_as_list(row))
    check(lib().dsa_matrix_delete_rows(matrix._h, _p(ids), C.c_int64(len(ids))))
    return True


def deletepartition(orientation, key):   # deletecolumn!(mpcsc, col) on one orientation is not exposed separately:
    raise ErrorException(_lib.DSA_ERR_ERROR, "deletepartition! on a single orientation would desynchronise the twin; "
                                             "use deletecolumn / deleterow on the matrix")


def nbpartitions(orientation):   # pcsr.jl:21-22
    return orientation.info()["nb_partitions"]


def nnz(x):   # pma.jl:163, pcsr.jl:11, matrix.jl:91
    if isinstance(x, DynamicSparseVector):
        return x.info()["nnz"]
    if isinstance(x, DynamicSparseMatrix):
        return x.info(_lib.ROWMAJOR)["nnz"]
    if isinstance(x, _Orientation):
        return x.info()["nnz"]
    if isinstance(x, SparseVector):
        return len(x.nzind)
    raise TypeError(type(x))
