"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

ctypes wrapper around the CPU oracle (``oracle/liboracle.so``), the literal C++
restatement of /root/reference/src/*.jl.  Imported only by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs, and only as the checker or the timed CPU baseline — never by the product.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

OK, ERR_ARGUMENT, ERR_BOUNDS, ERR_ERROR, ERR_ASSERT = 0, 1, 2, 3, 4
COMB_ADD, COMB_MUL, COMB_LAST, COMB_FIRST, COMB_MIN, COMB_MAX = range(6)

i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


class OracleError(Exception):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def build(force=False):
    """Compile liboracle.so with the committed Makefile (g++ only; no GPU needed)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp", "Makefile"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_last_error.restype = C.c_char_p
        for name in ("orc_vec_build", "orc_vec_clone", "orc_pcsc_build", "orc_pcsc_clone", "orc_mat_build", "orc_mat_empty",
                     "orc_mat_clone"):
            getattr(L, name).restype = C.c_void_p
        L.orc_vec_get.restype = C.c_double
        for name in ("orc_vec_shrink_size", "orc_mat_column", "orc_mat_row_scan", "orc_mat_mul"):
            getattr(L, name).restype = C.c_int64
        _LIB = L
    return _LIB


def _check(code):
    if code != 0:
        raise OracleError(code, lib().orc_last_error().decode())


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------------------------------
# raw gapped arrays: python lists of None | (key, val)
# ---------------------------------------------------------------------------------------
class Cells:
    """A gapped array as three numpy arrays (tag, key, val); built from a list of None|(k,v)."""

    def __init__(self, cells):
        n = len(cells)
        self.tag = np.zeros(n, np.uint8)
        self.key = np.zeros(n, np.int64)
        self.val = np.zeros(n, np.float64)
        for i, c in enumerate(cells):
            if c is not None:
                self.tag[i] = 1
                self.key[i] = c[0]
                self.val[i] = c[1]

    @property
    def n(self):
        return len(self.tag)

    def tolist(self):
        return [None if not t else (int(k), float(v)) for t, k, v in zip(self.tag, self.key, self.val)]

    def args(self):
        return _p(self.tag), _p(self.key), _p(self.val), C.c_int64(self.n)


def find(cells, key, frm=None, to=None):
    a = cells if isinstance(cells, Cells) else Cells(cells)
    frm = 1 if frm is None else frm
    to = a.n if to is None else to
    pos = C.c_int64()
    _check(lib().orc_find(*a.args(), C.c_int64(key), C.c_int64(frm), C.c_int64(to), C.byref(pos)))
    p = pos.value
    return (p, None if p == 0 else a.tolist()[p - 1])


def _sem_args(sem):
    if sem is None:
        return None, C.c_int64(0), None
    s = _i64([0 if x is None else x for x in sem])
    return _p(s), C.c_int64(len(s)), s


def insert(a, key, val, frm=None, to=None, sem=None):
    frm = 1 if frm is None else frm
    to = a.n if to is None else to
    sp, sn, s = _sem_args(sem)
    pos, isnew = C.c_int64(), C.c_int()
    _check(lib().orc_insert(*a.args(), C.c_int64(key), C.c_double(val), C.c_int64(frm), C.c_int64(to), sp, sn,
                            C.byref(pos), C.byref(isnew)))
    if sem is not None:
        sem[:] = [None if x == 0 else int(x) for x in s]
    return pos.value, bool(isnew.value)


def delete(a, key, frm=None, to=None):
    frm = 1 if frm is None else frm
    to = a.n if to is None else to
    pos, d = C.c_int64(), C.c_int()
    _check(lib().orc_delete(*a.args(), C.c_int64(key), C.c_int64(frm), C.c_int64(to), C.byref(pos), C.byref(d)))
    return pos.value, bool(d.value)


def purge(a, frm, to):
    mid, nb = C.c_int64(), C.c_int64()
    _check(lib().orc_purge(*a.args(), C.c_int64(frm), C.c_int64(to), C.byref(mid), C.byref(nb)))
    return mid.value, nb.value


def move(a, right, frm, to, sem=None):
    sp, sn, s = _sem_args(sem)
    _check(lib().orc_move(*a.args(), C.c_int(1 if right else 0), C.c_int64(frm), C.c_int64(to), sp, sn))
    if sem is not None:
        sem[:] = [None if x == 0 else int(x) for x in s]


def pack(a, ws, we, m):
    _check(lib().orc_pack(*a.args(), C.c_int64(ws), C.c_int64(we), C.c_int64(m)))


def spread(a, ws, we, m, sem=None, five_arg=False):
    sp, sn, s = _sem_args(sem)
    _check(lib().orc_spread(*a.args(), C.c_int64(ws), C.c_int64(we), C.c_int64(m), C.c_int(1 if (five_arg or sem is not None) else 0),
                            sp, sn))
    if sem is not None:
        sem[:] = [None if x == 0 else int(x) for x in s]


def arrays_equal(c1, c2):
    a, b = Cells(c1), Cells(c2)
    return bool(lib().orc_arrays_equal(*a.args(), *b.args()))


def geometry(n):
    out = np.zeros(4, np.int64)
    _check(lib().orc_geometry(C.c_int64(n), _p(out)))
    return dict(capacity=int(out[0]), segment_capacity=int(out[1]), nb_segments=int(out[2]), height=int(out[3]))


# ---------------------------------------------------------------------------------------
class _Handle:
    _free = None

    def __init__(self, h):
        self.h = C.c_void_p(h)

    def __del__(self):
        try:
            if self.h and self._free:
                getattr(lib(), self._free)(self.h)
                self.h = None
        except Exception:
            pass


class Vec(_Handle):
    """DynamicSparseVector (vector.jl)."""
    _free = "orc_vec_free"

    def __init__(self, I=(), V=(), combine=COMB_ADD, n=None, _h=None):
        if _h is not None:
            super().__init__(_h)
            return
        I, V = _i64(I), _f64(V)
        if len(I) != len(V):
            raise OracleError(ERR_ARGUMENT, "keys & nonzeros vectors must have same length.")
        err = C.c_int()
        h = lib().orc_vec_build(_p(I), _p(V), C.c_int64(len(I)), C.c_int(combine), C.c_int64(0 if n is None else n),
                                C.c_int(0 if n is None else 1), C.byref(err))
        _check(err.value)
        super().__init__(h)

    def clone(self):
        return Vec(_h=lib().orc_vec_clone(self.h))

    def __setitem__(self, key, val):
        _check(lib().orc_vec_set(self.h, C.c_int64(key), C.c_double(val)))

    def __getitem__(self, key):
        return lib().orc_vec_get(self.h, C.c_int64(key))

    def set_many(self, keys, vals):
        keys, vals = _i64(keys), _f64(vals)
        _check(lib().orc_vec_set_many(self.h, _p(keys), _p(vals), C.c_int64(len(keys))))

    def set_batch_policy(self, keys, vals):
        keys, vals = _i64(keys), _f64(vals)
        _check(lib().orc_vec_set_batch_policy(self.h, _p(keys), _p(vals), C.c_int64(len(keys))))

    def get_many(self, keys):
        keys = _i64(keys)
        out = np.zeros(len(keys), np.float64)
        lib().orc_vec_get_many(self.h, _p(keys), C.c_int64(len(keys)), _p(out))
        return out

    def info(self):
        out = np.zeros(6, np.int64)
        lib().orc_vec_info(self.h, _p(out))
        return dict(capacity=int(out[0]), segment_capacity=int(out[1]), nb_segments=int(out[2]), nnz=int(out[3]),
                    height=int(out[4]), n=int(out[5]))

    def export(self):
        cap = self.info()["capacity"]
        tag, key, val = np.zeros(cap, np.uint8), np.zeros(cap, np.int64), np.zeros(cap, np.float64)
        lib().orc_vec_export(self.h, _p(tag), _p(key), _p(val))
        return tag, key, val

    def items(self):
        tag, key, val = self.export()
        m = tag.astype(bool)
        return key[m], val[m]

    def shrink_size(self):
        return lib().orc_vec_shrink_size(self.h)

    def __len__(self):
        return self.info()["n"]

    def __eq__(self, other):
        return bool(lib().orc_vec_equal(self.h, other.h))

    __hash__ = None


class Pcsc(_Handle):
    """Raw PackedCSC (pcsr.jl:4-9), partitions addressed by integer id."""
    _free = "orc_pcsc_free"

    def __init__(self, keys=(), values=(), combine=COMB_ADD, _h=None):
        if _h is not None:
            super().__init__(_h)
            return
        offs = np.zeros(len(keys) + 1, np.int64)
        for i, k in enumerate(keys):
            offs[i + 1] = offs[i] + len(k)
        fk = _i64(np.concatenate([_i64(k) for k in keys]) if len(keys) else [])
        fv = _f64(np.concatenate([_f64(v) for v in values]) if len(values) else [])
        err = C.c_int()
        h = lib().orc_pcsc_build(_p(fk), _p(fv), _p(offs), C.c_int64(len(keys)), C.c_int(combine), C.byref(err))
        _check(err.value)
        super().__init__(h)

    def clone(self):
        return Pcsc(_h=lib().orc_pcsc_clone(self.h))

    def __setitem__(self, idx, val):
        key, part = idx
        _check(lib().orc_pcsc_set(self.h, C.c_int64(key), C.c_int64(part), C.c_double(val)))

    def __getitem__(self, idx):
        key, part = idx
        out = C.c_double()
        _check(lib().orc_pcsc_get(self.h, C.c_int64(key), C.c_int64(part), C.byref(out)))
        return out.value

    def deletepartition(self, part):
        _check(lib().orc_pcsc_deletepartition(self.h, C.c_int64(part)))

    def info(self):
        out = np.zeros(7, np.int64)
        lib().orc_pcsc_info(self.h, _p(out))
        return dict(capacity=int(out[0]), segment_capacity=int(out[1]), nb_segments=int(out[2]), nb_elements=int(out[3]),
                    height=int(out[4]), nb_partitions=int(out[5]), nb_semaphores=int(out[6]),
                    nnz=int(out[3]) - int(out[5]))

    def export(self):
        inf = self.info()
        cap = inf["capacity"]
        tag, key, val = np.zeros(cap, np.uint8), np.zeros(cap, np.int64), np.zeros(cap, np.float64)
        sem = np.zeros(max(inf["nb_semaphores"], 1), np.int64)
        lib().orc_pcsc_export(self.h, _p(tag), _p(key), _p(val), _p(sem))
        return tag, key, val, sem[:inf["nb_semaphores"]]


class Matrix(_Handle):
    """DynamicSparseMatrix (matrix.jl) = col-major + row-major MappedPackedCSC (+ fill-mode buffer)."""
    _free = "orc_mat_free"

    def __init__(self, I=None, J=None, V=None, m=None, n=None, fill_mode=None, combine=COMB_ADD, _h=None):
        if _h is not None:
            super().__init__(_h)
            return
        if I is None:
            super().__init__(lib().orc_mat_empty(C.c_int(1 if (fill_mode is None or fill_mode) else 0)))
            return
        I, J, V = _i64(I), _i64(J), _f64(V)
        if not (len(I) == len(J) == len(V)):
            raise OracleError(ERR_ARGUMENT, "rows, columns, and nonzeros do not have same length.")
        err = C.c_int()
        given = m is not None or n is not None
        if given:   # each missing dimension defaults on its own: m = _guess_length(I), n = _guess_length(J)  (matrix.jl:15)
            m = int(m) if m is not None else (int(I.max()) if len(I) else 0)
            n = int(n) if n is not None else (int(J.max()) if len(J) else 0)
        h = lib().orc_mat_build(_p(I), _p(J), _p(V), C.c_int64(len(I)), C.c_int64(m or 0), C.c_int64(n or 0),
                                C.c_int(1 if given else 0), C.c_int(combine), C.byref(err))
        _check(err.value)
        super().__init__(h)

    def clone(self):
        return Matrix(_h=lib().orc_mat_clone(self.h))

    def __setitem__(self, idx, val):
        r, c = idx
        _check(lib().orc_mat_set(self.h, C.c_int64(r), C.c_int64(c), C.c_double(val)))

    def __getitem__(self, idx):
        r, c = idx
        out = C.c_double()
        _check(lib().orc_mat_get(self.h, C.c_int64(r), C.c_int64(c), C.byref(out)))
        return out.value

    def set_many(self, rows, cols, vals):
        rows, cols, vals = _i64(rows), _i64(cols), _f64(vals)
        _check(lib().orc_mat_set_many(self.h, _p(rows), _p(cols), _p(vals), C.c_int64(len(rows))))

    def set_batch_policy(self, rows, cols, vals):
        rows, cols, vals = _i64(rows), _i64(cols), _f64(vals)
        _check(lib().orc_mat_set_batch_policy(self.h, _p(rows), _p(cols), _p(vals), C.c_int64(len(rows))))

    def delete_columns_policy(self, cols):
        cols = _i64(cols)
        _check(lib().orc_mat_delete_columns_policy(self.h, C.c_int(0), _p(cols), C.c_int64(len(cols))))

    def delete_rows_policy(self, rows):
        rows = _i64(rows)
        _check(lib().orc_mat_delete_columns_policy(self.h, C.c_int(1), _p(rows), C.c_int64(len(rows))))

    def get_many(self, rows, cols, which=0):
        rows, cols = _i64(rows), _i64(cols)
        out = np.zeros(len(rows), np.float64)
        _check(lib().orc_mat_get_many(self.h, C.c_int(which), _p(rows), _p(cols), C.c_int64(len(rows)), _p(out)))
        return out

    def deletecolumn(self, col):
        _check(lib().orc_mat_deletecolumn(self.h, C.c_int64(col)))

    def deleterow(self, row):
        _check(lib().orc_mat_deleterow(self.h, C.c_int64(row)))

    def addrow(self, row, colids, vals):
        colids, vals = _i64(colids), _f64(vals)
        _check(lib().orc_mat_addrow(self.h, C.c_int64(row), _p(colids), _p(vals), C.c_int64(len(colids))))

    def closefillmode(self):
        _check(lib().orc_mat_closefillmode(self.h))

    def info(self, which=0):
        out = np.zeros(10, np.int64)
        lib().orc_mat_info(self.h, C.c_int(which), _p(out))
        return dict(capacity=int(out[0]), segment_capacity=int(out[1]), nb_segments=int(out[2]), nb_elements=int(out[3]),
                    height=int(out[4]), nb_partitions=int(out[5]), nb_semaphores=int(out[6]), m=int(out[7]), n=int(out[8]),
                    fillmode=bool(out[9]), nnz=int(out[3]) - int(out[5]))

    @property
    def size(self):
        inf = self.info(0)
        return inf["m"], inf["n"]

    def nnz(self):
        return self.info(1)["nnz"]   # matrix.jl:91 nnz(rowmajor)

    def export(self, which=0):
        inf = self.info(which)
        cap, ns = inf["capacity"], inf["nb_semaphores"]
        tag, key, val = np.zeros(cap, np.uint8), np.zeros(cap, np.int64), np.zeros(cap, np.float64)
        sem, ck, cl = np.zeros(max(ns, 1), np.int64), np.zeros(max(ns, 1), np.int64), np.zeros(max(ns, 1), np.uint8)
        lib().orc_mat_export(self.h, C.c_int(which), _p(tag), _p(key), _p(val), _p(sem), _p(ck), _p(cl))
        return dict(tag=tag, key=key, val=val, semaphores=sem[:ns], col_keys=ck[:ns], col_live=cl[:ns], **inf)

    def column(self, col, which=0):
        n = lib().orc_mat_column(self.h, C.c_int(which), C.c_int64(col), None, None, C.c_int64(0))
        k, v = np.zeros(max(n, 1), np.int64), np.zeros(max(n, 1), np.float64)
        lib().orc_mat_column(self.h, C.c_int(which), C.c_int64(col), _p(k), _p(v), C.c_int64(n))
        return k[:n], v[:n]

    def row(self, row):
        """view(matrix, row, :) = column `row` of the row-major twin (matrix.jl:70-81)."""
        return self.column(row, which=1)

    def row_scan(self, row, which=0):
        n = lib().orc_mat_row_scan(self.h, C.c_int(which), C.c_int64(row), None, None, C.c_int64(0))
        k, v = np.zeros(max(n, 1), np.int64), np.zeros(max(n, 1), np.float64)
        lib().orc_mat_row_scan(self.h, C.c_int(which), C.c_int64(row), _p(k), _p(v), C.c_int64(n))
        return k[:n], v[:n]

    def mul(self, xk, xv, trans=False):
        """mat * x (trans=False) or transpose(mat) * x; x = ascending (keys, vals). Returns sparse (keys, vals)."""
        xk, xv = _i64(xk), _f64(xv)
        cap = 1 << 16
        while True:
            yk, yv = np.zeros(cap, np.int64), np.zeros(cap, np.float64)
            err = C.c_int()
            n = lib().orc_mat_mul(self.h, C.c_int(1 if trans else 0), _p(xk), _p(xv), C.c_int64(len(xk)), _p(yk), _p(yv),
                                  C.c_int64(cap), C.byref(err))
            _check(err.value)
            if n <= cap:
                return yk[:n], yv[:n]
            cap = int(n)

    def mul_dense(self, x, ny, trans=False):
        x = _f64(x)
        y = np.zeros(ny, np.float64)
        _check(lib().orc_mat_mul_dense(self.h, C.c_int(1 if trans else 0), _p(x), C.c_int64(len(x)), _p(y), C.c_int64(ny)))
        return y
