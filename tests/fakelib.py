"""TEST INFRASTRUCTURE — a test double for libdsa.so.

The CPU suite (`-m "not gpu"`) has no GPU, yet the host glue in dynamicsparsearrays.jl_b200/api.py (pending-write queue,
fill-mode buffer, key codecs, operand-order wrappers, exception mapping, two-call size queries) is logic worth testing there.
FakeLib answers the subset of include/dsa.h that api.py calls, with the same argument conventions (ctypes pointers + sizes,
int return code + dsa_last_error), executing on the CPU oracle.  It lives under tests/ and is installed only by the
`api` fixture of tests/test_zz_api_hostlogic.py through monkeypatching; the package itself never imports it, and the same
tests run against the real library under `-m gpu`.
"""
import ctypes as C

import numpy as np

from oracle import oracle as O

_COMBINE = {0: O.COMB_ADD, 1: O.COMB_MUL, 2: O.COMB_LAST, 3: O.COMB_FIRST, 4: O.COMB_MIN, 5: O.COMB_MAX}
# oracle error code -> DSA_ERR_* (include/dsa.h); the oracle's ERR_ASSERT (a reference bug reproduced) maps to an internal error
_ERR = {O.ERR_ARGUMENT: 1, O.ERR_BOUNDS: 2, O.ERR_ERROR: 3, O.ERR_ASSERT: 12}


def _val(x):
    return x.value if hasattr(x, "value") else x


def _arr(ptr, n, dtype):
    """numpy view of the caller's buffer behind a ctypes pointer"""
    n = int(_val(n))
    addr = _val(ptr)
    if n == 0 or not addr:
        return np.zeros(0, dtype)
    ct = {np.int64: C.c_int64, np.float64: C.c_double, np.uint8: C.c_uint8}[dtype]
    return np.ctypeslib.as_array((ct * n).from_address(addr))


def _set(ref, value):
    ref._obj.value = value   # ref = ctypes.byref(x)


class FakeLib:
    def __init__(self):
        self._objs = {}
        self._next = 1
        self._err = b""
        self.calls = []          # (entry point, batch length) of every mutating call: lets tests check the batching of the glue

    # ---- plumbing -------------------------------------------------------------------------------------------------
    def _new(self, obj, out):
        hid = self._next
        self._next += 1
        self._objs[hid] = obj
        _set(out, hid)
        return 0

    def _get(self, h):
        return self._objs[_val(h)]

    def _guard(self, fn):
        try:
            fn()
            return 0
        except O.OracleError as ex:
            self._err = str(ex).encode()
            return _ERR.get(ex.code, 12)

    def dsa_last_error(self):
        return self._err

    # ---- vectors ---------------------------------------------------------------------------------------------------
    def dsa_vec_build(self, keys, vals, n, combine, length, len_given, out):
        def run():
            v = O.Vec(_arr(keys, n, np.int64).copy(), _arr(vals, n, np.float64).copy(), _COMBINE[_val(combine)],
                      _val(length) if _val(len_given) else None)
            self._new(v, out)
        return self._guard(run)

    def dsa_vec_destroy(self, h):
        self._objs.pop(_val(h), None)
        return 0

    def dsa_vec_clone(self, h, out):
        return self._new(self._get(h).clone(), out)

    def dsa_vec_set_batch(self, h, keys, vals, n):
        self.calls.append(("dsa_vec_set_batch", int(_val(n))))
        return self._guard(lambda: self._get(h).set_batch_policy(_arr(keys, n, np.int64).copy(), _arr(vals, n, np.float64).copy()))

    def dsa_vec_get_batch(self, h, keys, n, out):
        _arr(out, n, np.float64)[:] = self._get(h).get_many(_arr(keys, n, np.int64).copy())
        return 0

    def dsa_vec_info(self, h, out6):
        i = self._get(h).info()
        _arr(out6, 6, np.int64)[:] = [i["capacity"], i["segment_capacity"], i["nb_segments"], i["nnz"], i["height"], i["n"]]
        return 0

    def dsa_vec_nonzeros(self, h, keys_out, vals_out, cap, count_out):
        k, v = self._get(h).items()
        _set(count_out, len(k))
        if _val(cap) >= len(k) and len(k):
            _arr(keys_out, len(k), np.int64)[:] = k
            _arr(vals_out, len(k), np.float64)[:] = v
        return 0

    def dsa_vec_shrink_size(self, h, n_out):
        _set(n_out, int(self._get(h).shrink_size()))
        return 0

    def dsa_vec_export(self, h, occ, keys, vals):
        t, k, v = self._get(h).export()
        _arr(occ, len(t), np.uint8)[:] = t
        _arr(keys, len(t), np.int64)[:] = k
        _arr(vals, len(t), np.float64)[:] = v
        return 0

    # ---- matrices --------------------------------------------------------------------------------------------------
    def dsa_matrix_create(self, out):
        return self._new(O.Matrix(fill_mode=False), out)

    def dsa_matrix_build_coo(self, rows, cols, vals, n, m, ncols, dims_given, combine, out):
        def run():
            given = bool(_val(dims_given))
            A = O.Matrix(_arr(rows, n, np.int64).copy(), _arr(cols, n, np.int64).copy(), _arr(vals, n, np.float64).copy(),
                         _val(m) if given else None, _val(ncols) if given else None, combine=_COMBINE[_val(combine)])
            self._new(A, out)
        return self._guard(run)

    def dsa_matrix_destroy(self, h):
        self._objs.pop(_val(h), None)
        return 0

    def dsa_matrix_clone(self, h, out):
        return self._new(self._get(h).clone(), out)

    def dsa_matrix_set_batch(self, h, rows, cols, vals, n):
        self.calls.append(("dsa_matrix_set_batch", int(_val(n))))
        return self._guard(lambda: self._get(h).set_batch_policy(_arr(rows, n, np.int64).copy(), _arr(cols, n, np.int64).copy(),
                                                                 _arr(vals, n, np.float64).copy()))

    def dsa_matrix_stage_batch(self, h, rows, cols, vals, n):
        A = self._get(h)
        if getattr(A, "_staged", None) is None:
            A._staged = []
        if len(A._staged) >= 2:
            self._err = b"both staging slots are in use"
            return 3
        A._staged.append((_arr(rows, n, np.int64).copy(), _arr(cols, n, np.int64).copy(), _arr(vals, n, np.float64).copy()))
        return 0

    def dsa_matrix_apply_staged(self, h):
        A = self._get(h)
        if not getattr(A, "_staged", None):
            self._err = b"no staged batch"
            return 3
        r, c, v = A._staged.pop(0)
        self.calls.append(("dsa_matrix_apply_staged", len(r)))
        return self._guard(lambda: A.set_batch_policy(r, c, v))

    def dsa_matrix_get_batch(self, h, which, rows, cols, n, out):
        def run():
            _arr(out, n, np.float64)[:] = self._get(h).get_many(_arr(rows, n, np.int64).copy(), _arr(cols, n, np.int64).copy(),
                                                                which=_val(which))
        return self._guard(run)

    def dsa_matrix_delete_columns(self, h, cols, n):
        self.calls.append(("dsa_matrix_delete_columns", int(_val(n))))
        return self._guard(lambda: self._get(h).delete_columns_policy(_arr(cols, n, np.int64).copy()))

    def dsa_matrix_delete_rows(self, h, rows, n):
        self.calls.append(("dsa_matrix_delete_rows", int(_val(n))))
        return self._guard(lambda: self._get(h).delete_rows_policy(_arr(rows, n, np.int64).copy()))

    def _span(self, k, v, keys_out, vals_out, cap, count_out):
        _set(count_out, len(k))
        if _val(cap) >= len(k) and len(k) and _val(keys_out):
            _arr(keys_out, len(k), np.int64)[:] = k
            _arr(vals_out, len(k), np.float64)[:] = v
        return 0

    def dsa_matrix_column(self, h, col, keys_out, vals_out, cap, count_out):
        k, v = self._get(h).column(_val(col), which=0)
        return self._span(k, v, keys_out, vals_out, cap, count_out)

    def dsa_matrix_row(self, h, row, keys_out, vals_out, cap, count_out):
        k, v = self._get(h).row(_val(row))
        return self._span(k, v, keys_out, vals_out, cap, count_out)

    def dsa_matrix_spmv(self, h, trans, xk, xv, nx, yk, yv, cap, count_out):
        def run():
            k, v = self._get(h).mul(_arr(xk, nx, np.int64).copy(), _arr(xv, nx, np.float64).copy(), trans=bool(_val(trans)))
            self._span(k, v, yk, yv, cap, count_out)
        return self._guard(run)

    def dsa_matrix_spmv_dense(self, h, trans, x, nx, y, ny):
        def run():
            if _val(ny):
                _arr(y, ny, np.float64)[:] = self._get(h).mul_dense(_arr(x, nx, np.float64).copy(), _val(ny), trans=bool(_val(trans)))
        return self._guard(run)

    def dsa_matrix_info(self, h, which, out10):
        i = self._get(h).info(_val(which))
        _arr(out10, 10, np.int64)[:] = [i["capacity"], i["segment_capacity"], i["nb_segments"], i["nb_elements"], i["height"],
                                       i["nb_partitions"], i["nb_semaphores"], i["m"], i["n"], i["nnz"]]
        return 0

    def dsa_matrix_export(self, h, which, occ, keys, vals, sem, ck, cl):
        e = self._get(h).export(_val(which))
        cap, ns = e["capacity"], e["nb_semaphores"]
        _arr(occ, cap, np.uint8)[:] = e["tag"]
        _arr(keys, cap, np.int64)[:] = e["key"]
        _arr(vals, cap, np.float64)[:] = e["val"]
        if ns:
            _arr(sem, ns, np.int64)[:] = e["semaphores"]
            _arr(ck, ns, np.int64)[:] = e["col_keys"]
            _arr(cl, ns, np.uint8)[:] = e["col_live"]
        return 0
