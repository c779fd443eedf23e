"""CPU tests of libdsa's host logic and ABI surface (no compute calls: there is no GPU in the build container)."""
import ctypes as C

import numpy as np
import pytest

import dsa_b200 as D
from oracle import oracle as O
from oracle.oracle import Cells


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_library_loads_and_exports_every_declared_symbol():
    L = D.lib()
    names = D.declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(L, n), f"include/dsa.h declares {n} but libdsa.so does not export it"
    assert L.dsa_version() >= 100


def test_compute_calls_fail_loudly_without_gpu():
    if D.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(D.CudaError):
        D.dynamicsparsevec([1, 2], [1.0, 2.0])
    with pytest.raises(D.CudaError):
        D.dynamicsparse([1], [1], [1.0])
    with pytest.raises(D.CudaError):
        D.dynamicsparse(fill_mode=False)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 5, 6, 9, 11, 12, 100, 179, 180, 1000, 45875, 45876, 1_000_000, 10_100_000, 510_000,
                               126_250_000, 1_010_000_000])
def test_geometry_matches_oracle_and_survey(n):   # pma.jl:42-55,64
    out = np.zeros(4, np.int64)
    assert D.lib().dsa_pma_geometry(C.c_int64(n), _p(out)) == 0
    if n <= 1_000_000:
        g = O.geometry(n)
        assert out.tolist() == [g["capacity"], g["segment_capacity"], g["nb_segments"], g["height"]]
    table = {0: (256, 8, 32, 5), 1_000_000: (1 << 21, 16, 131072, 17), 10_100_000: (1 << 24, 16, 1 << 20, 20),
             510_000: (1 << 20, 16, 65536, 16), 1_010_000_000: (1 << 31, 16, 1 << 27, 27), 126_250_000: (1 << 28, 16, 1 << 24, 24)}
    if n in table:   # SURVEY.md §8 size table
        assert tuple(out.tolist()) == table[n]


def test_level_bounds_match_survey():   # integer bounds for seg 16
    mn, mx = np.zeros(21, np.int64), np.zeros(21, np.int64)
    assert D.lib().dsa_level_bounds(C.c_int64(16), C.c_int64(17), _p(mn), _p(mx)) == 0
    assert (mn[0], mx[0]) == (2, 14)
    assert (mn[1], mx[1]) == (3, 29)
    assert (mn[2], mx[2]) == (7, 57)
    assert (mn[17], mx[17]) == (629_146, 1_468_006)
    assert D.lib().dsa_level_bounds(C.c_int64(16), C.c_int64(20), _p(mn), _p(mx)) == 0
    assert (mn[20], mx[20]) == (5_033_165, 11_744_051)


def _oracle_spread_positions(c, m):
    a = Cells([(i + 1, 1.0) for i in range(m)] + [None] * (c - m))
    O.spread(a, 1, c, m)
    return np.nonzero(a.tag)[0]


def _check_spread(c, m):
    pos = _oracle_spread_positions(c, m)
    L = D.lib()
    got = np.array([L.dsa_spread_dest(C.c_int64(c), C.c_int64(m), C.c_int64(r)) for r in range(m)], dtype=np.int64)
    assert np.array_equal(got, pos), (c, m)
    occ = np.zeros(c, bool)
    occ[pos] = True
    ranks = np.array([L.dsa_spread_rank(C.c_int64(c), C.c_int64(m), C.c_int64(p)) for p in range(c)], dtype=np.int64)
    assert np.array_equal(ranks >= 0, occ), (c, m)
    assert np.array_equal(ranks[occ], np.arange(m)), (c, m)


def test_spread_closed_form_equals_reference_loop_exhaustive_small():   # moves.jl:120-140
    for lg in range(1, 8):
        c = 1 << lg
        for m in range(0, c + 1):
            _check_spread(c, m)
    for c in (100, 497, 500, 855, 1000):
        for m in (0, 1, c // 3, c // 2, int(c * 0.7), c - 1, c):
            _check_spread(c, m)


def test_spread_closed_form_rounding_cases():
    # power-of-two c with fl(c/e)*e rounding below c (e.g. c=64, e=49): the top gap is not at the window end
    _check_spread(64, 64 - 49)
    rng = np.random.default_rng(0)
    for _ in range(60):
        c = 1 << int(rng.integers(8, 14))
        m = int(rng.integers(int(0.08 * c), int(0.92 * c)))
        _check_spread(c, m)


def test_spread_closed_form_large_window_sampled():
    c, m = 1 << 22, 2_517_321
    pos = _oracle_spread_positions(c, m)
    L = D.lib()
    rng = np.random.default_rng(1)
    for r in rng.integers(0, m, 3000).tolist() + [0, 1, m - 2, m - 1]:
        assert L.dsa_spread_dest(C.c_int64(c), C.c_int64(m), C.c_int64(r)) == pos[r]
    occ = np.zeros(c, bool)
    occ[pos] = True
    for p in rng.integers(0, c, 3000).tolist() + [0, 1, c - 2, c - 1]:
        rk = L.dsa_spread_rank(C.c_int64(c), C.c_int64(m), C.c_int64(p))
        assert (rk >= 0) == occ[p]
        if rk >= 0:
            assert pos[rk] == p


def _plan(slot_key, slot_live, new_keys):
    sk, sl, nk = np.asarray(slot_key, np.int64), np.asarray(slot_live, np.uint8), np.asarray(new_keys, np.int64)
    n = len(sk) + len(nk)
    ok, ol, oo = np.zeros(max(n, 1), np.int64), np.zeros(max(n, 1), np.uint8), np.zeros(max(n, 1), np.int64)
    cnt = D.lib().dsa_colmap_plan(_p(sk), _p(sl), C.c_int64(len(sk)), _p(nk), C.c_int64(len(nk)), _p(ok), _p(ol), _p(oo))
    return ok[:cnt], ol[:cnt], oo[:cnt]


def test_colmap_plan_matches_addcolumn_replay():   # pcsr.jl:148-169 replayed by the oracle's batch policy
    rng = np.random.default_rng(2)
    for trial in range(40):
        ncols = int(rng.integers(1, 40))
        cols = np.sort(rng.choice(np.arange(1, 400, 3), ncols, replace=False))
        M = O.Matrix(np.ones(ncols, np.int64), cols, np.ones(ncols))
        ndel = int(rng.integers(0, ncols))
        dead = rng.choice(cols, ndel, replace=False)
        if ndel:
            M.delete_columns_policy(dead)
        e = M.export(0)
        slot_key, slot_live = e["col_keys"].copy(), e["col_live"].copy()
        live = set(slot_key[slot_live == 1].tolist())
        cand = [int(k) for k in rng.permutation(np.arange(0, 401)) if int(k) not in live and int(k) >= 1][: int(rng.integers(1, 30))]
        M.set_batch_policy(np.ones(len(cand), np.int64), cand, np.ones(len(cand)))
        e2 = M.export(0)
        ok, ol, oo = _plan(slot_key, slot_live, cand)
        assert ol.tolist() == e2["col_live"].tolist()
        assert ok[ol == 1].tolist() == e2["col_keys"][e2["col_live"] == 1].tolist()
        # old slot mapping: every previously live slot keeps its key
        for s, o in enumerate(oo.tolist()):
            if o > 0:
                assert slot_live[o - 1] == 1 and slot_key[o - 1] == ok[s]
        assert sorted(o for o in oo.tolist() if o > 0) == (np.nonzero(slot_live)[0] + 1).tolist()


def test_colmap_plan_order_dependence():
    ok, ol, _ = _plan([1, 0, 0, 9], [1, 0, 0, 1], [3, 5])
    assert [int(k) if l else None for k, l in zip(ok, ol)] == [1, 3, 5, 9]
    ok, ol, _ = _plan([1, 0, 0, 9], [1, 0, 0, 1], [5, 3])
    assert [int(k) if l else None for k, l in zip(ok, ol)] == [1, 3, 5, None, 9]
    ok, ol, _ = _plan([], [], [7, 3, 5])
    assert ok.tolist() == [3, 5, 7] and ol.tolist() == [1, 1, 1]
    ok, ol, _ = _plan([4, 0], [1, 0], [9])   # reuse of a trailing tombstone (reference bug (ii) — supported here)
    assert [int(k) if l else None for k, l in zip(ok, ol)] == [4, 9]


def test_python_buffer_matches_reference_semantics():   # buffer.jl:10-50
    b = D.Buffer()
    b.addelem(1, 2, 1.0)
    b.addelem(2, 1, 2.0)
    b.addelem(1, 7, 3.0)
    b.addelem(1, 2, 0.5)
    k, v = b.row(1)                     # test/unit/views.jl:46-67: view(buffer, 1, :) combines with +
    assert k.tolist() == [2, 7] and v.tolist() == [1.5, 3.0]
    b.addrow(5, [9, 3], [1.0, 2.0])
    assert b.rowmajor_coo[5] == ([3, 9], [2.0, 1.0])
    with pytest.raises(D.ErrorException):
        b.addrow(5, [1], [1.0])
    I, J, V = b.get_rowids_colids_vals()
    assert len(I) == b.length == 6
    assert sorted(zip(I.tolist(), J.tolist(), V.tolist())) == sorted([(1, 2, 1.0), (2, 1, 2.0), (1, 7, 3.0), (1, 2, 0.5), (5, 3, 2.0),
                                                                      (5, 9, 1.0)])




def test_bench_workload_is_periodic_and_distinct():
    """bench.py's workload: the batches form a cycle (batch k deletes what batch k-1 inserted, the initial matrix holds the last
    batch's inserts), so nnz is stationary and after every full cycle the stored keys are those of the start — checked here on
    the CPU oracle at a reduced size with the same generator."""
    import bench as B
    from oracle import oracle as O
    old = B.N_OVER
    B.N_OVER = 200
    try:
        x, coo, batches, ins_only = B.make_workload(4, m=3000, n=3000, nnz0=20_000, nb=4000, seed=11)
    finally:
        B.N_OVER = old
    lin = lambda i, j: (np.asarray(i) - 1) * 3000 + (np.asarray(j) - 1)
    init = lin(coo[0], coo[1])
    assert len(np.unique(init)) == len(init) == 20_000 - 200                # distinct (i, j); nnz0 - N_OVER entries
    assert len(x) == 3000 and all(len(b[0]) == 4000 for b in batches)
    M = O.Matrix(coo[0], coo[1], coo[2], m=3000, n=3000)
    nnz0 = M.nnz()

    def stored():
        e = M.export(1)
        mk = e["tag"].astype(bool) & (e["key"] > 0)
        return int(mk.sum())

    for rep in range(2):
        for b in batches:
            M.set_many(*b)
            assert M.nnz() == nnz0                                          # stationary, batch after batch
    assert stored() == M.nnz()
    # the insert-only batch holds new entries only
    assert len(np.intersect1d(lin(ins_only[0], ins_only[1]), init)) == 0
    before = M.nnz()
    M.set_many(*ins_only)
    assert M.nnz() == before + 4000
