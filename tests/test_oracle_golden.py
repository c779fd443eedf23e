"""Pins the CPU oracle against every known-answer assertion the reference's own
test-suite holds for the hot path.  Each test cites the reference
test file:line it restates.  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle.oracle import Cells


def N(*cells):
    return list(cells)


# ------------------------------------------------------------------ test/unit/finds.jl
def test_find_whole_array():   # finds.jl:4-24
    array = [None, (3, 10), (4, 10), None, (8, 10), None, (9, 10)]
    look = [1, 2, 3, 4, 5, 7, 8, 9, 100]
    exp = [(0, None), (0, None), (2, (3, 10)), (3, (4, 10)), (3, (4, 10)), (3, (4, 10)), (5, (8, 10)), (7, (9, 10)),
           (7, (9, 10))]
    for k, e in zip(look, exp):
        assert O.find(array, k) == e


def test_find_subarray():   # finds.jl:26-49
    array = [None, (3, 10), (4, 10), None, (8, 10), None, (9, 10)]
    look = [1, 2, 3, 4, 5, 7, 8, 9, 100]
    exp = [(3, (4, 10))] * 6 + [(5, (8, 10))] * 3
    for k, e in zip(look, exp):
        assert O.find(array, k, 4, 6) == e


def test_find_semaphore_regression():   # finds.jl:62-108
    array = [None, (3, 10), None, None, (9, 10), None, (10, 10), None, (3, 10), None, None, (2, 1), None]
    assert O.find(array, 4, 2, 8) == (2, (3, 10))
    assert O.find(array, 2, 2, 8) == (0, None)
    assert O.find(array, 4, 3, 8) == (2, (3, 10))
    assert O.find(array, 2, 3, 8) == (2, (3, 10))
    array = [None, (3, 10), None, (9, 10), None, (10, 10), (3, 10), (3, 10), None, (2, 1), None]
    assert O.find(array, 2, 3, 6) == (2, (3, 10))
    assert O.find(array, 2, 8, 7) == (7, (3, 10))
    assert O.find(array, 1, 9, 11) == (8, (3, 10))


# ------------------------------------------------------------------ test/unit/writes.jl
def test_insert():   # writes.jl:5-47
    a = Cells([(2, 10), None, (3, 10), (5, 10), (6, 10), None, (7, 10)])
    O.insert(a, 4, 10)
    assert a.tolist() == [(2, 10), None, (3, 10), (4, 10), (5, 10), (6, 10), (7, 10)]
    O.insert(a, 1, 10)
    assert a.tolist() == [(1, 10), (2, 10), (3, 10), (4, 10), (5, 10), (6, 10), (7, 10)]
    O.insert(a, 2, 11)
    assert a.tolist() == [(1, 10), (2, 11), (3, 10), (4, 10), (5, 10), (6, 10), (7, 10)]
    with pytest.raises(O.OracleError) as e:
        O.insert(a, 8, 10)
    assert e.value.code == O.ERR_ERROR
    a = Cells([(2, 10), None, None, (3, 10), (5, 10), (6, 10), None, (7, 10)])
    O.insert(a, 1, 10, 3, 6)
    assert a.tolist() == [(2, 10), (1, 10), None, (3, 10), (5, 10), (6, 10), None, (7, 10)]
    O.insert(a, 1, 11, 3, 6)
    assert a.tolist() == [(2, 10), (1, 10), (1, 11), (3, 10), (5, 10), (6, 10), None, (7, 10)]
    O.insert(a, 4, 10, 3, 6)
    assert a.tolist() == [(2, 10), (1, 10), (1, 11), (3, 10), (4, 10), (5, 10), (6, 10), (7, 10)]


def test_delete_purge():   # writes.jl:50-71
    a = Cells([(2, 10), (3, 10), None, (8, 10), (9, 10), None, (10, 10)])
    assert O.delete(a, 2) == (1, True)
    assert a.tolist() == [None, (3, 10), None, (8, 10), (9, 10), None, (10, 10)]
    assert O.delete(a, 2) == (0, False)
    assert a.tolist() == [None, (3, 10), None, (8, 10), (9, 10), None, (10, 10)]
    assert O.purge(a, 3, 5) == (4, 2)
    assert a.tolist() == [None, (3, 10), None, None, None, None, (10, 10)]


# ------------------------------------------------------------------ test/unit/comparison.jl
def test_arrays_equal():   # comparison.jl:2-10
    a1 = [None, (1, 1), None, None, (2, 1), None, (3, 2)]
    assert O.arrays_equal(a1, [None, (1, 1), (2, 1), (3, 2), None])
    assert not O.arrays_equal(a1, [None, (1, 1), (2, 1), (3, 2), (4, 2)])


# ------------------------------------------------------------------ test/unit/moves.jl (properties; Julia RNG not reproducible)
def _array_factory(rng, capacity, expnbempty, k):   # test/utils.jl:12-28
    cells, i = [], 1
    for _ in range(capacity):
        if rng.random() < expnbempty / capacity:
            cells.append(None)
        else:
            cells.append((i, int(rng.integers(1, 151))))
            i += int(rng.integers(1, k + 1))
    nbempty = sum(c is None for c in cells)
    return cells, nbempty, capacity - nbempty


def _partitioned_array_factory(rng, capacity, expnbempty, prob=0.05):   # test/utils.jl:41-66
    cells, sems, i, k = [], [], 1, 1
    for j in range(1, capacity + 1):
        p = rng.random()
        if p < expnbempty / capacity:
            cells.append(None)
        elif p > 1 - prob:
            cells.append((0, k))
            sems.append(j)
            i, k = 1, k + 1
        else:
            cells.append((i, int(rng.integers(1, 151))))
            i += 1
    nbempty = sum(c is None for c in cells)
    return cells, sems, nbempty, capacity - nbempty


def check_semaphores(cells, sems):   # test/utils.jl:68-92
    off = 0
    for pid, pos in enumerate(sems, start=1):
        if pos is not None and pos != 0:
            assert cells[pos - 1] == (0, pid)
            off += 1
    inarr = 0
    for pos, c in enumerate(cells, start=1):
        if c is not None and c[0] == 0:
            assert pos == sems[int(c[1]) - 1]
            inarr += 1
    assert off == inarr
    return off


def check_key_order(cells):   # test/utils.jl:94-113 (strictly increasing inside a partition; stronger than the reference's)
    pred = None
    for c in cells:
        if c is None:
            continue
        if c[0] == 0:
            pred = None
        else:
            if pred is not None:
                assert pred < c[0]
            pred = c[0]


def test_movecells_left_right():   # moves.jl:1-116
    rng = np.random.default_rng(1)
    cells, _, _ = _array_factory(rng, 20, 5, 3)
    to = next(i for i, c in enumerate(cells, 1) if c is None)
    frm = min(to + 4, 20)
    a = Cells(cells)
    O.move(a, False, frm, to)
    out = a.tolist()
    assert out[frm - 1] is None
    assert cells[to:frm] == out[to - 1:frm - 1]
    assert cells[frm:] == out[frm:]
    # move onto a non-empty cell -> ArgumentError
    to2 = next(i for i, c in enumerate(cells, 1) if c is not None)
    with pytest.raises(O.OracleError) as e:
        O.move(Cells(cells), False, min(to2 + 4, 20), to2)
    assert e.value.code == O.ERR_ARGUMENT
    # bounds
    for frm_, to_ in ((-1, 100), (1, 100)):
        for right in (False, True):
            with pytest.raises(O.OracleError) as e:
                O.move(Cells(cells), right, frm_, to_)
            assert e.value.code == O.ERR_BOUNDS
    # right
    to = max(i for i, c in enumerate(cells, 1) if c is None)
    frm = max(to - 4, 1)
    a = Cells(cells)
    O.move(a, True, frm, to)
    out = a.tolist()
    assert out[frm - 1] is None
    assert cells[frm - 1:to - 1] == out[frm:to]
    assert cells[:frm - 1] == out[:frm - 1]


def test_movecells_with_semaphores():   # moves.jl:44-57,102-116
    rng = np.random.default_rng(2)
    for right in (False, True):
        cells, sems, _, _ = _partitioned_array_factory(rng, 50, 20, 0.2)
        check_semaphores(cells, sems)
        a = Cells(cells)
        s = list(sems)
        if right:
            to = max(i for i, c in enumerate(cells, 1) if c is None)
            frm = max(to - 25, 1)
        else:
            to = next(i for i, c in enumerate(cells, 1) if c is None)
            frm = min(to + 25, 50)
        O.move(a, right, frm, to, s)
        check_semaphores(a.tolist(), s)


@pytest.mark.parametrize("capacity,expnbempty", [(100, 10), (1000, 8), (500, 11), (497, 97), (855, 17), (100000, 5961)])
def test_pack_spread(capacity, expnbempty):   # moves.jl:118-141, unitests.jl:17-22
    rng = np.random.default_rng(capacity)
    cells, nbempty, nbcells = _array_factory(rng, capacity, expnbempty, 1)
    a = Cells(cells)
    O.pack(a, 1, capacity, nbcells)
    out = a.tolist()
    for i in range(nbcells):
        assert out[i][0] == i + 1
    O.spread(a, 1, capacity, nbcells)
    out = a.tolist()
    live = [c for c in out if c is not None]
    assert [c[0] for c in live] == list(range(1, nbcells + 1))
    assert sum(c is None for c in out) == nbempty


def test_pack_spread_empty():   # moves.jl:143-150
    a = Cells([None] * 20)
    O.pack(a, 1, 20, 0)
    O.spread(a, 1, 20, 0, five_arg=True)
    assert a.tolist() == [None] * 20


@pytest.mark.parametrize("capacity,expnbempty", [(100, 10), (1000, 8), (500, 11), (497, 97), (855, 17), (100000, 5961)])
def test_pack_spread_with_semaphores(capacity, expnbempty):   # moves.jl:152-185
    rng = np.random.default_rng(capacity + 1)
    cells, sems, nbempty, nbcells = _partitioned_array_factory(rng, capacity, expnbempty)
    a = Cells(cells)
    s = list(sems)
    O.pack(a, 1, capacity, nbcells)
    O.spread(a, 1, capacity, nbcells, s)
    out = a.tolist()
    assert sum(c is None for c in out) == nbempty
    i = 1
    for c in out:
        if c is None:
            continue
        if c[0] != 0:
            assert c[0] == i
            i += 1
        else:
            i = 1
    check_semaphores(out, s)


# ------------------------------------------------------------------ test/functional/sparsevector.jl
def test_dynsparsevec_simple_use():   # sparsevector.jl:2-79
    vec = O.Vec([], [])
    assert len(vec) == 0
    I = [1, 2, 5, 5, 3, 10, 1, 8, 1, 5]
    V = [1.0, 3.5, 2.1, 8.5, 2.1, 1.1, 5.0, 7.8, 1.1, 2.0]
    vec = O.Vec(I, V)
    assert vec[1] == 1.0 + 1.1 + 5.0          # pins left-to-right input-order fold
    assert vec[2] == 3.5
    assert vec[3] == 2.1
    assert vec[4] == 0.0
    assert vec[5] == 2.1 + 8.5 + 2.0
    assert vec[8] == 7.8
    assert vec[10] == 1.1
    vec2 = O.Vec(I, V, combine=O.COMB_MUL)
    assert vec2[1] == 1.0 * 1.1 * 5.0
    assert vec2[2] == 3.5
    assert vec2[3] == 2.1
    assert vec2[5] == 2.1 * 8.5 * 2.0
    assert vec2[6] == 0.0
    assert vec2[8] == 7.8
    assert vec2[10] == 1.1
    assert len(vec) == 10
    assert vec.info()["capacity"] == 16 and vec.info()["nnz"] == 6   # repr test (skipped upstream): 16-element, 6 stored
    vec[1] = 0
    vec[2] = 0
    vec[3] = 0
    vec[22] = 0
    vec[1001] = 1.8
    vec[987] = 4.7
    vec[2] = 15 / 3
    vec[4] = 42
    assert vec[1] == 0
    assert vec[2] == 15 / 3
    assert vec[3] == 0
    assert vec[4] == 42
    assert vec[1001] == 1.8
    assert vec[987] == 4.7
    k, v = vec.items()
    exp = [(2, 5), (4, 42), (5, 12.6), (8, 7.8), (10, 1.1), (987, 4.7), (1001, 1.8)]
    assert list(zip(k.tolist(), v.tolist())) == exp
    assert len(vec) == 1001
    # Test 5: equality after shrink_size!
    vec1 = O.Vec([1, 2, 3, 5, 6, 8, 9], [1.0, 1.0, 1.0, 2.0, 1.0, 1.0, 3.0])
    vec2 = O.Vec([1, 2, 3, 5, 6, 8, 9, 10, 11], [1.0, 1.0, 1.0, 2.0, 1.0, 1.0, 3.0, 2.0, 3.0])
    assert not (vec1 == vec2)
    vec2[10] = 0
    vec2[11] = 0
    vec1.shrink_size()
    vec2.shrink_size()
    assert vec1 == vec2


def test_dynsparsevec_fill_empty():   # sparsevector.jl:88-119 (own RNG; nnz after every op)
    rng = np.random.default_rng(3)
    vec = O.Vec([], [])
    for n in (20, 100, 1000, 10000):
        keys = np.unique(rng.integers(1, 10_000_000_000, n))
        vals = rng.integers(10, 100000, len(keys)) / 10.0
        cnt = 0
        for k, v in zip(keys, vals):
            vec[int(k)] = float(v)
            cnt += 1
            assert vec.info()["nnz"] == cnt
        for k in keys:
            vec[int(k)] = 0.0
            cnt -= 1
            assert vec.info()["nnz"] == cnt
    assert vec.info()["nnz"] == 0


def test_dynsparsevec_insertions_and_gets():   # sparsevector.jl:121-161 (reduced sizes)
    rng = np.random.default_rng(4)
    keys = np.unique(rng.integers(1, 10_000_000_000, 200_000))
    vals = rng.integers(10, 100000, len(keys)) / 10.0
    perm = rng.permutation(len(keys))
    vec = O.Vec(keys[perm], vals[perm])
    assert np.array_equal(vec.get_many(keys), vals)
    k2 = np.unique(rng.integers(1, 10_000_000_000, 200_000))
    v2 = rng.integers(10, 100000, len(k2)) / 10.0
    p2 = rng.permutation(len(k2))
    vec.set_many(k2[p2], v2[p2])
    d = dict(zip(keys.tolist(), vals.tolist()))
    d.update(zip(k2.tolist(), v2.tolist()))
    kk = np.array(sorted(d))
    assert np.array_equal(vec.get_many(kk), np.array([d[int(k)] for k in kk]))
    gk, gv = vec.items()
    assert np.array_equal(gk, kk)
    vec = O.Vec(rng.integers(1, 100000, 10), rng.random(10) + 1)
    vec.set_many(np.arange(1, 100001), np.full(100000, 10.0))
    assert np.all(vec.get_many(np.arange(1, 100001)) == 10.0)


# ------------------------------------------------------------------ test/functional/sparsematrix.jl : PackedCSC
def _cells(tag, key, val):
    return [None if not t else (int(k), float(v)) for t, k, v in zip(tag, key, val)]


def test_pcsc_simple_use():   # sparsematrix.jl:1-121
    keys = [[1, 2, 3], [2, 6, 7], [1, 6, 8]]
    values = [[2, 3, 4], [2, 4, 5], [3, 5, 7]]
    p = O.Pcsc(keys, values)
    assert p.info()["nb_partitions"] == 3
    tag, key, val, sem = p.export()
    check_semaphores(_cells(tag, key, val), sem.tolist())
    check_key_order(_cells(tag, key, val))
    assert p.info()["nnz"] == 9
    matrix = np.array([[2, 0, 3], [3, 2, 0], [4, 0, 0], [0, 0, 0], [0, 0, 0], [0, 4, 5], [0, 5, 0], [0, 0, 7]], float)
    for i in range(8):
        for j in range(3):
            assert p[i + 1, j + 1] == matrix[i, j]
    matrix[0, 0] = 4
    p[1, 1] = 4
    matrix[0, 1] += 3
    p[1, 2] = p[1, 2] + 3
    matrix[2, 0] = 0
    p[3, 1] = 0
    matrix[3, 1] = 1
    p[4, 2] = 1
    assert p.info()["nnz"] == 10
    assert p.info()["nb_partitions"] == 3
    for i in range(8):
        for j in range(3):
            assert p[i + 1, j + 1] == matrix[i, j]
    p3 = p.clone()
    for i in range(1, 4):
        p3[2, i] = 0
    for i in range(1, 9):
        p3[i, 2] = 0
    assert all(p3[i, 2] == 0 for i in range(1, 9))
    p[10, 5] = 9                                   # A.6: new element and 2 new partitions
    assert p.info()["nnz"] == 11
    assert p.info()["nb_partitions"] == 5
    p[1, 4] = 2
    assert p.info()["nnz"] == 12
    tag, key, val, sem = p.export()
    check_semaphores(_cells(tag, key, val), sem.tolist())
    check_key_order(_cells(tag, key, val))
    nb = p.info()["nnz"]
    in2 = sum(1 for i in range(1, 11) if p[i, 2] != 0)
    p.deletepartition(2)                           # A.7
    assert p.info()["nb_partitions"] == 4
    assert p.info()["nnz"] == nb - in2
    tag, key, val, sem = p.export()
    assert check_semaphores(_cells(tag, key, val), sem.tolist()) == 4
    with pytest.raises(O.OracleError) as e:
        p[1, 2] = 1
    assert e.value.code == O.ERR_ERROR
    # Test B
    keys = [[1, 2, 3, 1, 2], [], [2, 6, 7, 7, 5], [1, 6, 8, 2, 1]]
    values = [[2, 3, 4, 1, 1], [], [2, 4, 5, 1, 1], [3, 5, 7, 1, 1]]
    p2 = O.Pcsc(keys, values)
    assert p2.info()["nb_partitions"] == 4
    assert p2.info()["nnz"] == 11
    m2 = np.array([[3, 0, 0, 4], [4, 0, 2, 1], [4, 0, 0, 0], [0, 0, 0, 0], [0, 0, 1, 0], [0, 0, 4, 5], [0, 0, 6, 0],
                   [0, 0, 0, 7]], float)
    for i in range(8):
        for j in range(4):
            assert p2[i + 1, j + 1] == m2[i, j]


def test_pcsc_derived_layout():   # SURVEY.md §8c derived golden (transliteration, not Julia output)
    p = O.Pcsc([[1, 2, 3], [2, 6, 7], [1, 6, 8]], [[2, 3, 4], [2, 4, 5], [3, 5, 7]])
    inf = p.info()
    assert (inf["capacity"], inf["segment_capacity"], inf["height"]) == (32, 4, 3)
    tag, key, val, sem = p.export()
    _ = None
    exp = [_, (0, 1), _, _, (1, 2), _, (2, 3), _, _, (3, 4), _, _, (0, 2), _, (2, 2), _, _, (6, 4), _, _, (7, 5), _, (0, 3), _, _,
           (1, 3), _, _, (6, 5), _, (8, 7), _]
    assert _cells(tag, key, val) == exp
    assert sem.tolist() == [2, 13, 23]
    v = O.Vec([1, 10, 3, 5, 3], [1.0, 2.4, 7.1, 1.1, 1.0])
    inf = v.info()
    assert (inf["capacity"], inf["segment_capacity"], inf["height"]) == (8, 2, 2)
    tag, key, val = v.export()
    assert _cells(tag, key, val) == [(1, 1.0), _, (3, 8.1), _, (5, 1.1), _, (10, 2.4), _]


# ------------------------------------------------------------------ sparsematrix.jl : DynamicSparseMatrix
def _check_matrix_invariants(M):
    for which in (0, 1):
        e = M.export(which)
        cells = _cells(e["tag"], e["key"], e["val"])
        n = check_semaphores(cells, e["semaphores"].tolist())
        check_key_order(cells)
        assert n == e["nb_partitions"]
    return M.export(0)["nb_partitions"], M.export(1)["nb_partitions"]


def test_dynsparsematrix_simple_use():   # sparsematrix.jl:174-285
    J = [1, 1, 1, 2, 2, 2, 3, 3, 3]
    I = [1, 2, 3, 2, 6, 7, 1, 6, 8]
    V = [2, 3, 4, 2, 4, 5, 3, 5, 7]
    M = O.Matrix(I, J, V)
    _check_matrix_invariants(M)
    assert M.info(0)["nnz"] == M.info(1)["nnz"] == M.nnz() == 9
    assert M.size[1] == 3
    m2 = np.array([[2, 0, 3], [3, 2, 0], [4, 0, 0], [0, 0, 0], [0, 0, 0], [0, 4, 5], [0, 5, 0], [0, 0, 7]], float)
    for i in range(8):
        for j in range(3):
            assert M[i + 1, j + 1] == m2[i, j]
    m2[0, 0] = 4
    M[1, 1] = 4
    m2[0, 1] += 3
    M[1, 2] = M[1, 2] + 3
    m2[2, 0] = 0
    M[3, 1] = 0
    m2[3, 1] = 1
    M[4, 2] = 1
    assert M.info(0)["nnz"] == M.info(1)["nnz"] == 10
    assert M.size == (8, 3)
    for i in range(8):
        for j in range(3):
            assert M[i + 1, j + 1] == m2[i, j]
    k, v = M.row(2)                      # A.5.1  matrix[2, :]
    assert len(k) == 2
    for j in range(3):
        assert dict(zip(k.tolist(), v.tolist())).get(j + 1, 0.0) == m2[1, j]
    k, v = M.column(2)                   # A.5.2
    assert len(k) == 5
    for i in range(8):
        assert dict(zip(k.tolist(), v.tolist())).get(i + 1, 0.0) == m2[i, 1]
    M[10, 5] = 9                         # A.6
    assert M.nnz() == 11
    assert M.info(0)["nb_partitions"] == 4
    M[1, -1] = 1
    M[1, 4] = 2
    M[3, 4] = 5
    assert M[1, 4] == 2 and M[3, 4] == 5 and M[1, -1] == 1
    assert M.info(0)["nb_partitions"] == 6
    _check_matrix_invariants(M)
    M.deletecolumn(2)                    # A.7
    assert M.info(1)["nb_partitions"] == 8
    assert M.info(0)["nb_partitions"] == 5
    assert _check_matrix_invariants(M) == (5, 8)
    M[1, 2] = 1
    assert M[1, 2] == 1
    assert _check_matrix_invariants(M) == (6, 8)


def test_dynsparsematrix_combine_and_deletions():   # sparsematrix.jl:288-299, 386-409
    I = [1, 1, 2, 4, 3, 5, 1, 3, 1, 5, 1, 5, 4]
    J = [4, 3, 3, 7, 18, 9, 3, 18, 4, 2, 3, 1, 7]
    V = [1, 8, 10, 2, -5, 3, 2, 1, 1, 1, 5, 3, 2]
    M = O.Matrix(I, J, V)
    assert M[1, 4] == 1 + 1
    assert M[1, 3] == 8 + 2 + 5
    assert M[4, 7] == 2 + 2
    assert M[3, 18] == -5 + 1
    assert M[5, 9] == 3 and M[5, 2] == 1 and M[5, 1] == 3 and M[2, 3] == 10
    _check_matrix_invariants(M)
    M.deletecolumn(3)
    _check_matrix_invariants(M)
    for i in range(1, 6):
        assert M[i, 3] == 0


def test_dynsparsematrix_insertions_and_gets():   # sparsematrix.jl:341-383
    M = O.Matrix([1, 4, 3, 5], [4, 7, 18, 9], [1, 2, -5, 3])
    M[2, 7] = 8
    assert M[2, 7] == 8
    M[1, 2] = 21
    assert M[1, 2] == 21
    M[10, 33] = 21
    assert M[10, 33] == 21
    M[55, 54] = 53
    assert M[55, 54] == 53
    assert M[1, 4] == 1 and M[4, 7] == 2 and M[3, 18] == -5 and M[5, 9] == 3
    rng = np.random.default_rng(5)
    nb_rows, nb_cols = 340, 1000
    mask = rng.random((nb_rows, nb_cols)) <= 0.05
    I, J = np.nonzero(mask)
    I, J = I + 1, J + 1
    V = rng.integers(0, 10_000_000, len(I)) / 10000.0
    M = O.Matrix(I, J, V)
    assert np.array_equal(M.get_many(I, J), V)
    cols = np.arange(nb_cols, 2001)
    M.set_many(np.ones(len(cols), np.int64), cols, np.ones(len(cols)))
    assert np.all(M.get_many(np.ones(len(cols), np.int64), cols) == 1)
    _check_matrix_invariants(M)


def test_dynsparsematrix_fill_mode():   # sparsematrix.jl:412-519
    M = O.Matrix()
    values = np.array([[1, 0, 0, 2, 0, 7, 0, 0, 0, 9, 1, 2], [0, 3, 0, 0, 1, 1, 0, 0, 0, 1, 0, 2], [0, 0, 0, 1, 1, 2, 0, 0, 1, 2, 0, 0],
                       [0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 1], [1, 2, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0]], float)
    for i in range(5):
        colids = np.nonzero(values[i])[0] + 1
        M.addrow(i + 1, colids, values[i, colids - 1])
    with pytest.raises(O.OracleError) as e:   # buffer.jl:13
        M.addrow(1, [1], [1.0])
    assert e.value.code == O.ERR_ERROR
    M[1, 2] = 2
    M[1, 1] = 1   # in fill mode it adds to the current value
    values[0, 1] = 2
    values[0, 0] += 1
    M.closefillmode()
    for i in range(5):
        for j in range(12):
            assert M[i + 1, j + 1] == values[i, j]
            assert M.get_many([i + 1], [j + 1], which=1)[0] == values[i, j]
    M.addrow(7, [1, 3, 4, 5], [2, 3, 6, 7])
    for j, v in zip([1, 3, 4, 5], [2, 3, 6, 7]):
        assert M[7, j] == v
    # second test: fill mode sums duplicates like SparseArrays.sparse
    rng = np.random.default_rng(6)
    row, col = rng.integers(1, 101, 3000), rng.integers(1, 101, 3000)
    vals = rng.integers(1, 100001, 3000).astype(float)
    M = O.Matrix()
    dense = np.zeros((100, 100))
    for r, c, v in zip(row, col, vals):
        M[int(r), int(c)] = float(v)
        dense[r - 1, c - 1] += v
    M.closefillmode()
    rr, cc = np.meshgrid(np.arange(1, 101), np.arange(1, 101), indexing="ij")
    assert np.array_equal(M.get_many(rr.ravel(), cc.ravel()), dense.ravel())
    assert np.array_equal(M.get_many(rr.ravel(), cc.ravel(), which=1), dense.ravel())
    # third test: no fill mode -> last writer wins
    M = O.Matrix(fill_mode=False)
    dense = np.zeros((100, 100))
    for r, c, v in zip(row, col, vals):
        M[int(r), int(c)] = float(v)
        dense[r - 1, c - 1] = v
    assert np.array_equal(M.get_many(rr.ravel(), cc.ravel()), dense.ravel())
    assert np.array_equal(M.get_many(rr.ravel(), cc.ravel(), which=1), dense.ravel())
    # fourth / fifth
    M4 = O.Matrix()
    M4.closefillmode()
    assert M4.nnz() == 0 and M4[1, 1] == 0
    M5 = O.Matrix(fill_mode=False)
    with pytest.raises(O.OracleError) as e:
        M5.closefillmode()
    assert e.value.code == O.ERR_ERROR


def test_readme_example():   # README.md:19-40 + SURVEY.md §8c derived column structure
    v = O.Vec([1, 10, 3, 5, 3], [1.0, 2.4, 7.1, 1.1, 1.0])
    assert v[3] == 7.1 + 1.0
    v[78] = 1.5
    assert v[2] == 0
    v[2] = 0
    M = O.Matrix([1, 2, 3, 2, 6, 7, 1, 6, 8], [1, 1, 1, 2, 2, 2, 3, 3, 3], [2, 3, 4, 2, 4, 5, 3, 5, 7])
    e1 = M.export(1)
    assert e1["semaphores"].tolist() == [2, 8, 14, 19, 25, 29]
    assert e1["col_keys"].tolist() == [1, 2, 3, 6, 7, 8]
    M[4, 1] = 1
    M[2, 2] = 0
    M.deletecolumn(2)
    assert M[2, 6] == 0
    e0, e1 = M.export(0), M.export(1)
    assert [s if l else None for s, l in zip(e0["semaphores"].tolist(), e0["col_live"].tolist())] == [2, None, 19]
    assert [k if l else None for k, l in zip(e0["col_keys"].tolist(), e0["col_live"].tolist())] == [1, None, 3]
    assert e1["col_keys"].tolist() == [1, 2, 3, 4, 6, 7, 8]
    assert e1["semaphores"].tolist() == [2, 8, 14, 19, 21, 25, 29]


# ------------------------------------------------------------------ test/unit/views.jl
def test_views():   # views.jl:3-32,35-44
    I = [1, 1, 2, 4, 3, 5, 1, 4, 1, 5, 1, 5, 4, 4, 3, 9, 1]
    J = [4, 3, 3, 7, 18, 9, 3, 18, 4, 2, 3, 1, 7, 3, 3, 3, 18]
    V = [1, 8, 10, 2, -5, 3, 2, 1, 1, 1, 5, 3, 2, 1, 7, 8, 1]
    M = O.Matrix(I, J, V)
    k, v = M.row(5)
    assert k.tolist() == [1, 2, 9] and v.tolist() == [3, 1, 3]
    k, v = M.column(3)
    assert k.tolist() == [1, 2, 3, 4, 9] and v.tolist() == [15, 10, 7, 1, 8]
    k, v = M.column(18)
    assert k.tolist() == [1, 3, 4] and v.tolist() == [1, -5, 1]


# ------------------------------------------------------------------ test/unit/spmv.jl (integer-key cases) + math.jl
def _dict(k, v):
    return dict(zip(k.tolist(), v.tolist()))


def test_spmv_3():   # spmv.jl:61-81
    M = O.Matrix([1, 1, 3, 3, 4, 4, 4, 6, 6, 6], [2, 4, 1, 3, 1, 3, 6, 1, 3, 6], [1, 2, 1, 1, 1, 2, 1, 1, 1, 1])
    r = _dict(*M.mul([2, 5, 6], [1, 1, 1]))
    assert r[1] == 1 and r.get(2, 0.0) == 0 and r.get(3, 0.0) == 0 and r[4] == 1 and r.get(5, 0.0) == 0 and r[6] == 1


def test_spmv_4():   # spmv.jl:84-127
    M = O.Matrix([1, 1, 3, 3, 4, 4, 4, 6, 6, 6], [2, 4, 1, 3, 1, 3, 6, 1, 3, 6], [1, 2, 1, 1, 1, 1, 1, 1, 1, 1])
    x = ([2, 3, 5, 6], [1, 1, 1, 1])
    r = _dict(*M.mul(*x))
    assert [r.get(i, 0.0) for i in range(1, 7)] == [1, 0, 1, 2, 0, 2]
    M.deletecolumn(3)
    r = _dict(*M.mul(*x))
    assert [r.get(i, 0.0) for i in range(1, 7)] == [1, 0, 0, 1, 0, 1]
    M.deleterow(4)
    r = _dict(*M.mul(*x))
    assert [r.get(i, 0.0) for i in range(1, 7)] == [1, 0, 0, 0, 0, 1]


def test_spmv_transposed_mapped():   # spmv.jl:29-58 with Char rows mapped to ints a..e -> 1..5
    I = [1, 1, 1, 2, 2, 3, 4, 4, 4]
    J = [1, 3, 5, 2, 4, 4, 1, 4, 5]
    V = [1, 2, 1, 2, 1, 3, 3, 2, 2]
    M = O.Matrix(I, J, V)
    r = _dict(*M.mul([1, 3, 5], [1, 1, 1]))                # test_spmv_1
    assert r[1] == 4 and r.get(2, 0.0) == 0 and r.get(3, 0.0) == 0 and r[4] == 5
    r = _dict(*M.mul([1, 3, 5], [1, 1, 1], trans=True))    # test_spmv_2: rows a, c, e
    assert r[1] == 1 and r.get(2, 0.0) == 0 and r[3] == 2 and r[4] == 3 and r[5] == 1


def test_spmv_vs_dense():   # math.jl:1-51 (all operand orders reduce to these two products)
    rng = np.random.default_rng(7)
    row, col = rng.integers(1, 111, 50), rng.integers(1, 101, 50)
    vals = rng.integers(1, 11, 50).astype(float)
    M = O.Matrix(row, col, vals, m=110, n=100)
    D = np.zeros((110, 100))
    for r, c, v in zip(row, col, vals):
        D[r - 1, c - 1] += v
    xr = np.unique(rng.integers(1, 101, 25))
    xv = rng.integers(1, 11, len(xr)).astype(float)
    x = np.zeros(100)
    x[xr - 1] = xv
    yk, yv = M.mul(xr, xv)
    y = np.zeros(110)
    y[yk - 1] = yv
    assert np.array_equal(y, D @ x)
    xr2 = np.unique(rng.integers(1, 111, 25))
    xv2 = rng.integers(1, 11, len(xr2)).astype(float)
    x2 = np.zeros(110)
    x2[xr2 - 1] = xv2
    yk, yv = M.mul(xr2, xv2, trans=True)
    y = np.zeros(100)
    y[yk - 1] = yv
    assert np.array_equal(y, D.T @ x2)


def test_reference_bugs_are_reproduced():   # SURVEY.md §7 hard parts (i), (ii)
    M = O.Matrix([1, 1, 1], [1, 2, 3], [1.0, 1.0, 1.0])
    M.deletecolumn(3)
    with pytest.raises(O.OracleError) as e:    # (ii) append after a trailing tombstone -> BoundsError semaphores[0]
        M[1, 4] = 1.0
    assert e.value.code == O.ERR_BOUNDS
    M = O.Matrix([1, 1, 1, 1], [1, 3, 5, 7], [1.0, 1.0, 1.0, 1.0])
    M.deletecolumn(5)
    with pytest.raises(O.OracleError) as e:    # (i) mid-insert with a deleted partition to its right -> @assert
        M[1, 2] = 1.0
    assert e.value.code == O.ERR_ASSERT
