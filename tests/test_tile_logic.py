"""CPU checks of the two places where the tile-streamed pipeline (csrc/tile.cuh) does NOT run the reference's gapped binary
search literally, against the oracle's `find` (oracle/dsa_oracle.cpp, a restatement of finds.jl:29-57 pinned by the reference's
own vectors in test_oracle_golden.py):

  * `tile_find`: inside a tile the search first picks the leaf (first stored cell of the leaf in range and <= key: monotone),
    then scans the leaf's live mask from the right;
  * `k_tile_assign`: the tile of an op whose partition span straddles tile borders is found by probing the first stored cell at
    or after a border (binary search over the borders).

Both are restated here in plain Python, statement for statement, and compared with `find` on random gapped arrays: the answer
(hit or predecessor) is a property of the array's contents, so any correct search must agree with the reference's.
The CUDA code itself is compared with the oracle in tests/test_gpu_tile.py (-m gpu)."""
import numpy as np
import pytest

from oracle import oracle as O

GAP = None


def _gapped(rng, ncells, density, first_key=1):
    """sorted keys spread over ncells cells with gaps; returns the cell list (key or None) — keys strictly increasing"""
    live = rng.random(ncells) < density
    n = int(live.sum())
    keys = first_key + np.cumsum(rng.integers(1, 5, n))
    cells, it = [], iter(keys)
    for l in live:
        cells.append(int(next(it)) if l else GAP)
    return cells


def tile_find(cells, lgS, key, lo, hi):
    """csrc/tile.cuh: tile_find<LGS> (0-based, inclusive range); returns (pos, hit)"""
    S = 1 << lgS
    nl = len(cells) >> lgS
    live = [sum(1 << q for q in range(S) if cells[(l << lgS) + q] is not GAP) for l in range(nl)]
    fpos = []
    for l in range(nl):   # first stored cell at or after the leaf's start (len(cells) = none)
        p = l << lgS
        while p < len(cells) and cells[p] is GAP:
            p += 1
        fpos.append(p)
    if lo <= hi:
        a, b = lo >> lgS, hi >> lgS
        best = a
        a += 1
        while a <= b:
            mid = (a + b) >> 1
            p = fpos[mid]
            if p <= hi and cells[p] <= key:
                best, a = mid, mid + 1
            else:
                b = mid - 1
        c0 = best << lgS
        m = live[best]
        if lo > c0:
            m &= ~((1 << (lo - c0)) - 1)
        if hi < c0 + S - 1:
            m &= (2 << (hi - c0)) - 1
        while m:
            q = m.bit_length() - 1
            k = cells[c0 + q]
            if k <= key:
                return c0 + q, k == key
            m ^= 1 << q
        hi = lo - 1
    i = hi
    while i > 0 and cells[i] is GAP:
        i -= 1
    return max(i, 0), False


@pytest.mark.parametrize("lgS", [3, 4, 5])
@pytest.mark.parametrize("density", [0.15, 0.6, 0.92])
def test_leaf_first_search_equals_the_reference_find(lgS, density):
    rng = np.random.default_rng(100 * lgS + int(density * 100))
    ncells = 256
    for trial in range(60):
        cells = _gapped(rng, ncells, density)
        if cells[0] is GAP:
            cells[0] = 0   # a partition semaphore (key 0): every search of a tile has a stored cell at or before its range
        oc = O.Cells([None if c is GAP else (c, 1.0) for c in cells])
        stored = [c for c in cells if c is not GAP]
        for _ in range(40):
            lo = int(rng.integers(0, ncells))
            hi = int(rng.integers(lo, ncells)) if rng.random() < 0.9 else lo - 1   # lo > hi: an empty span
            key = int(rng.choice(stored)) if rng.random() < 0.5 else int(rng.integers(1, stored[-1] + 3))
            pos, hit = tile_find(cells, lgS, key, lo, hi)
            opos, ocell = O.find(oc, key, lo + 1, hi + 1)   # 1-based, inclusive
            if hi < lo:   # the reference is never asked for an empty range: the predecessor is the stored cell left of it
                exp = max((p for p in range(0, lo) if cells[p] is not GAP), default=0)
                assert pos == exp and not hit
                continue
            assert opos >= 1, "a stored cell exists at or before every range of this test"
            assert pos == opos - 1, (lgS, lo, hi, key, pos, opos)
            if pos >= lo:   # below the range (finds.jl:49-56: the nearest stored cell to its left) a hit is never reported: in
                assert hit == (ocell[0] == key)   # the callers' ranges that cell is the partition's semaphore (pcsr.jl:305-307)


def tile_of_op(cells, tile_lg, ps, pe, key, is_set):
    """csrc/tile.cuh: k_tile_assign — tile of the predecessor cell and the tile-local search range.
    cells: the whole array; [ps, pe) = span of the op's partition (ps = its semaphore)."""
    T = 1 << tile_lg
    frm = ps + 1 if is_set else ps
    to = pe - 1
    t, lo, hi = ps >> tile_lg, frm, to
    t_last = to >> tile_lg
    if t != t_last:
        blo, bhi = t + 1, t_last
        while blo <= bhi:
            u = (blo + bhi) >> 1
            p = u << tile_lg
            kk = cells[p]
            while kk is GAP and p < to:
                p += 1
                kk = cells[p]
            if kk is not GAP and kk <= key:
                t, blo = u, u + 1
            else:
                bhi = u - 1
        tb0 = t << tile_lg
        lo = max(frm, tb0)
        hi = min(to, tb0 + T - 1)
    tb = t << tile_lg
    return t, lo - tb, hi - tb


@pytest.mark.parametrize("tile_lg,lgS", [(6, 3), (7, 4)])
def test_border_probes_find_the_tile_of_the_predecessor(tile_lg, lgS):
    """partitions of 1 .. 6 tiles laid one after the other ([semaphore, keys...]); for every op the tile chosen by the border
    probes must hold the cell the reference's find returns over the whole span, and the tile-local search must return that cell"""
    rng = np.random.default_rng(7 + tile_lg)
    T = 1 << tile_lg
    for trial in range(25):
        cells, spans = [], []
        for part in range(int(rng.integers(2, 7))):
            length = int(rng.integers(3, 6 * T))
            body = _gapped(rng, length - 1, float(rng.choice([0.2, 0.6, 0.9])))
            spans.append((len(cells), len(cells) + length))
            cells += [0] + body   # key 0 = the partition's semaphore cell (pcsr.jl:23)
        cells += [GAP] * (-len(cells) % T)
        oc = O.Cells([None if c is GAP else (c, 1.0) for c in cells])
        for ps, pe in spans:
            stored = [c for c in cells[ps + 1:pe] if c is not GAP]
            top = (stored[-1] if stored else 5) + 3
            for _ in range(30):
                key = int(rng.choice(stored)) if stored and rng.random() < 0.5 else int(rng.integers(1, top))
                for is_set in (True, False):
                    frm = ps + 1 if is_set else ps
                    # reference: find over the span (pcsr.jl:305-307: (sem, end] for inserts, [sem, end] for deletes)
                    opos, ocell = O.find(oc, key, frm + 1, pe) if frm <= pe - 1 else (ps + 1, (0, 1.0))
                    ref = max(opos - 1, ps)   # an insert below every key of the span goes right after the semaphore
                    t, lo, hi = tile_of_op(cells, tile_lg, ps, pe, key, is_set)
                    assert t == ref >> tile_lg, (ps, pe, key, is_set, t, ref)
                    pos, hit = tile_find(cells[t << tile_lg:(t + 1) << tile_lg], lgS, key, lo, hi)
                    assert (t << tile_lg) + pos == ref
                    assert hit == (cells[ref] == key)
