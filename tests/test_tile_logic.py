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


# ----------------------------------------------------------------------------------------------------------------------------
# the leaf re-lay of k_tile_merge (phases C, D1, D2, E) restated on the oracle's exported arrays, against the oracle's batch policy
# ----------------------------------------------------------------------------------------------------------------------------
def _tile_relay(e, parts, keys, vals, spread_dest, mn0, mx0):
    """One orientation, one batch, the way csrc/tile.cuh applies it — for batches that create no partition and whose touched
    leaves all stay inside their own bounds (otherwise None: those leaves go to the density tree / window kernels, which are
    not restated here).  e = oracle export (tag, key, val, semaphores 1-based, col_keys, col_live)."""
    tag, key, val = e["tag"].copy(), e["key"].copy(), e["val"].copy()
    cap, S = len(tag), int(e["segment_capacity"])
    sem = e["semaphores"].astype(np.int64) - 1
    live_slots = [s for s in range(len(sem)) if e["col_live"][s]]
    slot_of = {int(e["col_keys"][s]): s for s in live_slots}
    order = sorted(live_slots, key=lambda s: sem[s])
    span_end = {s: (sem[order[i + 1]] if i + 1 < len(order) else cap) for i, s in enumerate(order)}
    # last writer wins per (partition, key), arrival order
    last = {}
    for arr, (p, k, v) in enumerate(zip(parts, keys, vals)):
        if int(p) not in slot_of:
            return None
        last[(slot_of[int(p)], int(k))] = (arr, float(v))
    per_leaf = {}
    for (s, k), (arr, v) in last.items():
        ps, pe = int(sem[s]), int(span_end[s])
        idx = ps + np.flatnonzero(tag[ps:pe])          # stored cells of the span, the semaphore (key 0) first
        j = int(np.searchsorted(key[idx], k, side="right")) - 1
        pos = int(idx[j])
        per_leaf.setdefault(pos // S, []).append((pos, key[pos] == k and pos != ps, k, v))
    new_sem = sem.copy()
    for leaf, ops in per_leaf.items():
        c0 = leaf * S
        dele = {pos for pos, hit, k, v in ops if hit and v == 0.0}
        for pos, hit, k, v in ops:
            if hit and v != 0.0:
                val[pos] = v                                  # writes.jl:16-19
        ins = sorted((pos, k, v) for pos, hit, k, v in ops if not hit and v != 0.0)   # by (predecessor cell, key)
        surv = [q for q in range(S) if tag[c0 + q] and (c0 + q) not in dele]
        m = len(surv) + len(ins)
        if dele or ins:
            if m < mn0 or m > mx0:
                return None
        if not ins:
            for pos in dele:                                  # writes.jl:65-68
                tag[pos] = 0
            continue
        # merged order: an insert goes after the survivors up to its predecessor cell and after the inserts ordered before it
        items = [((q, 0, 0), int(key[c0 + q]), float(val[c0 + q])) for q in surv]
        items += [((pos - c0, 1, k), int(k), float(v)) for pos, k, v in ins]
        items.sort(key=lambda it: it[0])
        tag[c0:c0 + S] = 0
        for r, (_, k, v) in enumerate(items):                 # pack! + spread! (moves.jl:94-172)
            d = c0 + spread_dest(S, m, r)
            tag[d], key[d], val[d] = 1, k, v
            if k == 0:
                new_sem[int(v) - 1] = d                       # moves.jl:160-166
    return tag, key, val, new_sem + 1


@pytest.mark.parametrize("seed", range(6))
def test_leaf_relay_of_the_tile_kernel_equals_the_batch_policy(seed):
    import ctypes as C

    import dsa_b200 as D
    L = D.lib()
    L.dsa_spread_dest.restype = C.c_int64
    spread_dest = lambda c, m, r: int(L.dsa_spread_dest(C.c_int64(c), C.c_int64(m), C.c_int64(r)))
    rng = np.random.default_rng(2000 + seed)
    m, n = [(60, 50), (300, 40), (40, 400), (500, 500), (200, 900), (1000, 30)][seed]
    nnz0 = int(rng.integers(2000, 9000))
    I, J = rng.integers(1, m + 1, nnz0), rng.integers(1, n + 1, nnz0)
    V = rng.integers(1, 100, nnz0).astype(float)
    pol = O.Matrix(I, J, V, m=m, n=n)
    checked = 0
    for rnd in range(6):
        nb = int(rng.integers(20, 300))
        src = rng.integers(0, nnz0, nb)
        I2 = np.where(rng.random(nb) < 0.5, I[src], rng.choice(I, nb))     # existing rows and columns only
        J2 = np.where(rng.random(nb) < 0.5, J[src], rng.choice(J, nb))
        V2 = np.where(rng.random(nb) < 0.4, 0.0, rng.integers(1, 100, nb).astype(float))
        I2[:4], J2[:4] = I2[4:8], J2[4:8]                                   # repeated keys: last writer wins
        before = [pol.export(w) for w in (0, 1)]
        pol.set_batch_policy(I2, J2, V2)
        for w in (0, 1):
            e = before[w]
            mn, mx = np.zeros(40, np.int64), np.zeros(40, np.int64)
            assert L.dsa_level_bounds(C.c_int64(int(e["segment_capacity"])), C.c_int64(int(e["height"])),
                                      mn.ctypes.data_as(C.c_void_p), mx.ctypes.data_as(C.c_void_p)) == 0
            parts, keys = (J2, I2) if w == 0 else (I2, J2)                  # matrix.jl:53-59
            got = _tile_relay(e, parts, keys, V2, spread_dest, int(mn[0]), int(mx[0]))
            if got is None:
                continue   # a leaf left its bounds: the density tree / window kernels take over (not restated here)
            after = pol.export(w)
            tag, key, val, sem = got
            assert np.array_equal(tag, after["tag"]), (seed, rnd, w)
            mk = tag.astype(bool)
            assert np.array_equal(key[mk], after["key"][mk]) and np.array_equal(val[mk], after["val"][mk])
            lv = after["col_live"].astype(bool)
            assert np.array_equal(sem[lv], after["semaphores"][lv])
            checked += 1
    assert checked >= 3, "the batches are meant to stay inside the leaves' bounds most of the time"
